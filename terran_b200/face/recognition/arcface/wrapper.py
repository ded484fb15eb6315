"""ArcFace (IR-ResNet-100) model class — the plugin ``Recognition`` calls.

Same protocol as the reference's ``ArcFace`` (``terran/face/recognition/arcface/
wrapper.py:102-184``): ``cls(device=...)`` and ``call(images, faces_per_image=None)``
returning L2-normalised 512-d float32 embeddings, split per image when faces
are given.  The ResNet forward, the final FC + BatchNorm1d and the L2
normalisation (done on the host with sklearn upstream) run in the native library.

Face alignment stays on the host in this round (PIL affine warp exactly as
upstream; the similarity transform is the closed-form Umeyama estimate that
``skimage.transform.SimilarityTransform.estimate`` implements).
"""
import ctypes as C

import numpy as np
import torch
from PIL import Image

from terran_b200 import _native as nat
from terran_b200.checkpoint import get_checkpoint_path
from terran_b200.defaults import cuda_index, default_device
from terran_b200.weights import Net, arcface_program

CLASS_PATH = 'terran_b200.face.recognition.arcface.ArcFace'

#: five-point template of the 112x96 ArcFace crop (wrapper.py:39-45); x is
#: shifted by 8 for the 112-wide crop (:47-48).
LANDMARK_TEMPLATE = np.array([
    [30.2946, 51.6963], [65.5318, 51.5014], [48.0252, 71.7366],
    [33.5493, 92.3655], [62.7299, 92.2041]], dtype=np.float32)


def umeyama_similarity(src, dst):
    """Least-squares similarity (rotation, uniform scale, translation) mapping
    ``src`` points onto ``dst`` — Umeyama 1991, the estimator behind
    ``SimilarityTransform.estimate``.  Returns the 3x3 homogeneous matrix."""
    src = np.asarray(src, np.float64)
    dst = np.asarray(dst, np.float64)
    n, dim = src.shape
    mu_s, mu_d = src.mean(0), dst.mean(0)
    sc, dc = src - mu_s, dst - mu_d
    cov = dc.T @ sc / n
    d = np.ones(dim)
    if np.linalg.det(cov) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1)
    U, S, Vt = np.linalg.svd(cov)
    rank = np.linalg.matrix_rank(cov)
    if rank == 0:
        return np.full((dim + 1, dim + 1), np.nan)
    if rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(Vt) > 0:
            T[:dim, :dim] = U @ Vt
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ Vt
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ Vt
    scale = 1.0 / sc.var(0).sum() * (S @ d)
    T[:dim, dim] = mu_d - scale * (T[:dim, :dim] @ mu_s)
    T[:dim, :dim] *= scale
    return T


def preprocess_face(image, landmark, image_size=(112, 112)):
    """Align one face with its 5 landmarks and return the (3,112,112) uint8 BGR
    crop (reference ``preprocess_face`` :22-72)."""
    template = LANDMARK_TEMPLATE.copy()
    if image_size[1] == 112:
        template[:, 0] += 8.0
    T = umeyama_similarity(np.asarray(landmark).astype(np.float32), template)
    coeffs = np.linalg.inv(T)[0:-1, :].flatten()
    warped = Image.fromarray(image).transform(
        size=(image_size[1], image_size[0]), method=Image.AFFINE, data=coeffs,
        resample=Image.BILINEAR, fillcolor=0)
    return np.array(warped).transpose([2, 0, 1])[::-1, ...]


def preprocess_face_no_landmarks(image, image_side=112):
    """Resize to max side 112 and centre-pad (reference :75-99)."""
    face = Image.fromarray(image)
    scale = image_side / max(face.size[0], face.size[1])
    face = face.resize((int(face.size[0] * scale), int(face.size[1] * scale)))
    x_min = int((image_side - face.size[0]) / 2)
    y_min = int((image_side - face.size[1]) / 2)
    out = np.zeros((3, image_side, image_side), dtype=np.uint8)
    out[:, y_min:y_min + face.size[1], x_min:x_min + face.size[0]] = (
        np.asarray(face).transpose([2, 0, 1])[::-1, ...])
    return out


class ArcFace:

    def __init__(self, device=default_device, image_side=112, state_dict=None):
        self.device = device
        self.device_index = cuda_index(device)
        self.image_side = image_side
        if state_dict is None:
            state_dict = torch.load(get_checkpoint_path(CLASS_PATH), map_location='cpu')
        program, self.roles = arcface_program(state_dict)
        with torch.cuda.device(self.device_index):
            self.net = Net(program, self.device_index)

    # -- device stages --------------------------------------------------------
    def embed_device(self, crops, layout='nhwc_rgb', normalise=True):
        """crops: CUDA uint8, (N,112,112,3) RGB (``nhwc_rgb``) or the reference
        model input (N,3,112,112) BGR (``nchw_bgr``).  Returns (N,512) fp32 CUDA."""
        N = crops.shape[0]
        S = self.image_side
        if layout == 'nhwc_rgb':
            assert tuple(crops.shape[1:]) == (S, S, 3)
            self.net.run(crops, N, S, S, (S * S * 3, S * 3, 3, -1), ptr_offset=2)
        else:
            assert tuple(crops.shape[1:]) == (3, S, S)
            self.net.run(crops, N, S, S, (3 * S * S, S, 1, S * S))
        raw = self.net.export_nchw_f32(self.roles['embedding'], 0, 512).reshape(N, 512)
        if not normalise:
            return raw
        out = torch.empty_like(raw)
        nat.check(nat.lib().tr_l2_normalize(C.c_void_p(raw.data_ptr()), C.c_void_p(out.data_ptr()),
                                            N, 512, nat.current_stream_ptr()))
        return out

    def call(self, images, faces_per_image=None):
        """Feature extraction (reference ``ArcFace.call`` :109-184)."""
        S = self.image_side
        fast = faces_per_image is None and all(
            getattr(im, 'shape', None) == (S, S, 3) for im in images)
        if fast and len(images):
            batch = np.ascontiguousarray(np.stack(images, 0))
            layout = 'nhwc_rgb'
            splits = []
        else:
            pre = []
            if faces_per_image is not None:
                for image, faces in zip(images, faces_per_image):
                    for face in faces:
                        pre.append(preprocess_face(image, face['landmarks']))
                splits = np.cumsum(list(map(len, faces_per_image)))[:-1]
            else:
                for image in images:
                    pre.append(preprocess_face_no_landmarks(image, S))
                splits = []
            if not pre:
                # upstream returns float64 here (np.empty default dtype)
                return [np.empty((0, 512)) for _ in images]
            batch = np.ascontiguousarray(np.stack(pre, axis=0))
            layout = 'nchw_bgr'

        with torch.cuda.device(self.device_index):
            dev = torch.from_numpy(batch).pin_memory().to(f'cuda:{self.device_index}',
                                                          non_blocking=True)
            features = self.embed_device(dev, layout).cpu().numpy()
        per_image = np.split(features, splits, axis=0)
        if faces_per_image is None:
            per_image = per_image[0]
        return per_image
