"""ArcFace (IR-ResNet-100) model class — the plugin ``Recognition`` calls.

Same protocol as the reference's ``ArcFace`` (``terran/face/recognition/arcface/
wrapper.py:102-184``): ``cls(device=...)`` and ``call(images, faces_per_image=None)``
returning L2-normalised 512-d float32 embeddings, split per image when faces
are given.  The ResNet forward, the final FC + BatchNorm1d and the L2
normalisation (done on the host with sklearn upstream) run in the native library.

Face alignment: the 2x3 similarity per face is estimated in closed form (what
``skimage.transform.SimilarityTransform.estimate`` computes for five 2-D points) — on the device
from the detection rows in the pipeline (``tr_face_similarity``), on the host from the face dicts
in ``call`` — and the affine bilinear warp to the 112x112 crop runs on the GPU (``tr_face_align``,
bit-exact with the PIL warp upstream uses), for batches and for lists of differently sized images.
Faces given without landmarks are resized and centre-padded on the GPU too
(``tr_face_letterbox``, bit-exact with PIL's ``Image.resize``): there is no host image path.
"""
import ctypes as C

import numpy as np
import torch

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


def similarity_coefficients(landmarks, image_size=(112, 112)):
    """Vectorised ``alignment_coefficients`` for (F,5,2) landmarks -> (F,6) float64, in closed
    form: in 2-D the least-squares similarity needs no SVD — with centred landmarks p and
    template points q, ``a = sum p.q``, ``b = sum p x q``, rotation ``atan2(b, a)``, scale
    ``hypot(a, b) / sum |p|^2`` (= Umeyama's ``S @ d / var``, reflections included).  The same
    arithmetic runs on the device in ``tr_face_similarity``.  Pinned against the oracle's SVD
    Umeyama + ``np.linalg.inv`` (``oracle/align.py``) on 10 000 random faces
    (tests/test_host_logic.py::test_closed_form_similarity_matches_umeyama)."""
    template = LANDMARK_TEMPLATE.copy()
    if image_size[1] == 112:
        template[:, 0] += 8.0
    p = np.asarray(landmarks).astype(np.float32).astype(np.float64).reshape(-1, 5, 2)
    q = template.astype(np.float64)
    mp, mq = p.mean(1, keepdims=True), q.mean(0, keepdims=True)
    u, v = p - mp, (q - mq)[None]
    a = (u * v).sum((1, 2))
    b = (u[..., 0] * v[..., 1] - u[..., 1] * v[..., 0]).sum(1)
    var = (u * u).sum((1, 2))
    nrm = np.hypot(a, b)
    with np.errstate(divide='ignore', invalid='ignore'):
        s, c, sn = nrm / var, a / nrm, b / nrm
        t0 = mq[0, 0] - s * (c * mp[:, 0, 0] - sn * mp[:, 0, 1])
        t1 = mq[0, 1] - s * (sn * mp[:, 0, 0] + c * mp[:, 0, 1])
        out = np.stack([c / s, sn / s, -(c * t0 + sn * t1) / s,
                        -sn / s, c / s, -(-sn * t0 + c * t1) / s], axis=1)
    out[~((var > 0) & (nrm > 0))] = np.nan
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
        crops = crops.contiguous()          # element strides below assume a dense tensor
        n_real = crops.shape[0]
        if n_real % 8:
            # batches are padded to a multiple of 8 crops: the final FC then runs as a
            # (1, N/8, 8, C) map on the resident-patch kernel with its K split over all SMs
            pad = torch.zeros((8 - n_real % 8,) + tuple(crops.shape[1:]), dtype=crops.dtype, device=crops.device)
            crops = torch.cat([crops, pad], 0)
        N = crops.shape[0]
        S = self.image_side
        if layout == 'nhwc_rgb':
            assert tuple(crops.shape[1:]) == (S, S, 3)
            self.net.run(crops, N, S, S, (S * S * 3, S * 3, 3, -1), ptr_offset=2)
        else:
            assert tuple(crops.shape[1:]) == (3, S, S)
            self.net.run(crops, N, S, S, (3 * S * S, S, 1, S * S))
        raw = self.net.export_nchw_f32(self.roles['embedding'], 0, 512).reshape(N, 512)[:n_real]
        N = n_real
        if not normalise:
            return raw.contiguous()
        raw = raw.contiguous()
        out = torch.empty_like(raw)
        nat.check(nat.lib().tr_l2_normalize(C.c_void_p(raw.data_ptr()), C.c_void_p(out.data_ptr()),
                                            N, 512, nat.current_stream_ptr()))
        return out

    def align_device(self, frames, faces_per_image):
        """frames: CUDA uint8 (N,H,W,3) RGB; faces: the detection dicts.  Warps every
        face to the (F,3,112,112) BGR crop on the GPU (``tr_face_align``, bit-exact
        with the PIL warp of the host path); the 2x3 matrices are estimated on the
        host (five points per face)."""
        S = self.image_side
        index = [i for i, faces in enumerate(faces_per_image) for _ in faces]
        F = len(index)
        coefs = similarity_coefficients(
            np.stack([face['landmarks'] for faces in faces_per_image for face in faces])
            if F else np.zeros((0, 5, 2)), (S, S))
        out = torch.empty((F, 3, S, S), dtype=torch.uint8, device=frames.device)
        if F:
            N, H, W, _ = frames.shape
            coef = torch.from_numpy(np.asarray(coefs, np.float64)).to(frames.device)
            idx = torch.from_numpy(np.asarray(index, np.int32)).to(frames.device)
            nat.check(nat.lib().tr_face_align(
                C.c_void_p(frames.data_ptr()), H, W, C.c_void_p(coef.data_ptr()),
                C.c_void_p(idx.data_ptr()), F, C.c_void_p(out.data_ptr()), S,
                nat.current_stream_ptr()))
        return out

    def letterbox_device(self, images):
        """Faces without landmarks (reference ``preprocess_face_no_landmarks`` :75-99): a list of
        (h,w,3) uint8 RGB arrays of any size -> (n,3,112,112) uint8 BGR CUDA crops, each image
        resized to longer side 112 (PIL's default antialiased bicubic, bit-exact) and centred on
        a zero canvas — one upload of all pixels, two kernels (``tr_face_letterbox``)."""
        S = self.image_side
        images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        for im in images:
            if im.ndim != 3 or im.shape[2] != 3:
                raise ValueError(f'expected (h, w, 3) uint8 RGB images, got shape {im.shape}')
            if min(int(im.shape[1] * (S / max(im.shape[:2]))), int(im.shape[0] * (S / max(im.shape[:2])))) < 1:
                raise ValueError('height and width must be > 0')        # (PIL's own error)
        n = len(images)
        if n == 0:
            return torch.empty((0, 3, S, S), dtype=torch.uint8, device=torch.device('cuda', self.device_index))
        sizes = np.array([im.shape[:2] for im in images], np.int32).reshape(n, 2)
        nbytes = np.array([im.size for im in images], np.int64)
        offsets = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
        blob = torch.empty(int(nbytes.sum()), dtype=torch.uint8).pin_memory()
        flat = blob.numpy()
        for im, off in zip(images, offsets):
            flat[off:off + im.size] = im.reshape(-1)
        dev = torch.device('cuda', self.device_index)
        pixels = blob.to(dev, non_blocking=True)
        need = nat.lib().tr_face_letterbox_workspace_bytes(sizes.ctypes.data, n, S)
        if need == 0:
            raise nat.NativeError(nat.lib().tr_last_error().decode())
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.empty((n, 3, S, S), dtype=torch.uint8, device=dev)
        nat.check(nat.lib().tr_face_letterbox(
            C.c_void_p(pixels.data_ptr()), offsets.ctypes.data, sizes.ctypes.data, n, S,
            C.c_void_p(ws.data_ptr()), C.c_void_p(out.data_ptr()), nat.current_stream_ptr()))
        return out

    def embed_detections(self, frames, pending, scale):
        """Device-resident detect -> align -> embed: ``pending`` is the ``PendingDetections`` of
        ``RetinaFace.detect_async`` on (the resized copy of) ``frames`` (CUDA uint8 (N,H,W,3)),
        ``scale`` the detector's resize factor.  The 5-point similarity of every face is fitted on
        the device from the detection rows (``tr_face_similarity``), the crops are warped on the
        device (``tr_face_align``) and embedded; the landmarks never visit the host — only the
        per-frame face COUNT does, because the embedding batch size has to be known to launch.
        Returns (features (F,512) fp32 CUDA, counts (N,) numpy); equals
        ``call(frames, faces)`` with the faces ``Detection`` returns."""
        S = self.image_side
        pending.done.synchronize()                        # the counts are in pinned host memory
        counts = np.minimum(pending.slot['count'].numpy(), pending.max_det).astype(np.int64)
        if pending.count_dev is None or int(counts.max(initial=0)) > pending.max_det:
            raise nat.NativeError('detections are not resident on the device')
        F = int(counts.sum())
        dev = frames.device
        if F == 0:
            return torch.empty((0, 512), dtype=torch.float32, device=dev), counts
        N, H, W, _ = frames.shape
        coef = torch.empty((F, 6), dtype=torch.float64, device=dev)
        idx = torch.empty(F, dtype=torch.int32, device=dev)
        total = torch.empty(1, dtype=torch.int32, device=dev)
        crops = torch.empty((F, 3, S, S), dtype=torch.uint8, device=dev)
        stream = nat.current_stream_ptr()
        nat.check(nat.lib().tr_face_similarity(
            C.c_void_p(pending.det_dev.data_ptr()), C.c_void_p(pending.count_dev.data_ptr()), N,
            pending.max_det, float(scale), F, C.c_void_p(coef.data_ptr()), C.c_void_p(idx.data_ptr()),
            C.c_void_p(total.data_ptr()), stream))
        nat.check(nat.lib().tr_face_align(
            C.c_void_p(frames.data_ptr()), H, W, C.c_void_p(coef.data_ptr()),
            C.c_void_p(idx.data_ptr()), F, C.c_void_p(crops.data_ptr()), S, stream))
        return self.embed_device(crops, 'nchw_bgr'), counts

    def call(self, images, faces_per_image=None):
        """Feature extraction (reference ``ArcFace.call`` :109-184)."""
        S = self.image_side
        # Uniform frame batch + faces: align on the GPU, no per-face host warp.
        batched = isinstance(images, (np.ndarray, torch.Tensor)) and images.ndim == 4
        if not batched and faces_per_image is not None and len(images) and all(
                isinstance(im, np.ndarray) and im.shape == images[0].shape for im in images):
            images, batched = np.stack(images, 0), True
        if batched and faces_per_image is not None:
            if not any(len(f) for f in faces_per_image):
                return [np.empty((0, 512)) for _ in images]
            from terran_b200.frames import to_device_u8
            with torch.cuda.device(self.device_index):
                frames = to_device_u8(images, self.device_index)
                crops = self.align_device(frames, faces_per_image)
                features = self.embed_device(crops, 'nchw_bgr').cpu().numpy()
            splits = np.cumsum(list(map(len, faces_per_image)))[:-1]
            return np.split(features, splits, axis=0)
        if faces_per_image is not None:
            # Differently sized images: each one is uploaded and its faces are warped on the
            # GPU (the same bit-exact ``tr_face_align`` as the batched path) — no host PIL warp.
            if not any(len(f) for f in faces_per_image):
                return [np.empty((0, 512)) for _ in images]
            from terran_b200.frames import to_device_u8
            with torch.cuda.device(self.device_index):
                crops = [self.align_device(to_device_u8(np.asarray(image)[None], self.device_index), [faces])
                         for image, faces in zip(images, faces_per_image) if len(faces)]
                features = self.embed_device(torch.cat(crops, 0), 'nchw_bgr').cpu().numpy()
            splits = np.cumsum(list(map(len, faces_per_image)))[:-1]
            return np.split(features, splits, axis=0)
        # No landmarks: the images ARE the faces, returned as one (n,512) array (:150-183).
        if not len(images):
            return [np.empty((0, 512)) for _ in images]      # (upstream: float64 empties, here [])
        with torch.cuda.device(self.device_index):
            if all(getattr(im, 'shape', None) == (S, S, 3) for im in images):
                # already 112x112: PIL's resize would be the identity; straight to the stem
                batch = np.ascontiguousarray(np.stack(images, 0))
                dev = torch.from_numpy(batch).pin_memory().to(f'cuda:{self.device_index}',
                                                              non_blocking=True)
                return self.embed_device(dev, 'nhwc_rgb').cpu().numpy()
            # any other size: resized + centre-padded on the GPU (``tr_face_letterbox``)
            return self.embed_device(self.letterbox_device(images), 'nchw_bgr').cpu().numpy()
