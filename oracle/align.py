"""ORACLE (test infrastructure, not product code): the reference's host-side face alignment,
``preprocess_face`` of ``terran/face/recognition/arcface/wrapper.py:22-72`` — the 5-point
similarity (``skimage.transform.SimilarityTransform.estimate``: skimage is an un-pinned
dependency of the reference and absent here, so ``umeyama_similarity`` restates its published
``_umeyama`` estimator, Umeyama 1991 — **parity-unpinned** against skimage itself; it is pinned
against exact synthetic similarities in ``tests/test_host_logic.py``), the inverse matrix and the
PIL ``Image.transform(AFFINE, BILINEAR)`` warp.  The product path computes the similarity in
closed form (``similarity_coefficients`` on the host, ``tr_face_similarity`` on the device) and
warps on the GPU (``tr_face_align``); both are tested against this file.
"""
import numpy as np
from PIL import Image

#: five-point template of the 112x96 ArcFace crop (wrapper.py:39-45); x is shifted by 8 for the
#: 112-wide crop (:47-48).
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


def alignment_coefficients(landmark, image_size=(112, 112)):
    """The 6 PIL ``AFFINE`` coefficients (first two rows of the inverse
    similarity landmarks -> template) of reference ``preprocess_face`` :39-61."""
    template = LANDMARK_TEMPLATE.copy()
    if image_size[1] == 112:
        template[:, 0] += 8.0
    T = umeyama_similarity(np.asarray(landmark).astype(np.float32), template)
    return np.linalg.inv(T)[0:-1, :].flatten()


def preprocess_face(image, landmark, image_size=(112, 112)):
    """Align one face with its 5 landmarks and return the (3,112,112) uint8 BGR
    crop (reference ``preprocess_face`` :22-72) — host path (PIL)."""
    coeffs = alignment_coefficients(landmark, image_size)
    warped = Image.fromarray(image).transform(
        size=(image_size[1], image_size[0]), method=Image.AFFINE, data=coeffs,
        resample=Image.BILINEAR, fillcolor=0)
    return np.array(warped).transpose([2, 0, 1])[::-1, ...]


