"""ctypes binding of ``libterran_b200.so`` (C ABI in ``include/terran_b200.h``).

The library is loaded lazily on first use and the load fails loudly: there is
no Python/CPU fallback for any of the entry points.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libterran_b200.so')

#: Every symbol ``include/terran_b200.h`` declares.
SYMBOLS = (
    'tr_version', 'tr_last_error', 'tr_init',
    'tr_net_create', 'tr_net_destroy', 'tr_net_set_mode', 'tr_net_run', 'tr_net_buffer',
    'tr_net_export_nchw', 'tr_net_export_nchw_f32', 'tr_net_stats', 'tr_net_set_profile',
    'tr_net_profile',
    'tr_program_build', 'tr_program_destroy', 'tr_program_info', 'tr_program_copy',
    'tr_net_create_from_program',
    'tr_retinaface_create', 'tr_retinaface_forward', 'tr_arcface_create', 'tr_arcface_forward',
    'tr_openpose_create', 'tr_openpose_forward', 'tr_model_net', 'tr_model_destroy',
    'tr_conv2d', 'tr_sepconv2d',
    'tr_detect_workspace_bytes', 'tr_retinaface_decode_nms', 'tr_retinaface_detect',
    'tr_l2_normalize', 'tr_face_align', 'tr_face_similarity',
    'tr_face_letterbox_workspace_bytes', 'tr_face_letterbox', 'tr_resample_table',
    'tr_pose_workspace_bytes', 'tr_openpose_parse', 'tr_bicubic_table',
    'tr_resize_bilinear_u8',
)

TR_OP_STEM, TR_OP_CONV, TR_OP_DWCONV, TR_OP_MAXPOOL, TR_OP_COPY, TR_OP_VIEW, TR_OP_SEPCONV = range(7)
TR_ENGINE_AUTO, TR_ENGINE_MMA = 0, 1
TR_ACT_NONE, TR_ACT_RELU, TR_ACT_PRELU = range(3)
TR_SYNC_FORK, TR_SYNC_JOIN = 1, 2
TR_PEAK_CAP, TR_CAND_CAP, TR_HUMAN_CAP = 512, 4096, 128


class BufferDesc(C.Structure):
    _fields_ = [('channels', C.c_int32), ('is_f32', C.c_int32)]


class OpDesc(C.Structure):
    _fields_ = [
        ('type', C.c_int32),
        ('in_', C.c_int32), ('in_coff', C.c_int32), ('in_c', C.c_int32),
        ('out', C.c_int32), ('out_coff', C.c_int32), ('out_c', C.c_int32),
        ('out2', C.c_int32), ('out2_coff', C.c_int32),
        ('res', C.c_int32), ('res_coff', C.c_int32), ('res_up2', C.c_int32),
        ('k', C.c_int32), ('stride', C.c_int32), ('pad', C.c_int32), ('act', C.c_int32),
        ('cout_pad', C.c_int32),
        ('cin_real', C.c_int32), ('cout_real', C.c_int32),
        ('force_direct', C.c_int32), ('lane', C.c_int32), ('sync', C.c_int32),
        ('w_off', C.c_int64), ('scale_off', C.c_int64), ('shift_off', C.c_int64),
        ('slope_off', C.c_int64), ('scale2_off', C.c_int64), ('shift2_off', C.c_int64),
        ('in_scale', C.c_float), ('in_shift', C.c_float),
        ('dw_w_off', C.c_int64), ('dw_scale_off', C.c_int64), ('dw_shift_off', C.c_int64),
        ('dw_w16_off', C.c_int64),
        ('engine', C.c_int32), ('groups', C.c_int32),
        ('shift9_off', C.c_int64),
    ]


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded library (built by ``terran_b200.build`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f'{LIB_PATH} is missing: run `python -m terran_b200.build` '
                '(there is no fallback path)')
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
        L.tr_version.restype = i32
        L.tr_last_error.restype = C.c_char_p
        L.tr_init.argtypes = [i32]
        L.tr_net_create.argtypes = [C.POINTER(BufferDesc), i32, C.POINTER(OpDesc), i32, vp,
                                    C.c_size_t, C.POINTER(vp)]
        L.tr_net_destroy.argtypes = [vp]
        L.tr_net_destroy.restype = None
        L.tr_net_set_mode.argtypes = [vp, i32]
        L.tr_net_run.argtypes = [vp, vp, i32, i32, i32, i64, i64, i64, i64, vp]
        L.tr_net_buffer.argtypes = [vp, i32, C.POINTER(vp)] + [C.POINTER(i32)] * 4
        L.tr_net_export_nchw.argtypes = [vp, i32, i32, i32, vp, vp]
        L.tr_net_export_nchw_f32.argtypes = [vp, i32, i32, i32, vp, i32, vp]
        L.tr_net_stats.argtypes = [vp, C.POINTER(f64), C.POINTER(i32), C.POINTER(i32)]
        L.tr_net_set_profile.argtypes = [vp, i32]
        L.tr_net_profile.argtypes = [vp, C.POINTER(f32), C.POINTER(i32), C.POINTER(f64), i32,
                                     C.POINTER(i32)]
        L.tr_program_build.argtypes = [C.c_char_p, vp, C.c_size_t, i32, C.POINTER(vp)]
        L.tr_program_destroy.argtypes = [vp]
        L.tr_program_destroy.restype = None
        L.tr_program_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_int32)]
        L.tr_program_copy.argtypes = [vp, C.POINTER(BufferDesc), C.POINTER(OpDesc), vp]
        L.tr_net_create_from_program.argtypes = [vp, C.POINTER(vp)]
        for name in ('tr_retinaface_create', 'tr_arcface_create', 'tr_openpose_create'):
            getattr(L, name).argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.tr_retinaface_forward.argtypes = [vp, vp, i32, i32, i32, f32, f64, i32, vp, vp, vp]
        L.tr_arcface_forward.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        L.tr_openpose_forward.argtypes = [vp, vp, i32, i32, i32, f64, vp, vp, vp, vp, vp]
        L.tr_model_net.argtypes = [vp]
        L.tr_model_net.restype = vp
        L.tr_model_destroy.argtypes = [vp]
        L.tr_model_destroy.restype = None
        L.tr_conv2d.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, i32,
                                i32, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, i32,
                                C.POINTER(f32), vp]
        L.tr_sepconv2d.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32,
                                   i32, i32, vp, i32, i32, vp, i32, i32, C.POINTER(f32), vp]
        L.tr_detect_workspace_bytes.argtypes = [i32, i32, i32]
        L.tr_detect_workspace_bytes.restype = C.c_size_t
        L.tr_retinaface_decode_nms.argtypes = [C.POINTER(vp), i32, i32, i32, f32, f64, i32, vp, vp,
                                               vp, vp, vp]
        L.tr_retinaface_detect.argtypes = [vp, C.POINTER(i32), f32, f64, i32, vp, vp, vp, vp, vp]
        L.tr_l2_normalize.argtypes = [vp, vp, i32, i32, vp]
        L.tr_face_align.argtypes = [vp, i32, i32, vp, vp, i32, vp, i32, vp]
        L.tr_face_similarity.argtypes = [vp, vp, i32, i32, f32, i32, vp, vp, vp, vp]
        L.tr_pose_workspace_bytes.argtypes = [i32]
        L.tr_pose_workspace_bytes.restype = C.c_size_t
        L.tr_openpose_parse.argtypes = [vp, vp, i32, i32, i32, f64, vp, vp, vp, vp, vp, vp]
        L.tr_face_letterbox_workspace_bytes.argtypes = [vp, i32, i32]
        L.tr_face_letterbox_workspace_bytes.restype = C.c_size_t
        L.tr_face_letterbox.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
        L.tr_resample_table.argtypes = [i32, i32, vp, vp]
        L.tr_bicubic_table.argtypes = [C.POINTER(f32)]
        L.tr_bicubic_table.restype = None
        L.tr_resize_bilinear_u8.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp]
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise NativeError(lib().tr_last_error().decode())


_initialised = set()


def init(device_index=0):
    if device_index not in _initialised:
        check(lib().tr_init(int(device_index)))
        _initialised.add(device_index)


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class nvtx_range:
    """``with nvtx_range('detect:net'):`` — an NVTX range around a pipeline stage (shows up in
    nsys / ncu --nvtx timelines next to the per-op ranges the library pushes itself)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        import torch
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        import torch
        torch.cuda.nvtx.range_pop()
