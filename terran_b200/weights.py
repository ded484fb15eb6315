"""Weight ingestion: reference ``state_dict`` -> layer program + packed blob.

Reads the checkpoint key layout of the reference models (SURVEY.md Appendix C;
``retinaface/model.py``, ``arcface/model.py``, ``openpose/model.py``), folds
every BatchNorm that FOLLOWS a convolution into an fp32 per-channel
scale/shift applied in the kernel epilogue, repacks filters to the
``[cout][kh][kw][cin]`` fp16 layout the tcgen05 kernel's TMA descriptors
expect, and emits the op list the native executor (``csrc/net.cu``) runs.

A BatchNorm that PRECEDES a zero-padded convolution (ArcFace ``Unit.body[0]``,
and the ``(x-127.5)*0.0078125`` input affine) cannot be folded into that
convolution — padding is applied after the affine — so it is applied by the
producer instead: the previous convolution's epilogue emits a second,
normalised tensor (``out2``), and the u8 stem kernels apply the input affine to
in-bounds taps only.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .synth import ARCFACE_CHANNELS, ARCFACE_UNITS, OPENPOSE_TRUNK, RETINAFACE_SCALES, \
    openpose_stage_layers


def _r(x, m):
    return (x + m - 1) // m * m


class Program:
    """Builder for the (buffers, ops, blob) triple ``tr_net_create`` takes."""

    def __init__(self):
        self.blob = bytearray()
        self.buffers = []
        self.ops = []

    # -- blob
    def add(self, arr, dtype):
        arr = np.ascontiguousarray(np.asarray(arr, dtype=dtype))
        pad = (-len(self.blob)) % 16
        self.blob += b'\0' * pad
        off = len(self.blob)
        self.blob += arr.tobytes()
        return off

    def add_vec(self, t, n_pad):
        if t is None:
            return -1
        v = np.zeros(n_pad, np.float32)
        t = np.asarray(t, np.float32)
        v[:len(t)] = t
        return self.add(v, np.float32)

    # -- buffers
    def buffer(self, channels, f32=False):
        self.buffers.append((int(channels), int(f32)))
        return len(self.buffers) - 1

    # -- ops
    def _op(self, **kw):
        d = dict(type=0, in_=-1, in_coff=0, in_c=0, out=-1, out_coff=0, out_c=0, out2=-1,
                 out2_coff=0, res=-1, res_coff=0, res_up2=0, k=1, stride=1, pad=0, act=0,
                 cout_pad=0, cin_real=0, cout_real=0, force_direct=0, lane=0, sync=0, w_off=-1, scale_off=-1, shift_off=-1, slope_off=-1,
                 scale2_off=-1, shift2_off=-1, in_scale=1.0, in_shift=0.0,
                 dw_w_off=-1, dw_scale_off=-1, dw_shift_off=-1, dw_w16_off=-1, engine=0, groups=0,
                 shift9_off=-1)
        d.update(kw)
        self.ops.append(nat.OpDesc(**d))

    def conv(self, w, scale, shift, in_, out, *, in_coff=0, in_map=None, cin_pad=None,
             out_coff=0, k=None, stride=1, pad=None, act=nat.TR_ACT_NONE, slope=None,
             res=-1, res_coff=0, res_up2=0, out2=-1, scale2=None, shift2=None,
             force_direct=0, lane=0, sync=0, engine=nat.TR_ENGINE_AUTO, shift9=None, groups=0):
        """w: (cout, cin, k, k) fp32 tensor.  ``in_map[c]`` is the position of
        reference input channel c inside the (padded) input view.  ``shift9``: (9, cout)
        border-class shifts replacing ``shift`` (see ``pre_bn_fold``)."""
        w = w.detach().float().numpy()
        cout, cin, kh, kw_ = w.shape
        k = kh if k is None else k
        pad = k // 2 if pad is None else pad
        if cin_pad is None:
            cin_pad = _r(cin, 8) if cin <= 8 else _r(cin, 16)
        in_map = np.arange(cin) if in_map is None else np.asarray(in_map)
        cout_pad = _r(cout, 16)
        packed = np.zeros((cout_pad, kh, kw_, cin_pad), np.float16)
        packed[:cout, :, :, in_map] = w.transpose(0, 2, 3, 1).astype(np.float16)
        self._op(type=nat.TR_OP_CONV, in_=in_, in_coff=in_coff, in_c=cin_pad, out=out,
                 out_coff=out_coff, out_c=_r(cout, 8), out2=out2, res=res, res_coff=res_coff,
                 res_up2=res_up2, k=k, stride=stride, pad=pad, act=act, cout_pad=cout_pad,
                 cin_real=cin, cout_real=cout, force_direct=force_direct, lane=lane, sync=sync,
                 engine=engine, groups=groups, w_off=self.add(packed, np.float16),
                 scale_off=self.add_vec(scale, cout_pad), shift_off=self.add_vec(shift, cout_pad),
                 slope_off=self.add_vec(slope, cout_pad),
                 scale2_off=self.add_vec(scale2, cout_pad), shift2_off=self.add_vec(shift2, cout_pad),
                 shift9_off=self.add_mat(shift9, cout_pad))

    def add_mat(self, m, n_pad):
        if m is None:
            return -1
        m = np.asarray(m, np.float32)
        v = np.zeros((m.shape[0], n_pad), np.float32)
        v[:, :m.shape[1]] = m
        return self.add(v, np.float32)

    def stem(self, w, scale, shift, out, *, stride, act, slope=None, in_scale=1.0, in_shift=0.0,
             out2=-1, scale2=None, shift2=None):
        w = w.detach().float().numpy()            # (cout, 3, 3, 3) -> [cout][kh][kw][c]
        cout = w.shape[0]
        self._op(type=nat.TR_OP_STEM, out=out, out_c=cout, out2=out2, k=3, stride=stride, pad=1,
                 act=act, cout_pad=cout, in_scale=in_scale, in_shift=in_shift,
                 w_off=self.add(w.transpose(0, 2, 3, 1), np.float32),
                 scale_off=self.add_vec(scale, cout), shift_off=self.add_vec(shift, cout),
                 slope_off=self.add_vec(slope, cout),
                 scale2_off=self.add_vec(scale2, cout), shift2_off=self.add_vec(shift2, cout))

    def dwconv(self, w, scale, shift, in_, out, *, stride):
        w = w.detach().float().numpy()            # (C, 1, 3, 3) -> [kh][kw][C]
        c = w.shape[0]
        self._op(type=nat.TR_OP_DWCONV, in_=in_, in_c=c, out=out, out_c=c, k=3, stride=stride,
                 pad=1, act=nat.TR_ACT_RELU, cout_pad=c,
                 w_off=self.add(w[:, 0].transpose(1, 2, 0), np.float32),
                 scale_off=self.add_vec(scale, c), shift_off=self.add_vec(shift, c))

    def sepconv(self, dw_w, dw_scale, dw_shift, w, scale, shift, in_, out, *, stride,
                act=nat.TR_ACT_RELU):
        """Depthwise 3x3 (pad 1, ``stride``) + BN + ReLU fused with the 1x1 conv + BN (+act)
        that consumes it.  dw_w: (C, 1, 3, 3); w: (cout, C, 1, 1)."""
        dw_w = dw_w.detach().float().numpy()
        w = w.detach().float().numpy()
        cout, cin = w.shape[:2]
        cin_pad = _r(cin, 8) if cin <= 8 else _r(cin, 16)
        cout_pad = _r(cout, 16)
        packed = np.zeros((cout_pad, cin_pad), np.float16)
        packed[:cout, :cin] = w[:, :, 0, 0].astype(np.float16)
        dwp = np.zeros((3, 3, cin_pad), np.float32)
        dwp[:, :, :cin] = dw_w[:, 0].transpose(1, 2, 0)
        self._op(type=nat.TR_OP_SEPCONV, in_=in_, in_c=cin_pad, out=out, out_c=_r(cout, 8), k=3,
                 stride=stride, pad=1, act=act, cout_pad=cout_pad, cin_real=cin, cout_real=cout,
                 w_off=self.add(packed, np.float16),
                 scale_off=self.add_vec(scale, cout_pad), shift_off=self.add_vec(shift, cout_pad),
                 dw_w_off=self.add(dwp, np.float32), dw_w16_off=self.add(dwp, np.float16),
                 dw_scale_off=self.add_vec(dw_scale, cin_pad), dw_shift_off=self.add_vec(dw_shift, cin_pad))

    def maxpool(self, in_, out, channels):
        self._op(type=nat.TR_OP_MAXPOOL, in_=in_, in_c=channels, out=out, out_c=channels, k=2,
                 stride=2)

    def copy(self, in_, in_coff, out, out_coff, channels):
        self._op(type=nat.TR_OP_COPY, in_=in_, in_coff=in_coff, in_c=channels, out=out,
                 out_coff=out_coff, out_c=channels)

    def view(self, in_, out):
        self._op(type=nat.TR_OP_VIEW, in_=in_, out=out)


def pack_state_dict(sd):
    """Serialise a reference ``state_dict`` into the flat "TRSD" blob the C ABI takes
    (``include/terran_b200.h``): named float32 / int64 tensors, 8-byte aligned data."""
    import struct
    out = bytearray(b'TRSD' + struct.pack('<II', 1, len(sd)))
    for name, t in sd.items():
        t = t.detach().cpu()
        if t.dtype == torch.int64:
            code, arr = 1, t.numpy()
        else:
            code, arr = 0, t.float().numpy()
        arr = np.ascontiguousarray(arr)
        nb = name.encode()
        out += struct.pack('<H', len(nb)) + nb + struct.pack('<BB', code, arr.ndim)
        out += struct.pack(f'<{arr.ndim}q', *arr.shape) + struct.pack('<Q', arr.nbytes)
        out += b'\0' * ((-len(out)) % 8)
        out += arr.tobytes()
    return bytes(out)


def native_program(model, sd, flags=0):
    """Build the layer program of ``model`` ('retinaface' | 'arcface' | 'openpose') from a
    reference ``state_dict`` with the library's own builder (``csrc/program.cu`` — the same code a
    non-Python host reaches through ``tr_*_create``).  Host-only: works without a GPU.
    Returns (Program, roles[8])."""
    blob = pack_state_dict(sd)
    buf = np.frombuffer(blob, dtype=np.uint8)            # 8-byte aligned host copy
    handle = C.c_void_p()
    nat.check(nat.lib().tr_program_build(model.encode(), C.c_void_p(buf.ctypes.data), len(blob),
                                         int(flags), C.byref(handle)))
    try:
        nb, no, nbytes = C.c_int(), C.c_int(), C.c_size_t()
        roles = (C.c_int32 * 8)()
        nat.check(nat.lib().tr_program_info(handle, C.byref(nb), C.byref(no), C.byref(nbytes), roles))
        bufs = (nat.BufferDesc * nb.value)()
        ops = (nat.OpDesc * no.value)()
        data = (C.c_uint8 * nbytes.value)()
        nat.check(nat.lib().tr_program_copy(handle, bufs, ops, data))
    finally:
        nat.lib().tr_program_destroy(handle)
    P = Program()
    P.buffers = [(b.channels, b.is_f32) for b in bufs]
    P.ops = list(ops)
    P.blob = bytearray(bytes(data))
    return P, list(roles)


def retinaface_program(sd, fused=None):
    """Program + roles for reference ``RetinaFace`` (retinaface/model.py:319-341).  The stem
    reads the frame in MODEL channel order (BGR); callers with RGB memory pass a pointer to
    channel 2 and a channel stride of -1.  ``fused`` (default: env ``TRB_RETINA_FUSED`` != 0):
    every depthwise 3x3 of the backbone runs inside the 1x1 conv that consumes it and the
    small-channel convs go to the warp-level mma.sync kernel; otherwise one op per layer."""
    import os
    if fused is None:
        fused = os.environ.get('TRB_RETINA_FUSED', '1') != '0'
    P, roles = native_program('retinaface', sd, flags=0 if fused else 1)
    return P, {'heads': roles[:3]}


def arcface_program(sd, units=None):
    """Program for reference ``FaceResNet100`` (arcface/model.py:38-97), input in model channel
    order (BGR).  The depth (units per stage) is read from the checkpoint's keys.  Every
    ``Unit`` starts with a BatchNorm in front of a zero-padded 3x3 conv (model.py:11-14); it is
    folded into that conv exactly — scale into the filters, shift into nine border-class shift
    vectors — so the residual stream is the only tensor a unit reads and writes."""
    P, roles = native_program('arcface', sd)
    return P, {'embedding': roles[0]}


#: position of reference concat channel c (cat[PAF 38, heat 19, trunk 128]) in
#: the padded 192-channel buffer [PAF 0..37 | pad | heat 40..58 | pad | trunk 64..191]
OPENPOSE_CAT_MAP = np.concatenate([np.arange(38), 40 + np.arange(19), 64 + np.arange(128)])


def openpose_program(sd):
    """Program for reference ``BodyPoseModel`` (openpose/model.py:27-141).  The frame is read as
    stored (RGB, no flip: wrapper.py:116-122)."""
    P, roles = native_program('openpose', sd)
    return P, {'maps': roles[0], 'paf_coff': roles[1], 'heat_coff': roles[2]}


def program_traffic(program, N, H, W):
    """Algorithmic HBM traffic of one run of ``program`` on an (N, H, W, 3) uint8 batch: per op
    the bytes it must read (input view, residual, filters + per-channel parameters) and write
    (output views), each tensor counted once per op that touches it — the model behind the
    HBM roofline of the RetinaFace stack (42.6 MB/frame layer-wise,
    20.1 MB/frame for a block-fused plan at 416x739).  Shapes follow the executor's inference
    (``csrc/net.cu::build_plan``).  Returns (total_bytes, [(op_index, read, written)])."""
    dims = [None] * len(program.buffers)
    per_op = []
    for i, d in enumerate(program.ops):
        if d.in_ < 0:
            n, h, w = N, H, W
        else:
            n, h, w = dims[d.in_]
        esz = lambda b: 4 if program.buffers[b][1] else 2
        if d.type in (nat.TR_OP_STEM, nat.TR_OP_CONV, nat.TR_OP_DWCONV, nat.TR_OP_SEPCONV):
            oh, ow = (h + 2 * d.pad - d.k) // d.stride + 1, (w + 2 * d.pad - d.k) // d.stride + 1
            dims[d.out] = (n, oh, ow)
            if d.out2 >= 0:
                dims[d.out2] = (n, oh, ow)
            rd = n * h * w * (3 if d.type == nat.TR_OP_STEM else d.in_c * esz(d.in_))
            if d.type == nat.TR_OP_STEM:
                rd += d.out_c * 27 * 4
            elif d.type == nat.TR_OP_CONV:
                rd += d.cout_pad * d.k * d.k * d.in_c * 2
            elif d.type == nat.TR_OP_DWCONV:
                rd += 9 * d.in_c * 4
            else:
                rd += 9 * d.in_c * 2 + d.cout_pad * d.in_c * 2 + 2 * d.in_c * 4
            rd += 2 * max(d.cout_pad, d.out_c) * 4                      # scale / shift
            if d.res >= 0:
                rn, rh, rw = dims[d.res]
                rd += n * oh * ow * d.out_c * 2 if not d.res_up2 else rn * rh * rw * d.out_c * 2
            wr = n * oh * ow * d.out_c * esz(d.out)
            if d.out2 >= 0:
                wr += n * oh * ow * d.out_c * 2
        elif d.type == nat.TR_OP_MAXPOOL:
            dims[d.out] = (n, h // 2, w // 2)
            rd, wr = n * h * w * d.in_c * 2, n * (h // 2) * (w // 2) * d.out_c * 2
        elif d.type == nat.TR_OP_COPY:
            dims[d.out] = (n, h, w)
            rd = wr = n * h * w * d.in_c * 2
        else:                                                            # view: no traffic
            dims[d.out] = (n, 1, 1)
            rd = wr = 0
        per_op.append((i, rd, wr))
    return sum(r + w for _, r, w in per_op), per_op


# ----------------------------------------------------------------------- Net

class Net:
    """A compiled layer program living on one CUDA device."""

    def __init__(self, program, device_index=0):
        nat.init(device_index)
        self.program = program
        bufs = (nat.BufferDesc * len(program.buffers))(*[
            nat.BufferDesc(c, f) for c, f in program.buffers])
        ops = (nat.OpDesc * len(program.ops))(*program.ops)
        blob = bytes(program.blob)
        handle = C.c_void_p()
        nat.check(nat.lib().tr_net_create(bufs, len(program.buffers), ops, len(program.ops),
                                          blob, len(blob), C.byref(handle)))
        self.handle = handle
        self.weight_bytes = len(blob)

    def __del__(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            try:
                nat.lib().tr_net_destroy(h)
            except Exception:
                pass

    def set_force_direct(self, flag):
        nat.check(nat.lib().tr_net_set_mode(self.handle, int(bool(flag))))

    def run(self, image, N, H, W, strides, ptr_offset=0):
        """image: torch uint8 CUDA tensor; strides = element strides (n,h,w,c)."""
        sn, sh, sw, sc = (int(v) for v in strides)
        nat.check(nat.lib().tr_net_run(self.handle, C.c_void_p(image.data_ptr() + ptr_offset),
                                       N, H, W, sn, sh, sw, sc, nat.current_stream_ptr()))

    def buffer_info(self, buf):
        ptr = C.c_void_p()
        n, h, w, c = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        nat.check(nat.lib().tr_net_buffer(self.handle, buf, C.byref(ptr), C.byref(n), C.byref(h),
                                          C.byref(w), C.byref(c)))
        return ptr.value, n.value, h.value, w.value, c.value

    def export_nchw(self, buf, coff, channels):
        _, n, h, w, _ = self.buffer_info(buf)
        out = torch.empty((n, channels, h, w), dtype=torch.float32, device='cuda')
        nat.check(nat.lib().tr_net_export_nchw(self.handle, buf, coff, channels,
                                               C.c_void_p(out.data_ptr()), nat.current_stream_ptr()))
        return out

    def export_nchw_f32(self, buf, coff, channels, softmax_pairs=False):
        _, n, h, w, _ = self.buffer_info(buf)
        out = torch.empty((n, channels, h, w), dtype=torch.float32, device='cuda')
        nat.check(nat.lib().tr_net_export_nchw_f32(
            self.handle, buf, coff, channels, C.c_void_p(out.data_ptr()), int(softmax_pairs),
            nat.current_stream_ptr()))
        return out

    def set_profile(self, flag):
        nat.check(nat.lib().tr_net_set_profile(self.handle, int(bool(flag))))

    def profile(self):
        """Per-op (ms, is_tc, algorithmic flops) of the last profiled run."""
        cap = len(self.program.ops)
        ms, tc, fl, n = (C.c_float * cap)(), (C.c_int32 * cap)(), (C.c_double * cap)(), C.c_int()
        nat.check(nat.lib().tr_net_profile(self.handle, ms, tc, fl, cap, C.byref(n)))
        return [(ms[i], bool(tc[i]), fl[i]) for i in range(n.value)]

    def stats(self):
        fl, tc, tot = C.c_double(), C.c_int(), C.c_int()
        nat.check(nat.lib().tr_net_stats(self.handle, C.byref(fl), C.byref(tc), C.byref(tot)))
        return {'tc_flops': fl.value, 'tc_launches': tc.value, 'launches': tot.value}
