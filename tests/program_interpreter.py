"""TEST INFRASTRUCTURE: a torch-CPU interpreter of the layer programs that
``terran_b200/weights.py`` emits for the native executor (``csrc/net.cu``).

It executes the (buffers, ops, blob) triple exactly as ``include/terran_b200.h``
specifies them — channel-slice views, folded BatchNorm, fused heads, the
depthwise+1x1 ``TR_OP_SEPCONV`` pairs, residual / up-sampled residual, second
outputs, flatten views — with fp32 arithmetic on the fp16-rounded weights of
the blob.  Comparing its result with the oracle's ``nn.Module`` restatement
checks the whole weight-ingestion path (key layout of the reference
checkpoints, BN folding, filter re-indexing, op order) on a machine without a
GPU.  It is not used by the product.
"""
import numpy as np
import torch
import torch.nn.functional as F

from terran_b200 import _native as nat


class Interpreter:
    def __init__(self, program, round_activations=True):
        self.p = program
        self.blob = bytes(program.blob)
        self.round = round_activations

    # -- blob access
    def arr(self, off, dtype, count):
        if off < 0:
            return None
        return torch.from_numpy(np.frombuffer(self.blob, dtype=dtype, count=count, offset=off).copy())

    def vec(self, off, n):
        return self.arr(off, np.float32, n)

    def _store(self, bufs, idx, coff, y, N, H, W):
        """y: (N, C, H, W) fp32 -> channel slice of NHWC buffer idx."""
        ch, f32 = self.p.buffers[idx]
        if bufs[idx] is None:
            bufs[idx] = torch.zeros((N, H, W, ch), dtype=torch.float32)
        assert bufs[idx].shape[:3] == (N, H, W), 'buffer written with inconsistent dims'
        v = y.permute(0, 2, 3, 1)
        if self.round and not f32:
            v = v.half().float()                  # activations live in fp16
        bufs[idx][..., coff:coff + v.shape[3]] = v

    @staticmethod
    def _act(y, act, slope):
        if act == nat.TR_ACT_RELU:
            return y.clamp_min(0)
        if act == nat.TR_ACT_PRELU:
            return torch.where(y >= 0, y, y * slope.view(1, -1, 1, 1))
        return y

    def run(self, image):
        """image: (N, H, W, 3) uint8/float tensor in MODEL channel order."""
        bufs = [None] * len(self.p.buffers)
        for d in self.p.ops:
            t = d.type
            if t == nat.TR_OP_STEM:
                x = image.float().permute(0, 3, 1, 2) * d.in_scale + d.in_shift
                co = d.out_c
                w = self.arr(d.w_off, np.float32, co * 27).view(co, 3, 3, 3).permute(0, 3, 1, 2)
                y = F.conv2d(x, w, stride=d.stride, padding=1)
                y = y * self.vec(d.scale_off, co).view(1, -1, 1, 1) + self.vec(d.shift_off, co).view(1, -1, 1, 1)
                y = self._act(y, d.act, self.vec(d.slope_off, co))
                N, _, H, W = y.shape
                self._store(bufs, d.out, d.out_coff, y, N, H, W)
                if d.out2 >= 0:
                    y2 = y * self.vec(d.scale2_off, co).view(1, -1, 1, 1) + self.vec(d.shift2_off, co).view(1, -1, 1, 1)
                    self._store(bufs, d.out2, d.out2_coff, y2, N, H, W)
            elif t == nat.TR_OP_CONV:
                G = d.groups if d.groups > 1 else 1       # block g reads channels in_coff + g * in_c
                x = bufs[d.in_][..., d.in_coff:d.in_coff + G * d.in_c].permute(0, 3, 1, 2)
                cp, k = d.cout_pad, d.k
                w = self.arr(d.w_off, np.float16, cp * k * k * d.in_c).float().view(cp, k, k, d.in_c).permute(0, 3, 1, 2)
                y = F.conv2d(x, w, stride=d.stride, padding=d.pad, groups=G)
                y = y * self.vec(d.scale_off, cp).view(1, -1, 1, 1)
                if getattr(d, 'shift9_off', -1) >= 0:
                    # border-class shifts: class = 3 * row class + column class (first/inner/last)
                    assert k == 3 and d.pad == 1 and d.stride == 1
                    s9 = self.arr(d.shift9_off, np.float32, 9 * cp).view(9, cp)
                    Hh, Ww = y.shape[2:]
                    rc = torch.ones(Hh, dtype=torch.long); rc[0] = 0; rc[-1] = 2
                    cc = torch.ones(Ww, dtype=torch.long); cc[0] = 0; cc[-1] = 2
                    cls = rc.view(-1, 1) * 3 + cc.view(1, -1)                 # (H, W)
                    y = y + s9[cls].permute(2, 0, 1).unsqueeze(0)
                else:
                    y = y + self.vec(d.shift_off, cp).view(1, -1, 1, 1)
                y = self._act(y, d.act, self.vec(d.slope_off, cp))[:, :d.out_c]
                N, _, H, W = y.shape
                if d.res >= 0:
                    r = bufs[d.res][..., d.res_coff:d.res_coff + d.out_c].permute(0, 3, 1, 2)
                    if d.res_up2:
                        r = F.interpolate(r, scale_factor=2)[:, :, :H, :W]
                    y = y + r
                self._store(bufs, d.out, d.out_coff, y, N, H, W)
                if d.out2 >= 0:
                    y2 = y * self.vec(d.scale2_off, cp)[:d.out_c].view(1, -1, 1, 1) + \
                        self.vec(d.shift2_off, cp)[:d.out_c].view(1, -1, 1, 1)
                    self._store(bufs, d.out2, d.out2_coff, y2, N, H, W)
            elif t in (nat.TR_OP_DWCONV, nat.TR_OP_SEPCONV):
                c = d.in_c
                x = bufs[d.in_][..., d.in_coff:d.in_coff + c].permute(0, 3, 1, 2)
                if t == nat.TR_OP_DWCONV:
                    dw = self.arr(d.w_off, np.float32, 9 * c)
                    ds, dt = self.vec(d.scale_off, c), self.vec(d.shift_off, c)
                else:                              # the fused kernel multiplies by the fp16 copy
                    dw = self.arr(d.dw_w16_off, np.float16, 9 * c).float()
                    assert torch.equal(dw, self.arr(d.dw_w_off, np.float32, 9 * c).half().float())
                    ds, dt = self.vec(d.dw_scale_off, c), self.vec(d.dw_shift_off, c)
                w = dw.view(3, 3, c).permute(2, 0, 1).unsqueeze(1)
                y = F.conv2d(x, w, stride=d.stride, padding=1, groups=c)
                y = (y * ds.view(1, -1, 1, 1) + dt.view(1, -1, 1, 1)).clamp_min(0)
                if t == nat.TR_OP_SEPCONV:
                    if self.round:
                        y = y.half().float()       # the A operand of the 1x1 is fp16
                    cp = d.cout_pad
                    pw = self.arr(d.w_off, np.float16, cp * c).float().view(cp, c, 1, 1)
                    y = F.conv2d(y, pw)
                    y = y * self.vec(d.scale_off, cp).view(1, -1, 1, 1) + self.vec(d.shift_off, cp).view(1, -1, 1, 1)
                    y = self._act(y, d.act, None)[:, :d.out_c]
                N, _, H, W = y.shape
                self._store(bufs, d.out, d.out_coff, y, N, H, W)
            elif t == nat.TR_OP_MAXPOOL:
                x = bufs[d.in_][..., d.in_coff:d.in_coff + d.in_c].permute(0, 3, 1, 2)
                y = F.max_pool2d(x, 2, 2, 0)
                N, _, H, W = y.shape
                self._store(bufs, d.out, d.out_coff, y, N, H, W)
            elif t == nat.TR_OP_COPY:
                x = bufs[d.in_][..., d.in_coff:d.in_coff + d.in_c].permute(0, 3, 1, 2)
                N, _, H, W = x.shape
                self._store(bufs, d.out, d.out_coff, x, N, H, W)
            elif t == nat.TR_OP_VIEW:
                x = bufs[d.in_]
                bufs[d.out] = x.reshape(x.shape[0], 1, 1, -1)
            else:
                raise AssertionError(f'unknown op type {t}')
        return bufs
