"""Helpers for the GPU tests: single-op launches through the C ABI."""
import ctypes as C

import numpy as np
import torch


def conv2d_native(nat, x, w, scale, shift, *, stride=1, pad=None, act=0, slope=None, res=None,
                  res_up2=False, use_tc=True, out_f32=False, cin_pad=None, repeat=0):
    """x: (N,H,W,Cin) fp16 CUDA NHWC; w: (Cout,Cin,k,k) fp32.  Returns
    (out (N,Ho,Wo,Cout) torch, ms)."""
    nat.init(0)
    N, H, W, cin = x.shape
    cout, _, k, _ = w.shape
    pad = k // 2 if pad is None else pad
    cin_pad = cin_pad or cin
    cout_pad = (cout + 15) // 16 * 16
    cout_store = (cout + 7) // 8 * 8
    assert x.shape[3] == cin_pad
    wp = torch.zeros((cout_pad, k, k, cin_pad), dtype=torch.float16)
    wp[:cout, :, :, :w.shape[1]] = w.permute(0, 2, 3, 1).half()
    wp = wp.cuda().contiguous()

    def vec(v):
        if v is None:
            return None
        t = torch.zeros(cout_pad, dtype=torch.float32)
        t[:cout] = torch.as_tensor(v, dtype=torch.float32)
        return t.cuda()

    sc, sh, sl = vec(scale), vec(shift), vec(slope)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    out = torch.full((N, Ho, Wo, cout_store), float('nan'),
                     dtype=torch.float32 if out_f32 else torch.float16, device='cuda')
    ms = C.c_float(0)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    nat.check(nat.lib().tr_conv2d(
        ptr(x), N, H, W, cin_pad, 0, cin_pad, ptr(wp), ptr(sc), ptr(sh), ptr(sl), cout_pad,
        cout_store, k, stride, pad, act, ptr(res), res.shape[3] if res is not None else 0,
        int(res_up2), ptr(out), cout_store, 0, int(out_f32), int(use_tc), repeat, C.byref(ms),
        nat.current_stream_ptr()))
    torch.cuda.synchronize()
    return out[..., :cout], ms.value


def conv2d_reference(x, w, scale, shift, *, stride=1, pad=None, act=0, slope=None, res=None,
                     res_up2=False):
    """fp64 CPU reference on the SAME fp16-rounded operands."""
    k = w.shape[2]
    pad = k // 2 if pad is None else pad
    xin = x.detach().cpu().double().permute(0, 3, 1, 2)[:, :w.shape[1]]
    y = torch.nn.functional.conv2d(xin, w.half().double(), stride=stride, padding=pad)
    y = y * torch.as_tensor(scale).double().view(1, -1, 1, 1) + torch.as_tensor(shift).double().view(1, -1, 1, 1)
    if act == 1:
        y = y.clamp_min(0)
    elif act == 2:
        s = torch.as_tensor(slope).double().view(1, -1, 1, 1)
        y = torch.where(y >= 0, y, y * s)
    if res is not None:
        r = res.detach().cpu().double().permute(0, 3, 1, 2)[:, :y.shape[1]]
        if res_up2:
            r = torch.nn.functional.interpolate(r, scale_factor=2)[:, :, :y.shape[2], :y.shape[3]]
        y = y + r
    return y.permute(0, 2, 3, 1).contiguous()


def describe_mismatch(out, ref, tol):
    """Diagnostics that localise a broken descriptor / layout."""
    err = (out.detach().cpu().double() - ref).abs()
    bad = err > tol
    msg = [f'max err {err.max():.4g}, bad {int(bad.sum())}/{bad.numel()}']
    if bad.any():
        N, H, W, Cc = err.shape
        msg.append('bad per channel%16: ' + str([int(bad[..., c::16].sum()) for c in range(min(16, Cc))]))
        flat = bad.reshape(-1, Cc).any(1).numpy()
        rows = np.flatnonzero(flat)
        msg.append(f'first bad pixels {rows[:12].tolist()} of {len(flat)}')
        msg.append('bad per pixel%8: ' + str([int(flat[i::8].sum()) for i in range(8)]))
        i = np.unravel_index(int(err.argmax()), err.shape)
        msg.append(f'worst at {tuple(int(v) for v in i)}: got {float(out[i]):.5g} want {float(ref[i]):.5g}')
    return '; '.join(msg)
