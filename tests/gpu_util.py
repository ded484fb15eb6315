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
    cout_pad = ((cout + 127) // 128 * 128 if cout != 64 else 64) if int(use_tc) == 3 else (cout + 15) // 16 * 16
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


def sepconv2d_native(nat, x, dw_w, dw_scale, dw_shift, w, scale, shift, *, stride=1, act=1,
                     fused=True, repeat=0):
    """Depthwise 3x3 + BN + ReLU -> 1x1 + scale/shift (+act) through ``tr_sepconv2d``.
    x: (N,H,W,C) fp16 CUDA NHWC; dw_w: (C,1,3,3); w: (Cout,C,1,1)."""
    nat.init(0)
    N, H, W, cin = x.shape
    cout = w.shape[0]
    cout_pad = (cout + 15) // 16 * 16
    cout_store = (cout + 7) // 8 * 8
    wp = torch.zeros((cout_pad, cin), dtype=torch.float16)
    wp[:cout] = w[:, :, 0, 0].half()
    wp = wp.cuda()
    dwp = dw_w[:, 0].permute(1, 2, 0).contiguous().float().cuda()
    dwp16 = dwp.half()

    def vec(v, n):
        t = torch.zeros(n, dtype=torch.float32)
        t[:len(v)] = torch.as_tensor(v, dtype=torch.float32)
        return t.cuda()

    ds, dt = vec(dw_scale, cin), vec(dw_shift, cin)
    sc, sh = vec(scale, cout_pad), vec(shift, cout_pad)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.full((N, Ho, Wo, cout_store), float('nan'), dtype=torch.float16, device='cuda')
    tmp = torch.empty((N, Ho, Wo, cin), dtype=torch.float16, device='cuda')
    ms = C.c_float(0)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    nat.check(nat.lib().tr_sepconv2d(
        ptr(x), N, H, W, cin, 0, cin, ptr(dwp), ptr(dwp16), ptr(ds), ptr(dt), stride, ptr(wp), ptr(sc), ptr(sh),
        cout_pad, cout_store, act, ptr(out), cout_store, 0, ptr(tmp), int(fused), repeat,
        C.byref(ms), nat.current_stream_ptr()))
    torch.cuda.synchronize()
    return out[..., :cout], ms.value


def sepconv2d_reference(x, dw_w, dw_scale, dw_shift, w, scale, shift, *, stride=1, act=1):
    """fp64 CPU reference; the depthwise result is rounded to fp16 like the kernels do."""
    xin = x.detach().cpu().double().permute(0, 3, 1, 2)
    C_ = xin.shape[1]
    y = torch.nn.functional.conv2d(xin, dw_w.half().double(), stride=stride, padding=1, groups=C_)
    y = y * torch.as_tensor(dw_scale).double().view(1, -1, 1, 1) + torch.as_tensor(dw_shift).double().view(1, -1, 1, 1)
    y = y.clamp_min(0).half().double()
    z = torch.nn.functional.conv2d(y, w.half().double())
    z = z * torch.as_tensor(scale).double().view(1, -1, 1, 1) + torch.as_tensor(shift).double().view(1, -1, 1, 1)
    if act == 1:
        z = z.clamp_min(0)
    return z.permute(0, 2, 3, 1).contiguous()


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
