// CUDA-core kernels: the generic direct convolution (cross-check for the
// tcgen05 path and the fallback for shapes it does not take), the u8 stem
// convolutions, depthwise 3x3, 2x2 max-pool, channel-slice copy, layout export,
// L2 normalisation and the cv2-exact bilinear resize.  All HBM-bound byte/half
// work: coalesced 16-byte accesses over the NHWC channel axis.
#include "common.cuh"

namespace trb {

namespace {

__device__ __forceinline__ float act_f(float y, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(y, 0.f);
  if (act == ACT_PRELU) return y >= 0.f ? y : y * slope;
  return y;
}

// ---------------------------------------------------------------- direct conv
// One thread = one output pixel x 8 output channels, fp32 accumulation.
struct DirectParams {
  const __half* in; int in_cs, in_coff, H, W, N;
  const __half* w; int cin_pad, kh, kw, stride, pad;
  const float* scale; const float* shift; const float* slope;
  const float* scale2; const float* shift2;
  const float* shift9; int cout_pad;
  int act, H_out, W_out, cout_store;
  __half* out; int out_cs, out_coff;
  __half* out2; int out2_cs, out2_coff;
  const __half* res; int res_cs, res_coff, res_up2, res_H, res_W;
  float* out_f32;
};

__global__ void __launch_bounds__(128) conv_direct_kernel(const DirectParams p) {
  const long npix = static_cast<long>(p.N) * p.H_out * p.W_out;
  const long pix = blockIdx.x * 128L + threadIdx.x;
  if (pix >= npix) return;
  const int c = blockIdx.y * 8;
  if (c >= p.cout_store) return;
  const int ow = pix % p.W_out;
  const int oh = (pix / p.W_out) % p.H_out;
  const int on = pix / (static_cast<long>(p.W_out) * p.H_out);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int taps = p.kh * p.kw;
  for (int r = 0; r < p.kh; ++r) {
    const int ih = oh * p.stride + r - p.pad;
    if (ih < 0 || ih >= p.H) continue;
    for (int s = 0; s < p.kw; ++s) {
      const int iw = ow * p.stride + s - p.pad;
      if (iw < 0 || iw >= p.W) continue;
      const __half* ip = p.in + ((static_cast<long>(on) * p.H + ih) * p.W + iw) * p.in_cs + p.in_coff;
      const __half* wp = p.w + (static_cast<long>(c) * taps + r * p.kw + s) * p.cin_pad;
      for (int k = 0; k < p.cin_pad; k += 8) {
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(ip + k));
        const __half2* xh = reinterpret_cast<const __half2*>(&xv);
        float x[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(xh[j]);
          x[2 * j] = f.x; x[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wp + static_cast<long>(o) * taps * p.cin_pad + k));
          const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(wh[j]);
            acc[o] = fmaf(x[2 * j], f.x, acc[o]);
            acc[o] = fmaf(x[2 * j + 1], f.y, acc[o]);
          }
        }
      }
    }
  }
  float y[8];
  const float* shp = p.shift;
  if (p.shift9) {
    const int rc = oh == 0 ? 0 : (oh >= p.H_out - 1 ? 2 : 1), cc = ow == 0 ? 0 : (ow >= p.W_out - 1 ? 2 : 1);
    shp = p.shift9 + (rc * 3 + cc) * p.cout_pad;
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float t = fmaf(acc[o], __ldg(p.scale + c + o), __ldg(shp + c + o));
    y[o] = act_f(t, p.act, p.act == ACT_PRELU ? __ldg(p.slope + c + o) : 0.f);
  }
  if (p.res) {
    long rpix = pix;
    if (p.res_up2) rpix = (static_cast<long>(on) * p.res_H + (oh >> 1)) * p.res_W + (ow >> 1);
    const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.res + rpix * p.res_cs + p.res_coff + c));
    const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(rh[j]);
      y[2 * j] += f.x; y[2 * j + 1] += f.y;
    }
  }
  if (p.out_f32) {
    float4* o = reinterpret_cast<float4*>(p.out_f32 + pix * p.out_cs + p.out_coff + c);
    o[0] = make_float4(y[0], y[1], y[2], y[3]);
    o[1] = make_float4(y[4], y[5], y[6], y[7]);
  } else {
    uint4 ov;
    __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh2[j] = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
    *reinterpret_cast<uint4*>(p.out + pix * p.out_cs + p.out_coff + c) = ov;
  }
  if (p.out2) {
    uint4 ov;
    __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = fmaf(y[2 * j], __ldg(p.scale2 + c + 2 * j), __ldg(p.shift2 + c + 2 * j));
      const float a1 = fmaf(y[2 * j + 1], __ldg(p.scale2 + c + 2 * j + 1), __ldg(p.shift2 + c + 2 * j + 1));
      oh2[j] = __floats2half2_rn(a0, a1);
    }
    *reinterpret_cast<uint4*>(p.out2 + pix * p.out2_cs + p.out2_coff + c) = ov;
  }
}

// ------------------------------------------------------------------ u8 stems
// 3x3, pad 1, 3 input channels read straight from the u8 frame (any strides,
// any channel order: the BGR flip of the reference is a negative channel
// stride).  x' = x*in_scale + in_shift is applied to IN-BOUNDS taps only, so
// the zero padding sees zeros exactly like the reference (which pads after the
// affine).  Register-blocked: one thread = 4 consecutive output pixels x all
// COUT channels (16 at a time), so each smem weight vector feeds 4 pixels.
template <int COUT, int STRIDE>
__global__ void __launch_bounds__(128) stem_kernel(const StemArgs a, int H_out, int W_out) {
  constexpr int PX = 4;
  constexpr int COLS = (PX - 1) * STRIDE + 3;
  constexpr int CB = COUT < 16 ? COUT : 16;          // couts per register block
  __shared__ __align__(16) float sw[27 * COUT];
  __shared__ float sp[5 * COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    // packed as [cout][kh][kw][c]; stored transposed [tap*3+c][cout] for vector reads
    const int o = i / 27, t = i % 27;
    sw[t * COUT + o] = a.w[i];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    sp[i] = a.scale[i];
    sp[COUT + i] = a.shift[i];
    sp[2 * COUT + i] = a.slope ? a.slope[i] : 0.f;
    sp[3 * COUT + i] = a.scale2 ? a.scale2[i] : 1.f;
    sp[4 * COUT + i] = a.shift2 ? a.shift2[i] : 0.f;
  }
  __syncthreads();
  const int wq = (W_out + PX - 1) / PX;
  const long nthreads = static_cast<long>(a.N) * H_out * wq;
  const long tid = blockIdx.x * 128L + threadIdx.x;
  if (tid >= nthreads) return;
  const int ow0 = static_cast<int>(tid % wq) * PX;
  const int oh = (tid / wq) % H_out;
  const int on = tid / (static_cast<long>(wq) * H_out);

  float x[3][COLS][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = oh * STRIDE + r - 1;
    const bool rok = ih >= 0 && ih < a.H;
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int iw = ow0 * STRIDE + c - 1;
      const bool ok = rok && iw >= 0 && iw < a.W;
      const uint8_t* ip = a.in + on * a.sn + (ok ? ih : 0) * a.sh + (ok ? iw : 0) * a.sw;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        x[r][c][ch] = ok ? fmaf(static_cast<float>(ip[ch * a.sc]), a.in_scale, a.in_shift) : 0.f;
    }
  }
  const long pix0 = (static_cast<long>(on) * H_out + oh) * W_out + ow0;
#pragma unroll 1
  for (int o0 = 0; o0 < COUT; o0 += CB) {
    float acc[PX][CB];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int j = 0; j < CB; ++j) acc[p][j] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float4* wv = reinterpret_cast<const float4*>(sw + ((r * 3 + s) * 3 + ch) * COUT + o0);
          float w[CB];
#pragma unroll
          for (int j = 0; j < CB / 4; ++j) {
            const float4 t = wv[j];
            w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
          }
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            const float xv = x[r][p * STRIDE + s][ch];
#pragma unroll
            for (int j = 0; j < CB; ++j) acc[p][j] = fmaf(xv, w[j], acc[p][j]);
          }
        }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (ow0 + p >= W_out) break;
      __half* op = a.out.ptr + (pix0 + p) * a.out.cs + a.out.coff + o0;
      __half* op2 = a.out2.ptr ? a.out2.ptr + (pix0 + p) * a.out2.cs + a.out2.coff + o0 : nullptr;
#pragma unroll
      for (int g = 0; g < CB; g += 8) {
        uint4 ov, ov2;
        __half2* h = reinterpret_cast<__half2*>(&ov);
        __half2* h2 = reinterpret_cast<__half2*>(&ov2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c0 = o0 + g + 2 * j;
          float y0 = fmaf(acc[p][g + 2 * j], sp[c0], sp[COUT + c0]);
          float y1 = fmaf(acc[p][g + 2 * j + 1], sp[c0 + 1], sp[COUT + c0 + 1]);
          y0 = act_f(y0, a.act, sp[2 * COUT + c0]);
          y1 = act_f(y1, a.act, sp[2 * COUT + c0 + 1]);
          h[j] = __floats2half2_rn(y0, y1);
          h2[j] = __floats2half2_rn(fmaf(y0, sp[3 * COUT + c0], sp[4 * COUT + c0]),
                                    fmaf(y1, sp[3 * COUT + c0 + 1], sp[4 * COUT + c0 + 1]));
        }
        *reinterpret_cast<uint4*>(op + g) = ov;
        if (op2) *reinterpret_cast<uint4*>(op2 + g) = ov2;
      }
    }
  }
}

// Tensor-core stem: the same 3x3 / 3-channel conv as an implicit GEMM on the warp-level
// MMA (mma.sync m16n8k16, fp16 x fp16 -> fp32): M = 16 consecutive output pixels of a row,
// N = COUT, K = 27 taps padded to 32.  The CUDA-core kernel above needs 27*COUT FMAs per
// pixel and is FMA-issue bound ~6x above the HBM time of the layer; here the MACs are 2*NT
// MMAs per 16 pixels and the kernel is bound by writing the activations.
//   A[m][k]  = x(pixel m, tap k) + in_shift/in_scale   (exact in fp16: u8 plus a half-integer)
//   B[k][n]  = w[n][k] * scale[n] * in_scale           (rounded to fp16 once)
//   acc init = shift[n]
// so the epilogue is activation + fp16 pack.  Zero padding: out-of-bounds taps are staged as
// 0, like the reference which pads after the input affine.  The block stages the u8 input
// window of an 8 x 64 pixel tile in smem as fp16; with k = r*9 + s*3 + c the 9 values of a
// filter row are contiguous there.
__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// OCC = CTAs per SM the register budget is cut for.  The kernel is latency-bound (one window
// fetch in flight per CTA while the previous tile is computed), so the 8-filter RetinaFace stem,
// which needs 80 registers without spilling, runs three CTAs per SM (TRB_STEM_OCC=2|3|4 selects
// the instantiation for A/B runs; 4 spills 88 bytes).
template <int COUT, int STRIDE, int ACT, int OCC = 2>
__global__ void __launch_bounds__(256, OCC) stem_mma_kernel(const StemArgs a, int H_out, int W_out, float in_off,
                                                          int tiles_w, int tiles_h, int total_tiles) {
  constexpr int NT = COUT / 8, TH = 8, TW = 64, SEGS = TW / 16;
  constexpr int IN_H = (TH - 1) * STRIDE + 3, IN_W = (TW - 1) * STRIDE + 3;
  constexpr int ROWB = IN_W * 3;                     // window row: ROWB contiguous (col, channel) values
  constexpr int PASSES = (ROWB + 255) / 256;
  constexpr int ZERO = IN_H * ROWB;                  // a slot that always holds 0 (padded K taps)
  constexpr int PITCH = COUT + 8;                    // halfs; conflict-free fragment stores
  __shared__ __half s_in[IN_H * ROWB + 2];
  __shared__ __align__(16) __half s_out[8][16 * PITCH];
  __shared__ float sp[5 * COUT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    sp[i] = a.scale[i];
    sp[COUT + i] = a.shift[i];
    sp[2 * COUT + i] = a.slope ? a.slope[i] : 0.f;
    sp[3 * COUT + i] = a.scale2 ? a.scale2[i] : 1.f;
    sp[4 * COUT + i] = a.shift2 ? a.shift2[i] : 0.f;
  }
  if (threadIdx.x < 2) s_in[ZERO + threadIdx.x] = __float2half(0.f);
  // B fragments (whole filter bank) stay in registers: b[nt][kstep][half] = w'[n = nt*8+g][k, k+1]
  uint32_t bfrag[NT][2][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int o = nt * 8 + g;
    const float fold = a.scale[o] * a.in_scale;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = ks * 16 + h * 8 + 2 * t;
        const float w0 = k < 27 ? a.w[o * 27 + k] * fold : 0.f;
        const float w1 = k + 1 < 27 ? a.w[o * 27 + k + 1] * fold : 0.f;
        const __half2 hw = __floats2half2_rn(w0, w1);
        bfrag[nt][ks][h] = *reinterpret_cast<const uint32_t*>(&hw);
      }
  }
  // smem offsets of this thread's 8 taps relative to the window origin of a pixel
  // (k = r*9 + s*3 + c; -1: a padded tap, read from the ZERO slot)
  int koff[2][2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = ks * 16 + h * 8 + 2 * t + e;
        koff[ks][h][e] = k < 27 ? (k / 9) * ROWB + k % 9 : -1;
      }
  // Staging map: thread x of pass p owns value x + 256 p of every window row.
  long goff[PASSES];
  int gcol[PASSES];
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const int x = threadIdx.x + p * 256;
    gcol[p] = x < ROWB ? x / 3 : -(1 << 20);
    goff[p] = (x / 3) * a.sw + (x % 3) * a.sc;
  }
  // The u8 window of the NEXT tile is fetched into registers (all loads in flight at once)
  // while the current tile is computed; 0xffff marks an out-of-bounds (zero) tap.
  uint16_t raw[IN_H][PASSES];
  auto fetch = [&](int tile) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int ih0 = th * TH * STRIDE - 1, iw0 = tw * TW * STRIDE - 1;
    const uint8_t* img = a.in + n * a.sn + ih0 * a.sh + iw0 * a.sw;
    bool cok[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) cok[p] = iw0 + gcol[p] >= 0 && iw0 + gcol[p] < a.W;
#pragma unroll
    for (int r = 0; r < IN_H; ++r) {
      const bool rok = ih0 + r >= 0 && ih0 + r < a.H;
#pragma unroll
      for (int p = 0; p < PASSES; ++p)
        raw[r][p] = rok && cok[p] ? static_cast<uint16_t>(img[r * a.sh + goff[p]]) : uint16_t(0xffff);
    }
  };
  if (blockIdx.x < total_tiles) fetch(blockIdx.x);
  // flush map: lane -> (row, 16-byte vector) of the 16 x COUT staging tile
  const int frow = lane / NT, fvec = lane % NT;
  constexpr int FROWS = 32 / NT > 16 ? 16 : 32 / NT;  // rows covered per flush instruction

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int oh0 = th * TH, ow0 = tw * TW;
    __syncthreads();                                 // the previous tile's window is consumed
#pragma unroll
    for (int r = 0; r < IN_H; ++r)
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
        const int x = threadIdx.x + p * 256;
        if (x < ROWB)
          s_in[r * ROWB + x] = __float2half_rn(raw[r][p] == 0xffff ? 0.f : static_cast<float>(raw[r][p]) + in_off);
      }
    __syncthreads();
    if (tile + gridDim.x < total_tiles) fetch(tile + gridDim.x);

    for (int seg = warp; seg < TH * SEGS; seg += 8) {
      const int r_l = seg / SEGS, c_l = (seg % SEGS) * 16;
      const int oh = oh0 + r_l, owb = ow0 + c_l;
      if (oh >= H_out || owb >= W_out) continue;     // warp-uniform
      const int base0 = r_l * STRIDE * ROWB + (c_l + g) * STRIDE * 3;
      const int base1 = base0 + 8 * STRIDE * 3;
      uint32_t afrag[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k0 = koff[ks][h][0], k1 = koff[ks][h][1];
          const __half2 lo = __halves2half2(s_in[k0 >= 0 ? base0 + k0 : ZERO], s_in[k1 >= 0 ? base0 + k1 : ZERO]);
          const __half2 hi = __halves2half2(s_in[k0 >= 0 ? base1 + k0 : ZERO], s_in[k1 >= 0 ? base1 + k1 : ZERO]);
          afrag[ks][2 * h] = *reinterpret_cast<const uint32_t*>(&lo);        // rows g     (a0 / a2)
          afrag[ks][2 * h + 1] = *reinterpret_cast<const uint32_t*>(&hi);    // rows g + 8 (a1 / a3)
        }
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 sh = *reinterpret_cast<const float2*>(sp + COUT + nt * 8 + 2 * t);
        acc[nt][0] = acc[nt][2] = sh.x;
        acc[nt][1] = acc[nt][3] = sh.y;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) mma_m16n8k16(acc[nt], afrag[ks], bfrag[nt][ks][0], bfrag[nt][ks][1]);
      }
      __half* so = s_out[warp];
      const long pix0 = (static_cast<long>(n) * H_out + oh) * W_out + owb;
#pragma unroll 1
      for (int pass = 0; pass < (a.out2.ptr ? 2 : 1); ++pass) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = nt * 8 + 2 * t;
          float y[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[j] = acc[nt][j];
            if (ACT == ACT_RELU) y[j] = fmaxf(y[j], 0.f);
            if (ACT == ACT_PRELU) y[j] = y[j] >= 0.f ? y[j] : y[j] * sp[2 * COUT + c + (j & 1)];
            if (pass) y[j] = fmaf(y[j], sp[3 * COUT + c + (j & 1)], sp[4 * COUT + c + (j & 1)]);
          }
          *reinterpret_cast<__half2*>(so + g * PITCH + c) = __floats2half2_rn(y[0], y[1]);
          *reinterpret_cast<__half2*>(so + (g + 8) * PITCH + c) = __floats2half2_rn(y[2], y[3]);
        }
        __syncwarp();
        const View& ov = pass ? a.out2 : a.out;
        __half* gp = ov.ptr + (pix0 + frow) * ov.cs + ov.coff + fvec * 8;
        const __half* sp_row = so + frow * PITCH + fvec * 8;
#pragma unroll
        for (int i = 0; i < 16 / FROWS; ++i) {
          const int row = frow + i * FROWS;
          if (row < 16 && owb + row < W_out)
            *reinterpret_cast<uint4*>(gp + static_cast<long>(i) * FROWS * ov.cs) =
                *reinterpret_cast<const uint4*>(sp_row + i * FROWS * PITCH);
        }
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------- depthwise 3x3
__global__ void __launch_bounds__(256) dwconv_kernel(const DwArgs a, int H_out, int W_out) {
  const int groups = a.in.C / 8;
  const long total = static_cast<long>(a.in.N) * H_out * W_out * groups;
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % groups;
  const long pix = idx / groups;
  const int ow = pix % W_out;
  const int oh = (pix / W_out) % H_out;
  const int on = pix / (static_cast<long>(W_out) * H_out);
  const int c = g * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < 3; ++r) {
    const int ih = oh * a.stride + r - 1;
    if (ih < 0 || ih >= a.in.H) continue;
    for (int s = 0; s < 3; ++s) {
      const int iw = ow * a.stride + s - 1;
      if (iw < 0 || iw >= a.in.W) continue;
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(
          a.in.ptr + ((static_cast<long>(on) * a.in.H + ih) * a.in.W + iw) * a.in.cs + a.in.coff + c));
      const __half2* xh = reinterpret_cast<const __half2*>(&xv);
      const float* wp = a.w + (r * 3 + s) * a.in.C + c;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(xh[j]);
        acc[2 * j] = fmaf(f.x, wv[2 * j], acc[2 * j]);
        acc[2 * j + 1] = fmaf(f.y, wv[2 * j + 1], acc[2 * j + 1]);
      }
    }
  }
  uint4 ov;
  __half2* h = reinterpret_cast<__half2*>(&ov);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float y0 = fmaxf(fmaf(acc[2 * j], __ldg(a.scale + c + 2 * j), __ldg(a.shift + c + 2 * j)), 0.f);
    const float y1 = fmaxf(fmaf(acc[2 * j + 1], __ldg(a.scale + c + 2 * j + 1), __ldg(a.shift + c + 2 * j + 1)), 0.f);
    h[j] = __floats2half2_rn(y0, y1);
  }
  *reinterpret_cast<uint4*>(a.out.ptr + pix * a.out.cs + a.out.coff + c) = ov;
}

// ---------------------------------------------------------------- 2x2 maxpool
__global__ void __launch_bounds__(256) maxpool2_kernel(const View in, const View out) {
  const int groups = out.C / 8;
  const long total = static_cast<long>(out.N) * out.H * out.W * groups;
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % groups;
  const long pix = idx / groups;
  const int ow = pix % out.W;
  const int oh = (pix / out.W) % out.H;
  const int on = pix / (static_cast<long>(out.W) * out.H);
  const __half* base = in.ptr + ((static_cast<long>(on) * in.H + oh * 2) * in.W + ow * 2) * in.cs + in.coff + g * 8;
  uint4 v00 = __ldg(reinterpret_cast<const uint4*>(base));
  const uint4 v01 = __ldg(reinterpret_cast<const uint4*>(base + in.cs));
  const uint4 v10 = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long>(in.W) * in.cs));
  const uint4 v11 = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long>(in.W) * in.cs + in.cs));
  __half2* a = reinterpret_cast<__half2*>(&v00);
  const __half2* b = reinterpret_cast<const __half2*>(&v01);
  const __half2* c = reinterpret_cast<const __half2*>(&v10);
  const __half2* d = reinterpret_cast<const __half2*>(&v11);
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = __hmax2(__hmax2(a[j], b[j]), __hmax2(c[j], d[j]));
  *reinterpret_cast<uint4*>(out.ptr + pix * out.cs + out.coff + g * 8) = v00;
}

__global__ void __launch_bounds__(256) copy_slice_kernel(const View in, const View out) {
  const int groups = in.C / 8;
  const long total = static_cast<long>(in.N) * in.H * in.W * groups;
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % groups;
  const long pix = idx / groups;
  *reinterpret_cast<uint4*>(out.ptr + pix * out.cs + out.coff + g * 8) =
      __ldg(reinterpret_cast<const uint4*>(in.ptr + pix * in.cs + in.coff + g * 8));
}

__global__ void __launch_bounds__(256) export_nchw_kernel(const View in, int C, float* out) {
  const long total = static_cast<long>(in.N) * C * in.H * in.W;
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % in.W;
  const int h = (idx / in.W) % in.H;
  const int c = (idx / (static_cast<long>(in.W) * in.H)) % C;
  const int n = idx / (static_cast<long>(in.W) * in.H * C);
  out[idx] = __half2float(in.ptr[((static_cast<long>(n) * in.H + h) * in.W + w) * in.cs + in.coff + c]);
}

// fp32 NHWC slice -> NCHW; with softmax_pairs the 4 channels are the class
// logits and channel a is soft-maxed against channel a^2 (model.py:283-286).
__global__ void __launch_bounds__(256)
export_nchw_f32_kernel(const float* in, int N, int H, int W, int cs, int coff, int C, float* out,
                       int softmax_pairs) {
  const long total = static_cast<long>(N) * C * H * W;
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= total) return;
  const int w = idx % W;
  const int h = (idx / W) % H;
  const int c = (idx / (static_cast<long>(W) * H)) % C;
  const int n = idx / (static_cast<long>(W) * H * C);
  const float* px = in + ((static_cast<long>(n) * H + h) * W + w) * cs + coff;
  float v = px[c];
  if (softmax_pairs) {
    const float o = px[c ^ 2];
    const float m = fmaxf(v, o);
    const float ev = expf(v - m), eo = expf(o - m);
    v = ev / (ev + eo);
  }
  out[idx] = v;
}

// x / max-safe L2 norm per row (sklearn normalize: zero rows divide by 1).
__global__ void __launch_bounds__(128) l2_normalize_kernel(const float* in, float* out, int D) {
  const float* x = in + static_cast<long>(blockIdx.x) * D;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += 128) s = fmaf(x[i], x[i], s);
  __shared__ float red[4];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = red[0] + red[1] + red[2] + red[3];
  float nrm = sqrtf(s);
  if (nrm == 0.f) nrm = 1.f;
  for (int i = threadIdx.x; i < D; i += 128) out[static_cast<long>(blockIdx.x) * D + i] = x[i] / nrm;
}

// cv2.resize(INTER_LINEAR) on uint8, bit-exact integer restatement
// (SURVEY.md Appendix A.5): 11-bit fixed-point taps, the intermediate row sum
// is >>4, the vertical blend is ((b0*r0)>>16) + ((b1*r1)>>16) + 2 >> 2.
// One thread = kResizePx consecutive output pixels of a row: all their source bytes are
// requested before the first blend, so a warp has 4 x 12 independent loads in flight instead of
// a dependent chain per pixel (the kernel is latency-, not bandwidth-bound).
constexpr int kResizePx = 4;
template <bool WIDE>
__global__ void __launch_bounds__(256)
resize_u8_kernel(const uint8_t* __restrict__ src, int N, int H, int W, uint8_t* __restrict__ dst, int h, int w,
                 double sy_d, double sx_d) {
  const int wq = (w + kResizePx - 1) / kResizePx;             // pixel quads per output row
  const long total = static_cast<long>(N) * h * wq;
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= total) return;
  const int oxq = idx % wq;
  const int oy = (idx / wq) % h;
  const int n = idx / (static_cast<long>(wq) * h);
  // OpenCV clamps index AND fraction along x, but along y only clips the row
  // index when fetching rows (the weights keep the unclamped fraction).
  auto tap = [](int o, double scale, int n_src, bool clamp_weights, int& i0, int& i1, int& a0,
                int& a1) {
    float f = static_cast<float>((o + 0.5) * scale - 0.5);
    int s = static_cast<int>(floorf(f));
    f -= s;
    if (clamp_weights) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
    }
    i0 = min(max(s, 0), n_src - 1);
    i1 = min(max(s + 1, 0), n_src - 1);
    a1 = static_cast<int>(rintf(f * 2048.f));
    a0 = static_cast<int>(rintf((1.f - f) * 2048.f));
  };
  int y0, y1, ay0, ay1;
  tap(oy, sy_d, H, false, y0, y1, ay0, ay1);
  const uint8_t* r0 = src + (static_cast<long>(n) * H + y0) * W * 3;
  const uint8_t* r1 = src + (static_cast<long>(n) * H + y1) * W * 3;
  const uint8_t* src_end = src + static_cast<long>(N) * H * W * 3;
  int x0[kResizePx], x1[kResizePx], ax0[kResizePx], ax1[kResizePx];
  uint8_t t00[kResizePx][3], t01[kResizePx][3], t10[kResizePx][3], t11[kResizePx][3];
#pragma unroll
  for (int j = 0; j < kResizePx; ++j) {
    const int ox = min(oxq * kResizePx + j, w - 1);            // (the tail quad repeats its last pixel)
    tap(ox, sx_d, W, true, x0[j], x1[j], ax0[j], ax1[j]);
  }
  // The two taps of a pixel are six CONTIGUOUS bytes of a source row (x1 = x0 + 1 except at the
  // clamped right border): two aligned 8-byte loads and a funnel shift fetch them, instead of
  // six byte loads — the kernel is bound by the number of load instructions (lg_throttle).
  // The second 8-byte word may reach up to 15 bytes past the taps: inside the frame batch
  // except at its very end, where the bytes are fetched one by one.
  auto six = [&](const uint8_t* a) -> unsigned long long {
    const unsigned long long addr = reinterpret_cast<unsigned long long>(a);
    const unsigned sh = static_cast<unsigned>(addr & 7u) * 8u;
    const unsigned long long* b = reinterpret_cast<const unsigned long long*>(addr & ~7ull);
    if (reinterpret_cast<const uint8_t*>(b) + 16 <= src_end) {
      const unsigned long long lo = __ldg(b), hi = __ldg(b + 1);
      return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
    }
    unsigned long long v = 0;
    for (int k = 0; k < 6; ++k)
      if (a + k < src_end) v |= static_cast<unsigned long long>(a[k]) << (8 * k);
    return v;
  };
  // (measured on 32 x 1080p: 27.5 -> 21.3 us at scale 5.9, but 69 -> 81 us at scale 2.6, where the
  // 16-byte windows of neighbouring pixels overlap and L1 bandwidth becomes the bound)
  if (WIDE) {
#pragma unroll
    for (int j = 0; j < kResizePx; ++j) {
      const unsigned long long v0 = six(r0 + x0[j] * 3), v1 = six(r1 + x0[j] * 3);
      const int s1 = x1[j] == x0[j] ? 0 : 24;                  // clamped border: both taps are pixel x0
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t00[j][c] = static_cast<uint8_t>(v0 >> (8 * c)); t01[j][c] = static_cast<uint8_t>(v0 >> (s1 + 8 * c));
        t10[j][c] = static_cast<uint8_t>(v1 >> (8 * c)); t11[j][c] = static_cast<uint8_t>(v1 >> (s1 + 8 * c));
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < kResizePx; ++j) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t00[j][c] = __ldg(r0 + x0[j] * 3 + c); t01[j][c] = __ldg(r0 + x1[j] * 3 + c);
        t10[j][c] = __ldg(r1 + x0[j] * 3 + c); t11[j][c] = __ldg(r1 + x1[j] * 3 + c);
      }
    }
  }
  uint8_t* o = dst + ((static_cast<long>(n) * h + oy) * w + static_cast<long>(oxq) * kResizePx) * 3;
#pragma unroll
  for (int j = 0; j < kResizePx; ++j) {
    if (oxq * kResizePx + j >= w) break;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int top = t00[j][c] * ax0[j] + t01[j][c] * ax1[j];
      const int bot = t10[j][c] * ax0[j] + t11[j][c] * ax1[j];
      const int v = (((ay0 * (top >> 4)) >> 16) + ((ay1 * (bot >> 4)) >> 16) + 2) >> 2;
      o[j * 3 + c] = static_cast<uint8_t>(min(max(v, 0), 255));
    }
  }
}

}  // namespace

void conv_direct_launch(const ConvArgs& a, cudaStream_t s) {
  TR_CHECK(a.cin_pad % 8 == 0 && a.cout_store % 8 == 0, "direct conv needs 8-channel granules");
  DirectParams p{};
  p.in = a.in.ptr; p.in_cs = a.in.cs; p.in_coff = a.in.coff; p.H = a.in.H; p.W = a.in.W; p.N = a.in.N;
  p.w = a.w; p.cin_pad = a.cin_pad; p.kh = a.kh; p.kw = a.kw; p.stride = a.stride; p.pad = a.pad;
  p.scale = a.scale; p.shift = a.shift; p.slope = a.slope; p.scale2 = a.scale2; p.shift2 = a.shift2;
  p.shift9 = a.shift9; p.cout_pad = a.cout_pad;
  p.act = a.act; p.H_out = a.H_out; p.W_out = a.W_out; p.cout_store = a.cout_store;
  p.out = a.out.ptr; p.out_cs = a.out.cs; p.out_coff = a.out.coff;
  p.out2 = a.out2.ptr; p.out2_cs = a.out2.cs; p.out2_coff = a.out2.coff;
  p.res = a.res.ptr; p.res_cs = a.res.cs; p.res_coff = a.res.coff; p.res_up2 = a.res_up2;
  p.res_H = a.res.H; p.res_W = a.res.W;
  p.out_f32 = a.out_f32;
  const long npix = static_cast<long>(p.N) * p.H_out * p.W_out;
  dim3 grid(static_cast<unsigned>((npix + 127) / 128), a.cout_store / 8);
  conv_direct_kernel<<<grid, 128, 0, s>>>(p);
  TR_CUDA(cudaGetLastError());
}

void stem_launch(const StemArgs& a, cudaStream_t s) {
  const int H_out = (a.H + 2 - 3) / a.stride + 1, W_out = (a.W + 2 - 3) / a.stride + 1;
  TR_CHECK((a.cout == 8 && a.stride == 2) || (a.cout == 64 && a.stride == 1),
           "stem conv supports (8 channels, stride 2) or (64 channels, stride 1)");
  // Tensor-core variant when the input affine folds exactly: x*s + b = s*(x + b/s) with
  // x + b/s representable in fp16 (an integer or half-integer below 2048).
  float in_off = a.in_shift / a.in_scale;
  if (fabsf(in_off * 2.f - rintf(in_off * 2.f)) < 1e-3f) in_off = rintf(in_off * 2.f) * 0.5f;   // 1/255 is inexact
  static const bool mma_ok = [] { const char* e = getenv("TRB_STEM_MMA"); return !e || atoi(e) != 0; }();
  if (mma_ok && a.use_mma && in_off * 2.f == rintf(in_off * 2.f) && fabsf(in_off) <= 1024.f) {
    const int tiles_w = (W_out + 63) / 64, tiles_h = (H_out + 7) / 8;
    const long total = static_cast<long>(a.N) * tiles_w * tiles_h;
    TR_CHECK(total < (1L << 31), "stem: too many tiles");
    int sms = 148;
    int dev = 0;
    TR_CUDA(cudaGetDevice(&dev));
    TR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static const int occ8 = [] { const char* e = getenv("TRB_STEM_OCC"); return e ? atoi(e) : 3; }();
    const unsigned grid = static_cast<unsigned>(std::min<long>(total, (a.cout == 8 ? 2L * occ8 : 4L) * sms));
    bool done = true;
    if (a.cout == 8 && a.act == ACT_RELU && occ8 == 3)
      stem_mma_kernel<8, 2, ACT_RELU, 3><<<grid, 256, 0, s>>>(a, H_out, W_out, in_off, tiles_w, tiles_h, int(total));
    else if (a.cout == 8 && a.act == ACT_RELU && occ8 == 4)
      stem_mma_kernel<8, 2, ACT_RELU, 4><<<grid, 256, 0, s>>>(a, H_out, W_out, in_off, tiles_w, tiles_h, int(total));
    else if (a.cout == 8 && a.act == ACT_RELU)
      stem_mma_kernel<8, 2, ACT_RELU><<<grid, 256, 0, s>>>(a, H_out, W_out, in_off, tiles_w, tiles_h, int(total));
    else if (a.cout == 64 && a.act == ACT_RELU)
      stem_mma_kernel<64, 1, ACT_RELU><<<grid, 256, 0, s>>>(a, H_out, W_out, in_off, tiles_w, tiles_h, int(total));
    else if (a.cout == 64 && a.act == ACT_PRELU)
      stem_mma_kernel<64, 1, ACT_PRELU><<<grid, 256, 0, s>>>(a, H_out, W_out, in_off, tiles_w, tiles_h, int(total));
    else
      done = false;
    if (done) {
      TR_CUDA(cudaGetLastError());
      return;
    }
  }
  const long nthreads = static_cast<long>(a.N) * H_out * ((W_out + 3) / 4);
  const unsigned grid = static_cast<unsigned>((nthreads + 127) / 128);
  if (a.cout == 8 && a.stride == 2) stem_kernel<8, 2><<<grid, 128, 0, s>>>(a, H_out, W_out);
  else stem_kernel<64, 1><<<grid, 128, 0, s>>>(a, H_out, W_out);
  TR_CUDA(cudaGetLastError());
}

void dwconv_launch(const DwArgs& a, cudaStream_t s) {
  const int H_out = (a.in.H + 2 - 3) / a.stride + 1, W_out = (a.in.W + 2 - 3) / a.stride + 1;
  TR_CHECK(H_out == a.out.H && W_out == a.out.W, "depthwise output dims");
  const long total = static_cast<long>(a.in.N) * H_out * W_out * (a.in.C / 8);
  dwconv_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(a, H_out, W_out);
  TR_CUDA(cudaGetLastError());
}

void maxpool2_launch(const View& in, const View& out, cudaStream_t s) {
  const long total = static_cast<long>(out.N) * out.H * out.W * (out.C / 8);
  maxpool2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out);
  TR_CUDA(cudaGetLastError());
}

void copy_slice_launch(const View& in, const View& out, cudaStream_t s) {
  const long total = static_cast<long>(in.N) * in.H * in.W * (in.C / 8);
  copy_slice_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out);
  TR_CUDA(cudaGetLastError());
}

void export_nchw_launch(const View& in, int C, float* out, cudaStream_t s) {
  const long total = static_cast<long>(in.N) * C * in.H * in.W;
  export_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, C, out);
  TR_CUDA(cudaGetLastError());
}

void export_nchw_f32_launch(const float* in, int N, int H, int W, int cs, int coff, int C,
                            float* out, int softmax_pairs, cudaStream_t s) {
  const long total = static_cast<long>(N) * C * H * W;
  export_nchw_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
      in, N, H, W, cs, coff, C, out, softmax_pairs);
  TR_CUDA(cudaGetLastError());
}

void l2_normalize_launch(const float* in, float* out, int N, int D, cudaStream_t s) {
  if (N == 0) return;
  l2_normalize_kernel<<<N, 128, 0, s>>>(in, out, D);
  TR_CUDA(cudaGetLastError());
}

void resize_bilinear_u8_launch(const uint8_t* src, int N, int H, int W, uint8_t* dst, int h,
                               int w, cudaStream_t s) {
  const long total = static_cast<long>(N) * h * ((w + kResizePx - 1) / kResizePx);
  const double sy = 1.0 / (double(h) / H), sx = 1.0 / (double(w) / W);     // OpenCV: 1/inv_scale
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (sx >= 4.0) resize_u8_kernel<true><<<grid, 256, 0, s>>>(src, N, H, W, dst, h, w, sy, sx);
  else resize_u8_kernel<false><<<grid, 256, 0, s>>>(src, N, H, W, dst, h, w, sy, sx);
  TR_CUDA(cudaGetLastError());
}

}  // namespace trb
