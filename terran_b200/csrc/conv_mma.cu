// Warp-level tensor-core kernels (mma.sync m16n8k16, fp16 x fp16 -> fp32) for the layers
// of RetinaFace's mobilenet-0.25 backbone, FPN refiner, context modules and heads
// (retinaface/model.py:6-316).  Those layers have 8..256 channels on maps of 13x24 ..
// 208x370: a few MB in and out each, i.e. HBM/L2-bound byte shuffling with a thin GEMM in
// the middle, where a tcgen05 launch (TMEM allocation, mbarrier rings, 128-row tiles) costs
// more than the layer itself.  Two shapes of work, one kernel template:
//
//   SEP    depthwise 3x3 (stride 1|2) + BN + ReLU  ->  1x1 conv + BN + ReLU   in ONE pass:
//          ConvSepBlock's sep_block followed by the next block's conv_block
//          (model.py:6-50); the depthwise result goes straight into the A operand of the
//          1x1 GEMM through shared memory and never touches HBM.
//   DENSE  1x1 or 3x3 stride-1 conv + scale/shift (+ReLU) (+residual, optionally read at
//          (h/2, w/2): the FPN's nearest x2 up-sample-add, model.py:213-236) with fp16 or
//          fp32 (heads) output.
//
// Every warp owns a strip of output pixels (1 x 16/32/64 for SEP, 2 x 16 for DENSE) and runs
// load -> (depthwise) -> MMA -> epilogue on its own slice of shared memory with warp-level
// synchronisation only, so the SM overlaps the phases of different warps; the filters of the
// layer are staged once per CTA (persistent grid).  Activations are read with plain (L1
// cached) 16-byte loads over the NHWC channel axis; the nine taps of neighbouring rows hit L1.
#include <algorithm>

#include "common.cuh"

namespace trb {

namespace {

#ifndef TRB_MMA_SMALL_CTAS
#define TRB_MMA_SMALL_CTAS 2   /* 3 measured slower (profiles/r01_retinaface_mma.txt) */
#endif

enum { MODE_DENSE1 = 0, MODE_DENSE3 = 1, MODE_SEP1 = 2, MODE_SEP2 = 3 };

struct MmaParams {
  View in, out, res;
  float* out_f32;
  const __half* w;                  // [COUT][taps][CIN] fp16
  const float* scale; const float* shift;
  const __half* dw_w;               // [3][3][CIN] fp16 (SEP)
  const float* dw_scale; const float* dw_shift;
  int res_up2, act, H_out, W_out, cout_store;
  int units_per_image;              // warp strips per image along H
  int total_units;                  // N * units_per_image
  int tiles_w, total_tiles;         // CTA tiles: WARPS consecutive strips x one column block
};

template <int CIN, int COUT, int MODE>
struct Geo {
  static constexpr bool SEP = MODE >= MODE_SEP1;
  static constexpr int K = MODE == MODE_DENSE3 ? 3 : 1;      // filter size of the GEMM stage
  static constexpr int TAPS = K * K;
  static constexpr int KP = CIN < 16 ? 16 : CIN;             // GEMM K per tap (zero padded)
  static constexpr int TR = SEP ? 1 : 2;                     // strip rows
  static constexpr int TC = SEP ? (CIN == 8 ? 64 : CIN == 16 ? 32 : 16) : 16;   // strip columns
  static constexpr int MT = TR * TC / 16;                    // m16 tiles per strip
  static constexpr int NT = COUT / 8;                        // n8 tiles
  static constexpr int NCH = (NT * MT > 16) ? 16 / MT : NT;  // n8 tiles per accumulator pass
  static constexpr int WARPS = (CIN >= 256 && COUT >= 256) ? 4 : 8;
  static constexpr int AR = TR + K - 1, AC = TC + K - 1;     // A tile incl. filter halo (pixels)
  static constexpr int AP = KP + 8;                          // pitches in halfs: +16 B keeps
  static constexpr int OP = COUT + 8;                        // ldmatrix / fragment accesses
  static constexpr int WP = TAPS * KP + 8;                   // free of bank conflicts
  static constexpr int A_HALFS = AR * AC * AP;
  static constexpr int O_HALFS = TR * TC * OP;
  // The output staging tile may reuse the A tile once the last MMA has read it.
  static constexpr bool UNION = NT <= NCH && CIN >= 16;
  static constexpr int WARP_HALFS = UNION ? (A_HALFS > O_HALFS ? A_HALFS : O_HALFS) : A_HALFS + O_HALFS;
  static constexpr int PAR_FLOATS = 2 * COUT + (SEP ? ((2 * CIN + (9 * CIN + 1) / 2 + 3) & ~3) : 0);
  static constexpr size_t SMEM = size_t(COUT) * WP * 2 + size_t(PAR_FLOATS) * 4 + size_t(WARPS) * WARP_HALFS * 2;
  // Two resident CTAs (16 warps) per SM where shared memory allows: caps registers at 128.
  // (3 resident CTAs for the small-channel fused layers were measured slower than 2)
  static constexpr int MIN_CTAS = (SEP && CIN <= 32 && SMEM <= 72 * 1024) ? TRB_MMA_SMALL_CTAS
                                  : (WARPS == 8 && SMEM <= 110 * 1024) ? 2 : 1;
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Four 8x8 b16 matrices; lane l supplies the row address of matrix l/8, row l%8.  With
// row = l & 15 and k offset (l >> 4) * 8 the four results are exactly the A fragment
// (a0: rows 0-7 k 0-7, a1: rows 8-15 k 0-7, a2: rows 0-7 k 8-15, a3: rows 8-15 k 8-15).
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const __half* p) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// 16-byte asynchronous global -> shared copy; src_bytes = 0 zero-fills the destination (the
// conv's zero padding).  No registers are held while the copy is in flight, so a warp has its
// whole tile outstanding at once.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Out-of-bounds depthwise taps read this granule instead of branching around the load.
__device__ uint4 g_zero_granule;

// acc += x * w on a packed fp16 pair with fp32 accumulation: sm_100's mixed-precision FMA
// (SASS FHFMA with .H0/.H1 operand selectors).  The fp16 product is exact in fp32, so this is
// bit-identical to convert-then-FFMA at half the instructions.
__device__ __forceinline__ void fhfma2(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n .reg .b16 xl, xh, wl, wh;\n mov.b32 {xl, xh}, %2;\n mov.b32 {wl, wh}, %3;\n"
      " fma.rn.f32.f16 %0, xl, wl, %0;\n fma.rn.f32.f16 %1, xh, wh, %1;\n}"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}

template <bool V>
struct BoolC { static constexpr bool value = V; };

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(Geo<CIN, COUT, MODE>::WARPS * 32, Geo<CIN, COUT, MODE>::MIN_CTAS)
conv_mma_kernel(const MmaParams p) {
  using G = Geo<CIN, COUT, MODE>;
  constexpr int THREADS = G::WARPS * 32;
  extern __shared__ __align__(16) uint8_t smem[];
  __half* s_w = reinterpret_cast<__half*>(smem);
  float* s_par = reinterpret_cast<float*>(smem + size_t(COUT) * G::WP * 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  __half* s_a = reinterpret_cast<__half*>(s_par + G::PAR_FLOATS) + warp * G::WARP_HALFS;
  __half* s_o = G::UNION ? s_a : s_a + G::A_HALFS;
  // SEP parameter block after scale/shift: dw scale [CIN], dw shift [CIN] (fp32), dw filter [9][CIN] (fp16)
  float* s_dws = s_par + 2 * COUT;
  __half* s_dww = reinterpret_cast<__half*>(s_dws + 2 * CIN);

  // ---- prologue: filters and per-channel parameters.  Constants of the net, so this part
  // may overlap the tail of the previous kernel (programmatic dependent launch).
  {
    constexpr int GK = CIN / 8, GP = G::KP / 8;       // 16-byte granules per (cout, tap): global / smem
    for (int i = threadIdx.x; i < COUT * G::TAPS * GP; i += THREADS) {
      const int gch = i % GP, tap = (i / GP) % G::TAPS, n = i / (GP * G::TAPS);
      __half* dst = s_w + n * G::WP + tap * G::KP + gch * 8;
      if (gch < GK) cp_async16(dst, p.w + (size_t(n) * G::TAPS + tap) * CIN + gch * 8, 16);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = threadIdx.x; i < COUT / 4; i += THREADS) {
      cp_async16(s_par + 4 * i, p.scale + 4 * i, 16);
      cp_async16(s_par + COUT + 4 * i, p.shift + 4 * i, 16);
    }
    if constexpr (G::SEP) {
      for (int i = threadIdx.x; i < CIN / 4; i += THREADS) {
        cp_async16(s_dws + 4 * i, p.dw_scale + 4 * i, 16);
        cp_async16(s_dws + CIN + 4 * i, p.dw_shift + 4 * i, 16);
      }
      for (int i = threadIdx.x; i < 9 * CIN / 8; i += THREADS) cp_async16(s_dww + 8 * i, p.dw_w + 8 * i, 16);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");   // activations are only read below

  const float* s_scale = s_par;
  const float* s_shift = s_par + COUT;
  const float act_lo = p.act == ACT_RELU ? 0.f : -INFINITY;
  // per-lane bases of the fragment accesses; everything else is a compile-time offset
  const __half* a_lane = s_a + (lane & 15) * G::AP + (lane >> 4) * 8;
  const __half* b_lane = s_w + ((lane >> 4) * 8 + (lane & 7)) * G::WP + ((lane >> 3) & 1) * 8;
  __half* o_lane = s_o + g * G::OP + 2 * t;

  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    // Let the next kernel's CTAs become resident only while this CTA works on its last tile.
    if (tile + static_cast<int>(gridDim.x) >= p.total_tiles)
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tw = tile % p.tiles_w;
    const int unit = (tile / p.tiles_w) * G::WARPS + warp;
    if (unit >= p.total_units) continue;               // warp-uniform
    const int n = unit / p.units_per_image;
    const int oh0 = (unit - n * p.units_per_image) * G::TR;
    const int ow0 = tw * G::TC;
    const __half* img = p.in.ptr + size_t(n) * p.in.H * p.in.W * p.in.cs + p.in.coff;

    // ------------------------------------------------------------ phase 1: the A tile
    if constexpr (!G::SEP) {
      constexpr int GK = CIN / 8, V = G::AR * G::AC * GK, PAD = G::K / 2;
#pragma unroll 4
      for (int i = 0; i < (V + 31) / 32; ++i) {
        const int v = lane + 32 * i;
        const int gch = v % GK, px = v / GK, col = px % G::AC, row = px / G::AC;
        const int ih = oh0 + row - PAD, iw = ow0 + col - PAD;
        const bool ok = ih >= 0 && ih < p.in.H && iw >= 0 && iw < p.in.W;   // else: zero padding
        if (v < V)
          cp_async16(s_a + (row * G::AC + col) * G::AP + gch * 8,
                     img + (ok ? (ih * p.in.W + iw) * p.in.cs + gch * 8 : 0), ok ? 16 : 0);
      }
      cp_async_wait_all();
    } else {
      // Depthwise 3x3 + BN + ReLU straight into the A tile.  The input is a dense NHWC tensor
      // (channel stride == CIN, checked by the launcher), so every tap of a lane is a
      // compile-time offset from one pointer per filter row.
      constexpr int S = MODE == MODE_SEP2 ? 2 : 1;
      constexpr int GK = CIN / 8, PL = 32 / GK, J = G::TC / PL, JC = J < 4 ? J : 4;
      constexpr int RB = JC <= 2 ? 3 : 1;              // filter rows per batch of loads
      const int gch = lane % GK, pl = lane / GK;       // lane = (pixel lane, 8-channel granule)
      const __half* src = img + gch * 8;
      const float4 sc0 = *reinterpret_cast<const float4*>(s_dws + gch * 8);
      const float4 sc1 = *reinterpret_cast<const float4*>(s_dws + gch * 8 + 4);
      const float4 sh0 = *reinterpret_cast<const float4*>(s_dws + CIN + gch * 8);
      const float4 sh1 = *reinterpret_cast<const float4*>(s_dws + CIN + gch * 8 + 4);
      const float dsc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
      const float dsh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
      const bool interior = oh0 * S >= 1 && oh0 * S + 1 < p.in.H && ow0 * S >= 1 &&
                            (ow0 + G::TC - 1) * S + 1 < p.in.W;   // warp-uniform
      auto depthwise = [&](auto interior_c) {
        constexpr bool INTERIOR = decltype(interior_c)::value;
#pragma unroll 1
        for (int j0 = 0; j0 < J; j0 += JC) {
          float acc[JC][8];
#pragma unroll
          for (int j = 0; j < JC; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
          const int owl = ow0 + pl + PL * j0;            // this lane's first pixel of the chunk
#pragma unroll
          for (int r0 = 0; r0 < 3; r0 += RB) {
            // All loads of a batch are unconditional and issued before the first use, so the
            // warp has RB*3*JC requests in flight.
            uint4 xv[RB][3][JC];
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
              const int ih = oh0 * S + r0 + rr - 1;
              const __half* rowp = src + (ih * p.in.W + owl * S - 1) * CIN;
              const bool rok = ih >= 0 && ih < p.in.H;
#pragma unroll
              for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int j = 0; j < JC; ++j) {
                  const __half* ptr = rowp + (PL * j * S + s) * CIN;
                  if constexpr (!INTERIOR) {
                    const int iw = (owl + PL * j) * S + s - 1;
                    if (!(rok && iw >= 0 && iw < p.in.W)) ptr = reinterpret_cast<const __half*>(&g_zero_granule);
                  }
                  xv[rr][s][j] = *reinterpret_cast<const uint4*>(ptr);
                }
            }
#pragma unroll
            for (int rr = 0; rr < RB; ++rr)
#pragma unroll
              for (int s = 0; s < 3; ++s) {
                const uint4 wv = *reinterpret_cast<const uint4*>(s_dww + ((r0 + rr) * 3 + s) * CIN + gch * 8);
                const uint32_t wq[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int j = 0; j < JC; ++j) {
                  const uint32_t xq[4] = {xv[rr][s][j].x, xv[rr][s][j].y, xv[rr][s][j].z, xv[rr][s][j].w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) fhfma2(acc[j][2 * q], acc[j][2 * q + 1], xq[q], wq[q]);
                }
              }
          }
#pragma unroll
          for (int j = 0; j < JC; ++j) {
            uint4 ov;
            __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              oh2[q] = __floats2half2_rn(fmaxf(fmaf(acc[j][2 * q], dsc[2 * q], dsh[2 * q]), 0.f),
                                         fmaxf(fmaf(acc[j][2 * q + 1], dsc[2 * q + 1], dsh[2 * q + 1]), 0.f));
            __half* dst = s_a + (pl + PL * (j0 + j)) * G::AP + gch * 8;
            *reinterpret_cast<uint4*>(dst) = ov;
            if constexpr (CIN < 16)     // K is padded to one 16-wide MMA step
              *reinterpret_cast<uint4*>(dst + 8) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      };
      if (interior) depthwise(BoolC<true>{});
      else depthwise(BoolC<false>{});
    }
    __syncwarp();

    // ------------------------------------------------------------ phase 2 + 3: GEMM, epilogue
#pragma unroll
    for (int nc = 0; nc < G::NT; nc += G::NCH) {
      float acc[G::MT][G::NCH][4];
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
        for (int j = 0; j < G::NCH; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;
#pragma unroll 1
      for (int tap = 0; tap < G::TAPS; ++tap) {
        const int r = tap / G::K, s = tap - r * G::K;
        const __half* a_tap = a_lane + (r * G::AC + s) * G::AP;
        const __half* b_tap = b_lane + nc * 8 * G::WP + tap * G::KP;
#pragma unroll
        for (int ks = 0; ks < G::KP / 16; ++ks) {
          uint32_t a[G::MT][4];
#pragma unroll
          for (int mt = 0; mt < G::MT; ++mt) {
            constexpr int CM = G::TC / 16;
            ldmatrix_x4(a[mt], a_tap + ((mt / CM) * G::AC + (mt % CM) * 16) * G::AP + ks * 16);
          }
#pragma unroll
          for (int j = 0; j < G::NCH; j += 2) {
            uint32_t b[4];                               // (b0, b1) of n-tiles j and j + 1
            ldmatrix_x4(b, b_tap + j * 8 * G::WP + ks * 16);
#pragma unroll
            for (int mt = 0; mt < G::MT; ++mt) {
              mma16816(acc[mt][j], a[mt], b[0], b[1]);
              mma16816(acc[mt][j + 1], a[mt], b[2], b[3]);
            }
          }
        }
      }
      if constexpr (G::UNION) __syncwarp();            // every lane's A reads are done
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt) {
        constexpr int CM = G::TC / 16;
        const int row_m = mt / CM, col0 = (mt % CM) * 16;
        const int oh = oh0 + row_m;
#pragma unroll
        for (int j = 0; j < G::NCH; ++j) {
          const int c = (nc + j) * 8 + 2 * t;
          const float2 sc = *reinterpret_cast<const float2*>(s_scale + c);
          const float2 sh = *reinterpret_cast<const float2*>(s_shift + c);
          const float y[4] = {fmaxf(fmaf(acc[mt][j][0], sc.x, sh.x), act_lo), fmaxf(fmaf(acc[mt][j][1], sc.y, sh.y), act_lo),
                              fmaxf(fmaf(acc[mt][j][2], sc.x, sh.x), act_lo), fmaxf(fmaf(acc[mt][j][3], sc.y, sh.y), act_lo)};
#pragma unroll
          for (int h = 0; h < 2; ++h) {                  // fragment rows g and g + 8
            float y0 = y[2 * h], y1 = y[2 * h + 1];
            if constexpr (!G::SEP) {
              const int ow = ow0 + col0 + g + 8 * h;
              const bool ok = oh < p.H_out && ow < p.W_out;
              if (p.res.ptr && ok) {
                const size_t rpix = p.res_up2 ? (size_t(n) * p.res.H + (oh >> 1)) * p.res.W + (ow >> 1)
                                              : (size_t(n) * p.H_out + oh) * p.W_out + ow;
                const float2 rf = __half22float2(
                    *reinterpret_cast<const __half2*>(p.res.ptr + rpix * p.res.cs + p.res.coff + c));
                y0 += rf.x; y1 += rf.y;
              }
              if (p.out_f32) {
                if (ok && c < p.cout_store)
                  *reinterpret_cast<float2*>(p.out_f32 + ((size_t(n) * p.H_out + oh) * p.W_out + ow) * p.out.cs +
                                             p.out.coff + c) = make_float2(y0, y1);
                continue;
              }
            }
            *reinterpret_cast<__half2*>(o_lane + (mt * 16 + 8 * h) * G::OP + (nc + j) * 8) = __floats2half2_rn(y0, y1);
          }
        }
      }
    }
    // ------------------------------------------------------------ coalesced copy-out
    if constexpr (G::SEP) {
      // The strip's TC pixels x COUT channels are one contiguous run of the dense output.
      __syncwarp();
      constexpr int GO = COUT / 8, VO = G::MT * 16 * GO;
      __half* dst = p.out.ptr + ((size_t(n) * p.H_out + oh0) * p.W_out + ow0) * COUT + lane * 8;
      const __half* srcv = s_o + (lane / GO) * G::OP + (lane % GO) * 8;
      const int valid = p.W_out - ow0;                   // pixels of the strip inside the map
#pragma unroll
      for (int i = 0; i < VO / 32; ++i) {
        constexpr int MSTEP = 32 / GO > 0 ? 32 / GO : 1;   // pixels advanced per iteration (GO <= 32)
        if (lane / GO + i * MSTEP < valid)
          *reinterpret_cast<uint4*>(dst + i * 256) = *reinterpret_cast<const uint4*>(srcv + i * MSTEP * G::OP);
      }
    } else if (!p.out_f32) {
      __syncwarp();
      constexpr int GO = COUT / 8, VO = G::MT * 16 * GO;
#pragma unroll 4
      for (int i = 0; i < (VO + 31) / 32; ++i) {
        const int v = lane + 32 * i;
        const int cg = v % GO, m = v / GO, mt = m / 16;
        const int oh = oh0 + mt / (G::TC / 16), ow = ow0 + (mt % (G::TC / 16)) * 16 + m % 16;
        if (v < VO && oh < p.H_out && ow < p.W_out && cg * 8 < p.cout_store)
          *reinterpret_cast<uint4*>(p.out.ptr + ((size_t(n) * p.H_out + oh) * p.W_out + ow) * p.out.cs +
                                    p.out.coff + cg * 8) =
              *reinterpret_cast<const uint4*>(s_o + m * G::OP + cg * 8);
      }
    }
    __syncwarp();                                       // the tile is reused by the next strip
  }
}

template <int CIN, int COUT, int MODE>
void launch_t(MmaParams p, int N, cudaStream_t s) {
  using G = Geo<CIN, COUT, MODE>;
  auto kernel = conv_mma_kernel<CIN, COUT, MODE>;
  p.units_per_image = ceil_div(p.H_out, G::TR);
  p.total_units = N * p.units_per_image;
  p.tiles_w = ceil_div(p.W_out, G::TC);
  const long tiles = static_cast<long>(ceil_div(p.total_units, G::WARPS)) * p.tiles_w;
  TR_CHECK(tiles > 0 && tiles < (1L << 30), "conv_mma: tile count");
  p.total_tiles = static_cast<int>(tiles);
  int dev = 0;
  TR_CUDA(cudaGetDevice(&dev));
  static int ctas[16] = {};                             // resident CTAs per device
  TR_CHECK(dev < 16, "conv_mma: device ordinal");
  if (!ctas[dev]) {
    TR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(G::SMEM)));
    int per_sm = 0, sms = 0;
    TR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, G::WARPS * 32, G::SMEM));
    TR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TR_CHECK(per_sm > 0, "conv_mma: kernel does not fit on an SM");
    ctas[dev] = per_sm * sms;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(std::min<long>(tiles, ctas[dev])));
  cfg.blockDim = dim3(G::WARPS * 32);
  cfg.dynamicSmemBytes = G::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TR_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
}

// (cin_pad, cout_pad) pairs of the reference nets; anything else stays on the other kernels.
#define TRB_DENSE_CASES(X) \
  X(256, 64, 1) X(128, 64, 1) X(64, 64, 1) X(64, 32, 1) X(64, 64, 3) X(64, 32, 3) X(64, 16, 3) X(16, 16, 3)
#define TRB_SEP_CASES(X) \
  X(8, 16, 1) X(16, 32, 2) X(32, 32, 1) X(32, 64, 2) X(64, 64, 1) X(64, 128, 2) X(128, 128, 1) X(128, 256, 2) \
  X(256, 256, 1)

}  // namespace

bool conv_mma_eligible(const ConvArgs& a) {
  if (a.kh != a.kw || a.stride != 1 || a.pad != a.kh / 2) return false;
  if (a.act != ACT_NONE && a.act != ACT_RELU) return false;
  if (a.out2.ptr || a.slope || a.scale2) return false;
  if (a.in.cs % 8 || a.in.coff % 8 || a.out.cs % 8 || a.out.coff % 8) return false;
  if (a.res.ptr && (a.res.cs % 2 || a.res.coff % 2)) return false;
#define X(ci, co, k) if (a.cin_pad == ci && a.cout_pad == co && a.kh == k) return true;
  TRB_DENSE_CASES(X)
#undef X
  return false;
}

void conv_mma_launch(const ConvArgs& a, cudaStream_t s) {
  TR_CHECK(conv_mma_eligible(a), "conv_mma: unsupported layer");
  MmaParams p{};
  p.in = a.in; p.out = a.out; p.res = a.res; p.res_up2 = a.res_up2; p.out_f32 = a.out_f32;
  p.w = a.w; p.scale = a.scale; p.shift = a.shift; p.act = a.act;
  p.H_out = a.H_out; p.W_out = a.W_out; p.cout_store = a.cout_store;
#define X(ci, co, k)                                                              \
  if (a.cin_pad == ci && a.cout_pad == co && a.kh == k)                            \
    return launch_t<ci, co, k == 3 ? MODE_DENSE3 : MODE_DENSE1>(p, a.in.N, s);
  TRB_DENSE_CASES(X)
#undef X
}

bool sep_mma_eligible(const SepArgs& a) {
  // dense input and output tensors: the kernel folds the channel strides into immediates
  if (a.in.cs != a.cin_pad || a.in.coff || a.out.cs != a.cout_pad || a.out.coff || a.cout_store != a.cout_pad)
    return false;
  if (a.act != ACT_NONE && a.act != ACT_RELU) return false;
#define X(ci, co, st) if (a.cin_pad == ci && a.cout_pad == co && a.stride == st) return true;
  TRB_SEP_CASES(X)
#undef X
  return false;
}

void sep_mma_launch(const SepArgs& a, cudaStream_t s) {
  TR_CHECK(sep_mma_eligible(a), "sep_mma: unsupported (channels, stride) combination");
  const int H_out = (a.in.H + 2 - 3) / a.stride + 1, W_out = (a.in.W + 2 - 3) / a.stride + 1;
  TR_CHECK(H_out == a.out.H && W_out == a.out.W, "sep_mma: output dims");
  MmaParams p{};
  p.in = a.in; p.out = a.out; p.w = a.w; p.scale = a.scale; p.shift = a.shift; p.act = a.act;
  p.dw_w = a.dw_w16; p.dw_scale = a.dw_scale; p.dw_shift = a.dw_shift;
  p.H_out = H_out; p.W_out = W_out; p.cout_store = a.cout_store;
#define X(ci, co, st)                                                             \
  if (a.cin_pad == ci && a.cout_pad == co && a.stride == st)                       \
    return launch_t<ci, co, st == 2 ? MODE_SEP2 : MODE_SEP1>(p, a.in.N, s);
  TRB_SEP_CASES(X)
#undef X
}

}  // namespace trb
