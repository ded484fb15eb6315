// Implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a).
//
//   D[pixels, cout] = sum over taps (r,s) and channel chunks of
//                     A_tap[pixels, KC] * W_tap[cout, KC]^T
//
// Activations are NHWC fp16.  For every filter tap the A operand is just a
// shifted window of the input, so a plain *tiled* TMA load of the box
// {KC channels, bw, bh, bn} at coordinates shifted by the tap gives the
// im2col tile directly in the canonical K-major swizzled layout, and TMA's
// out-of-bounds zero fill implements the convolution's zero padding.  Stride-2
// convolutions view the input as {2C, W/2, 2, H/2, N} so a tap selects a
// (row parity, column parity) plane and the window stays dense.
//
// Warp roles (320 threads, one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer  (whole warp converged, one elected lane issues)
//   warp 1      TMEM allocator + tcgen05.mma issuer (same)
//   warps 2-9   epilogue: tcgen05.ld -> scale/shift/activation/residual -> HBM
//               (two warps per TMEM lane quarter, alternating 16-column chunks)
// Pipelines: smem ring (full/empty mbarriers) between TMA and MMA, and a
// double-buffered TMEM accumulator (tmem_full/tmem_empty) between MMA and the
// epilogue, so the epilogue of tile i overlaps the main loop of tile i+1.
// One ring stage holds `sub` (A,B) k-blocks: the issue loops run on a single
// warp whose per-iteration latency (~400 cycles of dependent instructions and
// mbarrier round trips, measured) must stay below the MMA time queued per
// iteration, so a stage carries >= 512 cycles of tensor work where smem allows.
//
// Replaces the cuDNN/ATen convolutions behind every nn.Conv2d of the
// reference's models (retinaface/model.py, arcface/model.py, openpose/model.py).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

#include <cstdlib>
#include <vector>

namespace trb {

namespace {

constexpr int kThreads = 320;
constexpr int kThreads1 = 352;              // single-CTA kernel: + a second MMA-issuing warp (warp 10)
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 8;
constexpr int kMaxSub = 4;
constexpr uint32_t kSmemBudget = 196 * 1024;
constexpr int kMaxParamChannels = 1024;     // per-channel epilogue params staged in smem

struct TcParams {
  int bw, bh, bn, rows;
  int tiles_w, tiles_h, tiles_n, n_tiles, total_tiles;
  int N_tile;
  int N, H_out, W_out;
  int kh, kw, pad, stride;
  int KC, kchunks, cin_pad;
  int in_coff, in_cs;
  int k_blocks;            // taps * kchunks
  int sub, iters;          // k-blocks per stage, stage iterations per tile
  int stages;
  uint32_t a_bytes, b_bytes, sub_bytes, stage_bytes;
  uint32_t sbo_bytes, layout_type, idesc;
  uint32_t tmem_cols;
  int cout_pad;
  // epilogue
  const float* scale; const float* shift; const float* slope;
  const float* scale2; const float* shift2;
  // Border-class shifts (9 x cout_pad, class = 3 * row class + column class; 0 first, 1 inner,
  // 2 last row / column): a 3x3 pad-1 conv whose input carries a folded per-channel affine has
  // an input-independent term that depends on which taps are in bounds (arcface/model.py:11-35).
  const float* shift9;
  int param_rows;            // rows of cout_pad floats staged in smem: 5, or 14 with shift9
  int act;
  __half* out; int out_cs, out_coff, cout_store;
  __half* out2; int out2_cs, out2_coff;
  const __half* res; int res_cs, res_coff, res_up2, res_H, res_W;
  float* out_f32;
  int* err;   // device flag set on a pipeline timeout
  // Stream-K: the CTAs split the launch's (tile, stage-iteration) sequence into equal
  // contiguous ranges instead of whole tiles, so no SM idles in the last scheduling round.
  // A range may start inside a tile (its TAIL: the raw fp32 accumulators go to sk_ws[cta] and
  // sk_flags[cta] is raised) and end inside one (its HEAD: the CTA adds the tail partial of
  // CTA + 1 — computed first thing by that CTA — and runs the normal epilogue).
  int sk;
  float* sk_ws;
  int* sk_flags;
  int pdl_late;   // 1: let the next kernel start when this CTA begins its last tile, not at once
  int debug;  // timing experiments only: 1 skip TMA loads, 2 skip MMAs, 4 skip epilogue stores, 8 skip epilogue
  // halo mode: the input patch of a tile (+ filter halo) is loaded ONCE per channel chunk and
  // every tap's A operand is a shifted UMMA descriptor into it.  1: tile 8w x 16h, 2: 16w x 8h.
  int halo, taps, iters_kc;
  // swap mode (128 output channels): the FILTER tile is the UMMA A operand (M = 128 couts) and
  // the pixel tile the B operand (N = rows <= 256 pixels, a band of full-width rows), so one
  // instruction does N = 240 instead of 128 columns of work per 128 x 64 filter block read from
  // shared memory.  TMEM holds D^T: lane = cout, column = pixel.
  // 2: warps 1 and 10 issue alternate ring stages into two accumulators the epilogue adds
  // (non-halo path, N_tile <= 128).  The tensor pipe queues only ~3 instructions, so while ONE
  // issuer does its per-stage barrier/commit work (~600 cycles) the pipe runs dry; a second
  // issuer's MMAs fill that time.
  int issuers;
  int swap;
  int rotate;                // per-tile rotation of the tap / k-block order (L2 hot-spot avoidance)
  int acc_cols;              // TMEM column stride between the two accumulators
  int cta2, pair_units;      // CTA-pair kernel: units = ceil(m_tiles / 2) * n_tiles
  uint32_t idesc2;
  uint32_t patch_bytes, patch_tx, ring_off;
  // Resident-filter halo mode (one channel chunk, one filter tile, taps * b_bytes small — the
  // 64 -> 64 3x3 layers): the whole filter bank is loaded ONCE per CTA into the ring area and
  // never released; the only stream is one input patch per tile through `npatch` buffers
  // (requested up to npatch - 1 tiles ahead), and a tile costs the issuing thread one patch
  // wait + taps * KSTEPS MMAs + two commits — no ring hand-shake.  Measured before (clock64
  // instrumentation, profiles/r02_resident_filters.txt): 5 080 cycles per 128-pixel tile, of
  // which ~4 500 is the latency of the single patch in flight and 2 700 the MMA issue.
  int resident, npatch;
  // Fused 2x2 / stride-2 max-pool (halo tiles only: 8 x 16 / 16 x 8 pixels, the four pixels of a
  // window are lanes l, l ^ 1, l ^ 8, l ^ 9 of one epilogue warp): the pooled tensor is written
  // INSTEAD of the full-resolution one.
  int pw;      // halo patch: pixels per patch line along the 8-pixel axis (8 + 2 pad, or 16)
  int pool;
  __half* pool_out; int pool_cs, pool_coff, pool_H, pool_W;
};

// Walks the tile segments of one CTA: whole tiles with the static stride, or the CTA's
// contiguous share of all stage iterations under stream-K.
struct TileWalk {
  long long pos, end;
  int tile;
};
__device__ __forceinline__ TileWalk walk_begin(const TcParams& p) {
  TileWalk w;
  const long long total = static_cast<long long>(p.total_tiles) * p.iters;
  w.pos = p.sk ? total * blockIdx.x / gridDim.x : 0;
  w.end = p.sk ? total * (blockIdx.x + 1) / gridDim.x : 0;
  w.tile = blockIdx.x;
  return w;
}
__device__ __forceinline__ bool walk_next(const TcParams& p, TileWalk& w, int& tile, int& it0, int& it1) {
  if (p.sk) {
    if (w.pos >= w.end) return false;
    tile = static_cast<int>(w.pos / p.iters);
    it0 = static_cast<int>(w.pos - static_cast<long long>(tile) * p.iters);
    it1 = static_cast<int>(min(static_cast<long long>(p.iters), it0 + (w.end - w.pos)));
    w.pos += it1 - it0;
    return true;
  }
  if (w.tile >= p.total_tiles) return false;
  tile = w.tile; it0 = 0; it1 = p.iters;
  w.tile += gridDim.x;
  return true;
}
__device__ __forceinline__ bool walk_done(const TcParams& p, const TileWalk& w) {
  return p.sk ? w.pos >= w.end : w.tile >= p.total_tiles;
}

struct EpiCtx {
  const float* sp;       // smem params: [5][cout_pad] scale, shift, slope, scale2, shift2 (+ [9] border shifts)
  const float* shp;      // this thread's shift row
  int cpad;
  bool valid;
  long pix, rpix;
  bool pool_store;       // this lane writes the pooled pixel of its 2x2 window
  long ppix;
};

// Scale/shift/activation/residual/store of 16 consecutive output channels.
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, const EpiCtx& e,
                                               const uint32_t (&v)[16], int cbase) {
  // (fused pooling shuffles across the warp: lanes of invalid pixels take part, with -inf)
  if ((!e.valid && !p.pool) || cbase >= p.cout_store || (p.debug & 8)) return;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int c = cbase + g * 8;
    if (c >= p.cout_store) break;
    float y[8];
    const float4* sc4 = reinterpret_cast<const float4*>(e.sp + c);
    const float4* sh4 = reinterpret_cast<const float4*>(e.shp + c);
    const float4 s0 = sc4[0], s1 = sc4[1], b0 = sh4[0], b1 = sh4[1];
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = fmaf(__uint_as_float(v[g * 8 + j]), sc[j], sh[j]);
    if (p.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
    } else if (p.act == ACT_PRELU) {
      const float4* sl4 = reinterpret_cast<const float4*>(e.sp + 2 * e.cpad + c);
      const float4 l0 = sl4[0], l1 = sl4[1];
      const float sl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = y[j] >= 0.f ? y[j] : y[j] * sl[j];
    }
    if (p.pool) {
      // max over the window's four lanes on packed halves (rounding is monotonic: the same
      // value as pooling the rounded full-resolution tensor), then one lane of the four stores
      uint32_t h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 t2 = __floats2half2_rn(e.valid ? y[2 * j] : -65504.f, e.valid ? y[2 * j + 1] : -65504.f);
        h[j] = *reinterpret_cast<const uint32_t*>(&t2);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t o = __shfl_xor_sync(0xffffffffu, h[j], 1);
        __half2 m = __hmax2(*reinterpret_cast<__half2*>(&h[j]), *reinterpret_cast<__half2*>(&o));
        h[j] = *reinterpret_cast<uint32_t*>(&m);
        o = __shfl_xor_sync(0xffffffffu, h[j], 8);
        m = __hmax2(*reinterpret_cast<__half2*>(&h[j]), *reinterpret_cast<__half2*>(&o));
        h[j] = *reinterpret_cast<uint32_t*>(&m);
      }
      if (e.pool_store && !(p.debug & 4))
        *reinterpret_cast<uint4*>(p.pool_out + e.ppix * p.pool_cs + p.pool_coff + c) = make_uint4(h[0], h[1], h[2], h[3]);
      continue;
    }
    if (p.res) {
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.res + e.rpix * p.res_cs +
                                                            p.res_coff + c));
      const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(rh[j]);
        y[2 * j] += f.x;
        y[2 * j + 1] += f.y;
      }
    }
    if (p.debug & 4) continue;
    if (p.out_f32) {
      float4* o = reinterpret_cast<float4*>(p.out_f32 + e.pix * p.out_cs + p.out_coff + c);
      o[0] = make_float4(y[0], y[1], y[2], y[3]);
      o[1] = make_float4(y[4], y[5], y[6], y[7]);
    } else {
      uint4 ov;
      __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh2[j] = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
      *reinterpret_cast<uint4*>(p.out + e.pix * p.out_cs + p.out_coff + c) = ov;
    }
    if (p.out2) {
      const float4* a4 = reinterpret_cast<const float4*>(e.sp + 3 * e.cpad + c);
      const float4* t4 = reinterpret_cast<const float4*>(e.sp + 4 * e.cpad + c);
      const float4 a0 = a4[0], a1 = a4[1], t0 = t4[0], t1 = t4[1];
      const float s2[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float h2[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
      uint4 ov;
      __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        oh2[j] = __floats2half2_rn(fmaf(y[2 * j], s2[2 * j], h2[2 * j]),
                                   fmaf(y[2 * j + 1], s2[2 * j + 1], h2[2 * j + 1]));
      *reinterpret_cast<uint4*>(p.out2 + e.pix * p.out2_cs + p.out2_coff + c) = ov;
    }
  }
}

// Swap mode: this thread owns ONE output channel (TMEM lane) and v holds 16 consecutive
// accumulator columns = pixels of it.  A warp's 32 lanes write 32 consecutive channels of a
// pixel (64 B runs).  Column n of the tile is pixel pix0 + n of a band of full-width rows, or,
// with the halo patch (columns of 8 rows), pixel (h0 + (n & 7), w0 + (n >> 3)).
struct SwapTile {
  long pix0;          // first pixel of the band / of image `on`
  int valid;          // band: pixels inside the map
  int h0, w0;         // halo patch tile origin
};
__device__ __forceinline__ void epilogue_chunk_swap(const TcParams& p, const float* sp, int cpad,
                                                    const uint32_t (&v)[16], int cout,
                                                    const SwapTile& t, int n0) {
  if (cout >= p.cout_store || (p.debug & 8)) return;
  const float sc = sp[cout], sh = sp[cpad + cout], sl = sp[2 * cpad + cout];
  __half* o = p.out + p.out_coff + cout;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float y = fmaf(__uint_as_float(v[j]), sc, sh);
    if (p.act == ACT_RELU) y = fmaxf(y, 0.f);
    else if (p.act == ACT_PRELU) y = y >= 0.f ? y : y * sl;
    long pix;
    bool ok;
    if (p.halo) {
      const int h = t.h0 + (j & 7), w = t.w0 + ((n0 + j) >> 3);     // n0 is a multiple of 16
      ok = h < p.H_out && w < p.W_out;
      pix = t.pix0 + static_cast<long>(h) * p.W_out + w;
    } else {
      ok = n0 + j < t.valid;
      pix = t.pix0 + n0 + j;
    }
    if (ok && !(p.debug & 4)) o[pix * p.out_cs] = __float2half_rn(y);
  }
}

// One of TWO MMA-issuing warps (non-halo path): issuer `w` owns the ring stages of every
// other GLOBAL iteration (g = w, w + 2, ...), waits for their full barriers, issues their MMAs
// and commits their empty barriers itself (tcgen05.commit tracks the issuing thread's own
// instructions only).  Each issuer accumulates into ITS OWN TMEM tile (columns w * N_tile of
// the accumulator buffer) and the epilogue adds the two: no instruction of one thread ever
// depends on one of the other, and the result does not depend on how the two streams
// interleave in the tensor pipe (bit-reproducible).  tmem_full is initialised with count 2;
// an issuer without work in a (one-iteration) segment arrives plainly.
template <int KSTEPS>
__device__ __forceinline__ void mma_issuer_alternate(const TcParams& p, int w, uint32_t ring,
                                                     uint32_t tmem_base, uint32_t bars, int* prog) {
  const int lane_id = threadIdx.x & 31;
  int mine = 0;                                  // my full-barrier waits that have passed
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t desc_hi = umma_desc_hi(p.sbo_bytes, p.layout_type);
  const uint32_t a_lo0 = umma_desc_lo(ring);
  const uint32_t stage_step = p.stage_bytes >> 4, sub_step = p.sub_bytes >> 4, b_off = p.a_bytes >> 4;
  // my stage / phase follow the global iteration counter in steps of two
  int stage = w % p.stages;
  uint32_t phase = (w / p.stages) & 1u;
  int g = 0, tile_it = 0;                       // global iteration index of the segment start
  TileWalk walk = walk_begin(p);
  int tile, it0, it1;
  for (; walk_next(p, walk, tile, it0, it1); ++tile_it) {
    const int acc = tile_it & 1;
    const uint32_t acc_phase = (tile_it >> 1) & 1u;
    const int n_it = it1 - it0;
    const int first = ((g & 1) == w) ? 0 : 1;   // my first iteration within the segment
    const int gstart = g;
    g += n_it;
    mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err, 2);   // (also: never lap the previous use)
    if (first >= n_it) {                         // a one-iteration segment owned by the other issuer
      if (elect_one()) mbar_arrive(tfull_bar(acc));
      __syncwarp();
      continue;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t d_tmem = tmem_base + acc * p.acc_cols + w * p.N_tile;
    for (int i = first; i < n_it; i += 2) {
      const int gi = gstart + i;                 // global iteration index (ring position)
      const int nsub = min(p.sub, p.k_blocks - (it0 + i) * p.sub);
      // With an odd number of stages a stage changes owner every ring cycle: the stage's
      // previous phase (global iteration gi - stages) belongs to the OTHER issuer.  If that one
      // has not seen its data yet, the barrier is still one phase behind and a parity wait for
      // this phase would pass spuriously.  Each issuer therefore publishes how many of its
      // full-barrier waits have passed, and waits for the other's count where it matters.
      if ((p.stages & 1) && gi >= p.stages) {
        const int need = (gi - p.stages - (w ^ 1)) / 2 + 1;       // the other's iteration ordinal
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile int*>(prog + (w ^ 1)) < need) {
          if (clock64() - t0 > 4000000000LL) {
            if (p.err) atomicExch(p.err, 8);
            __threadfence_system();
            __trap();
          }
        }
        __threadfence_block();
      }
      mbar_wait(full_bar(stage), phase, p.err, 3);
      if (p.stages & 1) {
        __threadfence_block();
        if (lane_id == 0) *reinterpret_cast<volatile int*>(prog + w) = ++mine;
        else ++mine;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        uint32_t a_lo = a_lo0 + stage * stage_step;
        uint32_t accumulate = i == first ? 0u : 1u;
        if (!(p.debug & 2)) {
          for (int j = 0; j < nsub; ++j, a_lo += sub_step) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              if (p.swap)
                umma_f16(d_tmem, a_lo + b_off + 2 * k, desc_hi, a_lo + 2 * k, desc_hi, p.idesc, accumulate);
              else
                umma_f16(d_tmem, a_lo + 2 * k, desc_hi, a_lo + b_off + 2 * k, desc_hi, p.idesc, accumulate);
              accumulate = 1;
            }
          }
        }
        umma_commit(empty_bar(stage));
        if (i + 2 >= n_it) umma_commit(tfull_bar(acc));     // my last iteration of this segment
      }
      __syncwarp();
      stage += 2;
      if (stage >= p.stages) { stage -= p.stages; phase ^= 1u; }
    }
  }
}

// The halo-mode counterpart of mma_issuer_alternate: the A (or, in swap mode, B) operand of
// every tap is a shifted window of the resident patch, the ring carries filter blocks only.
// A tile is kchunks chunk-steps of iters_kc ring iterations; issuer `w` owns every other GLOBAL
// iteration.  Both issuers wait for the step's patch (pfull, non-consuming) and BOTH release it
// (pempty is initialised with count 2; an issuer without an iteration in a step arrives plainly
// after its pfull wait, so it can never lap the patch's previous use).
template <int KSTEPS>
__device__ __forceinline__ void mma_issuer_alternate_halo(const TcParams& p, int w, uint32_t base,
                                                          uint32_t ring, uint32_t tmem_base,
                                                          uint32_t bars, uint32_t tab_s, int* prog) {
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + 2 + a); };
  auto pfull_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 6 + b); };
  auto pempty_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 8 + b); };
  const int lane_id = threadIdx.x & 31;
  const uint32_t desc_hi = umma_desc_hi(p.sbo_bytes, p.layout_type);
  const uint32_t halo_hi = umma_desc_hi(static_cast<uint32_t>(p.pw) * 128u, p.layout_type);
  const uint32_t a_lo0 = umma_desc_lo(ring);
  const uint32_t stage_step = p.stage_bytes >> 4;
  int stage = w % p.stages;
  uint32_t phase = (w / p.stages) & 1u;
  int mine = 0, g = 0, q = 0, tile_it = 0;       // g: global ring iteration, q: global chunk-step
  const int n_tile = p.kchunks * p.iters_kc;     // ring iterations per tile
  TileWalk walk = walk_begin(p);
  int tile, it0, it1;
  for (; walk_next(p, walk, tile, it0, it1); ++tile_it) {
    const int acc = tile_it & 1;
    const uint32_t acc_phase = (tile_it >> 1) & 1u;
    const int first_tile = ((g & 1) == w) ? 0 : 1;
    mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err, 2);
    if (first_tile >= n_tile) {                   // a one-iteration tile owned by the other issuer
      mbar_wait(pfull_bar(q & 1), (q >> 1) & 1u, p.err, 6);
      if (elect_one()) { mbar_arrive(pempty_bar(q & 1)); mbar_arrive(tfull_bar(acc)); }
      __syncwarp();
      g += n_tile; q += p.kchunks;
      continue;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t d_tmem = tmem_base + acc * p.acc_cols + w * p.N_tile;
    const int g_tile = g;
    int last_mine = -1;                           // my last iteration index within the tile
    for (int i = first_tile; i < n_tile; i += 2) last_mine = i;
    bool started = false;
    for (int kc = 0; kc < p.kchunks; ++kc, ++q) {
      const int pb = q & 1;
      mbar_wait(pfull_bar(pb), (q >> 1) & 1u, p.err, 6);
      const uint32_t patch_lo = umma_desc_lo(base + pb * p.patch_bytes);
      const int gq = g_tile + kc * p.iters_kc;    // global index of the step's first iteration
      const int first = ((gq & 1) == w) ? 0 : 1;
      int last_step = -1;
      for (int it = first; it < p.iters_kc; it += 2) last_step = it;
      if (last_step < 0) {                        // no iteration of mine in this step
        if (elect_one()) mbar_arrive(pempty_bar(pb));
        __syncwarp();
        continue;
      }
      for (int it = first; it < p.iters_kc; it += 2) {
        const int gi = gq + it;
        const int tap = it * p.sub;
        const int nsub = min(p.sub, p.taps - tap);
        if ((p.stages & 1) && gi >= p.stages) {   // see mma_issuer_alternate
          const int need = (gi - p.stages - (w ^ 1)) / 2 + 1;
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(prog + (w ^ 1)) < need) {
            if (clock64() - t0 > 4000000000LL) {
              if (p.err) atomicExch(p.err, 8);
              __threadfence_system();
              __trap();
            }
          }
          __threadfence_block();
        }
        mbar_wait(full_bar(stage), phase, p.err, 3);
        if (p.stages & 1) {
          __threadfence_block();
          if (lane_id == 0) *reinterpret_cast<volatile int*>(prog + w) = ++mine;
          else ++mine;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t b_lo0 = a_lo0 + stage * stage_step;
          uint32_t accumulate = started ? 1u : 0u;
          for (int j = 0; j < nsub; ++j) {
            uint32_t tap_off;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tap_off) : "r"(tab_s + 4u * (tap + j)));
            const uint32_t a_lo = patch_lo + tap_off;
            const uint32_t b_lo = b_lo0 + j * (p.b_bytes >> 4);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              if (p.swap)
                umma_f16(d_tmem, b_lo + 2 * k, desc_hi, a_lo + 2 * k, halo_hi, p.idesc, accumulate);
              else
                umma_f16(d_tmem, a_lo + 2 * k, halo_hi, b_lo + 2 * k, desc_hi, p.idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(empty_bar(stage));
          if (it == last_step) umma_commit(pempty_bar(pb));
          if (kc * p.iters_kc + it == last_mine) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        started = true;
        stage += 2;
        if (stage >= p.stages) { stage -= p.stages; phase ^= 1u; }
      }
    }
    g += n_tile;
  }
}

// MINB = 2: two CTAs co-resident per SM (<= 96 registers, <= ~110 KB smem, <= 256 TMEM
// columns each).  tcgen05.mma issue costs the single issuing thread ~50-66 cycles per
// instruction plus ~500 cycles of barrier/commit latency per ring stage (measured with the
// clock64 instrumentation, profiles/r01_swap_modes.txt), which is MORE than the tensor time of
// an N <= 128 instruction (64 cycles): two CTAs give the SM's tensor pipe two issuing threads.
// MODE: 0 = every path compiled in (181 KB of SASS); 1 = resident-filter halo layers only;
// 2 = plain per-tap layers (no halo / swap / second output / debug instrumentation).  The lean
// instantiations keep the role loops and the epilogue of the layers that use them inside the
// instruction cache — the same effect as in conv_patch.cu (`no_instructions` was one of the
// top three stall reasons of the fat kernel, profiles/r02_resident_filters.txt).
template <int KSTEPS>
__device__ __forceinline__ void conv_tc_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // Manual 1024-byte alignment (SWIZZLE_128B atoms repeat every 1024 bytes).
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring = base + p.ring_off;              // halo mode: two patch buffers come first
  const uint32_t params_s = ring + p.stages * p.stage_bytes;
  const int cpad = p.cout_pad <= kMaxParamChannels ? p.cout_pad : 0;
  const uint32_t tab_s = params_s + static_cast<uint32_t>(p.param_rows) * cpad * 4u;   // halo: per-tap descriptor offset of the shifted patch window
  const uint32_t bars = tab_s + 256u;
  // barrier layout: full[stages], empty[stages], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 4);
  auto pfull_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 6 + b); };
  auto pempty_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 8 + b); };
  // two-issuer progress words (see mma_issuer_alternate)
  int* prog = reinterpret_cast<int*>(smem_raw + (bars + 8u * (2 * kMaxStages + 10) - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), p.issuers);
      mbar_init(tempty_bar(a), kEpiWarps);
      mbar_init(pfull_bar(a), 1);
      mbar_init(pempty_bar(a), p.issuers);
    }
    if (p.resident)      // patch buffers use ring barriers 1..npatch (the ring itself is one resident stage)
      for (int b = 0; b < p.npatch; ++b) {
        mbar_init(full_bar(1 + b), 1);
        mbar_init(empty_bar(1 + b), 1);
      }
    prog[0] = 0; prog[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && p.halo) {
    // rows_off * 128 B >> 4: what tap (r, s) adds to the patch's descriptor start address
    for (int t = lane; t < p.taps && t < 64; t += 32) {
      const int r = t / p.kw, sx = t - r * p.kw;
      const uint32_t v = (p.halo == 1 ? r * p.pw + sx : sx * p.pw + r) * 8u;
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(tab_s + 4u * t), "r"(v) : "memory");
    }
  }
  if (warp >= 2 && cpad) {
    // Stage the per-channel epilogue parameters once per CTA.
    float* sp = reinterpret_cast<float*>(smem_raw + (params_s - smem_u32(smem_raw)));
    for (int i = threadIdx.x - 64; i < cpad; i += kThreads1 - 64) {
      sp[i] = p.scale[i];
      sp[cpad + i] = p.shift[i];
      sp[2 * cpad + i] = p.slope ? p.slope[i] : 0.f;
      sp[3 * cpad + i] = p.scale2 ? p.scale2[i] : 1.f;
      sp[4 * cpad + i] = p.shift2 ? p.shift2[i] : 0.f;
      if (p.shift9)
        for (int k = 0; k < 9; ++k) sp[(5 + k) * cpad + i] = p.shift9[k * cpad + i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);     // provably warp-uniform
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation,
  // parameter staging — none of it reads the previous layer's output) may overlap the
  // tail of the previous kernel; activations are only touched after this wait.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!p.pdl_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // The producer and MMA loops run with the WHOLE warp converged and every
  // operand warp-uniform; one elected lane issues the TMA / tcgen05 instructions.
  // (Issuing from inside an `if (lane == 0)` region makes the compiler wrap each
  // uniform-datapath instruction in a divergence "waterfall" loop — measured at
  // ~240 cycles per tcgen05.mma, 4x the MMA's own execution time.)
  // Resident-filter mode: one MMA-issuing warp per tile parity (warp 1: even tiles into
  // accumulator 0, warp 10: odd tiles into accumulator 1 when p.resident == 2).  The two never
  // touch the same barrier phase: a tile's accumulator, patch buffer and commits belong to the
  // warp that issues it, so the barrier protocol is the single-issuer one split by parity.  An
  // N = 64 MMA is 32 cycles of tensor work but costs its issuing thread ~75: two threads keep
  // the pipe fed.
  auto resident_issuer = [&](int first) {
    const uint32_t desc_hi = umma_desc_hi(p.sbo_bytes, p.layout_type);
    const uint32_t halo_hi = umma_desc_hi(static_cast<uint32_t>(p.pw) * 128u, p.layout_type);
    const uint32_t a_lo0 = umma_desc_lo(ring);
    const int step = p.resident;                            // 1 or 2 issuers
    TileWalk walk = walk_begin(p);
    int tile, it0, it1;
    mbar_wait(full_bar(0), 0, p.err, 3);                    // the filter bank has landed
    for (int tile_it = 0; walk_next(p, walk, tile, it0, it1); ++tile_it) {
      if (step == 2 && (tile_it & 1) != first) continue;
      const int acc = tile_it & 1;
      mbar_wait(tempty_bar(acc), ((tile_it >> 1) & 1u) ^ 1u, p.err, 2);
      const int buf = tile_it % p.npatch;
      mbar_wait(full_bar(1 + buf), (tile_it / p.npatch) & 1u, p.err, 6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * p.acc_cols;
        const uint32_t patch_lo = umma_desc_lo(base + buf * p.patch_bytes);
        uint32_t accumulate = 0;
        if (!(p.debug & 2)) {
          for (int t = 0; t < p.taps; ++t) {
            uint32_t tap_off;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tap_off) : "r"(tab_s + 4u * t));
            const uint32_t a_lo = patch_lo + tap_off;
            const uint32_t b_lo = a_lo0 + t * (p.b_bytes >> 4);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              umma_f16(d_tmem, a_lo + 2 * k, halo_hi, b_lo + 2 * k, desc_hi, p.idesc, accumulate);
              accumulate = 1;
            }
          }
        }
        umma_commit(empty_bar(1 + buf));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
    }
  };
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0, pb = 0;
    uint32_t phase = 0, pphase = 0;
    const int w_rows = p.swap ? 128 : p.N_tile;             // filter rows per k-block
    const uint32_t sub_tx = p.rows * p.KC * 2 + w_rows * p.KC * 2;
    TileWalk walk = walk_begin(p);
    int tile, it0, it1;
    long long dbg_wait = 0, dbg_pwait = 0, dbg_t0 = clock64();
    int dbg_iters = 0;
    int pq = 0, pq_issued = 0;                   // halo: chunk-steps produced / patches requested
    if (p.resident) {
      if (elect_one()) {
        mbar_expect_tx(full_bar(0), p.taps * w_rows * p.KC * 2);
        for (int t = 0; t < p.taps; ++t)
          tma_load_2d(ring + t * p.b_bytes, &tmB, full_bar(0), t * p.cin_pad, 0);
      }
      __syncwarp();
      for (int q = 0; walk_next(p, walk, tile, it0, it1); ++q) {
        if (p.pdl_late && walk_done(p, walk))
          asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        const int buf = q % p.npatch;
        mbar_wait(empty_bar(1 + buf), ((q / p.npatch) & 1u) ^ 1u, p.err, 5);
        if (elect_one()) {
          int pm = tile / p.n_tiles;
          const int pwb = pm % p.tiles_w; pm /= p.tiles_w;
          const int phb = pm % p.tiles_h; pm /= p.tiles_h;
          const int cw = pwb * p.bw - p.pad, ch = phb * p.bh - p.pad;
          mbar_expect_tx(full_bar(1 + buf), p.patch_tx);
          tma_load_5d(base + buf * p.patch_bytes, &tmA, full_bar(1 + buf), p.in_coff,
                      p.halo == 1 ? cw : ch, p.halo == 1 ? ch : cw, pm * p.bn, 0);
        }
        __syncwarp();
      }
    }
    while (!p.resident && walk_next(p, walk, tile, it0, it1)) {
      // A dependent CTA that is resident early only spins in griddepcontrol.wait while holding
      // an SM that a kernel of another stream could use: release it late.
      if (p.pdl_late && walk_done(p, walk))
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int wb = mt % p.tiles_w; mt /= p.tiles_w;
      const int hb = mt % p.tiles_h; mt /= p.tiles_h;
      const int nb = mt;
      const int w0 = wb * p.bw, h0 = hb * p.bh, n0 = nb * p.bn;
      // Every CTA streams the SAME filter blocks; started in lockstep they would all hit the
      // same L2 lines at the same time.  Each tile walks the taps / k-blocks from its own
      // starting point (a sum, so the order is free; both CTAs of a stream-K split tile agree).
      const int rot = p.rotate ? (tile * 13) % (p.halo ? p.taps : p.k_blocks) : 0;
      if (p.halo) {
        // Patches are requested ONE chunk-step ahead (the next channel chunk of this tile, or
        // the first chunk of the CTA's next tile): a patch is tens of KB gathered from 16 x
        // (bw + 2 pad) pixel rows and takes several microseconds to land, far longer than the
        // few ring stages the filter stream runs ahead of the MMAs.
        auto issue_patch = [&](int ptile, int pkc) {
          const int buf = pq_issued & 1;
          const uint32_t ph = (pq_issued >> 1) & 1u;
          const long long tp0 = (p.debug & 32) ? clock64() : 0;
          mbar_wait(pempty_bar(buf), ph ^ 1u, p.err, 5);
          if (p.debug & 32) dbg_pwait += clock64() - tp0;
          if (elect_one()) {
            int pm = ptile / p.n_tiles;
            const int pwb = pm % p.tiles_w; pm /= p.tiles_w;
            const int phb = pm % p.tiles_h; pm /= p.tiles_h;
            const int cw = pwb * p.bw - p.pad, ch = phb * p.bh - p.pad;
            mbar_expect_tx(pfull_bar(buf), p.patch_tx);
            tma_load_5d(base + buf * p.patch_bytes, &tmA, pfull_bar(buf), p.in_coff + pkc * p.KC,
                        p.halo == 1 ? cw : ch, p.halo == 1 ? ch : cw, pm * p.bn, 0);
          }
          __syncwarp();
          ++pq_issued;
        };
        const int pf_it = min(p.stages, p.iters_kc - 1);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (pq_issued == pq) issue_patch(tile, kc);       // first step of this CTA
          int tap = 0;
          for (int it = 0; it < p.iters_kc; ++it) {
            if (it == pf_it && pq_issued == pq + 1 && !(p.debug & 64)) {
              if (kc + 1 < p.kchunks) issue_patch(tile, kc + 1);
              else if (!walk_done(p, walk)) issue_patch(walk.tile, 0);   // no stream-K in halo mode
            }
            const int nsub = min(p.sub, p.taps - tap);
            const long long tq0 = (p.debug & 32) ? clock64() : 0;
            mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 1);
            if (p.debug & 32) { dbg_wait += clock64() - tq0; ++dbg_iters; }
            const uint32_t sb = ring + stage * p.stage_bytes;
            if (p.debug & 1) {                       // timing only: no filter loads
              if (elect_one()) mbar_arrive(full_bar(stage));
            } else if (elect_one()) {
              mbar_expect_tx(full_bar(stage), nsub * w_rows * p.KC * 2);
              for (int j = 0; j < nsub; ++j) {
                int t = tap + j + rot;
                if (t >= p.taps) t -= p.taps;
                tma_load_2d(sb + j * p.b_bytes, &tmB, full_bar(stage), t * p.cin_pad + kc * p.KC, nt * w_rows);
              }
            }
            __syncwarp();
            tap += nsub;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          ++pq;
        }
        continue;
      }
      int kb = it0 * p.sub;                       // running (tap row, tap col, channel chunk)
      const int kb0 = (kb + rot) % p.k_blocks;
      int kc = kb0 % p.kchunks, s = (kb0 / p.kchunks) % p.kw, r = kb0 / (p.kchunks * p.kw);
      for (int it = it0; it < it1; ++it) {
        const int nsub = min(p.sub, p.k_blocks - kb);
        mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 1);
        const uint32_t sa = ring + stage * p.stage_bytes;
        if (p.debug & 1) {
          if (elect_one()) mbar_arrive(full_bar(stage));
          kb += nsub;
        } else {
          if (elect_one()) mbar_expect_tx(full_bar(stage), sub_tx * nsub);
          __syncwarp();
          for (int j = 0; j < nsub; ++j, ++kb) {
            int c0 = p.in_coff + kc * p.KC, c1, c2, c3, c4;
            if (p.stride == 1) {
              c1 = w0 + s - p.pad; c2 = h0 + r - p.pad; c3 = n0; c4 = 0;
            } else {
              const int oy = r - p.pad, ox = s - p.pad;
              const int py = oy & 1, px = ox & 1;
              c0 += px * p.in_cs;
              c1 = w0 + ((ox - px) >> 1); c2 = py; c3 = h0 + ((oy - py) >> 1); c4 = n0;
            }
            const uint32_t dst = sa + j * p.sub_bytes;
            if (elect_one()) {
              tma_load_5d(dst, &tmA, full_bar(stage), c0, c1, c2, c3, c4);
              tma_load_2d(dst + p.a_bytes, &tmB, full_bar(stage),
                          (r * p.kw + s) * p.cin_pad + kc * p.KC, nt * w_rows);
            }
            if (++kc == p.kchunks) { kc = 0; if (++s == p.kw) { s = 0; if (++r == p.kh) r = 0; } }
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    if ((p.debug & 32) && blockIdx.x == 0 && lane == 0)
      printf("producer: total %lld clk, %d iters, wait(empty) %lld, wait(patch empty) %lld\n",
             clock64() - dbg_t0, dbg_iters, dbg_wait, dbg_pwait);
  } else if (warp == 10 && p.issuers == 2) {
    // ------------------------------------------------- second MMA issuer (odd ring stages)
    if (p.halo) mma_issuer_alternate_halo<KSTEPS>(p, 1, base, ring, tmem_base, bars, tab_s, prog);
    else mma_issuer_alternate<KSTEPS>(p, 1, ring, tmem_base, bars, prog);
  } else if (warp == 10) {
    // single-issuer launch: nothing to do — except the odd tiles of resident-filter mode
    if (p.resident == 2) resident_issuer(1);
  } else if (warp == 1 && p.issuers == 2) {
    if (p.halo) mma_issuer_alternate_halo<KSTEPS>(p, 0, base, ring, tmem_base, bars, tab_s, prog);
    else mma_issuer_alternate<KSTEPS>(p, 0, ring, tmem_base, bars, prog);
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    int stage = 0, pb = 0;
    uint32_t phase = 0, pphase = 0;
    int tile_it = 0;
    const uint32_t desc_hi = umma_desc_hi(p.sbo_bytes, p.layout_type);
    const uint32_t halo_hi = umma_desc_hi(static_cast<uint32_t>(p.pw) * 128u, p.layout_type);   // 16 patch rows per 8-row group step
    const uint32_t a_lo0 = umma_desc_lo(ring);
    const uint32_t stage_step = p.stage_bytes >> 4, sub_step = p.sub_bytes >> 4,
                   b_off = p.a_bytes >> 4;
    TileWalk walk = walk_begin(p);
    int tile, it0, it1;
    long long dbg_wf = 0, dbg_te = 0, dbg_pw = 0, dbg_last = 0, dbg_t0 = clock64(), dbg_issue = 0, dbg_commit = 0;
    int dbg_n = 0;
    if (p.resident) resident_issuer(0);
    for (; !p.resident && walk_next(p, walk, tile, it0, it1); ++tile_it) {
      const int acc = tile_it & 1;
      const uint32_t acc_phase = (tile_it >> 1) & 1u;
      if (p.debug & 32) dbg_last = clock64();
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (p.debug & 32) dbg_te += clock64() - dbg_last, dbg_last = clock64();
      const uint32_t d_tmem = tmem_base + acc * p.acc_cols;
      int kb = it0 * p.sub;
      uint32_t accumulate = 0;
      const int rot = p.rotate && p.halo ? (tile * 13) % p.taps : 0;
      if (p.halo) {
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (p.debug & 32) dbg_last = clock64();
          mbar_wait(pfull_bar(pb), pphase, p.err, 6);
          if (p.debug & 32) dbg_pw += clock64() - dbg_last;
          const uint32_t patch_lo = umma_desc_lo(base + pb * p.patch_bytes);
          int tap = 0;
          for (int it = 0; it < p.iters_kc; ++it) {
            const int nsub = min(p.sub, p.taps - tap);
            const long long tq0 = (p.debug & 32) ? clock64() : 0;
            mbar_wait(full_bar(stage), phase, p.err, 3);
            if (p.debug & 32) { dbg_wf += clock64() - tq0; ++dbg_n; }
            if (!(p.debug & 16)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              const long long ti0 = (p.debug & 32) ? clock64() : 0;
              const uint32_t b_lo0 = a_lo0 + stage * stage_step;
              for (int j = 0; j < nsub; ++j) {
                int t = tap + j + rot;
                if (t >= p.taps) t -= p.taps;
                uint32_t tap_off;                // no divisions on the issuing thread's critical path
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tap_off) : "r"(tab_s + 4u * t));
                const uint32_t a_lo = patch_lo + tap_off;
                const uint32_t a_hi = halo_hi;   // base_offset stays 0: the swizzle XOR uses absolute smem address bits
                const uint32_t b_lo = b_lo0 + j * (p.b_bytes >> 4);
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  if (p.debug & 2) break;
                  if (p.swap)   // filters on M; the shifted patch window (bw columns of 8 rows) on N
                    umma_f16(d_tmem, b_lo + 2 * k, desc_hi, a_lo + 2 * k, a_hi, p.idesc, accumulate);
                  else
                    umma_f16(d_tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, desc_hi, p.idesc, accumulate);
                  accumulate = 1;
                }
              }
              const long long ti1 = (p.debug & 32) ? clock64() : 0;
              umma_commit(empty_bar(stage));
              if (it == p.iters_kc - 1) {
                umma_commit(pempty_bar(pb));
                if (kc == p.kchunks - 1) umma_commit(tfull_bar(acc));
              }
              if (p.debug & 32) { dbg_issue += ti1 - ti0; dbg_commit += clock64() - ti1; }
            }
            __syncwarp();
            tap += nsub;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          if (++pb == 2) { pb = 0; pphase ^= 1u; }
        }
        continue;
      }
      for (int it = it0; it < it1; ++it) {
        const int nsub = min(p.sub, p.k_blocks - kb);
        kb += nsub;
        const long long tq0 = (p.debug & 32) ? clock64() : 0;
        mbar_wait(full_bar(stage), phase, p.err, 3);
        if (p.debug & 32) { dbg_wf += clock64() - tq0; ++dbg_n; }
        if (!(p.debug & 16)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const long long ti0 = (p.debug & 32) ? clock64() : 0;
          uint32_t a_lo = a_lo0 + stage * stage_step;
          if (!(p.debug & 2)) {
            for (int j = 0; j < nsub; ++j, a_lo += sub_step) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {      // +32 B (2 x 16 B units) per K step
                if (p.swap)
                  umma_f16(d_tmem, a_lo + b_off + 2 * k, desc_hi, a_lo + 2 * k, desc_hi, p.idesc, accumulate);
                else
                  umma_f16(d_tmem, a_lo + 2 * k, desc_hi, a_lo + b_off + 2 * k, desc_hi, p.idesc, accumulate);
                accumulate = 1;
              }
            }
          }
          const long long ti1 = (p.debug & 32) ? clock64() : 0;
          umma_commit(empty_bar(stage));          // frees the smem slot when the MMAs retire
          if (it == it1 - 1) umma_commit(tfull_bar(acc));
          if (p.debug & 32) { dbg_issue += ti1 - ti0; dbg_commit += clock64() - ti1; }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    if (p.debug & 32) {
      for (int o = 16; o; o >>= 1) {
        dbg_issue += __shfl_xor_sync(0xffffffffu, dbg_issue, o);
        dbg_commit += __shfl_xor_sync(0xffffffffu, dbg_commit, o);
      }
      if (blockIdx.x == 0 && lane == 0)
        printf("mma: total %lld clk, %d iters, wait(full) %lld, wait(tmem empty) %lld, wait(patch) %lld, "
               "mma issue %lld, commit %lld\n",
               clock64() - dbg_t0, dbg_n, dbg_wf, dbg_te, dbg_pw, dbg_issue, dbg_commit);
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = ew >> 2;               // which 16-column chunks (even / odd) it takes
    const int row = q * 32 + lane;
    // halo == 2: MMA row m = (w_l, h_l) with h_l fastest (16w x 8h tile)
    const int w_l = p.halo == 2 ? row >> 3 : row % p.bw;
    const int h_l = p.halo == 2 ? row & 7 : (row / p.bw) % p.bh;
    const int n_l = p.halo == 2 ? 0 : row / (p.bw * p.bh);
    const int nchunks = p.N_tile >> 4;
    EpiCtx e;
    e.cpad = cpad;
    e.sp = reinterpret_cast<const float*>(smem_raw + (params_s - smem_u32(smem_raw)));
    int tile_it = 0;
    int g_it = 0;                               // global iteration counter (two-issuer accounting)
    TileWalk walk = walk_begin(p);
    int tile, it0, it1;
    for (; walk_next(p, walk, tile, it0, it1); ++tile_it) {
      const int acc = tile_it & 1;
      const uint32_t acc_phase = (tile_it >> 1) & 1u;
      // two issuers: the segment's first iteration went to issuer (g & 1); the other one has a
      // partial accumulator too iff the segment has at least two iterations
      const int seg_iters = p.halo ? p.kchunks * p.iters_kc : it1 - it0;
      const int first_w = g_it & 1;
      const bool both = p.issuers == 2 && seg_iters >= 2;
      g_it += seg_iters;
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int wb = mt % p.tiles_w; mt /= p.tiles_w;
      const int hb = mt % p.tiles_h; mt /= p.tiles_h;
      const int nb = mt;
      const int ow = wb * p.bw + w_l, oh = hb * p.bh + h_l, on = nb * p.bn + n_l;
      e.valid = row < p.rows && ow < p.W_out && oh < p.H_out && on < p.N;
      e.pix = (static_cast<long>(on) * p.H_out + oh) * p.W_out + ow;
      e.pool_store = p.pool && e.valid && !(lane & 9) && oh + 1 < p.H_out && ow + 1 < p.W_out;
      e.ppix = p.pool ? (static_cast<long>(on) * p.pool_H + (oh >> 1)) * p.pool_W + (ow >> 1) : 0;
      // swap mode: the tile is a band of full-width rows of image nb (consecutive pixels), or
      // bw columns x 8 rows of the halo patch
      SwapTile st;
      st.pix0 = p.halo ? static_cast<long>(nb) * p.H_out * p.W_out
                       : (static_cast<long>(nb) * p.H_out + hb * p.bh) * p.W_out;
      st.valid = min(p.rows, (p.H_out - hb * p.bh) * p.W_out);
      st.h0 = hb * p.bh; st.w0 = wb * p.bw;
      const int sw_cout = nt * 128 + row;
      e.rpix = e.pix;
      if (p.res && p.res_up2)
        e.rpix = (static_cast<long>(on) * p.res_H + (oh >> 1)) * p.res_W + (ow >> 1);
      e.shp = e.sp + e.cpad;
      if (p.shift9) {
        const int rc = oh == 0 ? 0 : (oh >= p.H_out - 1 ? 2 : 1), cc = ow == 0 ? 0 : (ow >= p.W_out - 1 ? 2 : 1);
        e.shp = e.sp + (5 + rc * 3 + cc) * e.cpad;
      }

      mbar_wait(tfull_bar(acc), acc_phase, p.err, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.acc_cols +
                             (p.issuers == 2 ? first_w * p.N_tile : 0);
      const uint32_t taddr2 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.acc_cols +
                              (first_w ^ 1) * p.N_tile;           // the other issuer's accumulator
      const int cn0 = nt * p.N_tile;
      uint32_t va[16], vb[16];
      int c = half;
      // two issuers: v = chunk of issuer A (+ chunk of issuer B), synchronous
      auto load_sum = [&](uint32_t (&v)[16], int chunk) {
        __syncwarp();
        tmem_ld16_async(taddr + chunk * 16, v);
        if (both) {
          uint32_t t2[16];
          tmem_ld16_async(taddr2 + chunk * 16, t2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t2[j]));
        } else {
          tmem_ld_wait();
        }
      };
      if (it0 > 0) {
        // TAIL of a tile whose first iterations belong to the previous CTA: park the raw
        // accumulators, laid out [chunk][row][16] so that a warp writes 2 KB runs.
        float4* ws = reinterpret_cast<float4*>(p.sk_ws + static_cast<size_t>(blockIdx.x) * 128 * p.N_tile);
        for (; c < nchunks; c += 2) {
          load_sum(va, c);
          float4* dst = ws + (c * 128 + row) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            __stcg(dst + j, make_float4(__uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]),
                                        __uint_as_float(va[4 * j + 2]), __uint_as_float(va[4 * j + 3])));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tempty_bar(acc));
          atomicAdd(p.sk_flags + blockIdx.x, 1);
        }
        continue;
      }
      const float4* partial = nullptr;
      if (it1 < p.iters) {
        // HEAD: the rest of this tile is the tail partial of the next CTA (the first thing
        // that CTA computes).  All kEpiWarps warps of both CTAs take part; the last reader
        // re-arms the flag for the next launch.
        int* flag = p.sk_flags + blockIdx.x + 1;
        if (lane == 0) {
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(flag) < kEpiWarps) {
            if (clock64() - t0 > 4000000000LL) {
              if (p.err) atomicExch(p.err, 9);
              __threadfence_system();
              __trap();
            }
          }
          __threadfence();
        }
        __syncwarp();
        partial = reinterpret_cast<const float4*>(p.sk_ws + static_cast<size_t>(blockIdx.x + 1) * 128 * p.N_tile);
      }
      auto add_partial = [&](uint32_t (&v)[16], int chunk) {
        if (!partial) return;
        const float4* src = partial + (chunk * 128 + row) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = __ldcg(src + j);
          v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + f.x);
          v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + f.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + f.z);
          v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + f.w);
        }
      };
      if (p.issuers == 2) {
        for (; c < nchunks; c += 2) {
          load_sum(va, c);
          add_partial(va, c);
          if (p.swap) epilogue_chunk_swap(p, e.sp, cpad, va, sw_cout, st, c * 16);
          else epilogue_chunk(p, e, va, cn0 + c * 16);
        }
      }
      // software-pipelined TMEM reads: chunk i+2 is in flight while chunk i is stored
      __syncwarp();
      if (c < nchunks) tmem_ld16_async(taddr + c * 16, va);
      while (c < nchunks) {
        __syncwarp();                       // tcgen05.ld / wait::ld are warp-collective
        tmem_ld_wait();
        if (c + 2 < nchunks) tmem_ld16_async(taddr + (c + 2) * 16, vb);
        add_partial(va, c);
        if (p.swap) epilogue_chunk_swap(p, e.sp, cpad, va, sw_cout, st, c * 16);
        else epilogue_chunk(p, e, va, cn0 + c * 16);
        c += 2;
        if (c >= nchunks) break;
        __syncwarp();
        tmem_ld_wait();
        if (c + 2 < nchunks) tmem_ld16_async(taddr + (c + 2) * 16, va);
        add_partial(vb, c);
        if (p.swap) epilogue_chunk_swap(p, e.sp, cpad, vb, sw_cout, st, c * 16);
        else epilogue_chunk(p, e, vb, cn0 + c * 16);
        c += 2;
      }
      // Release the accumulator back to the MMA warp.
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(tempty_bar(acc));
        if (partial) {
          int* flag = p.sk_flags + blockIdx.x + 1;
          if (atomicAdd(flag, 1) == 2 * kEpiWarps - 1) atomicExch(flag, 0);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(p.tmem_cols) : "memory");
  }
}

template <int KSTEPS, int MINB, int MODE = 0>
__global__ void __launch_bounds__(kThreads1, MINB)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams pk) {
  if (MODE == 0) {
    conv_tc_body<KSTEPS>(tmA, tmB, pk);
    return;
  }
  // the features a lean instantiation leaves out become compile-time constants
  TcParams p = pk;
  p.swap = 0; p.debug = 0; p.out2 = nullptr; p.rotate = 0; p.cta2 = 0;
  if (MODE == 1) {            // resident filters: one issuer, whole tiles
    p.resident = 1; p.issuers = 1; p.sk = 0; p.out_f32 = nullptr;
    if (p.halo != 2) p.halo = 1;
  }
  if (MODE == 2) {            // plain per-tap loads
    p.halo = 0; p.resident = 0; p.pool = 0;
  }
  conv_tc_body<KSTEPS>(tmA, tmB, p);
}

// ---------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2).  Two CTAs of a cluster compute two
// adjacent pixel tiles against the SAME filter tile: each CTA loads its own A
// tile and only HALF of the B tile; one tcgen05.mma (M = 256) issued by the
// leader CTA drives both SMs, reading A and its B half from each CTA's shared
// memory and writing each CTA's 128 accumulator rows to its own TMEM.  Per SM
// this halves the filter traffic (L2 -> smem and smem -> tensor core) and the
// number of MMA instructions per flop.  Barriers: both CTAs' TMA loads signal
// the leader's `full` barrier; tcgen05.commit multicasts to both CTAs' `empty`
// and `tmem_full` barriers; both CTAs' epilogue warps arrive on the leader's
// `tmem_empty` barrier.  Plain (per-tap) A loads only, KC = 64.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                             int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                             int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                          uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n"
      "}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}\n"
      ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t params_s = base + p.stages * p.stage_bytes;
  const int cpad = p.cout_pad;
  const uint32_t bars = params_s + static_cast<uint32_t>(p.param_rows) * cpad * 4u;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);          // leader's expect_tx arrive (+ tx bytes of both CTAs)
      mbar_init(empty_bar(s), 1);         // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);                 // multicast commit
      mbar_init(tempty_bar(a), 2 * kEpiWarps);    // leader's copy is the one used
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    float* sp = reinterpret_cast<float*>(smem_raw + (params_s - smem_u32(smem_raw)));
    for (int i = threadIdx.x - 64; i < cpad; i += kThreads - 64) {
      sp[i] = p.scale[i];
      sp[cpad + i] = p.shift[i];
      sp[2 * cpad + i] = p.slope ? p.slope[i] : 0.f;
      sp[3 * cpad + i] = p.scale2 ? p.scale2[i] : 1.f;
      sp[4 * cpad + i] = p.shift2 ? p.shift2[i] : 0.f;
      if (p.shift9)
        for (int k = 0; k < 9; ++k) sp[(5 + k) * cpad + i] = p.shift9[k * cpad + i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                 // peer barriers are initialised before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int units = p.pair_units;     // (pixel-tile pairs) x n_tiles
  const int half_n = p.N_tile >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t cta_tx = p.rows * p.KC * 2 + half_n * p.KC * 2;
    for (int u = pair; u < units; u += npairs) {
      const int nt = u % p.n_tiles;
      int mt = (u / p.n_tiles) * 2 + static_cast<int>(rank);
      const int wb = mt % p.tiles_w; mt /= p.tiles_w;
      const int hb = mt % p.tiles_h; mt /= p.tiles_h;
      const int nb = mt;                  // may be >= tiles_n for the odd tail: loads zero-fill
      const int w0 = wb * p.bw, h0 = hb * p.bh, n0 = nb * p.bn;
      int r = 0, s = 0, kc = 0, kb = 0;
      for (int it = 0; it < p.iters; ++it) {
        const int nsub = min(p.sub, p.k_blocks - kb);
        mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 1);
        const uint32_t sa = base + stage * p.stage_bytes;
        const uint32_t lead_full = mapa_u32(full_bar(stage), 0);
        if (rank == 0 && elect_one()) mbar_expect_tx(full_bar(stage), 2u * cta_tx * nsub);
        __syncwarp();
        for (int j = 0; j < nsub; ++j, ++kb) {
          int c0 = p.in_coff + kc * p.KC, c1, c2, c3, c4;
          if (p.stride == 1) {
            c1 = w0 + s - p.pad; c2 = h0 + r - p.pad; c3 = n0; c4 = 0;
          } else {
            const int oy = r - p.pad, ox = s - p.pad;
            const int py = oy & 1, px = ox & 1;
            c0 += px * p.in_cs;
            c1 = w0 + ((ox - px) >> 1); c2 = py; c3 = h0 + ((oy - py) >> 1); c4 = n0;
          }
          const uint32_t dst = sa + j * p.sub_bytes;
          if (elect_one()) {
            tma2_load_5d(dst, &tmA, lead_full, c0, c1, c2, c3, c4);
            tma2_load_2d(dst + p.a_bytes, &tmB, lead_full, (r * p.kw + s) * p.cin_pad + kc * p.KC,
                         nt * p.N_tile + static_cast<int>(rank) * half_n);
          }
          if (++kc == p.kchunks) { kc = 0; if (++s == p.kw) { s = 0; ++r; } }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int tile_it = 0;
      const uint32_t desc_hi = umma_desc_hi(p.sbo_bytes, p.layout_type);
      const uint32_t a_lo0 = umma_desc_lo(base);
      const uint32_t stage_step = p.stage_bytes >> 4, sub_step = p.sub_bytes >> 4,
                     b_off = p.a_bytes >> 4;
      for (int u = pair; u < units; u += npairs, ++tile_it) {
        const int acc = tile_it & 1;
        const uint32_t acc_phase = (tile_it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * p.N_tile;
        int kb = 0;
        uint32_t accumulate = 0;
        for (int it = 0; it < p.iters; ++it) {
          const int nsub = min(p.sub, p.k_blocks - kb);
          kb += nsub;
          mbar_wait(full_bar(stage), phase, p.err, 3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            uint32_t a_lo = a_lo0 + stage * stage_step;
            if (!(p.debug & 2)) {
              for (int j = 0; j < nsub; ++j, a_lo += sub_step) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma2_f16(d_tmem, a_lo + 2 * k, a_lo + b_off + 2 * k, desc_hi, p.idesc2, accumulate);
                  accumulate = 1;
                }
              }
            }
            umma2_commit_both(empty_bar(stage));
            if (it == p.iters - 1) umma2_commit_both(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    const int w_l = row % p.bw;
    const int h_l = (row / p.bw) % p.bh;
    const int n_l = row / (p.bw * p.bh);
    const int nchunks = p.N_tile >> 4;
    EpiCtx e;
    e.cpad = cpad;
    e.sp = reinterpret_cast<const float*>(smem_raw + (params_s - smem_u32(smem_raw)));
    int tile_it = 0;
    for (int u = pair; u < units; u += npairs, ++tile_it) {
      const int acc = tile_it & 1;
      const uint32_t acc_phase = (tile_it >> 1) & 1u;
      const int nt = u % p.n_tiles;
      int mt = (u / p.n_tiles) * 2 + static_cast<int>(rank);
      const int wb = mt % p.tiles_w; mt /= p.tiles_w;
      const int hb = mt % p.tiles_h; mt /= p.tiles_h;
      const int nb = mt;
      const int ow = wb * p.bw + w_l, oh = hb * p.bh + h_l, on = nb * p.bn + n_l;
      e.valid = row < p.rows && ow < p.W_out && oh < p.H_out && on < p.N;
      e.pix = (static_cast<long>(on) * p.H_out + oh) * p.W_out + ow;
      e.pool_store = false; e.ppix = 0;
      // swap mode: the tile is a band of full-width rows of image nb (consecutive pixels), or
      // bw columns x 8 rows of the halo patch
      SwapTile st;
      st.pix0 = p.halo ? static_cast<long>(nb) * p.H_out * p.W_out
                       : (static_cast<long>(nb) * p.H_out + hb * p.bh) * p.W_out;
      st.valid = min(p.rows, (p.H_out - hb * p.bh) * p.W_out);
      st.h0 = hb * p.bh; st.w0 = wb * p.bw;
      const int sw_cout = nt * 128 + row;
      e.rpix = e.pix;
      if (p.res && p.res_up2)
        e.rpix = (static_cast<long>(on) * p.res_H + (oh >> 1)) * p.res_W + (ow >> 1);
      e.shp = e.sp + e.cpad;
      if (p.shift9) {
        const int rc = oh == 0 ? 0 : (oh >= p.H_out - 1 ? 2 : 1), cc = ow == 0 ? 0 : (ow >= p.W_out - 1 ? 2 : 1);
        e.shp = e.sp + (5 + rc * 3 + cc) * e.cpad;
      }

      mbar_wait(tfull_bar(acc), acc_phase, p.err, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.N_tile;
      const int cn0 = nt * p.N_tile;
      uint32_t va[16], vb[16];
      int c = half;
      __syncwarp();
      if (c < nchunks) tmem_ld16_async(taddr + c * 16, va);
      while (c < nchunks) {
        __syncwarp();
        tmem_ld_wait();
        if (c + 2 < nchunks) tmem_ld16_async(taddr + (c + 2) * 16, vb);
        epilogue_chunk(p, e, va, cn0 + c * 16);
        c += 2;
        if (c >= nchunks) break;
        __syncwarp();
        tmem_ld_wait();
        if (c + 2 < nchunks) tmem_ld16_async(taddr + (c + 2) * 16, va);
        epilogue_chunk(p, e, vb, cn0 + c * 16);
        c += 2;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                 // the peer may still signal our barriers / read our smem
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(p.tmem_cols) : "memory");
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    TR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    TR_CHECK(q == cudaDriverEntryPointSuccess && ptr, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// Pipeline-timeout code written by a trapping kernel.  Mapped pinned host memory: the host can
// still read it after the trap has killed the context (device allocations are gone by then).
int* g_err_host = nullptr;
int* tc_error_flag() {
  static int* flag[kMaxDevices] = {};
  const int dev = current_device();
  if (!g_err_host) {
    TR_CUDA(cudaHostAlloc(&g_err_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *g_err_host = 0;
  }
  if (!flag[dev]) TR_CUDA(cudaHostGetDevicePointer(&flag[dev], g_err_host, 0));
  return flag[dev];
}

int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) TR_CUDA(cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev));
  return n[dev];
}

using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const TcParams);

int env_npatch() {
  static const int n = [] { const char* e = getenv("TRB_TC_NPATCH"); return e ? atoi(e) : 4; }();   // (6 measured no better than 4)
  return n;
}

KernelFn kernel_for(int kc, int ctas_per_sm, int mode = 0) {
  KernelFn fn = ctas_per_sm == 2
                    ? (kc == 64 ? conv_tc_kernel<4, 2> : (kc == 32 ? conv_tc_kernel<2, 2> : conv_tc_kernel<1, 2>))
                    : (kc == 64 ? (mode == 1 ? conv_tc_kernel<4, 1, 1> : mode == 2 ? conv_tc_kernel<4, 1, 2> : conv_tc_kernel<4, 1>)
                                : (kc == 32 ? conv_tc_kernel<2, 1> : conv_tc_kernel<1, 1>));
  if (ctas_per_sm == 2 || kc != 64) mode = 0;
  static bool attr_set[kMaxDevices][8] = {};
  const int slot = mode ? 5 + mode : (kc == 64 ? 0 : (kc == 32 ? 1 : 2)) + (ctas_per_sm == 2 ? 3 : 0);
  const int dev = current_device();
  if (!attr_set[dev][slot]) {
    TR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev][slot] = true;
  }
  return fn;
}

}  // namespace

// Stream-K scratch: one flag per CTA (+1), then per CTA a 128 x 256 fp32 accumulator tile.
constexpr size_t kSkFlagBytes = 2048;
size_t conv_tc_sk_scratch_bytes() { return kSkFlagBytes + size_t(num_sms()) * 128 * 256 * 4; }

struct ConvTcPlan {
  CUtensorMap tmA, tmB;
  TcParams p;
  void* sk_own = nullptr;
  int grid;
  int ctas_per_sm = 1;
  int mode = 0;                 // kernel instantiation (see MODE of conv_tc_kernel)
  uint32_t smem;
  double flops;
};

bool conv_tc_eligible(const ConvArgs& a) {
  if (a.cin_pad % 16 || a.cout_pad % 16 || a.cout_store % 8) return false;
  if (a.stride != 1 && a.stride != 2) return false;
  if (a.stride == 2 && ((a.in.H | a.in.W) & 1)) return false;
  if (a.in.cs % 8 || a.in.coff % 8 || a.out.cs % 8 || a.out.coff % 8) return false;
  if (a.cin_pad > 64 && a.cin_pad % 64) return false;
  if (a.cin_pad != 16 && a.cin_pad != 32 && a.cin_pad % 64) return false;
  if (a.cout_pad > 256 && a.cout_pad % 256) return false;
  if (a.cout_pad > kMaxParamChannels) return false;
  return true;
}

ConvTcPlan* conv_tc_plan_create(const ConvArgs& a) {
  TR_CHECK(conv_tc_eligible(a), "convolution not eligible for the tcgen05 path");
  auto* plan = new ConvTcPlan();
  TcParams& p = plan->p;
  p = TcParams{};
  p.N = a.in.N; p.H_out = a.H_out; p.W_out = a.W_out;
  p.kh = a.kh; p.kw = a.kw; p.pad = a.pad; p.stride = a.stride;
  p.cin_pad = a.cin_pad;
  p.KC = a.cin_pad >= 64 ? 64 : a.cin_pad;
  p.kchunks = a.cin_pad / p.KC;
  p.k_blocks = a.kh * a.kw * p.kchunks;
  p.in_coff = a.in.coff; p.in_cs = a.in.cs;
  p.N_tile = a.cout_pad > 256 ? 256 : a.cout_pad;
  p.n_tiles = a.cout_pad / p.N_tile;
  p.cout_pad = a.cout_pad;
  p.shift9 = a.shift9;
  p.param_rows = a.shift9 ? 14 : 5;

  // Pick the pixel box {bw, bh, bn} (<= 128 rows) that wastes the fewest MMA rows.
  double best = -1.0;
  for (int bw = 1; bw <= std::min(p.W_out, 128); ++bw) {
    for (int bh = 1; bh <= std::min(p.H_out, 128 / bw); ++bh) {
      for (int bn = 1; bn <= std::min(p.N, 128 / (bw * bh)); ++bn) {
        const double tiles = double(ceil_div(p.W_out, bw)) * ceil_div(p.H_out, bh) * ceil_div(p.N, bn);
        double eff = double(p.W_out) * p.H_out * p.N / (tiles * 128.0);
        eff += 1e-6 * bw;      // prefer long contiguous runs on ties
        if (eff > best) { best = eff; p.bw = bw; p.bh = bh; p.bn = bn; }
      }
    }
  }
  // Halo mode (stride-1 KxK, 64-channel chunks): tile 8x16 or 16x8 pixels, the input patch
  // incl. filter halo is 16 pixels along the 8-pixel axis so an 8-row UMMA group never
  // straddles a patch row and the group stride (SBO) is a whole swizzle repeat (2048 B).
  p.taps = a.kh * a.kw;
  {
    int want = 3;                                   // 0 off, 1 / 2 forced orientation, 3 auto
    if (const char* h = getenv("TRB_TC_HALO")) want = atoi(h);
    const bool ok = a.stride == 1 && p.KC == 64 && a.kh == a.kw && a.kh >= 3 && a.pad == a.kh / 2 &&
                    8 + 2 * a.pad <= 16;
    if (want && ok) {
      const long t1 = long(ceil_div(p.W_out, 8)) * ceil_div(p.H_out, 16) * p.N;
      const long t2 = long(ceil_div(p.W_out, 16)) * ceil_div(p.H_out, 8) * p.N;
      const long t0 = long(ceil_div(p.W_out, p.bw)) * ceil_div(p.H_out, p.bh) * ceil_div(p.N, p.bn);
      int mode = want == 3 ? (t2 < t1 ? 2 : 1) : want;
      const long th = mode == 1 ? t1 : t2;
      const int sms = num_sms();
      // Measured (profiles/r01_conv_microbench_*): the halo patch pays where re-reading
      // the input once per tap through L2 is the bound — 3x3 layers with K = 9*64/9*128.
      const bool worth = a.cin_pad <= 128 && a.kh == 3 &&
                         ceil_div(int(th * p.n_tiles), sms) * 100 <= ceil_div(int(t0 * p.n_tiles), sms) * 115;
      if (want != 3 || worth) {
        p.halo = mode;
        p.bw = mode == 1 ? 8 : 16; p.bh = mode == 1 ? 16 : 8; p.bn = 1;
      }
    }
  }
  {
    // Swap mode: 128-cout layers whose K is long enough for the main loop to dominate and whose
    // epilogue is the plain one (per-channel affine + activation, fp16 store).
    //   band variant:  N = bh full-width rows (<= 256 pixels), one TMA box per tap (like the
    //                  plain mode, 46 KB of operands per 128x240x64 block);
    //   patch variant: N = bw columns x 8 rows read through shifted descriptors from the
    //                  resident halo patch (orientation 2), so only the 16 KB filter block
    //                  streams per k-block — the one that pays for 5x5 / 7x7 filters.
    // Default off.  Measured on the 7x7 128->128 layers (profiles/r01_swap_modes.txt): the band
    // variant is 3-6 % faster alone and no faster with both OpenPose branches in flight; the patch
    // variant is bound by the ISSUING thread (~66 cycles per tcgen05.mma against 80 cycles of
    // tensor work at N = 160, plus the per-stage handshake), not by operand traffic.  Both are
    // parity-tested (tests/test_gpu_ops.py::test_conv_swap_mode) and kept for the next round.
    int want = 0;                                   // 0 off, 1 auto, 2 whenever possible, 3 band variant only
    if (const char* e = getenv("TRB_TC_SWAP")) want = atoi(e);
    const bool plain_epi = !a.res.ptr && !a.out2.ptr && !a.out_f32 && !a.shift9;
    const bool base_ok = a.stride == 1 && p.KC == 64 && a.cout_pad % 128 == 0 && plain_epi;
    const bool patch_ok = base_ok && a.kh == a.kw && a.kh >= 3 && a.pad == a.kh / 2 && 8 + 2 * a.pad <= 16;
    const bool patch_worth = a.cout_pad == 128 && p.taps >= 25;
    if (want && want != 3 && patch_ok && (want == 2 || patch_worth)) {
      const int tw = ceil_div(p.W_out, 32);
      int bw = ceil_div(p.W_out, tw);
      bw += bw & 1;                                 // N = 8 * bw must be a multiple of 16
      p.swap = 1; p.halo = 2;
      p.bw = bw; p.bh = 8; p.bn = 1;
      p.N_tile = 8 * bw;
      p.n_tiles = a.cout_pad / 128;
    } else if (want && !p.halo && base_ok && p.W_out <= 256) {
      // band of bh full-width rows: N = W_out * bh pixels, a multiple of 16, at most 256
      int best_bh = 0;
      double best_eff = 0;
      for (int bh = 1; bh <= p.H_out && bh * p.W_out <= 256; ++bh) {
        if ((bh * p.W_out) % 16) continue;
        const double eff = double(p.H_out) / (ceil_div(p.H_out, bh) * bh) * (bh * p.W_out >= 128 ? 1.0 : 0.5);
        if (eff >= best_eff) { best_eff = eff; best_bh = bh; }
      }
      const bool worth = a.cout_pad == 128 && p.k_blocks >= 8 && best_bh * p.W_out >= 192 && best_eff >= 0.9;
      if (best_bh && (want >= 2 || worth)) {
        p.swap = 1;
        p.bw = p.W_out; p.bh = best_bh; p.bn = 1;
        p.N_tile = p.bw * p.bh;                      // UMMA N = pixels of the band
        p.n_tiles = a.cout_pad / 128;                // filter (M) tiles
      }
    }
  }
  // Patch lines hold exactly the 8 + 2 pad pixels a tap's 8-pixel group can touch (TRB_TC_PW=16:
  // lines padded to 16 pixels = a whole number of 1024-byte swizzle repeats per line, the
  // round-1 layout).  The UMMA group stride then is (8 + 2 pad) * 128 B — not a multiple of the
  // swizzle repeat, which is fine because the 128B swizzle XORs ABSOLUTE address bits on both
  // the TMA write and the MMA read (lesson in `Halo mode`): 36 % fewer patch bytes for 3x3.
  p.pw = 16;
  {
    int want = 0;
    if (const char* e = getenv("TRB_TC_PW")) want = atoi(e);
    if (p.halo && !p.swap && want != 16) p.pw = 8 + 2 * a.pad;
  }
  p.rows = p.bw * p.bh * p.bn;
  p.tiles_w = ceil_div(p.W_out, p.bw);
  p.tiles_h = ceil_div(p.H_out, p.bh);
  p.tiles_n = ceil_div(p.N, p.bn);
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  {
    // CTA-pair kernel (cta_group::2): big-K layers with wide filter tiles, when pairing the
    // pixel tiles does not add a scheduling round.
    // Default off: measured slower than the single-CTA kernel on every BASELINE layer
    // (profiles/r01_conv_microbench_cta2.txt) — those layers are bound by the tensor pipe and
    // by tile/wave quantisation, not by filter traffic.  Kept (and covered by the parity
    // tests under TRB_TC_CTA2=1) as the base for B-multicast / larger-N work.
    int want = 0;                                   // 0 off, 1 forced, 2 auto
    if (const char* c = getenv("TRB_TC_CTA2")) want = atoi(c);
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int units = ceil_div(m_tiles, 2) * p.n_tiles;
    const int sms = num_sms();
    const bool ok = !p.halo && !p.swap && p.KC == 64 && p.N_tile % 32 == 0 && m_tiles >= 2;
    const bool worth = p.N_tile >= 128 && p.k_blocks >= 8 &&
                       ceil_div(units, sms / 2) <= ceil_div(p.total_tiles, sms);
    if (ok && (want == 1 || (want == 2 && worth))) {
      p.cta2 = 1;
      p.pair_units = units;
      p.idesc2 = (1u << 4) | (uint32_t(p.N_tile >> 3) << 17) | (uint32_t(256 >> 4) << 24);
    }
  }

  p.a_bytes = round_up((p.swap ? p.rows : 128) * p.KC * 2, 1024);   // pixel tile (the MMA reads 128 rows as A)
  p.b_bytes = round_up((p.swap ? 128 : p.cta2 ? p.N_tile / 2 : p.N_tile) * p.KC * 2, 1024);   // filter tile
  p.sub_bytes = p.halo ? p.b_bytes : p.a_bytes + p.b_bytes;
  // k-blocks per ring stage: enough tensor work (>= ~512 cycles = 8 MMAs of N=128) to
  // cover the single-warp issue latency, within ~64 KB per stage.
  const int mma_cycles = (p.KC / 16) * std::max(p.N_tile / 2, 8);     // per k-block
  const int k_per_group = p.halo ? p.taps : p.k_blocks;
  auto clamp_sub = [&](int v) { return std::max(1, std::min(std::min(kMaxSub, v), k_per_group)); };
  // A stage carries ~1024 cycles of tensor work within 96 KB (two stages still fit): the
  // issuing thread's per-stage cost (~500-600 cycles, lesson 5) is amortised over more MMAs.
  // Measured against the 512-cycle / 64 KB rule (profiles/r01_stage_size.txt): N = 256 layers
  // +2.5..5 %, halo layers at 92x163 +10..13 %, nothing slower.
  int stage_cycles = 1024;
  if (const char* e = getenv("TRB_TC_STAGE_CYCLES")) stage_cycles = atoi(e);
  const int stage_cap = stage_cycles > 512 ? 98304 : 65536;
  p.sub = clamp_sub(std::min(ceil_div(stage_cycles, mma_cycles), int(stage_cap / p.sub_bytes)));
  {
    // Two CTAs per SM for narrow filter tiles (see conv_tc_kernel): one k-block per stage so
    // that three stages fit in half of the SM's shared memory.
    int want = 1;                                   // 0 off, 1 auto, 2 whenever possible
    if (const char* e = getenv("TRB_TC_CTAS")) want = atoi(e);
    // (two issuing warps in ONE CTA keep full-size ring stages and measured faster — 53.7 vs
    // 66.4 us on the 7x7 128->128 layer — but need all 512 TMEM columns at N = 128: exclusive)
    int want_issuers = 1;
    if (const char* e = getenv("TRB_TC_ISSUERS")) want_issuers = atoi(e);
    const bool two_issuers = want_issuers && !p.halo && !p.cta2 && p.N_tile <= 128 &&
                             ceil_div(p.k_blocks, p.sub) >= (want_issuers == 2 ? 2 : 4);
    const bool ok = !two_issuers && !p.halo && !p.cta2 && !p.swap && p.N_tile <= 128 &&
                    3 * p.sub_bytes <= 100 * 1024;
    const bool worth = p.k_blocks >= 8 && p.total_tiles >= num_sms() + num_sms() / 2;
    if (ok && (want == 2 || (want == 1 && worth))) {
      plan->ctas_per_sm = 2;
      p.sub = 1;
    }
    // two issuers own alternate stages: keep at least three of them (<= 64 KB each)
    if (two_issuers) p.sub = clamp_sub(std::min(p.sub, int(65536 / p.sub_bytes)));
  }
  if (const char* s = getenv("TRB_TC_SUB")) p.sub = clamp_sub(atoi(s));
  {
    // two stages (+ the halo patches, parameters, barriers) must fit in the SM's 227 KB
    const uint32_t patches = p.halo ? 2 * round_up(((p.swap ? p.bw : 16) + 2 * a.pad) * p.pw * 128, 1024) : 0;
    const uint32_t limit = plan->ctas_per_sm == 2 ? 108u * 1024 : 224u * 1024;
    while (p.sub > 1 && patches + 2u * p.sub * p.sub_bytes + uint32_t(p.param_rows) * p.cout_pad * 4u + 2048u > limit) --p.sub;
  }
  p.iters = ceil_div(p.k_blocks, p.sub);
  {
    // Measured (profiles/r01_two_issuers.txt): N = 128 long-K layers 12-22 % faster (7x7 128->128
    // 68.8 -> 60.7 us), OpenPose forward 4.70 -> 4.47 ms, ArcFace 10.4 -> 9.7 ms, bit-reproducible.
    // Two hazards found and closed on the way: (1) combined with two CTAs per SM each CTA wanted
    // all 512 TMEM columns (now exclusive); (2) with an odd stage count a stage changes owner
    // every ring cycle and a parity wait passes spuriously when the other issuer lags a whole
    // phase (now ordered by the progress words in mma_issuer_alternate).
    int want = 1;                                   // 0 one issuer, 1 auto, 2 whenever possible
    if (const char* e = getenv("TRB_TC_ISSUERS")) want = atoi(e);
    // Halo layers: implemented (mma_issuer_alternate_halo), opt-in through TRB_TC_ISSUERS_HALO
    // until it has had the stress runs the non-halo path had.
    static const bool halo_ok = [] {
      const char* e = getenv("TRB_TC_ISSUERS_HALO");
      const char* r = getenv("TRB_TC_ROTATE");      // (the halo issuers walk the taps unrotated)
      return e && atoi(e) != 0 && !(r && atoi(r) != 0);
    }();
    const bool ok = (!p.halo || (halo_ok && !p.swap)) && !p.cta2 && p.iters >= 2 &&
                    p.N_tile <= 128 &&            // 4 accumulators in 512 TMEM columns,
                    plan->ctas_per_sm == 1;       // which one CTA per SM can have
    const int per_tile = p.halo ? p.kchunks * ceil_div(p.taps, p.sub) : p.iters;
    p.issuers = ok && (want == 2 || (want == 1 && per_tile >= 4)) ? 2 : 1;
  }
  const uint32_t param_bytes = uint32_t(p.param_rows) * p.cout_pad * 4u;
  if (p.halo) {
    p.patch_tx = ((p.swap ? p.bw : 16) + 2 * a.pad) * p.pw * 128;
    p.patch_bytes = round_up(p.patch_tx, 1024);
    p.ring_off = 2 * p.patch_bytes;
    // Resident filters (see TcParams::resident): whole filter bank + >= 3 patch buffers in 227 KB.
    int want = 1;      // (2 = a second issuing warp on the odd tiles: 203 vs 195 us on conv1_2, not faster)
    if (const char* e = getenv("TRB_TC_RESIDENT")) want = atoi(e);
    const uint32_t bank = uint32_t(p.taps) * p.b_bytes;
    const uint32_t room = 227u * 1024 - 1024 /*alignment*/ - 256 - 8 * (2 * kMaxStages + 12) - 16 - param_bytes;
    const int npatch = bank < room ? std::min<int>(kMaxStages - 1, std::min<int>(env_npatch(), (room - bank) / p.patch_bytes)) : 0;
    if (want && !p.swap && !p.cta2 && p.kchunks == 1 && p.n_tiles == 1 && plan->ctas_per_sm == 1 &&
        p.taps <= 64 && npatch >= 3) {
      p.resident = want >= 2 ? 2 : 1;      // TRB_TC_RESIDENT: 0 off, 1 one issuer (default), 2 two
      p.npatch = npatch;
      p.issuers = 1;
      p.sub = p.taps;                      // the "ring" is one stage that holds every tap
      p.iters = 1;
      p.ring_off = npatch * p.patch_bytes;
    }
  }
  p.iters_kc = ceil_div(p.taps, p.sub);
  p.stage_bytes = p.sub * p.sub_bytes;
  // the patch variant of swap mode keeps two (bw + 2 pad) x 16 pixel patches resident: use all 227 KB
  const uint32_t budget = p.swap && p.halo ? 224u * 1024 : plan->ctas_per_sm == 2 ? 108u * 1024 : kSmemBudget;
  p.stages = std::min(kMaxStages, int((budget - param_bytes - p.ring_off) / p.stage_bytes));
  p.stages = std::max(2, std::min(p.stages, (p.halo ? p.iters_kc * p.kchunks : p.iters) + 1));
  if (p.resident) p.stages = 1;
  p.sbo_bytes = 8u * p.KC * 2u;
  CUtensorMapSwizzle swz;
  if (p.KC == 64) { p.layout_type = 2; swz = CU_TENSOR_MAP_SWIZZLE_128B; }
  else if (p.KC == 32) { p.layout_type = 4; swz = CU_TENSOR_MAP_SWIZZLE_64B; }
  else { p.layout_type = 6; swz = CU_TENSOR_MAP_SWIZZLE_32B; }
  // kind::f16 instruction descriptor: D=f32, A=B=f16, K-major, N>>3, M>>4.
  p.idesc = (1u << 4) | (uint32_t(p.N_tile >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  p.acc_cols = p.issuers == 2 ? 2 * p.N_tile : p.swap ? 256 : p.N_tile;
  uint32_t cols = 32;
  while (cols < uint32_t(2 * p.acc_cols)) cols <<= 1;
  p.tmem_cols = cols;

  p.scale = a.scale; p.shift = a.shift; p.slope = a.slope;
  p.scale2 = a.scale2; p.shift2 = a.shift2;
  p.act = a.act;
  p.out = a.out.ptr; p.out_cs = a.out.cs; p.out_coff = a.out.coff; p.cout_store = a.cout_store;
  p.out2 = a.out2.ptr; p.out2_cs = a.out2.cs; p.out2_coff = a.out2.coff;
  p.res = a.res.ptr; p.res_cs = a.res.cs; p.res_coff = a.res.coff; p.res_up2 = a.res_up2;
  p.res_H = a.res.H; p.res_W = a.res.W;
  p.out_f32 = a.out_f32;
  p.pool = a.pool_out.ptr && p.halo && !p.swap && !p.cta2 && a.act == ACT_RELU && !a.res.ptr && !a.out2.ptr &&
           !a.out_f32 && !a.shift9 && a.pool_out.H == a.H_out / 2 && a.pool_out.W == a.W_out / 2;
  if (p.pool) {
    p.pool_out = a.pool_out.ptr; p.pool_cs = a.pool_out.cs; p.pool_coff = a.pool_out.coff;
    p.pool_H = a.pool_out.H; p.pool_W = a.pool_out.W;
  }
  p.err = tc_error_flag();
  if (const char* dbg = getenv("TRB_TC_DEBUG")) p.debug = atoi(dbg);
  p.pdl_late = 1;
  p.rotate = 0;   // measured: no effect (profiles/r01_swap_modes.txt) — the filter stream is not L2 hot-spot bound
  if (const char* e = getenv("TRB_TC_ROTATE")) p.rotate = atoi(e);
  if (const char* e = getenv("TRB_TC_PDL_LATE")) p.pdl_late = atoi(e);
  TR_CHECK(a.scale && a.shift, "epilogue scale/shift are required");
  TR_CHECK(a.act != ACT_PRELU || a.slope, "PReLU needs slopes");
  TR_CHECK(!a.out2.ptr || (a.scale2 && a.shift2), "second output needs scale2/shift2");
  TR_CHECK(!a.shift9 || (a.kh == 3 && a.pad == 1 && a.stride == 1 && !p.cta2), "border-class shifts are for 3x3 pad-1 stride-1 layers");

  // ---- tensor maps
  auto encode = encode_fn();
  const cuuint64_t cs = a.in.cs, W = a.in.W, H = a.in.H, N = a.in.N;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  if (p.halo == 2) {        // dims (C, H, W, N): patch rows run along H
    gdim[0] = cs; gdim[1] = H; gdim[2] = W; gdim[3] = N; gdim[4] = 1;
    gstr[0] = W * cs * 2; gstr[1] = cs * 2; gstr[2] = H * W * cs * 2; gstr[3] = N * H * W * cs * 2;
    box[0] = p.KC; box[1] = p.pw; box[2] = (p.swap ? p.bw : 16) + 2 * a.pad; box[3] = 1; box[4] = 1;
  } else if (a.stride == 1) {
    gdim[0] = cs; gdim[1] = W; gdim[2] = H; gdim[3] = N; gdim[4] = 1;
    gstr[0] = cs * 2; gstr[1] = W * cs * 2; gstr[2] = H * W * cs * 2; gstr[3] = N * H * W * cs * 2;
    box[0] = p.KC; box[1] = p.bw; box[2] = p.bh; box[3] = p.bn; box[4] = 1;
    if (p.halo == 1) { box[1] = p.pw; box[2] = 16 + 2 * a.pad; box[3] = 1; }
  } else {
    gdim[0] = 2 * cs; gdim[1] = W / 2; gdim[2] = 2; gdim[3] = H / 2; gdim[4] = N;
    gstr[0] = 2 * cs * 2; gstr[1] = W * cs * 2; gstr[2] = 2 * W * cs * 2; gstr[3] = H * W * cs * 2;
    box[0] = p.KC; box[1] = p.bw; box[2] = 1; box[3] = p.bh; box[4] = p.bn;
  }
  CUresult r = encode(&plan->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, a.in.ptr, gdim, gstr, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: " + std::to_string(int(r)));
  const cuuint64_t ktot = cuuint64_t(a.kh) * a.kw * a.cin_pad;
  cuuint64_t wdim[2] = {ktot, cuuint64_t(a.cout_pad)};
  cuuint64_t wstr[1] = {ktot * 2};
  cuuint32_t wbox[2] = {cuuint32_t(p.KC), cuuint32_t(p.swap ? 128 : p.cta2 ? p.N_tile / 2 : p.N_tile)};
  r = encode(&plan->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(a.w), wdim, wstr,
             wbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: " + std::to_string(int(r)));

  plan->grid = p.cta2 ? 2 * std::min(p.pair_units, num_sms() / 2)
                      : std::min(p.total_tiles, plan->ctas_per_sm * num_sms());
  plan->smem = p.ring_off + p.stages * p.stage_bytes + param_bytes + 1024 /*alignment*/ +
               256 /*tap table*/ + 8 * (2 * kMaxStages + 12) + 16;
  {
    // Stream-K where whole-tile scheduling leaves SMs idle in the last round.  Measured
    // (profiles/r01_stream_k.txt): the partial-tile hand-over costs ~8 us per launch with
    // N = 128 tiles and ~18 us with N = 256, so it pays only where the idle tail is longer.
    int want = 1;                                     // 0 off, 1 auto, 2 whenever possible
    if (const char* e = getenv("TRB_TC_SK")) want = atoi(e);
    const int rounds = ceil_div(p.total_tiles, plan->grid);
    const double tile_us = 2.0 * 128 * p.N_tile * p.k_blocks * p.KC / 7.5e6;   // ~7.5 TFLOP/s per SM
    const double saved_us = (rounds - double(p.total_tiles) / plan->grid) * tile_us;
    const bool ok = !p.halo && !p.cta2 && p.total_tiles > plan->grid && p.iters >= 2 &&
                    plan->grid < int(kSkFlagBytes / 4);
    const bool worth = p.iters >= 4 && saved_us >= 14.0;
    if (ok && (want == 2 || (want == 1 && worth))) {
      void* scratch = a.sk_scratch;
      if (!scratch) {
        TR_CUDA(cudaMalloc(&plan->sk_own, conv_tc_sk_scratch_bytes()));
        TR_CUDA(cudaMemset(plan->sk_own, 0, kSkFlagBytes));
        scratch = plan->sk_own;
      }
      p.sk = 1;
      p.sk_flags = static_cast<int*>(scratch);
      p.sk_ws = reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + kSkFlagBytes);
    }
  }
  plan->flops = 2.0 * p.N * p.H_out * p.W_out * double(a.cout_pad) * a.kh * a.kw * a.cin_pad;
  // lean instantiations (see MODE of conv_tc_kernel)
  {
    static const bool lean = [] { const char* e = getenv("TRB_TC_LEAN"); return !e || atoi(e) != 0; }();
    const bool base_ok = lean && p.KC == 64 && plan->ctas_per_sm == 1 && !p.swap && !p.debug && !p.out2 && !p.rotate && !p.cta2;
    plan->mode = 0;
    if (base_ok && p.resident == 1 && !p.out_f32) plan->mode = 1;
    else if (base_ok && !p.halo && !p.resident) plan->mode = 2;
  }
  kernel_for(p.KC, plan->ctas_per_sm, plan->mode);
  return plan;
}

void conv_tc_plan_destroy(ConvTcPlan* p) {
  if (p->sk_own) cudaFree(p->sk_own);
  delete p;
}

double conv_tc_plan_flops(const ConvTcPlan* p) { return p->flops; }

bool conv_tc_plan_pooled(const ConvTcPlan* p) { return p->p.pool != 0; }

int* conv_tc_error_flag() { return tc_error_flag(); }

// 0, or the code of the mbarrier wait that timed out (1 ring empty, 2 tmem empty, 3 ring full,
// 4 tmem full, 5 patch empty, 6 patch full, 9 stream-K partial) in a kernel that then trapped.
int conv_tc_last_timeout() { return g_err_host ? *reinterpret_cast<volatile int*>(g_err_host) : 0; }

void conv_tc_launch(const ConvTcPlan* plan, cudaStream_t s) {
  if (plan->p.cta2) {
    static bool attr_set[kMaxDevices] = {};
    const int dev = current_device();
    if (!attr_set[dev]) {
      TR_CUDA(cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
      attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = plan->smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    TR_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel, plan->tmA, plan->tmB, plan->p));
    return;
  }
  static const bool pdl = [] { const char* e = getenv("TRB_TC_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(kThreads1);
  cfg.dynamicSmemBytes = plan->smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  TR_CUDA(cudaLaunchKernelEx(&cfg, kernel_for(plan->p.KC, plan->ctas_per_sm, plan->mode), plan->tmA, plan->tmB, plan->p));
}

}  // namespace trb
