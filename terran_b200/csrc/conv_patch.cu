// Implicit-GEMM convolution on tcgen05, "filters on M, pixels from a resident patch on N".
//
//   D^T[cout, pixel] = sum over channel chunks kc and taps (r,s) of
//                      W_tap,kc[cout, 64] * X_patch(r,s)[pixel, 64]^T
//
// Why a second tensor-core kernel (measured in round 1, profiles/r01_swap_modes.txt and
// r01_conv_microbench_v2_debugmodes.txt): conv_tc_kernel loads one im2col window per filter
// tap, so a 7x7 layer pulls every input pixel 49 times through L2 -> shared memory
// (738 MB per launch for 9 MB of operands); with the TMA half or the MMA half of the pipeline
// alone the 7x7 128->128 layer takes the same 47 us — the L2 -> SM operand supply, not the tensor
// pipe, bounds every layer whose per-SM tile is 128 x 128.  Here
//   * the input patch of an output tile INCLUDING the filter halo is loaded once per
//     64-channel chunk and stays resident; every tap's B operand is a UMMA descriptor whose
//     start address is shifted inside that patch (rows of 16 pixels = 2048 B = the
//     descriptor's group stride, so an 8-pixel group never straddles a patch row);
//   * the FILTER block (128 couts x 64 channels = 16 KB per tap and chunk) is the only operand
//     that streams, through a deep ring;
//   * the 128 output channels sit on the UMMA M dimension and the pixels on N, so N = 8 R can
//     be chosen per layer (R rows of 8 pixels, N <= 256): one instruction carries up to twice
//     the work of the N = 128 tile, and L2 -> SM traffic per flop drops with N.
// TMEM holds D^T (lane = cout, column = pixel); the epilogue thread of a lane owns one output
// channel, so per-channel parameters are registers and a warp writes 64-byte runs.
//
// Warp roles (352 threads, one CTA per SM, persistent over tiles):
//   warp 0      filter producer (TMA 2-D, ring of `stages` x `sub` filter blocks) — does NOT wait
//               for the previous layer (weights are constant): the ring fills during the
//               previous kernel's tail under programmatic dependent launch
//   warp 1      TMEM allocator + tcgen05.mma issuer
//   warps 2-9   epilogue (tcgen05.ld -> scale/shift/act/residual -> HBM), double-buffered TMEM
//   warp 10     patch producer (TMA 4-D, two patch buffers; waits for the previous layer)
//
// Replaces the cuDNN convolutions behind the stride-1 nn.Conv2d layers with >= 128 filters of
// openpose/model.py:41-95 (and arcface/model.py:11-35 where eligible).
#include <cudaTypedefs.h>

#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace trb {

namespace {

constexpr int kPThreads = 352;            // 8 epilogue warps, one CTA per SM
constexpr int kPThreadsDual = 224;        // 4 epilogue warps, two CTAs per SM
constexpr int kPThreadsWide = 480;        // 12 epilogue warps (three teams): short-K layers whose epilogue is the bound
constexpr int kPMaxStages = 12;
constexpr int kPMaxPatchBufs = 8;         // patch buffers (2 normally; more for small 1x1 patches)
constexpr uint32_t kFilterBlock = 128 * 64 * 2;      // one (tap, chunk) filter block: 16 KB
// Epilogue staging: per epilogue warp a ring of kStageSlots slots of 16 pixels x 32 couts fp16
// (1 KB), so that kStageSlots TMA stores of a warp are in flight and a chunk only waits for
// the store issued kStageSlots chunks earlier.
constexpr int kStageSlots = 2;
constexpr uint32_t kOutStage = 12 * kStageSlots * 1024;      // (up to 12 epilogue warps)

struct PatchParams {
  int N, H, W;                 // output == input dims (stride 1, "same" padding)
  int k, pad, taps;
  int cin_pad, kchunks, in_coff;
  int cout_tiles, tiles_per_group;   // grouped conv: cout tile ct reads channels in_coff + (ct / tiles_per_group) * cin_pad
  int axis;                    // 0: 8-pixel groups along w, R rows along h;  1: groups along h, R along w
  int R, NP, PA;               // NP = 8 R pixels per tile (UMMA N); PA = patch pixels per row (8 | 16)
  int tiles_a, tiles_b, pix_tiles, total_tiles;
  int sub, iters, stages;      // taps per ring stage, stages per chunk, ring depth
  uint32_t stage_bytes, patch_bytes, patch_tx, ring_off;
  int tma_store;               // 1: epilogue transposes through smem and stores with TMA (UTMASTG)
  // Stacked images (small maps: ArcFace 14x14, 7x7): a tile takes `stack` consecutive images
  // along the R axis.  The patch holds their (B + 2 pad)-line blocks one after the other, so
  // the B operand keeps ONE group stride; R = stack * bstride - 2 pad lines are computed, the
  // 2 pad lines between two images are garbage nobody stores.  N = 8 R instead of 8 B columns
  // per instruction: 240 instead of 112 for 14x14 maps (whose MMAs were issue-bound).
  int stack, bstride;          // stack = 1: off; bstride = B + 2 pad
  int pool;                    // 1: 2x2 / stride-2 max-pool fused into the (staged) epilogue: the output
                               // maps describe the POOLED tensor (floor(H/2) x floor(W/2))
  // Stream-K: the CTAs split the launch's (tile, ring iteration) sequence into equal contiguous
  // ranges instead of whole tiles, so no SM idles in a last partial round and a launch with
  // fewer tiles than SMs still uses all of them (split-K).  A range that starts inside a tile
  // computes that tile's remaining iterations FIRST, parks the raw fp32 accumulator in its
  // workspace slot and raises its flag; the CTA that owns the tile's first iteration reaches
  // it LAST in its own range, adds the parked partials in CTA order (deterministic) and runs
  // the epilogue; the last of the 2 x epi_warps warps that touch a flag re-arms it.
  // Two CTAs per SM ("dual"): each with ONE patch buffer, ONE accumulator (<= 256 TMEM columns),
  // four epilogue warps and half the shared memory.  Everything a CTA does outside its MMA loop —
  // launch, barrier / TMEM set-up, waiting for the previous layer, the first patch, the chunk
  // switch, draining the tensor pipe, the epilogue — runs while the other CTA's MMAs keep the
  // SM's tensor pipe busy.
  int epi_warps, nbuf, nacc;
  int sk, ipt;                 // ipt: ring iterations per tile = kchunks * iters
  float* sk_ws; int* sk_flags;
  uint32_t idesc, tmem_cols;
  const float* scale; const float* shift; const float* slope;
  const float* shift9; int cout_pad;   // optional [9][cout_pad] border-class shifts (see ConvArgs)
  int act;
  __half* out; int out_cs, out_coff, cout_store;
  const __half* res; int res_cs, res_coff;
  float* out_f32;
  int* err;
  int pdl_late;
  unsigned long long* trace;   // debug & 32: globaltimer stamps of CTA 0, 16 per launch
  int debug;                   // timing experiments: 1 no filter TMA, 2 no MMA, 4 no stores,
                               // 8 no ring-release handshake (with 1), 16 no full-barrier waits (with 1),
                               // 64 no epilogue (with TRB_PT_SK=0)
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7])
      : "r"(taddr));
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// debug & 32: one lane of CTA 0 stamps event `ev` of this launch
#define PT_STAMP(ev)                                                             \
  do {                                                                           \
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0)                   \
      p.trace[1 + 16 * (trace_slot & 63) + (ev)] = global_ns();                  \
  } while (0)

struct TileCoord {
  int ct, n, a0, b0;
};
__device__ __forceinline__ TileCoord tile_coord(const PatchParams& p, int tile) {
  TileCoord t;
  t.ct = tile / p.pix_tiles;
  int pt = tile - t.ct * p.pix_tiles;
  const int ta = pt % p.tiles_a; pt /= p.tiles_a;
  const int tb = pt % p.tiles_b;
  t.n = pt / p.tiles_b * p.stack;
  t.a0 = ta * 8;
  t.b0 = tb * p.R;
  return t;
}

// One CTA's sequence of tile segments [i0, i1) (in ring iterations of the tile).  32-bit
// arithmetic: the plan only enables stream-K when total_tiles * ipt * grid < 2^32 (a 64-bit
// division is ~100 inlined instructions, and every role walks).
struct Walk {
  unsigned pos, end;
  int tile;
};
struct Seg {
  int tile, i0, i1;
};
__device__ __forceinline__ unsigned walk_bound(const PatchParams& p, unsigned cta) {
  return static_cast<unsigned>(p.total_tiles) * static_cast<unsigned>(p.ipt) * cta / gridDim.x;
}
__device__ __forceinline__ Walk walk_begin(const PatchParams& p) {
  Walk w;
  w.pos = p.sk ? walk_bound(p, blockIdx.x) : 0u;
  w.end = p.sk ? walk_bound(p, blockIdx.x + 1) : 0u;
  w.tile = blockIdx.x;
  return w;
}
__device__ __forceinline__ bool walk_next(const PatchParams& p, Walk& w, Seg& s) {
  if (p.sk) {
    if (w.pos >= w.end) return false;
    const unsigned ipt = static_cast<unsigned>(p.ipt);
    const unsigned tile = w.pos / ipt;
    s.tile = static_cast<int>(tile);
    s.i0 = static_cast<int>(w.pos - tile * ipt);
    s.i1 = static_cast<int>(min(ipt, static_cast<unsigned>(s.i0) + (w.end - w.pos)));
    w.pos += static_cast<unsigned>(s.i1 - s.i0);
    return true;
  }
  if (w.tile >= p.total_tiles) return false;
  s.tile = w.tile; s.i0 = 0; s.i1 = p.ipt;
  w.tile += gridDim.x;
  return true;
}
__device__ __forceinline__ bool walk_last(const PatchParams& p, const Walk& w) {
  return p.sk ? w.pos >= w.end : w.tile >= p.total_tiles;
}

// MODE: what the EPILOGUE of this instantiation can do (bits: 1 stream-K partials, 2 border-class
// shifts, 4 residual, 8 everything else: the unstaged store path, fp32 output, the timing /
// trace debug modes).  The epilogue is ~900 SASS instructions per 16-pixel chunk with every
// path compiled in; a layer launches the smallest instantiation that covers it, so the hot
// loop of the common layers (plain fp16 output through the staged TMA store) stays short.
constexpr int kModeSK = 1, kModeS9 = 2, kModeRes = 4, kModeAll = 15, kModePool = 16;
template <int THREADS, int MINB, int MODE>
__global__ void __launch_bounds__(THREADS, MINB)
conv_patch_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO32,
                  const PatchParams pk) {
  // the features this instantiation leaves out become compile-time constants
  PatchParams p = pk;
  if (!(MODE & kModeSK)) p.sk = 0;
  if (!(MODE & kModeS9)) p.shift9 = nullptr;
  if (!(MODE & kModeRes)) p.res = nullptr;
  if (!(MODE & 8)) { p.tma_store = 1; p.debug = 0; p.trace = nullptr; p.out_f32 = nullptr; }
  p.pool = (MODE & kModePool) ? 1 : 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // two patch buffers first
  const uint32_t ring = base + p.ring_off;
  const uint32_t stage_out = ring + p.stages * p.stage_bytes;       // epilogue staging: 2 teams x 4 KB
  const uint32_t tab_s = stage_out + kOutStage;                     // per-tap descriptor offsets
  const uint32_t bars = tab_s + 512u;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kPMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kPMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kPMaxStages + 2 + a); };
  auto pfull_bar = [&](int b) { return bars + 8u * (2 * kPMaxStages + 4 + b); };
  auto pempty_bar = [&](int b) { return bars + 8u * (2 * kPMaxStages + 4 + kPMaxPatchBufs + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * kPMaxStages + 4 + 2 * kPMaxPatchBufs);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  volatile unsigned* trace_slot_s =
      reinterpret_cast<volatile unsigned*>(smem_raw + (bars + 8u * (2 * kPMaxStages + 5 + 2 * kPMaxPatchBufs) - smem_u32(smem_raw)));
  if (p.trace && threadIdx.x == 0) {
    *trace_slot_s = blockIdx.x == 0 ? static_cast<unsigned>(atomicAdd(p.trace, 1ULL)) : 0u;
    if (blockIdx.x == 0) p.trace[1 + 16 * (*trace_slot_s & 63) + 0] = global_ns();
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (p.tma_store) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO32) : "memory");
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), p.epi_warps);
    }
    for (int b = 0; b < p.nbuf; ++b) {
      mbar_init(pfull_bar(b), 1);
      mbar_init(pempty_bar(b), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2) {
    // what tap (r, s) adds to the patch's descriptor start address, in 16-byte units
    for (int t = lane; t < p.taps; t += 32) {
      const int r = t / p.k, s = t - r * p.k;
      const int ta = p.axis == 0 ? s : r, tb = p.axis == 0 ? r : s;
      const uint32_t v = static_cast<uint32_t>(tb * p.PA + ta) * 8u;
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(tab_s + 4u * t), "r"(v) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const unsigned trace_slot = p.trace ? *trace_slot_s : 0u;
  if (warp == 3) PT_STAMP(1);                               // prologue done

  const int grid = gridDim.x;
  // Every role loop runs with the whole warp converged; one elected lane issues (see
  // conv_tc.cu, lesson 1: issuing from a divergent region costs ~240 cycles per instruction).
  if (warp == 0) {
    // ---------------------------------------------------------- filter producer
    int stage = 0;
    uint32_t phase = 0;
    if (!p.pdl_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    Walk w = walk_begin(p);
    Seg sg;
    while (walk_next(p, w, sg)) {
      if (p.pdl_late && walk_last(p, w))
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      const int ct = sg.tile / p.pix_tiles;
      int kc = sg.i0 / p.iters, it = sg.i0 - kc * p.iters;
      for (int i = sg.i0; i < sg.i1; ++i) {
        const int tap = it * p.sub;
        const int nsub = min(p.sub, p.taps - tap);
        mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 1);
        const uint32_t dst = ring + stage * p.stage_bytes;
        if (p.debug & 1) {
          if (elect_one()) mbar_arrive(full_bar(stage));
        } else if (elect_one()) {
          mbar_expect_tx(full_bar(stage), nsub * kFilterBlock);
          for (int j = 0; j < nsub; ++j)
            tma_load_2d(dst + j * kFilterBlock, &tmW, full_bar(stage),
                        (tap + j) * p.cin_pad + kc * 64, ct * 128);
        }
        __syncwarp();
        if (++it == p.iters) { it = 0; ++kc; }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 2 + p.epi_warps) {
    // ----------------------------------------------------------- patch producer
    asm volatile("griddepcontrol.wait;" ::: "memory");      // activations of the previous layer
    PT_STAMP(2);                                            // previous grid complete
    int q = 0;
    Walk w = walk_begin(p);
    Seg sg;
    while (walk_next(p, w, sg)) {
      const TileCoord t = tile_coord(p, sg.tile);
      const int kc0 = sg.i0 / p.iters, kc1 = (sg.i1 - 1) / p.iters;   // chunks the segment touches
      for (int kc = kc0; kc <= kc1; ++kc, ++q) {
        const int buf = q % p.nbuf;
        mbar_wait(pempty_bar(buf), ((q / p.nbuf) & 1u) ^ 1u, p.err, 5);
        if (elect_one()) {
          mbar_expect_tx(pfull_bar(buf), p.patch_tx);
          const int c0 = p.in_coff + (t.ct / p.tiles_per_group) * p.cin_pad + kc * 64;
          if (p.stack == 1) {
            tma_load_4d(base + buf * p.patch_bytes, &tmX, pfull_bar(buf), c0, t.a0 - p.pad, t.b0 - p.pad, t.n);
          } else {
            // one box per image (an image index beyond the batch is zero-filled and still counted)
            for (int si = 0; si < p.stack; ++si)
              tma_load_4d(base + buf * p.patch_bytes + static_cast<uint32_t>(si * p.bstride * p.PA) * 128u, &tmX,
                          pfull_bar(buf), c0, t.a0 - p.pad, -p.pad, t.n + si);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer
    int stage = 0, q = 0, tile_it = 0;
    uint32_t phase = 0;
    const uint32_t w_hi = umma_desc_hi(1024, 2);                       // filters: 8 rows x 128 B groups
    const uint32_t x_hi = umma_desc_hi(static_cast<uint32_t>(p.PA) * 128u, 2);   // one patch row per group
    const uint32_t ring_lo = umma_desc_lo(ring);
    const uint32_t stage_step = p.stage_bytes >> 4;
    Walk w = walk_begin(p);
    Seg sg;
    for (; walk_next(p, w, sg); ++tile_it) {
      const int acc = tile_it % p.nacc;
      mbar_wait(tempty_bar(acc), ((tile_it / p.nacc) & 1u) ^ 1u, p.err, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + acc * p.NP;
      uint32_t accumulate = 0;
      int kc = sg.i0 / p.iters, it = sg.i0 - kc * p.iters;
      int buf = 0;
      uint32_t patch_lo = 0;
      bool need_patch = true;
      for (int i = sg.i0; i < sg.i1; ++i) {
        if (need_patch) {                                   // first iteration of a chunk in this segment
          buf = q % p.nbuf;
          mbar_wait(pfull_bar(buf), (q / p.nbuf) & 1u, p.err, 6);
          if (q == 0) PT_STAMP(3);                          // first patch landed
          patch_lo = umma_desc_lo(base + buf * p.patch_bytes);
          need_patch = false;
        }
        const int tap = it * p.sub;
        const int nsub = min(p.sub, p.taps - tap);
        const bool chunk_end = it == p.iters - 1 || i == sg.i1 - 1;
        mbar_wait(full_bar(stage), phase, p.err, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t w_lo0 = ring_lo + stage * stage_step;
          if (!(p.debug & 2)) {
            for (int j = 0; j < nsub; ++j) {
              uint32_t tap_off;
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tap_off) : "r"(tab_s + 4u * (tap + j)));
              const uint32_t x_lo = patch_lo + tap_off;
              const uint32_t w_lo = w_lo0 + j * (kFilterBlock >> 4);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {            // +32 B (two 16-byte units) per K = 16 step
                umma_f16(d_tmem, w_lo + 2 * ks, w_hi, x_lo + 2 * ks, x_hi, p.idesc, accumulate);
                accumulate = 1;
              }
            }
          }
          umma_commit(empty_bar(stage));
          if (chunk_end) umma_commit(pempty_bar(buf));
          if (i == sg.i1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (chunk_end) { ++q; need_patch = true; }
        if (++it == p.iters) { it = 0; ++kc; }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      PT_STAMP(4);                                          // all MMAs of a segment issued
    }
  } else {
    // ----------------------------------------------------------------- epilogue
    asm volatile("griddepcontrol.wait;" ::: "memory");      // residual reads / buffer re-use
    const int ew = warp - 2;
    const int qd = warp & 3;                 // TMEM lane quarter this warp may access
    const int team = ew >> 2;                // teams of four warps split the pixel groups
    // the R / 2 chunks (pairs of lines) are dealt to the teams as evenly as possible
    const int n_teams = p.epi_warps >> 2, chunks = p.R >> 1;
    const int c_base = chunks / n_teams, c_rem = chunks - c_base * n_teams;
    const int g_begin = 2 * (team * c_base + min(team, c_rem));
    const int g_end = g_begin + 2 * (c_base + (team < c_rem ? 1 : 0));
    const int A_dim = p.axis == 0 ? p.W : p.H, B_dim = p.axis == 0 ? p.H : p.W;
    // line g of a tile -> (image offset, position along the R axis); a stacked tile has
    // garbage lines (b >= B_dim) between its images
    auto line_img = [&](int g) { return p.stack == 1 ? 0 : g / p.bstride; };
    auto line_b = [&](int b0, int g) { return p.stack == 1 ? b0 + g : g - (g / p.bstride) * p.bstride; };
    const int a_step = p.axis == 0 ? 1 : p.W;             // pixel-index step along the group axis
    const int b_step = p.axis == 0 ? p.W : 1;
    unsigned chunk_ctr = 0;                               // staging slot = chunk_ctr % kStageSlots (across tiles)
    int tile_it = 0;
    Walk w = walk_begin(p);
    Seg sg;
    int par_ct = -1;                                        // cout tile whose parameters are in registers
    float sc = 0.f, sh = 0.f, slope = 1.f, s9[9];
    for (; walk_next(p, w, sg); ++tile_it) {
      const int acc = tile_it % p.nacc;
      const TileCoord t = tile_coord(p, sg.tile);
      const int cl = qd * 32 + lane;                       // cout within the tile
      const int cout = t.ct * 128 + cl;
      const bool c_in = cout < p.cout_pad;                 // (a 64-filter layer fills half a tile)
      // Per-channel parameters are re-read only when the cout tile changes (consecutive tiles
      // of a CTA almost always share it): when the epilogue is the bound, the L2 round trip of
      // these loads sits exposed at the head of every tile.
      // Border-class shifts: the three candidates of a pixel group (its position along the R
      // axis is fixed) are picked once per group, the one of a pixel by its position along the
      // 8-pixel axis.  Without shift9 all nine are the plain shift.
      if (t.ct != par_ct) {
        par_ct = t.ct;
        sc = c_in ? p.scale[cout] : 0.f;
        sh = c_in ? p.shift[cout] : 0.f;
        slope = (p.act == ACT_PRELU && c_in) ? p.slope[cout] : 1.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) s9[k] = (p.shift9 && c_in) ? p.shift9[k * p.cout_pad + cout] : sh;
      }
      auto group_shifts = [&](int g, float& first, float& inner, float& last) {
        const int b = line_b(t.b0, g);
        const int bc = b == 0 ? 0 : (b >= B_dim - 1 ? 2 : 1);
        if (p.axis == 0) {               // b = output row, the 8-pixel axis runs along columns
          first = bc == 0 ? s9[0] : (bc == 1 ? s9[3] : s9[6]);
          inner = bc == 0 ? s9[1] : (bc == 1 ? s9[4] : s9[7]);
          last = bc == 0 ? s9[2] : (bc == 1 ? s9[5] : s9[8]);
        } else {                         // b = output column, the 8-pixel axis runs along rows
          first = bc == 0 ? s9[0] : (bc == 1 ? s9[1] : s9[2]);
          inner = bc == 0 ? s9[3] : (bc == 1 ? s9[4] : s9[5]);
          last = bc == 0 ? s9[6] : (bc == 1 ? s9[7] : s9[8]);
        }
      };
      auto pixel_shift = [&](int i, float first, float inner, float last) {
        const int a = t.a0 + i;
        return a == 0 ? first : (a >= A_dim - 1 ? last : inner);
      };
      // Residual (staged paths): the residual tile is first copied into the staging tile with
      // coalesced 16-byte loads ([pixel][cout] layout, like the output), each thread then adds
      // its channel's value in fp32 before the single rounding to fp16.  Row r of a staging
      // block that starts at group g0 is pixel (g0 + r / 8, r % 8).
      auto res_row_ptr = [&](int g0, int r) -> const __half* {
        const int gg = g0 + (r >> 3), i = r & 7;
        const int b = line_b(t.b0, gg), a = t.a0 + i, img = t.n + line_img(gg);
        if (gg >= g_end || b >= B_dim || a >= A_dim || img >= p.N) return nullptr;
        const unsigned pix = static_cast<unsigned>(img) * p.H * p.W + a * a_step + b * b_step;
        return p.res + static_cast<size_t>(pix) * p.res_cs + p.res_coff + t.ct * 128;
      };
      // one activation formula: y = max(y, 0) + neg * min(y, 0)   (ReLU 0, PReLU slope, none 1)
      const float neg = p.act == ACT_RELU ? 0.f : slope;
      mbar_wait(tfull_bar(acc), (tile_it / p.nacc) & 1u, p.err, 4);
      if (warp == 2) PT_STAMP(5);                           // accumulator complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + acc * p.NP;
      if (p.debug & 64) {                                   // timing only (with TRB_PT_SK=0): no epilogue at all
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        continue;
      }
      if (sg.i0 > 0) {
        // Stream-K: a tile whose first iterations belong to an earlier CTA.  Park the raw
        // accumulator as [group][lane][8]: a warp writes 1 KB runs.
        float4* ws = reinterpret_cast<float4*>(p.sk_ws + static_cast<size_t>(blockIdx.x) * 128 * p.NP);
        for (int g = g_begin; g < g_end; ++g) {
          uint32_t v[8];
          __syncwarp();
          tmem_ld8_async(taddr + g * 8, v);
          tmem_ld_wait();
          float4* dst = ws + (g * 128 + cl) * 2;
          __stcg(dst, make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]),
                                  __uint_as_float(v[3])));
          __stcg(dst + 1, make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]),
                                      __uint_as_float(v[6]), __uint_as_float(v[7])));
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
      int n_parts = 0;
      if (sg.i1 < p.ipt) {
        // The rest of this tile was the FIRST thing the next CTA(s) computed.
        const unsigned tile_end = static_cast<unsigned>(sg.tile + 1) * static_cast<unsigned>(p.ipt);
        for (int b = blockIdx.x + 1; b < grid && walk_bound(p, b) < tile_end; ++b) ++n_parts;
        if (lane == 0) {
          for (int k = 1; k <= n_parts; ++k) {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile int*>(p.sk_flags + blockIdx.x + k) < p.epi_warps) {
              if (clock64() - t0 > 4000000000LL) {
                if (p.err) atomicExch(p.err, 9);
                __threadfence_system();
                __trap();
              }
            }
          }
          __threadfence();
        }
        __syncwarp();
      }
      const float* part0 = p.sk_ws + static_cast<size_t>(blockIdx.x + 1) * 128 * p.NP;
      // v[0..7] += the parked partials of group g, in CTA order
      auto add_parts = [&](uint32_t* v, int g) {
        // four partials in flight at a time (a long split-K tail would otherwise pay one L2
        // round trip per partial), added in CTA order
        for (int k0 = 0; k0 < n_parts; k0 += 4) {
          float4 f[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (k0 + j < n_parts) {
              const float4* src = reinterpret_cast<const float4*>(part0 + static_cast<size_t>(k0 + j) * 128 * p.NP) +
                                  (g * 128 + cl) * 2;
              f[j][0] = __ldcg(src); f[j][1] = __ldcg(src + 1);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (k0 + j < n_parts) {
              v[0] = __float_as_uint(__uint_as_float(v[0]) + f[j][0].x);
              v[1] = __float_as_uint(__uint_as_float(v[1]) + f[j][0].y);
              v[2] = __float_as_uint(__uint_as_float(v[2]) + f[j][0].z);
              v[3] = __float_as_uint(__uint_as_float(v[3]) + f[j][0].w);
              v[4] = __float_as_uint(__uint_as_float(v[4]) + f[j][1].x);
              v[5] = __float_as_uint(__uint_as_float(v[5]) + f[j][1].y);
              v[6] = __float_as_uint(__uint_as_float(v[6]) + f[j][1].z);
              v[7] = __float_as_uint(__uint_as_float(v[7]) + f[j][1].w);
            }
          }
        }
      };
      if (p.tma_store) {
        // WARP-LOCAL staged epilogue.  A warp owns 32 output channels (its TMEM lane quarter)
        // of the team's pixel groups: 16 pixels at a time go TMEM -> registers -> a private
        // [16 pixels][32 couts] fp16 slot in shared memory (a warp writes the 64 contiguous bytes
        // of a pixel) -> two TMA stores of {32 channels x 8 pixels} (64-byte runs; the tensor map
        // clips the ragged border).  No barrier between warps: the team-wide version (one
        // [16][128] tile, three bar.sync per chunk) cost 8.7 us per 256-pixel tile against
        // 2.3-4.7 us of MMA time (profiles/r02_patch_epilogue.txt; reading the slot back and
        // storing with st.global.v4 instead of TMA measured 20-25 % slower).  The TMEM read, the
        // residual and the first stream-K partial of the NEXT chunk are in flight while this
        // chunk is staged; kStageSlots stores in flight per warp.
        // The CTA's LAST segment (`whole`): nothing overlaps this epilogue and every MMA of the
        // CTA has retired, so the patch buffers are free — the same per-warp code stages the
        // whole [pixel][128 couts] tile there (no per-chunk store, fence or slot wait), then all
        // epilogue warps meet once and issue the tile's TMA stores.
        const bool whole = walk_last(p, w);
        float4 pf[4];
        auto fetch_part0 = [&](int g) {
          const float4* src = reinterpret_cast<const float4*>(part0) + (g * 128 + cl) * 2;
          pf[0] = __ldcg(src); pf[1] = __ldcg(src + 1);
          if (g + 1 < g_end) { pf[2] = __ldcg(src + 256); pf[3] = __ldcg(src + 257); }
        };
        uint4 rv[2];
        auto fetch_res = [&](int g) {          // this warp's 16 pixels x 64 bytes of the residual tile
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k = lane + 32 * j;
            const __half* src = res_row_ptr(g, k >> 2);
            rv[j] = src ? __ldg(reinterpret_cast<const uint4*>(src + qd * 32) + (k & 3)) : make_uint4(0, 0, 0, 0);
          }
        };
        const uint32_t warp_stage = stage_out + static_cast<uint32_t>(ew) * (kStageSlots * 1024u);
        // (`parts_c`: compile-time "this tile has parked stream-K partials" — the loop below is
        // instantiated twice so that the common no-partials chunk code is short and contiguous)
        auto finish = [&](uint32_t (&v)[16], int g, auto parts_c) {
          if (decltype(parts_c)::value) {
            const float* f = reinterpret_cast<const float*>(pf);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + f[j]);
            if (g + 2 < g_end) fetch_part0(g + 2);
            for (int k = 1; k < n_parts; ++k) {           // split-K over more than two CTAs
              const float* pk = part0 + static_cast<size_t>(k) * 128 * p.NP;
              for (int gg = 0; gg < 2 && g + gg < g_end; ++gg) {
                const float4* src = reinterpret_cast<const float4*>(pk) + ((g + gg) * 128 + cl) * 2;
                const float4 f0 = __ldcg(src), f1 = __ldcg(src + 1);
                uint32_t* u = v + 8 * gg;
                u[0] = __float_as_uint(__uint_as_float(u[0]) + f0.x);
                u[1] = __float_as_uint(__uint_as_float(u[1]) + f0.y);
                u[2] = __float_as_uint(__uint_as_float(u[2]) + f0.z);
                u[3] = __float_as_uint(__uint_as_float(u[3]) + f0.w);
                u[4] = __float_as_uint(__uint_as_float(u[4]) + f1.x);
                u[5] = __float_as_uint(__uint_as_float(u[5]) + f1.y);
                u[6] = __float_as_uint(__uint_as_float(u[6]) + f1.z);
                u[7] = __float_as_uint(__uint_as_float(u[7]) + f1.w);
              }
            }
          }
          float y[16];
          if (p.shift9) {                   // border-class shifts (kernel parameter: warp-uniform branch)
            float f0, m0, l0, f1, m1, l1;
            group_shifts(g, f0, m0, l0);
            group_shifts(g + 1, f1, m1, l1);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              y[i] = fmaf(__uint_as_float(v[i]), sc,
                          i < 8 ? pixel_shift(i, f0, m0, l0) : pixel_shift(i - 8, f1, m1, l1));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaf(__uint_as_float(v[i]), sc, sh);
          }
          // (the epilogue is instruction-issue bound — two epilogue warps per scheduler — so the
          // common activations take their short forms)
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.f);
          } else if (p.act == ACT_PRELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaf(fminf(y[i], 0.f), neg, fmaxf(y[i], 0.f));
          }
          if (p.pool) {
            // Fused 2x2 / stride-2 max-pool (VGG: conv -> ReLU -> pool): the chunk holds two
            // adjacent lines of eight pixels of this lane's channel, so the four windows are
            // in-lane maxima; four pooled pixels are staged and stored instead of sixteen.
            float q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              q[j] = fmaxf(fmaxf(y[2 * j], y[2 * j + 1]), fmaxf(y[8 + 2 * j], y[9 + 2 * j]));
            const uint32_t pslot = whole ? base + static_cast<uint32_t>(g >> 1) * 1024u + qd * 64u
                                         : warp_stage + (chunk_ctr % kStageSlots) * 1024u;
            const uint32_t ppitch = whole ? 256u : 64u;
            if (!whole) {
              ++chunk_ctr;
              if (elect_one()) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStageSlots - 1) : "memory");
              __syncwarp();
            }
            const uint32_t sel = (lane & 1) ? 0x3276u : 0x5410u;
            const uint32_t dst0 = pslot + (lane & 1) * ppitch + (lane & ~1) * 2u;
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
              const __half2 h2 = __floats2half2_rn(q[i], q[i + 1]);
              const uint32_t own = *reinterpret_cast<const uint32_t*>(&h2);
              const uint32_t oth = __shfl_xor_sync(0xffffffffu, own, 1);
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst0 + static_cast<uint32_t>(i) * ppitch),
                           "r"(__byte_perm(own, oth, sel)) : "memory");
            }
            if (whole) return;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const int bp = (t.b0 + g) >> 1, ap = t.a0 >> 1;          // pooled coordinates
            const bool ok = t.b0 + g + 1 < B_dim;                     // warp-uniform (floor pooling)
            const int cw = p.axis == 0 ? ap : bp, ch = p.axis == 0 ? bp : ap;
            if (ok && elect_one())
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                  ::"l"(&tmO32), "r"(pslot), "r"(p.out_coff + t.ct * 128 + qd * 32), "r"(cw), "r"(ch), "r"(t.n)
                  : "memory");
            __syncwarp();
            if (elect_one()) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            __syncwarp();
            return;
          }
          // 16 pixel rows of `pitch` bytes: a private slot, or this warp's 64-byte column of the
          // whole-tile staging area
          const uint32_t slot = whole ? base + static_cast<uint32_t>(g) * 2048u + qd * 64u
                                      : warp_stage + (chunk_ctr % kStageSlots) * 1024u;
          const uint32_t pitch = whole ? 256u : 64u;
          // an odd number of groups per team: the second group of the last chunk is the OTHER
          // team's first one — its rows must not be written in the shared whole-tile area
          const int nrow = g + 1 < g_end ? 16 : 8;
          if (!whole) {
            ++chunk_ctr;
            // the store that used this slot kStageSlots chunks ago has finished READING it
            // (TMA / bulk-group instructions run on the uniform datapath: issued from a divergent
            // `lane == 0` region each costs a ~200-cycle waterfall — measured 445 cycles per chunk
            // for two stores and a commit; the warp stays converged and one elected lane issues)
            if (elect_one()) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStageSlots - 1) : "memory");
            __syncwarp();
          }
          if (p.res) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int k = lane + 32 * j;
              if ((k >> 2) < nrow)
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                             ::"r"(slot + static_cast<uint32_t>(k >> 2) * pitch + (k & 3) * 16u),
                               "r"(rv[j].x), "r"(rv[j].y), "r"(rv[j].z), "r"(rv[j].w) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i >= nrow) break;
              unsigned short rh;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(rh) : "r"(slot + lane * 2u + i * pitch));
              y[i] += __half2float(__ushort_as_half(rh));
            }
            __syncwarp();                                   // (the packed stores below write other lanes' cells)
            if (g + 2 < g_end) fetch_res(g + 2);            // in flight during the next chunk's TMEM read
          }
          // The MMAs read their operands from shared memory at ~96 of its 128 bytes per cycle
          // (N = 256), so every staging wavefront is taken from the tensor pipe: two lanes swap
          // halves so that each lane writes 4 bytes (two couts of one pixel) and a warp store
          // fills a whole 128-byte wavefront (pixels i and i + 1) instead of 64 bytes.
          {
            // even lane: pixel i, couts (lane, lane + 1) = (lo(own), lo(oth));
            // odd lane: pixel i + 1, couts (lane - 1, lane) = (hi(oth), hi(own))
            const uint32_t sel = (lane & 1) ? 0x3276u : 0x5410u;
            const uint32_t dst0 = slot + (lane & 1) * pitch + (lane & ~1) * 2u;
            const bool second = nrow == 16;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const __half2 h2 = __floats2half2_rn(y[i], y[i + 1]);     // lo = pixel i, hi = pixel i + 1
              const uint32_t own = *reinterpret_cast<const uint32_t*>(&h2);
              const uint32_t oth = __shfl_xor_sync(0xffffffffu, own, 1);
              const uint32_t val = __byte_perm(own, oth, sel);
              if (i < 8 || second)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst0 + static_cast<uint32_t>(i) * pitch), "r"(val)
                             : "memory");
            }
          }
          if (whole) return;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int b = line_b(t.b0, g + gg), img = t.n + line_img(g + gg);
            const bool ok = g + gg < g_end && b < B_dim && img < p.N && !(p.debug & 4);       // warp-uniform
            const int cw = p.axis == 0 ? t.a0 : b, ch = p.axis == 0 ? b : t.a0;
            if (ok && elect_one())
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                  ::"l"(&tmO32), "r"(slot + gg * 512u), "r"(p.out_coff + t.ct * 128 + qd * 32), "r"(cw),
                    "r"(ch), "r"(img) : "memory");
            __syncwarp();
          }
          if (elect_one()) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          __syncwarp();
        };
        if (p.res) fetch_res(g_begin);
        auto run = [&](auto parts_c) {
          if (decltype(parts_c)::value) fetch_part0(g_begin);
          uint32_t va[16], vb[16];
          int g = g_begin;
          __syncwarp();
          tmem_ld16_async(taddr + g * 8, va);
          while (g < g_end) {
            __syncwarp();                       // tcgen05.ld / wait::ld are warp-collective
            tmem_ld_wait();
            if (g + 2 < g_end) tmem_ld16_async(taddr + (g + 2) * 8, vb);
            finish(va, g, parts_c);
            g += 2;
            if (g >= g_end) break;
            __syncwarp();
            tmem_ld_wait();
            if (g + 2 < g_end) tmem_ld16_async(taddr + (g + 2) * 8, va);
            finish(vb, g, parts_c);
            g += 2;
          }
        };
        if (n_parts) run(std::true_type{});
        else run(std::false_type{});
        if (whole) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync 3, %0;" ::"r"(p.epi_warps * 32) : "memory");     // all epilogue warps
          // every epilogue warp stores the groups ew, ew + epi_warps, ...: converged loop, one
          // elected lane per instruction
          const int n_rows = p.pool ? p.R >> 1 : p.R;             // staged lines: 4 pooled / 8 pixels each
          for (int r = ew; r < n_rows; r += p.epi_warps) {
            const int b = p.pool ? (t.b0 >> 1) + r : line_b(t.b0, r), img = t.n + (p.pool ? 0 : line_img(r));
            const bool ok = (p.pool ? t.b0 + 2 * r + 1 < B_dim : b < B_dim && img < p.N) && !(p.debug & 4);   // warp-uniform
            const int a = p.pool ? t.a0 >> 1 : t.a0;
            const int cw = p.axis == 0 ? a : b, ch = p.axis == 0 ? b : a;
            if (ok && elect_one())
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                  ::"l"(&tmO), "r"(base + static_cast<uint32_t>(r) * (p.pool ? 1024u : 2048u)),
                    "r"(p.out_coff + t.ct * 128), "r"(cw), "r"(ch), "r"(img) : "memory");
            __syncwarp();
          }
          if (elect_one()) {
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          }
          __syncwarp();
        }
      } else {
        const bool c_ok = cout < p.cout_store;
        const unsigned pix00 = static_cast<unsigned>(t.n) * p.H * p.W + t.a0 * a_step + t.b0 * b_step;
        const int na = min(8, A_dim - t.a0);
        for (int g = g_begin; g < g_end; ++g) {
          uint32_t v[8];
          __syncwarp();
          tmem_ld8_async(taddr + g * 8, v);
          tmem_ld_wait();
          if (t.b0 + g >= B_dim || !c_ok || (p.debug & 4)) continue;
          if (n_parts) add_parts(v, g);
          const unsigned pix0 = pix00 + g * b_step;
          const unsigned o0 = pix0 * p.out_cs + p.out_coff + cout, os = a_step * p.out_cs;
          const unsigned r0 = pix0 * p.res_cs + p.res_coff + cout, rs = a_step * p.res_cs;
          float f0, m0, l0;
          group_shifts(g, f0, m0, l0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < na) {
              float y = fmaf(__uint_as_float(v[i]), sc, pixel_shift(i, f0, m0, l0));
              y = fmaf(fminf(y, 0.f), neg, fmaxf(y, 0.f));
              if (p.res) y += __half2float(p.res[r0 + i * rs]);
              if (p.out_f32) p.out_f32[o0 + i * os] = y;
              else p.out[o0 + i * os] = __float2half_rn(y);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(tempty_bar(acc));
        for (int k = 1; k <= n_parts; ++k) {              // last of the 16 warps re-arms the flag
          int* flag = p.sk_flags + blockIdx.x + k;
          if (atomicAdd(flag, 1) == 2 * p.epi_warps - 1) atomicExch(flag, 0);
        }
      }
      if (warp == 2) PT_STAMP(6);                           // epilogue of a tile done
    }
    if (p.tma_store && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 3) PT_STAMP(7);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(p.tmem_cols) : "memory");
  }
}

PFN_cuTensorMapEncodeTiled_v12000 patch_encode_fn() {
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

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

using PatchKernel = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, PatchParams);
#define TRB_PATCH_MODES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(7) X(15) X(16) X(17)
PatchKernel patch_kernel_for(int mode, bool dual, bool wide = false) {
  if (dual) return conv_patch_kernel<kPThreadsDual, 2, kModeAll>;
  if (wide) return mode == kModePool ? conv_patch_kernel<kPThreadsWide, 1, kModePool> : conv_patch_kernel<kPThreadsWide, 1, 0>;
  switch (mode) {
#define X(m) case m: return conv_patch_kernel<kPThreads, 1, m>;
    TRB_PATCH_MODES(X)
#undef X
  }
  return conv_patch_kernel<kPThreads, 1, kModeAll>;
}
template <typename F>
void for_each_patch_kernel(F f) {
#define X(m) f(reinterpret_cast<const void*>(conv_patch_kernel<kPThreads, 1, m>), false);
  TRB_PATCH_MODES(X)
#undef X
  f(reinterpret_cast<const void*>(conv_patch_kernel<kPThreadsDual, 2, kModeAll>), true);
  f(reinterpret_cast<const void*>(conv_patch_kernel<kPThreadsWide, 1, 0>), false);
  f(reinterpret_cast<const void*>(conv_patch_kernel<kPThreadsWide, 1, kModePool>), false);
}

}  // namespace

struct ConvPatchPlan {
  bool dual = false;
  bool wide = false;                // 12 epilogue warps (kPThreadsWide)
  int mode = 15;                    // epilogue instantiation (see MODE of conv_patch_kernel)
  unsigned long long* trace = nullptr;
  void* sk_own = nullptr;
  CUtensorMap tmX, tmW, tmO, tmO32;     // tmO: {128 ch x 8 px} boxes (whole-tile path), tmO32: {32 ch x 8 px} (warp-local path)
  PatchParams p;
  int grid;
  uint32_t smem;
  double flops;
};

// Stream-K scratch: 2 KB of flags, then one 128 x 256 fp32 accumulator slot per co-resident CTA
// (two per SM in the dual configuration).
size_t conv_patch_scratch_bytes() {
  int sms = 0;
  TR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device()));
  return 2048 + size_t(2 * sms) * 128 * 256 * 4;
}

bool conv_patch_eligible(const ConvArgs& a) {
  if (a.stride != 1 || a.kh != a.kw || !(a.kh & 1) || a.pad != a.kh / 2 || a.kh > 9) return false;
  // 64 filters run as half of a 128-row tile: the filter map's rows 64..127 are out of bounds
  // (TMA zero fill, no traffic) and the output map clips the upper 64 channels
  if (a.cin_pad % 64 || (a.cout_pad % 128 && a.cout_pad != 64)) return false;
  if (a.in.cs % 8 || a.in.coff % 8) return false;
  if (a.out2.ptr || a.res_up2) return false;
  if (a.groups > 1 && (a.cout_pad % (128 * a.groups) || a.res.ptr || a.shift9)) return false;
  if (a.shift9 && (a.kh != 3 || a.in.H < 2 || a.in.W < 2)) return false;
  if (a.H_out != a.in.H || a.W_out != a.in.W) return false;
  return true;
}

ConvPatchPlan* conv_patch_plan_create(const ConvArgs& a, int* err_flag) {
  TR_CHECK(conv_patch_eligible(a), "convolution not eligible for the patch kernel");
  auto* plan = new ConvPatchPlan();
  PatchParams& p = plan->p;
  p = PatchParams{};
  p.N = a.in.N; p.H = a.H_out; p.W = a.W_out;
  p.k = a.kh; p.pad = a.pad; p.taps = a.kh * a.kw;
  p.cin_pad = a.cin_pad; p.kchunks = a.cin_pad / 64; p.in_coff = a.in.coff;
  p.cout_tiles = ceil_div(a.cout_pad, 128);
  p.tiles_per_group = std::max(1, p.cout_tiles / std::max(1, a.groups));
  // patch lines hold exactly the 8 + 2 pad pixels a tap's 8-pixel group can touch (the group
  // stride of the B descriptor need not be a multiple of the 1024-byte swizzle repeat: the
  // swizzle XORs absolute address bits; TRB_PT_PA16=1 restores 16-pixel lines)
  p.PA = a.pad == 0 ? 8 : (env_int("TRB_PT_PA16", 0) ? 16 : 8 + 2 * a.pad);

  int sms = 0;
  const int dev = current_device();
  TR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  // Tile geometry: 8-pixel groups along one axis, R (even, <= 32) rows of groups along the
  // other; pick the (axis, R) that wastes the fewest MMA columns, preferring wide tiles
  // (N = 8 R >= 192 amortises the filter stream: 16 KB of L2 -> SM traffic per 4 N cycles).
  const int force_axis = env_int("TRB_PT_AXIS", -1), force_r = env_int("TRB_PT_R", a.patch_rows);
  double best = -1.0;
  const int halo = 2 * a.pad;
  for (int axis = 0; axis < 2; ++axis) {
    if (force_axis >= 0 && axis != force_axis) continue;
    const int A = axis == 0 ? p.W : p.H, B = axis == 0 ? p.H : p.W;
    for (int R = 2; R <= 32; R += 2) {
      if (force_r && R != force_r) continue;
      if (a.pool2 && R % 4) continue;          // a team's groups pair up into pooling windows
      const uint32_t patch = round_up(p.PA * (R + halo) * 128, 1024);
      if (2 * patch + 3 * kFilterBlock + kOutStage + 4096 > 227u * 1024) continue;
      const double util = double(A) / (8.0 * ceil_div(A, 8)) * double(B) / (double(R) * ceil_div(B, R));
      const double wide = std::min(1.0, (8.0 * R + 64.0) / 256.0);     // N = 192 and up count as full
      const double score = util * wide + 1e-4 * R;
      if (score > best) { best = score; p.axis = axis; p.R = R; }
    }
  }
  TR_CHECK(best > 0, "no patch tile fits in shared memory");
  // Stacked images (see PatchParams::stack): small maps whose whole R axis fits several times
  // into 32 lines.  Staged fp16 epilogue only (the unstaged path and the pooled one keep one
  // image per tile).
  p.stack = 1; p.bstride = 0;
  const bool staged_out = env_int("TRB_PT_TMA_STORE", 1) && !a.out_f32 && !env_int("TRB_PT_DEBUG", 0) &&
                          !env_int("TRB_PT_GENERIC", 0) && a.out.cs % 8 == 0 && a.out.coff % 8 == 0 &&
                          (a.cout_store % 128 == 0 || a.out.coff + a.cout_store == a.out.cs) &&
                          (!a.res.ptr || (a.res.cs % 8 == 0 && a.res.coff % 8 == 0));
  if (env_int("TRB_PT_STACK", 1) && !env_int("TRB_PT_DUAL", 0) && a.pad > 0 && !a.pool2 && !force_r && staged_out &&
      p.N >= 2) {
    for (int axis = 0; axis < 2; ++axis) {
      if (force_axis >= 0 && axis != force_axis) continue;
      const int A = axis == 0 ? p.W : p.H, B = axis == 0 ? p.H : p.W;
      const int bs = B + halo;
      const int S = std::min(p.N, (32 + halo) / bs);
      if (S < 2) continue;
      int R = S * bs - halo;
      R += R & 1;                                   // (one more garbage line keeps R even)
      const uint32_t patch = round_up(p.PA * (R + halo) * 128, 1024);
      if (R > 32 || 2 * patch + 3 * kFilterBlock + kOutStage + 4096 > 227u * 1024) continue;
      const double util = double(A) / (8.0 * ceil_div(A, 8)) * double(S * B) / double(R);
      const double wide = std::min(1.0, (8.0 * R + 64.0) / 256.0);
      const double score = util * wide + 1e-4 * R;
      if (score > best + 0.02) { best = score; p.axis = axis; p.R = R; p.stack = S; p.bstride = bs; }
    }
  }
  p.NP = 8 * p.R;
  const int A = p.axis == 0 ? p.W : p.H, B = p.axis == 0 ? p.H : p.W;
  p.tiles_a = ceil_div(A, 8);
  p.tiles_b = p.stack > 1 ? 1 : ceil_div(B, p.R);
  p.pix_tiles = p.tiles_a * p.tiles_b * ceil_div(p.N, p.stack);
  p.total_tiles = p.pix_tiles * p.cout_tiles;
  p.patch_tx = p.stack > 1 ? p.stack * p.PA * p.bstride * 128 : p.PA * (p.R + halo) * 128;
  p.patch_bytes = round_up(p.PA * (p.R + halo) * 128, 1024);
  const uint32_t misc = kOutStage + 512 + 8 * (2 * kPMaxStages + 6 + 2 * kPMaxPatchBufs) + 1024 /*alignment*/;
  // Two CTAs per SM when one patch buffer + a filter ring of >= 2 x 16 KB fit in half of the SM's
  // shared memory (TRB_PT_DUAL: 0 never, 1 auto, 2 whenever it fits).
  {
    const int want = env_int("TRB_PT_DUAL", 0);     // measured slower on every layer (profiles/r02_patch_dual.txt)
    const uint32_t half = 113u * 1024;
    const bool fits = p.patch_bytes + 2 * kFilterBlock + misc <= half;
    const bool worth = p.kchunks * p.taps >= 8;              // enough MMA work to hide the other CTA behind
    plan->dual = fits && (want == 2 || (want == 1 && worth));
  }
  p.nbuf = plan->dual ? 1 : 2;
  // A 1x1 layer consumes one (small) patch per ring iteration, so two buffers leave one TMA
  // load in flight and the K loop runs at one L2/HBM latency per chunk (the 25088-channel
  // ArcFace FC: 392 chunks).  Up to eight patches of <= 8 KB keep that latency covered.
  if (!plan->dual && p.taps == 1 && p.kchunks >= 8)
    p.nbuf = std::max(2, std::min<int>(kPMaxPatchBufs, (64u * 1024) / p.patch_bytes));
  p.nbuf = std::max(1, std::min(p.nbuf, env_int("TRB_PT_NBUF", kPMaxPatchBufs)));
  p.nacc = plan->dual ? 1 : 2;
  p.epi_warps = plan->dual ? 4 : 8;
  p.ring_off = p.nbuf * p.patch_bytes;
  const uint32_t budget = plan->dual ? 113u * 1024 : 227u * 1024;
  const uint32_t fixed = p.ring_off + misc;
  p.sub = std::max(1, std::min(env_int("TRB_PT_SUB", 3), p.taps));
  while (p.sub > 1 && fixed + 2u * p.sub * kFilterBlock > budget) --p.sub;
  p.iters = ceil_div(p.taps, p.sub);
  p.stage_bytes = p.sub * kFilterBlock;
  p.stages = std::min(kPMaxStages, int((budget - fixed) / p.stage_bytes));
  p.stages = std::min(p.stages, std::max(2, env_int("TRB_PT_STAGES", kPMaxStages)));
  TR_CHECK(p.stages >= 2, "filter ring does not fit");
  p.idesc = (1u << 4) | (uint32_t(p.NP >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  uint32_t cols = 32;
  while (cols < uint32_t(p.nacc * p.NP)) cols <<= 1;
  p.tmem_cols = cols;

  p.scale = a.scale; p.shift = a.shift; p.slope = a.slope; p.act = a.act;
  p.shift9 = a.shift9; p.cout_pad = a.cout_pad;
  p.out = a.out.ptr; p.out_cs = a.out.cs; p.out_coff = a.out.coff; p.cout_store = a.cout_store;
  p.res = a.res.ptr; p.res_cs = a.res.cs; p.res_coff = a.res.coff;
  p.out_f32 = a.out_f32;
  p.err = err_flag;
  p.pdl_late = env_int("TRB_TC_PDL_LATE", 1);
  p.debug = env_int("TRB_PT_DEBUG", 0);
  if (p.debug & 32) {
    TR_CUDA(cudaMalloc(&plan->trace, (1 + 16 * 64) * 8));
    TR_CUDA(cudaMemset(plan->trace, 0, (1 + 16 * 64) * 8));
    p.trace = plan->trace;
  }
  TR_CHECK(a.scale && a.shift, "epilogue scale/shift are required");
  TR_CHECK(a.act != ACT_PRELU || a.slope, "PReLU needs slopes");
  TR_CHECK(double(p.N) * p.H * p.W * std::max(a.out.cs, a.res.cs) < 4.0e9,
           "activation tensor too large for 32-bit element offsets");

  auto encode = patch_encode_fn();
  const cuuint64_t cs = a.in.cs, W = a.in.W, H = a.in.H, N = a.in.N;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4] = {64, cuuint32_t(p.PA), cuuint32_t(p.stack > 1 ? p.bstride : p.R + halo), 1}, estr[4] = {1, 1, 1, 1};
  gdim[0] = cs; gdim[3] = N;
  gstr[2] = H * W * cs * 2;
  if (p.axis == 0) { gdim[1] = W; gdim[2] = H; gstr[0] = cs * 2; gstr[1] = W * cs * 2; }
  else             { gdim[1] = H; gdim[2] = W; gstr[0] = W * cs * 2; gstr[1] = cs * 2; }
  CUresult r = encode(&plan->tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.in.ptr, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(patch) failed: " + std::to_string(int(r)));
  const cuuint64_t ktot = cuuint64_t(p.taps) * a.cin_pad;
  cuuint64_t wdim[2] = {ktot, cuuint64_t(a.cout_pad)};
  cuuint64_t wstr[1] = {ktot * 2};
  cuuint32_t wbox[2] = {64, 128};
  r = encode(&plan->tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(a.w), wdim, wstr, wbox,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(filters) failed: " + std::to_string(int(r)));

  // TMA-store epilogue: plain fp16 output of whole 128-channel tiles (no residual); the 4-D map
  // {C, W, H, N} clips the ragged border of the 8-pixel groups.
  const bool whole = a.cout_store % 128 == 0 || a.out.coff + a.cout_store == a.out.cs;   // the map clips
  p.pool = a.pool2 ? 1 : 0;
  p.tma_store = env_int("TRB_PT_TMA_STORE", 1) && !a.out_f32 && whole &&
                a.out.cs % 8 == 0 && a.out.coff % 8 == 0 &&
                (!a.res.ptr || (a.res.cs % 8 == 0 && a.res.coff % 8 == 0));
  plan->tmO = plan->tmW;
  plan->tmO32 = plan->tmW;
  if (p.tma_store) {
    const cuuint64_t ocs = a.out.cs;
    // (fused pooling: the maps describe the pooled tensor, four pooled pixels per staged line)
    const cuuint64_t oW = p.pool ? cuuint64_t(a.out.W) : W, oH = p.pool ? cuuint64_t(a.out.H) : H;
    const cuuint32_t line = p.pool ? 4 : 8;
    cuuint64_t odim[4] = {ocs, oW, oH, N};
    cuuint64_t ostr[3] = {ocs * 2, oW * ocs * 2, oH * oW * ocs * 2};
    cuuint32_t obox[4] = {128, cuuint32_t(p.axis == 0 ? line : 1), cuuint32_t(p.axis == 0 ? 1 : line), 1};
    r = encode(&plan->tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.out.ptr, odim, ostr, obox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(output) failed: " + std::to_string(int(r)));
    obox[0] = 32;
    r = encode(&plan->tmO32, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, a.out.ptr, odim, ostr, obox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(output, 32 channels) failed: " + std::to_string(int(r)));
  }

  const int slots = sms * (plan->dual ? 2 : 1);               // co-resident CTAs
  plan->grid = std::min(p.total_tiles, slots);
  {
    // Stream-K where whole-tile scheduling leaves SMs idle: a last partial round (160 tiles on
    // 148 SMs run as long as 296) or fewer tiles than CTA slots.
    const int want = env_int("TRB_PT_SK", 1);               // 0 off, 1 auto, 2 whenever possible
    p.ipt = p.kchunks * p.iters;
    const int rounds = ceil_div(p.total_tiles, slots);
    const double idle = 1.0 - double(p.total_tiles) / (double(rounds) * slots);
    const bool ok = p.ipt >= 2 && slots < int(2048 / 4) - 1 &&
                    double(p.total_tiles) * p.ipt * (slots + 1) < 4.0e9;      // 32-bit walk arithmetic
    const bool worth = p.ipt >= 4 && idle >= 0.08;
    if (ok && (want == 2 || (want == 1 && worth))) {
      void* scratch = a.sk_scratch;
      if (!scratch) {
        TR_CUDA(cudaMalloc(&plan->sk_own, conv_patch_scratch_bytes()));
        TR_CUDA(cudaMemset(plan->sk_own, 0, 2048));
        scratch = plan->sk_own;
      }
      p.sk = 1;
      p.sk_flags = static_cast<int*>(scratch);
      p.sk_ws = reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + 2048);
      plan->grid = int(std::min<long long>(slots, static_cast<long long>(p.total_tiles) * p.ipt));
    }
  }
  plan->smem = fixed + p.stages * p.stage_bytes;
  plan->flops = 2.0 * p.N * p.H * p.W * double(a.cout_pad) * p.taps * a.cin_pad;   // (cin_pad is per group)
  // the smallest instantiation that covers the layer (see MODE above)
  {
    const int need = (p.sk ? kModeSK : 0) | (p.shift9 ? kModeS9 : 0) | (p.res ? kModeRes : 0);
    const bool generic = plan->dual || !p.tma_store || p.debug || p.trace || env_int("TRB_PT_GENERIC", 0);
    plan->mode = generic ? kModeAll : (need == 6 ? 7 : need);
    TR_CHECK(p.stack == 1 || !generic, "stacked tiles need the staged epilogue");
    // Twelve epilogue warps (three teams) for plain short-K layers: their epilogue (~3 us per
    // 192..256-pixel tile with eight warps) is longer than their MMAs.
    const int wide_k = env_int("TRB_PT_WIDE_K", 18);
    if (!generic && env_int("TRB_PT_WIDE", 1) && need == 0 && p.R >= 6 && p.kchunks * p.taps <= wide_k) {
      plan->wide = true;
      p.epi_warps = 12;
    }
    if (p.pool) {
      TR_CHECK(!generic && !(need & (kModeS9 | kModeRes)) && a.act == ACT_RELU,
               "fused max-pool needs the staged plain ReLU epilogue");
      TR_CHECK(a.out.H == a.H_out / 2 && a.out.W == a.W_out / 2, "fused max-pool: pooled output dims");
      plan->mode = kModePool | (need & kModeSK);
    }
  }
  static bool attr_set[kMaxDevices] = {};
  if (!attr_set[dev]) {
    for_each_patch_kernel([](const void* fn, bool dual) {
      TR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, dual ? 113 * 1024 : 227 * 1024));
      if (dual) TR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    });
    attr_set[dev] = true;
  }
  return plan;
}

void conv_patch_plan_destroy(ConvPatchPlan* p) {
  if (p->sk_own) cudaFree(p->sk_own);
  if (p->trace) {
    // CTA 0's event times relative to its kernel entry, and the gap to the previous launch's end
    std::vector<unsigned long long> h(1 + 16 * 64);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), p->trace, h.size() * 8, cudaMemcpyDeviceToHost);
    const int n = int(std::min<unsigned long long>(h[0], 64));
    printf("patch trace (CTA 0, us from kernel entry): gap_prev prologue dep_wait patch0 mma_issued acc_done epi_done end\n");
    for (int i = 0; i < n; ++i) {
      const unsigned long long* e = &h[1 + 16 * i];
      const double gap = i ? (double(e[0]) - double(h[1 + 16 * (i - 1) + 7])) / 1e3 : 0.0;
      printf("  launch %2d: %7.2f", i, gap);
      for (int k = 1; k <= 7; ++k) printf(" %7.2f", e[k] ? (double(e[k]) - double(e[0])) / 1e3 : -1.0);
      printf("\n");
    }
    fflush(stdout);
    cudaFree(p->trace);
  }
  delete p;
}

void conv_patch_plan_describe(const ConvPatchPlan* plan, int* axis, int* R, int* tiles, int* stages) {
  *axis = plan->p.axis; *R = plan->p.R; *tiles = plan->p.total_tiles; *stages = plan->p.stages + (plan->dual ? 100 : 0);
}

void conv_patch_launch(const ConvPatchPlan* plan, cudaStream_t s) {
  static const bool pdl = env_int("TRB_TC_PDL", 1) != 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(plan->dual ? kPThreadsDual : plan->wide ? kPThreadsWide : kPThreads);
  cfg.dynamicSmemBytes = plan->smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  TR_CUDA(cudaLaunchKernelEx(&cfg, patch_kernel_for(plan->mode, plan->dual, plan->wide), plan->tmX, plan->tmW, plan->tmO,
                             plan->tmO32, plan->p));
}

}  // namespace trb
