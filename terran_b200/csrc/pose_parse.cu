// OpenPose heat-map / PAF parse on the GPU.  The x8 bicubic up-sampling of the
// reference (215 MB materialised per 16 frames) is never written to HBM: peaks
// are found on band tiles staged in shared memory and PAF samples are evaluated
// on the fly from the 1/8-resolution maps (0.21 MB/frame read).
//
// Replaces terran/pose/openpose/wrapper.py:214-483 (bicubic interpolate, peak
// extraction, limb line integrals, greedy matching, human assembly,
// get_keypoints).  Compiled with -fmad=false; the fp32 operation order is the
// one fixed by oracle/pose.py, which this file reproduces bit for bit.
#include "detect_post.cuh"

#include <algorithm>
#include <cmath>

namespace trb {

namespace {

__constant__ float c_bicubic[8][4];
__constant__ int c_map_idx[19][2] = {
    {31, 32}, {39, 40}, {33, 34}, {35, 36}, {41, 42}, {43, 44}, {19, 20}, {21, 22}, {23, 24},
    {25, 26}, {27, 28}, {29, 30}, {47, 48}, {49, 50}, {53, 54}, {51, 52}, {55, 56}, {37, 38},
    {45, 46}};
__constant__ int c_limbseq[19][2] = {
    {2, 3}, {2, 6}, {3, 4}, {4, 5}, {6, 7}, {7, 8}, {2, 9}, {9, 10}, {10, 11}, {2, 12},
    {12, 13}, {13, 14}, {2, 1}, {1, 15}, {15, 17}, {1, 16}, {16, 18}, {3, 17}, {6, 18}};

struct PoseParams {
  const float* paf;    // (N,38,h,w)
  const float* heat;   // (N,19,h,w)
  int N, h, w, Hu, Wu;
  double scale;
  float bicubic_gain;              // (max over phases of sum |w|)^2 * (1 + eps)
  unsigned long long* peak_keys;   // [N][18][kPeakCap]  (pos << 32 | score bits)
  int* peak_cnt;                   // [N][18]
  int* conn_src;                   // [N][19][kPeakCap]
  int* conn_dst;
  float* conn_score;
  int* conn_cnt;                   // [N][19]  (-1: limb missing)
  PoseOut out;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// floor((o + 0.5)/8 - 0.5) for o >= 0
__device__ __forceinline__ int src_floor(int o) { return (o + 4) / 8 - 1; }

// Bicubic x8 value of one channel map (h x w) at up-sampled pixel (y, x).
// (`tab` = the 8 x 4 tap table in SHARED memory: the phase differs from thread to thread, and a
// divergent index into __constant__ memory is serialised by the constant cache)
__device__ __forceinline__ float bicubic_at(const float* __restrict__ m, const float (*tab)[4], int h, int w,
                                            int y, int x) {
  const int fy = src_floor(y), fx = src_floor(x);
  const float* wy = tab[y & 7];
  const float* wx = tab[x & 7];
  int ix[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) ix[j] = clampi(fx - 1 + j, 0, w - 1);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* row = m + clampi(fy - 1 + i, 0, h - 1) * w;
    float r = row[ix[0]] * wx[0];
    r = r + row[ix[1]] * wx[1];
    r = r + row[ix[2]] * wx[2];
    r = r + row[ix[3]] * wx[3];
    acc = (i == 0) ? r * wy[0] : acc + r * wy[i];
  }
  return acc;
}

// ---- 1. peaks: block = (band of 8 up-sampled rows, part, frame)
// A thread owns COLUMNS of the band (x = tid, tid + 256, ...): the ten up-sampled rows
// 8*band-1 .. 8*band+8 of a column are combinations of the same five horizontally interpolated
// source rows, with weights and row slots that are compile-time constants per row (row k has
// phase (k - 1) & 7 and starts at slot k >= 5), so there is no integer division, no clamp and
// no dynamically indexed constant-memory read in the inner loops.
__global__ void __launch_bounds__(256) pose_peaks_kernel(const PoseParams p) {
  extern __shared__ float sm[];
  const int band = blockIdx.x, part = blockIdx.y, n = blockIdx.z;
  const int Wu = p.Wu, Hu = p.Hu, w = p.w;
  float* rows_h = sm;                 // [5][Wu] horizontally interpolated source rows band-2..band+2
  float* up = sm + 5 * Wu;            // [10][Wu] up-sampled rows 8*band-1 .. 8*band+8
  float* src = sm + 15 * Wu;          // [5][w] source rows band-2..band+2 (clamped)
  __shared__ __align__(16) float wtab[8][4];
  const float* m = p.heat + (static_cast<long>(n) * 19 + part) * p.h * w;
  if (threadIdx.x < 32) wtab[threadIdx.x >> 2][threadIdx.x & 3] = c_bicubic[threadIdx.x >> 2][threadIdx.x & 3];
  // Early out: every up-sampled value of this band is a bicubic combination of the
  // source rows band-2..band+2, so |value| <= (max sum|w|)^2 * max|source|.  If that bound
  // is below the 0.1 peak threshold the band cannot contain a peak (exact, not a heuristic).
  {
    float mx = 0.f;
    for (int k = 0; k < 5; ++k) {
      const float* row = m + clampi(band - 2 + k, 0, p.h - 1) * w;
      for (int x = threadIdx.x; x < w; x += 256) {
        const float v = row[x];
        src[k * w + x] = v;
        mx = fmaxf(mx, fabsf(v));
      }
    }
    if (__syncthreads_count(mx * p.bicubic_gain >= 0.1f) == 0) return;
  }
  for (int x = threadIdx.x; x < Wu; x += 256) {
    const int fx = src_floor(x);
    const float4 wx = *reinterpret_cast<const float4*>(wtab[x & 7]);
    const int i0 = clampi(fx - 1, 0, w - 1), i1 = clampi(fx, 0, w - 1), i2 = clampi(fx + 1, 0, w - 1),
              i3 = clampi(fx + 2, 0, w - 1);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float* row = src + k * w;
      float r = row[i0] * wx.x;
      r = r + row[i1] * wx.y;
      r = r + row[i2] * wx.z;
      r = r + row[i3] * wx.w;
      rows_h[k * Wu + x] = r;
    }
  }
  __syncthreads();
  const int y_first = 8 * band - 1;
  unsigned long long* keys = p.peak_keys + (static_cast<long>(n) * 18 + part) * kPeakCap;
  int* cnt = p.peak_cnt + n * 18 + part;
  // pass 1: the ten rows of each owned column -> shared memory (neighbouring columns need them)
  for (int x = threadIdx.x; x < Wu; x += 256) {
    float r[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) r[k] = rows_h[k * Wu + x];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      // row y = 8 band - 1 + k: phase (k + 7) & 7, source rows start at slot (k >= 5) (the slots
      // hold clamped rows, so the border clamp of the reference is already applied)
      const int ph = (k + 7) & 7, s0 = k >= 5 ? 1 : 0;
      float acc = r[s0] * c_bicubic[ph][0];
      acc = acc + r[s0 + 1] * c_bicubic[ph][1];
      acc = acc + r[s0 + 2] * c_bicubic[ph][2];
      acc = acc + r[s0 + 3] * c_bicubic[ph][3];
      up[k * Wu + x] = acc;
    }
  }
  __syncthreads();
  // pass 2: 4-neighbour maxima of rows 1..8
  for (int x = threadIdx.x; x < Wu; x += 256) {
    if (x < 1 || x > Wu - 2) continue;
    float c[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) c[k] = up[k * Wu + x];
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
      const int y = y_first + k;
      const float v = c[k];
      if (y < 1 || y > Hu - 2 || !(v >= 0.1f) || !(v >= c[k - 1]) || !(v >= c[k + 1])) continue;
      if (v >= up[k * Wu + x - 1] && v >= up[k * Wu + x + 1]) {
        const int slot = atomicAdd(cnt, 1);
        if (slot < kPeakCap)
          keys[slot] = (static_cast<unsigned long long>(y * Wu + x) << 32) | __float_as_uint(v);
        else
          atomicOr(p.out.status + n, 1);
      }
    }
  }
}

__device__ void bitonic_sort_sm(unsigned long long* k, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool upw = (lo & size) == 0;
        const unsigned long long a = k[lo], b = k[hi];
        if ((a > b) == upw) { k[lo] = b; k[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// ---- 2. order each (frame, part) peak list row-major (torch.nonzero order)
__global__ void __launch_bounds__(256) pose_sort_peaks_kernel(const PoseParams p) {
  __shared__ unsigned long long sk[kPeakCap];
  const int part = blockIdx.x, n = blockIdx.y;
  unsigned long long* keys = p.peak_keys + (static_cast<long>(n) * 18 + part) * kPeakCap;
  int* cnt = p.peak_cnt + n * 18 + part;
  const int K = min(*cnt, kPeakCap);
  int P = 1;
  while (P < K) P <<= 1;
  for (int i = threadIdx.x; i < P; i += 256) sk[i] = i < K ? keys[i] : ~0ull;
  __syncthreads();
  if (K > 1) bitonic_sort_sm(sk, P);
  for (int i = threadIdx.x; i < K; i += 256) keys[i] = sk[i];
  __syncthreads();
  if (threadIdx.x == 0) *cnt = K;
}

// torch.linspace(a, b, 10)[k] in fp32, truncated (oracle/pose.py segment_points)
__device__ __forceinline__ int seg_point(float a, float b, float step, int k) {
  const float v = k < 5 ? a + step * static_cast<float>(k) : b - step * static_cast<float>(9 - k);
  return static_cast<int>(v);
}

// ---- 3. limbs: block = (limb, frame): score all pairs, sort, greedy match
__global__ void __launch_bounds__(256) pose_limbs_kernel(const PoseParams p) {
  __shared__ unsigned long long cand[kCandCap];
  __shared__ int cand_n;
  __shared__ unsigned seen[kPeakCap / 32];
  __shared__ float wtab[8][4];
  if (threadIdx.x < 32) wtab[threadIdx.x >> 2][threadIdx.x & 3] = c_bicubic[threadIdx.x >> 2][threadIdx.x & 3];
  const int limb = blockIdx.x, n = blockIdx.y;
  const int ks = c_limbseq[limb][0] - 1, kd = c_limbseq[limb][1] - 1;
  const int* cnts = p.peak_cnt + n * 18;
  const int ns = cnts[ks], nd = cnts[kd];
  int* out_cnt = p.conn_cnt + n * 19 + limb;
  if (ns == 0 || nd == 0) {
    if (threadIdx.x == 0) *out_cnt = -1;      // missing limb
    return;
  }
  int id_s = 0, id_d = 0;
  for (int q = 0; q < ks; ++q) id_s += cnts[q];
  for (int q = 0; q < kd; ++q) id_d += cnts[q];
  const unsigned long long* src = p.peak_keys + (static_cast<long>(n) * 18 + ks) * kPeakCap;
  const unsigned long long* dst = p.peak_keys + (static_cast<long>(n) * 18 + kd) * kPeakCap;
  const float* mx = p.paf + (static_cast<long>(n) * 38 + (c_map_idx[limb][0] - 19)) * p.h * p.w;
  const float* my = p.paf + (static_cast<long>(n) * 38 + (c_map_idx[limb][1] - 19)) * p.h * p.w;
  if (threadIdx.x == 0) cand_n = 0;
  for (int i = threadIdx.x; i < kPeakCap / 32; i += 256) seen[i] = 0u;
  __syncthreads();

  const float half_h = 0.5f * static_cast<float>(p.Hu);
  for (int t = threadIdx.x; t < ns * nd; t += 256) {
    const int i = t / nd, j = t % nd;
    const int ps = static_cast<int>(src[i] >> 32), pd = static_cast<int>(dst[j] >> 32);
    const int sy = ps / p.Wu, sx = ps % p.Wu, dy = pd / p.Wu, dx = pd % p.Wu;
    const float vy = static_cast<float>(dy - sy), vx = static_cast<float>(dx - sx);
    const float nrm = sqrtf(vy * vy + vx * vx);
    const float uy = vy / nrm, ux = vx / nrm;
    const float fsy = static_cast<float>(sy), fdy = static_cast<float>(dy);
    const float fsx = static_cast<float>(sx), fdx = static_cast<float>(dx);
    const float step_y = (fdy - fsy) / 9.f, step_x = (fdx - fsx) / 9.f;
    float total = 0.f;
    int above = 0;
    for (int k = 0; k < 10; ++k) {
      const int py = seg_point(fsy, fdy, step_y, k);
      const int px = seg_point(fsx, fdx, step_x, k);
      const float a = bicubic_at(mx, wtab, p.h, p.w, py, px) * ux;
      const float b = bicubic_at(my, wtab, p.h, p.w, py, px) * uy;
      const float mscore = a + b;
      if (mscore > 0.05f) ++above;
      total = total + mscore;
    }
    const float pen = fminf(half_h / nrm - 1.f, 0.f);
    const float reg = total / 10.f + pen;
    if (above > 8 && reg > 0.f) {
      const int slot = atomicAdd(&cand_n, 1);
      if (slot < kCandCap)
        cand[slot] = (static_cast<unsigned long long>(~__float_as_uint(reg)) << 32) |
                     static_cast<unsigned>(t);
      else
        atomicOr(p.out.status + n, 2);
    }
  }
  __syncthreads();
  const int K = min(cand_n, kCandCap);
  int P = 1;
  while (P < K) P <<= 1;
  for (int i = K + threadIdx.x; i < P; i += 256) cand[i] = ~0ull;
  __syncthreads();
  if (K > 1) bitonic_sort_sm(cand, P);

  if (threadIdx.x == 0) {
    // greedy: ONE `seen` set shared by source and destination indices, and the
    // early break happens before the last accepted pair is inserted
    // (wrapper.py:336-359).
    int* cs = p.conn_src + (static_cast<long>(n) * 19 + limb) * kPeakCap;
    int* cd = p.conn_dst + (static_cast<long>(n) * 19 + limb) * kPeakCap;
    float* cf = p.conn_score + (static_cast<long>(n) * 19 + limb) * kPeakCap;
    const int limit = min(ns, nd);
    int made = 0;
    for (int c = 0; c < K; ++c) {
      const unsigned long long key = cand[c];
      const int t = static_cast<int>(key & 0xffffffffull);
      const int i = t / nd, j = t % nd;
      if (!((seen[i >> 5] >> (i & 31)) & 1u) && !((seen[j >> 5] >> (j & 31)) & 1u)) {
        cs[made] = id_s + i;
        cd[made] = id_d + j;
        cf[made] = __uint_as_float(~static_cast<unsigned>(key >> 32));
        ++made;
        if (made >= limit) break;
        seen[i >> 5] |= 1u << (i & 31);
        seen[j >> 5] |= 1u << (j & 31);
      }
    }
    *out_cnt = made;
  }
}

// ---- 4. assembly: one warp per frame, sequential over limbs / connections in
// float64 exactly like the reference's numpy code (wrapper.py:380-478).
__global__ void __launch_bounds__(32) pose_assemble_kernel(const PoseParams p) {
  __shared__ double hum[kHumanCap][20];
  __shared__ int offs[19];
  const int n = blockIdx.x, lane = threadIdx.x;
  const int* cnts = p.peak_cnt + n * 18;
  if (lane == 0) {
    int o = 0;
    for (int q = 0; q < 18; ++q) { offs[q] = o; o += cnts[q]; }
    offs[18] = o;
  }
  __syncwarp();
  auto peak_key = [&](int id) -> unsigned long long {
    int q = 0;
    while (q < 17 && id >= offs[q + 1]) ++q;
    return p.peak_keys[(static_cast<long>(n) * 18 + q) * kPeakCap + (id - offs[q])];
  };
  auto peak_score = [&](int id) -> double {
    return static_cast<double>(__uint_as_float(static_cast<unsigned>(peak_key(id) & 0xffffffffull)));
  };
  int nh = 0;          // warp-uniform
  bool overflow = false;
  for (int limb = 0; limb < 19; ++limb) {
    const int nc = p.conn_cnt[n * 19 + limb];
    if (nc < 0) continue;
    const int ks = c_limbseq[limb][0] - 1, kd = c_limbseq[limb][1] - 1;
    const int* cs = p.conn_src + (static_cast<long>(n) * 19 + limb) * kPeakCap;
    const int* cd = p.conn_dst + (static_cast<long>(n) * 19 + limb) * kPeakCap;
    const float* cf = p.conn_score + (static_cast<long>(n) * 19 + limb) * kPeakCap;
    for (int c = 0; c < nc; ++c) {
      const double ps = cs[c], pd = cd[c], sc = static_cast<double>(cf[c]);
      // first two matching humans, in row order
      int m1 = -1, m2 = -1, nm = 0;
      for (int b = 0; b < nh; b += 32) {
        const int hI = b + lane;
        const bool hit = hI < nh && (hum[hI][ks] == ps || hum[hI][kd] == pd);
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        nm += __popc(bal);
        while (bal && m2 < 0) {
          const int bit = __ffs(bal) - 1;
          bal &= bal - 1;
          if (m1 < 0) m1 = b + bit; else m2 = b + bit;
        }
      }
      __syncwarp();
      if (nm == 1) {
        if (lane == 0 && hum[m1][kd] != pd) {
          hum[m1][kd] = pd;
          hum[m1][19] += 1.0;
          hum[m1][18] += peak_score(static_cast<int>(pd)) + sc;
        }
      } else if (nm == 2) {
        bool overlap = false;
        if (lane < 18) overlap = hum[m1][lane] >= 0.0 && hum[m2][lane] >= 0.0;
        const bool any = __ballot_sync(0xffffffffu, overlap) != 0u;
        if (!any) {
          if (lane < 18) hum[m1][lane] += hum[m2][lane] + 1.0;
          if (lane == 18) hum[m1][18] = (hum[m1][18] + hum[m2][18]) + sc;
          if (lane == 19) hum[m1][19] += hum[m2][19];
          __syncwarp();
          for (int r = m2; r < nh - 1; ++r) {      // np.delete(humans, m2)
            if (lane < 20) hum[r][lane] = hum[r + 1][lane];
            __syncwarp();
          }
          --nh;
        } else if (lane == 0) {
          hum[m1][kd] = pd;
          hum[m1][19] += 1.0;
          hum[m1][18] += peak_score(static_cast<int>(pd)) + sc;
        }
      } else if (nm == 0 && limb < 17) {
        if (nh < kHumanCap) {
          if (lane < 18) hum[nh][lane] = -1.0;
          __syncwarp();
          if (lane == 0) {
            hum[nh][ks] = ps;
            hum[nh][kd] = pd;
            hum[nh][19] = 2.0;
            hum[nh][18] = ((0.0 + peak_score(static_cast<int>(ps))) + peak_score(static_cast<int>(pd))) + sc;
          }
          ++nh;
        } else {
          overflow = true;
        }
      }
      __syncwarp();
    }
  }
  // filter + keypoints (get_keypoints, wrapper.py:37-90)
  int outn = 0;
  for (int hI = 0; hI < nh; ++hI) {
    const double cntv = hum[hI][19], sumv = hum[hI][18];
    if (cntv < 4.0 || sumv / cntv < 0.4) continue;
    int* kp = p.out.keypoints + (static_cast<long>(n) * kHumanCap + outn) * 54;
    if (lane < 18) {
      const int pid = static_cast<int>(hum[hI][lane]);
      if (pid != -1) {
        const int pos = static_cast<int>(peak_key(pid) >> 32);
        const double y = static_cast<double>(pos / p.Wu), x = static_cast<double>(pos % p.Wu);
        kp[lane * 3 + 0] = static_cast<int>(x / p.scale);
        kp[lane * 3 + 1] = static_cast<int>(y / p.scale);
        kp[lane * 3 + 2] = 1;
      } else {
        kp[lane * 3 + 0] = 0; kp[lane * 3 + 1] = 0; kp[lane * 3 + 2] = 0;
      }
    }
    if (lane == 0) p.out.score[static_cast<long>(n) * kHumanCap + outn] = sumv / cntv;
    ++outn;
  }
  if (lane == 0) {
    p.out.count[n] = outn;
    if (overflow) atomicOr(p.out.status + n, 4);
  }
}

}  // namespace

void bicubic_table_host(float out[32]) {
  // Same fp32 operation order as oracle/pose.py::bicubic_table; volatile keeps
  // the host compiler from contracting or widening the intermediates.
  const volatile float A = -0.75f;
  auto c1 = [&](float v) {
    volatile float t = (A + 2.f);
    t = t * v;
    volatile float u = (A + 3.f);
    t = t - u;
    t = t * v;
    t = t * v;
    t = t + 1.f;
    return static_cast<float>(t);
  };
  auto c2 = [&](float v) {
    volatile float t = A * v;
    volatile float a5 = 5.f * A;
    t = t - a5;
    t = t * v;
    volatile float a8 = 8.f * A;
    t = t + a8;
    t = t * v;
    volatile float a4 = 4.f * A;
    t = t - a4;
    return static_cast<float>(t);
  };
  for (int ph = 0; ph < 8; ++ph) {
    const float src = static_cast<float>((ph + 0.5) / 8.0 - 0.5);
    const float t = src - floorf(src);
    out[ph * 4 + 0] = c2(t + 1.f);
    out[ph * 4 + 1] = c1(t);
    out[ph * 4 + 2] = c1(1.f - t);
    out[ph * 4 + 3] = c2(2.f - t);
  }
}

size_t pose_workspace_bytes(int N) {
  size_t b = 0;
  b += size_t(N) * 18 * kPeakCap * 8;   // peak keys
  b += size_t(N) * 19 * kPeakCap * 12;  // connections
  b += size_t(N) * (18 + 19) * 4;       // counts
  return b + 1024;
}

void pose_parse_launch(const float* paf, const float* heat, int N, int h, int w, double scale,
                       void* workspace, const PoseOut& out, cudaStream_t s) {
  if (N == 0) return;
  // __constant__ memory and function attributes live per device / context
  const int dev = current_device();
  static bool table_set[kMaxDevices] = {};
  static float gain = 0.f;
  if (!table_set[dev]) {
    float tab[32];
    bicubic_table_host(tab);
    TR_CUDA(cudaMemcpyToSymbol(c_bicubic, tab, sizeof(tab)));
    float gmax = 0.f;
    for (int ph = 0; ph < 8; ++ph) {
      float a = 0.f;
      for (int k = 0; k < 4; ++k) a += fabsf(tab[ph * 4 + k]);
      gmax = std::max(gmax, a);
    }
    gain = gmax * gmax * 1.001f;
    table_set[dev] = true;
  }
  PoseParams p{};
  p.bicubic_gain = gain;
  p.paf = paf; p.heat = heat; p.N = N; p.h = h; p.w = w; p.Hu = 8 * h; p.Wu = 8 * w;
  p.scale = scale;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  p.peak_keys = reinterpret_cast<unsigned long long*>(ws); ws += size_t(N) * 18 * kPeakCap * 8;
  p.conn_src = reinterpret_cast<int*>(ws); ws += size_t(N) * 19 * kPeakCap * 4;
  p.conn_dst = reinterpret_cast<int*>(ws); ws += size_t(N) * 19 * kPeakCap * 4;
  p.conn_score = reinterpret_cast<float*>(ws); ws += size_t(N) * 19 * kPeakCap * 4;
  p.peak_cnt = reinterpret_cast<int*>(ws); ws += size_t(N) * 18 * 4;
  p.conn_cnt = reinterpret_cast<int*>(ws);
  p.out = out;
  TR_CHECK(long(p.Hu) * p.Wu < (1L << 31), "up-sampled map too large");
  TR_CUDA(cudaMemsetAsync(p.peak_cnt, 0, size_t(N) * 18 * 4, s));
  TR_CUDA(cudaMemsetAsync(out.status, 0, size_t(N) * 4, s));
  const size_t smem = (size_t(15) * p.Wu + size_t(5) * w) * sizeof(float);
  static size_t smem_set[kMaxDevices] = {};
  if (smem > 48 * 1024 && smem > smem_set[dev]) {
    TR_CUDA(cudaFuncSetAttribute(pose_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(smem)));
    smem_set[dev] = smem;
  }
  pose_peaks_kernel<<<dim3(h, 18, N), 256, smem, s>>>(p);
  TR_CUDA(cudaGetLastError());
  pose_sort_peaks_kernel<<<dim3(18, N), 256, 0, s>>>(p);
  TR_CUDA(cudaGetLastError());
  pose_limbs_kernel<<<dim3(19, N), 256, 0, s>>>(p);
  TR_CUDA(cudaGetLastError());
  pose_assemble_kernel<<<N, 32, 0, s>>>(p);
  TR_CUDA(cudaGetLastError());
}

}  // namespace trb
