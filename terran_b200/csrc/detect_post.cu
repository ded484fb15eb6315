// RetinaFace post-processing on the GPU, batched over images:
//   score threshold -> candidate keys -> descending sort (stable by anchor
//   index) -> anchor/bbox/landmark decode -> greedy NMS -> survivors.
// Replaces the per-image Python loop of the reference
// (terran/face/detection/retinaface/wrapper.py:154-236, anchors.py:7-51) and
// torchvision.ops.nms.  Compiled with -fmad=false: every fp32 product and sum is
// rounded separately, exactly like the numpy oracle (oracle/detect.py), and the
// IoU is compared against the threshold as a double like torchvision's CPU nms.
#include "detect_post.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace trb {

namespace {

constexpr int kSelThreads = 512;
constexpr int kSmemKeys = 4096;

struct Level { int stride, fh, fw, off; };

struct DetParams {
  DetHeads heads;
  int N, H, W;
  Level lv[3];
  int A;           // anchors per image
  int capP;        // power-of-two key capacity per image (>= A)
  float thr;
  double nms_thr;
  int max_det;
  unsigned long long* keys;   // [N][capP]
  int* counts;                // [N] candidate counts
  float* sbox;                // [N][A][4] decoded boxes in sorted order
  int* kept;                  // [N][A] kept positions
  // outputs
  int* out_count;             // [N] survivors (may exceed max_det: truncated rows)
  int* out_cand;              // [N] candidates
  float* out_det;             // [N][max_det][16]
};

__device__ __forceinline__ void locate(const DetParams& p, int idx, int& l, int& pos, int& a) {
  l = idx >= p.lv[2].off ? 2 : (idx >= p.lv[1].off ? 1 : 0);
  const int local = idx - p.lv[l].off;
  a = local & 1;
  pos = local >> 1;
}

__device__ __forceinline__ float head_score(const DetParams& p, int n, int l, int pos, int a) {
  const int hw = p.lv[l].fh * p.lv[l].fw;
  if (p.heads.fused) {
    const float* px = p.heads.cls[l] + (static_cast<long>(n) * hw + pos) * 32;
    const float l0 = px[a], l1 = px[2 + a];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    return e1 / (e0 + e1);
  }
  return p.heads.cls[l][(static_cast<long>(n) * 4 + 2 + a) * hw + pos];
}

__device__ __forceinline__ float head_bbox(const DetParams& p, int n, int l, int pos, int ch) {
  const int hw = p.lv[l].fh * p.lv[l].fw;
  if (p.heads.fused) return p.heads.cls[l][(static_cast<long>(n) * hw + pos) * 32 + 4 + ch];
  return p.heads.bbox[l][(static_cast<long>(n) * 8 + ch) * hw + pos];
}

__device__ __forceinline__ float head_lmk(const DetParams& p, int n, int l, int pos, int ch) {
  const int hw = p.lv[l].fh * p.lv[l].fw;
  if (p.heads.fused) return p.heads.cls[l][(static_cast<long>(n) * hw + pos) * 32 + 12 + ch];
  return p.heads.lmk[l][(static_cast<long>(n) * 20 + ch) * hw + pos];
}

// ---- 1. threshold scan: every anchor of every image, warp-aggregated append.
__global__ void __launch_bounds__(256) det_scan_kernel(const DetParams p) {
  const int n = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  bool pass = false;
  float score = 0.f;
  if (idx < p.A) {
    int l, pos, a;
    locate(p, idx, l, pos, a);
    score = head_score(p, n, l, pos, a);
    pass = score >= p.thr;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, pass);
  if (!ballot) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(p.counts + n, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (pass) {
    const int slot = base + __popc(ballot & ((1u << lane) - 1u));
    // ascending key order = descending score, then ascending anchor index
    const unsigned long long key =
        (static_cast<unsigned long long>(~__float_as_uint(score)) << 32) | static_cast<unsigned>(idx);
    p.keys[static_cast<long>(n) * p.capP + slot] = key;
  }
}

__device__ void bitonic_sort(unsigned long long* k, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = k[lo], b = k[hi];
        if ((a > b) == up) { k[lo] = b; k[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float iou_f(const float4 a, const float4 b) {
  const float area_a = (a.z - a.x) * (a.w - a.y);
  const float area_b = (b.z - b.x) * (b.w - b.y);
  const float w = fmaxf(0.f, fminf(a.z, b.z) - fmaxf(a.x, b.x));
  const float h = fmaxf(0.f, fminf(a.w, b.w) - fmaxf(a.y, b.y));
  const float inter = w * h;
  return inter / (area_a + area_b - inter);
}

// ---- 2. one block per image: sort, decode, blocked greedy NMS, output.
__global__ void __launch_bounds__(kSelThreads) det_select_kernel(const DetParams p) {
  __shared__ unsigned long long skeys[kSmemKeys];
  __shared__ unsigned long long chunk_mask[64];
  __shared__ unsigned long long alive_s;
  __shared__ int kept_n_s;
  const int n = blockIdx.x;
  const int tid = threadIdx.x;
  const int K = min(p.counts[n], p.A);
  unsigned long long* gkeys = p.keys + static_cast<long>(n) * p.capP;
  int P = 1;
  while (P < K) P <<= 1;
  unsigned long long* keys = gkeys;
  if (P <= kSmemKeys) {
    for (int i = tid; i < P; i += kSelThreads) skeys[i] = i < K ? gkeys[i] : ~0ull;
    keys = skeys;
  } else {
    for (int i = K + tid; i < P; i += kSelThreads) gkeys[i] = ~0ull;
  }
  __syncthreads();
  if (K > 1) bitonic_sort(keys, P);

  // decode boxes of the sorted candidates
  float4* sbox = reinterpret_cast<float4*>(p.sbox) + static_cast<long>(n) * p.A;
  for (int i = tid; i < K; i += kSelThreads) {
    const int idx = static_cast<int>(keys[i] & 0xffffffffull);
    int l, pos, a;
    locate(p, idx, l, pos, a);
    const Level lv = p.lv[l];
    const int h = pos / lv.fw, w = pos % lv.fw;
    const float lo = p.heads.anchor_lo[l][a], hi = p.heads.anchor_hi[l][a];
    const float x1 = lo + static_cast<float>(w * lv.stride), y1 = lo + static_cast<float>(h * lv.stride);
    const float x2 = hi + static_cast<float>(w * lv.stride), y2 = hi + static_cast<float>(h * lv.stride);
    const float aw = x2 - x1 + 1.f, ah = y2 - y1 + 1.f;
    const float cx = x1 + 0.5f * (aw - 1.f), cy = y1 + 0.5f * (ah - 1.f);
    const float d0 = head_bbox(p, n, l, pos, a * 4 + 0), d1 = head_bbox(p, n, l, pos, a * 4 + 1);
    const float d2 = head_bbox(p, n, l, pos, a * 4 + 2), d3 = head_bbox(p, n, l, pos, a * 4 + 3);
    const float pcx = d0 * aw + cx, pcy = d1 * ah + cy;
    // correctly rounded fp32 exp (see oracle/detect.py exp_f32)
    const float pw = static_cast<float>(exp(static_cast<double>(d2))) * aw;
    const float ph = static_cast<float>(exp(static_cast<double>(d3))) * ah;
    sbox[i] = make_float4(pcx - 0.5f * (pw - 1.f), pcy - 0.5f * (ph - 1.f),
                          pcx + 0.5f * (pw - 1.f), pcy + 0.5f * (ph - 1.f));
  }
  if (tid == 0) kept_n_s = 0;
  __syncthreads();

  int* kept = p.kept + static_cast<long>(n) * p.A;
  for (int base = 0; base < K; base += 64) {
    const int m = min(64, K - base);
    if (tid < 64) chunk_mask[tid] = 0ull;
    if (tid == 0) alive_s = m == 64 ? ~0ull : ((1ull << m) - 1ull);
    __syncthreads();
    const int kept_n = kept_n_s;
    // A: suppress chunk members overlapping an already kept box
    for (int t = tid; t < m * kept_n; t += kSelThreads) {
      const int c = t % m, k = t / m;
      const float ovr = iou_f(sbox[kept[k]], sbox[base + c]);
      if (static_cast<double>(ovr) > p.nms_thr) atomicAnd(&alive_s, ~(1ull << c));
    }
    // B: pairwise overlaps inside the chunk (a < b)
    for (int t = tid; t < m * m; t += kSelThreads) {
      const int a = t / m, b = t % m;
      if (a < b) {
        const float ovr = iou_f(sbox[base + a], sbox[base + b]);
        if (static_cast<double>(ovr) > p.nms_thr) atomicOr(&chunk_mask[a], 1ull << b);
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long alive = alive_s;
      int kn = kept_n;
      for (int a = 0; a < m; ++a) {
        if ((alive >> a) & 1ull) {
          kept[kn++] = base + a;
          alive &= ~chunk_mask[a];
        }
      }
      kept_n_s = kn;
    }
    __syncthreads();
  }

  // output survivors in score-descending order
  const int S = kept_n_s;
  if (tid == 0) { p.out_count[n] = S; p.out_cand[n] = K; }
  float* out = p.out_det + static_cast<long>(n) * p.max_det * 16;
  for (int i = tid; i < min(S, p.max_det); i += kSelThreads) {
    const int pos_sorted = kept[i];
    const unsigned long long key = keys[pos_sorted];
    const int idx = static_cast<int>(key & 0xffffffffull);
    const float score = __uint_as_float(~static_cast<unsigned>(key >> 32));
    int l, pos, a;
    locate(p, idx, l, pos, a);
    const Level lv = p.lv[l];
    const int h = pos / lv.fw, w = pos % lv.fw;
    const float lo = p.heads.anchor_lo[l][a], hi = p.heads.anchor_hi[l][a];
    const float x1 = lo + static_cast<float>(w * lv.stride), y1 = lo + static_cast<float>(h * lv.stride);
    const float x2 = hi + static_cast<float>(w * lv.stride), y2 = hi + static_cast<float>(h * lv.stride);
    const float aw = x2 - x1 + 1.f, ah = y2 - y1 + 1.f;
    const float cx = x1 + 0.5f * (aw - 1.f), cy = y1 + 0.5f * (ah - 1.f);
    float* o = out + static_cast<long>(i) * 16;
    const float4 b = sbox[pos_sorted];
    o[0] = score; o[1] = b.x; o[2] = b.y; o[3] = b.z; o[4] = b.w;
    for (int q = 0; q < 5; ++q) {
      o[5 + 2 * q] = head_lmk(p, n, l, pos, a * 10 + 2 * q) * aw + cx;
      o[6 + 2 * q] = head_lmk(p, n, l, pos, a * 10 + 2 * q + 1) * ah + cy;
    }
    o[15] = __int_as_float(idx);
  }
}

}  // namespace

size_t detect_workspace_bytes(int N, int H, int W) {
  int A = 0;
  for (int s : {32, 16, 8}) A += ceil_div(H, s) * ceil_div(W, s) * 2;
  int capP = 1;
  while (capP < A) capP <<= 1;
  size_t b = 0;
  b += size_t(N) * capP * 8;        // keys
  b += size_t(N) * 4;               // counts
  b += size_t(N) * A * 16;          // sorted boxes
  b += size_t(N) * A * 4;           // kept
  return b + 1024;
}

void detect_post_launch(const DetHeads& heads, int N, int H, int W, float thr, double nms_thr,
                        int max_det, void* workspace, int* out_count, int* out_cand,
                        float* out_det, cudaStream_t s) {
  if (N == 0) return;
  DetParams p{};
  p.heads = heads;
  p.N = N; p.H = H; p.W = W;
  int off = 0, i = 0;
  for (int st : {32, 16, 8}) {
    p.lv[i].stride = st;
    p.lv[i].fh = ceil_div(H, st);
    p.lv[i].fw = ceil_div(W, st);
    p.lv[i].off = off;
    off += p.lv[i].fh * p.lv[i].fw * 2;
    ++i;
  }
  p.A = off;
  p.capP = 1;
  while (p.capP < p.A) p.capP <<= 1;
  p.thr = thr; p.nms_thr = nms_thr; p.max_det = max_det;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  p.keys = reinterpret_cast<unsigned long long*>(ws); ws += size_t(N) * p.capP * 8;
  p.sbox = reinterpret_cast<float*>(ws); ws += size_t(N) * p.A * 16;
  p.kept = reinterpret_cast<int*>(ws); ws += size_t(N) * p.A * 4;
  p.counts = reinterpret_cast<int*>(ws);
  p.out_count = out_count; p.out_cand = out_cand; p.out_det = out_det;
  TR_CUDA(cudaMemsetAsync(p.counts, 0, size_t(N) * 4, s));
  dim3 grid(ceil_div(p.A, 256), N);
  det_scan_kernel<<<grid, 256, 0, s>>>(p);
  TR_CUDA(cudaGetLastError());
  det_select_kernel<<<N, kSelThreads, 0, s>>>(p);
  TR_CUDA(cudaGetLastError());
}

}  // namespace trb

// ---------------------------------------------------------------------------
// Face alignment: inverse-affine bilinear warp of a detected face to the 112x112
// ArcFace crop, bit-exact with PIL's Image.transform(AFFINE, BILINEAR,
// fillcolor=0) that the reference uses (arcface/wrapper.py:22-72; Pillow
// Geometry.c affine_transform + bilinear_filter32RGB: sample at pixel centres,
// neighbours clamped to the image, lerp in double along x then y, truncate).
// Output is the reference model input layout: (F,3,S,S) uint8, channels BGR.
// This file is compiled with -fmad=false, so the double arithmetic is not contracted.
// ---------------------------------------------------------------------------
namespace trb {
namespace {

__global__ void __launch_bounds__(128)
face_align_kernel(const uint8_t* __restrict__ frames, int H, int W, const double* __restrict__ coef,
                  const int* __restrict__ image_index, uint8_t* __restrict__ out, int S) {
  const int f = blockIdx.y, y = blockIdx.x, x = threadIdx.x;
  if (x >= S) return;
  const double* a = coef + 6 * f;
  const uint8_t* im = frames + static_cast<long>(image_index[f]) * H * W * 3;
  const double xc = x + 0.5, yc = y + 0.5;
  double xin = a[0] * xc + a[1] * yc + a[2];
  double yin = a[3] * xc + a[4] * yc + a[5];
  uint8_t px[3] = {0, 0, 0};
  if (!(xin < 0.0 || xin >= W || yin < 0.0 || yin >= H)) {
    xin -= 0.5;
    yin -= 0.5;
    const int x0 = xin < 0.0 ? static_cast<int>(floor(xin)) : static_cast<int>(xin);
    const int y0 = yin < 0.0 ? static_cast<int>(floor(yin)) : static_cast<int>(yin);
    const double dx = xin - x0, dy = yin - y0;
    const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
    const int ya = min(max(y0, 0), H - 1);
    const bool has_next = y0 + 1 >= 0 && y0 + 1 < H;
    const uint8_t* r0 = im + static_cast<long>(ya) * W * 3;
    const uint8_t* r1 = im + static_cast<long>(has_next ? y0 + 1 : ya) * W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double p00 = r0[xa * 3 + c], p01 = r0[xb * 3 + c];
      double v1 = p00 + (p01 - p00) * dx;
      double v2 = v1;
      if (has_next) {
        const double p10 = r1[xa * 3 + c], p11 = r1[xb * 3 + c];
        v2 = p10 + (p11 - p10) * dx;
      }
      v1 = v1 + (v2 - v1) * dy;
      px[c] = static_cast<uint8_t>(v1);
    }
  }
  // RGB -> BGR planes
  uint8_t* o = out + static_cast<long>(f) * 3 * S * S + static_cast<long>(y) * S + x;
  o[0] = px[2];
  o[static_cast<long>(S) * S] = px[1];
  o[2L * S * S] = px[0];
}

}  // namespace

// Five-point similarity per detected face, on the device: the inverse 2x3 matrix (PIL AFFINE
// data) that maps the 112x112 crop onto the frame, from the landmarks of the detection rows.
// In two dimensions the least-squares similarity (Umeyama 1991, what skimage's
// SimilarityTransform.estimate computes; arcface/wrapper.py:47-61) has a closed form — no SVD:
// with centred landmarks p and template points q,
//     a = sum p.q,  b = sum p x q,  rotation = atan2(b, a),  scale = hypot(a, b) / sum |p|^2
// (hypot(a, b) = S0 + d S1 of Umeyama's diag(1, d), also when the best orthogonal fit would
// be a reflection).  Landmarks are first mapped back to frame pixels exactly like
// Detection.resize_out does (float32 value / float32 scale, round half to even), because
// that is what the reference's extract_features is given.
__global__ void face_similarity_kernel(const float* __restrict__ det, const int* __restrict__ count,
                                       int N, int max_det, float scale, int cap,
                                       double* __restrict__ coef, int* __restrict__ image_index,
                                       int* __restrict__ total) {
  // float32 template, x shifted by 8 IN float32 for the 112-wide crop (wrapper.py:39-48)
  const float txf[5] = {30.2946f + 8.0f, 65.5318f + 8.0f, 48.0252f + 8.0f, 33.5493f + 8.0f, 62.7299f + 8.0f};
  const float tyf[5] = {51.6963f, 51.5014f, 71.7366f, 92.3655f, 92.2041f};
  const int n = blockIdx.x;
  int base = 0;
  for (int i = 0; i < n; ++i) base += min(count[i], max_det);
  const int cnt = min(count[n], max_det);
  if (n == N - 1 && threadIdx.x == 0) *total = base + cnt;
  for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
    const int f = base + k;
    if (f >= cap) continue;
    const float* row = det + (static_cast<long>(n) * max_det + k) * 16;
    double px[5], py[5], mx = 0, my = 0, qx = 0, qy = 0;
    for (int j = 0; j < 5; ++j) {
      px[j] = static_cast<double>(rintf(row[5 + 2 * j] / scale));
      py[j] = static_cast<double>(rintf(row[6 + 2 * j] / scale));
      mx += px[j]; my += py[j];
      qx += static_cast<double>(txf[j]);
      qy += static_cast<double>(tyf[j]);
    }
    mx /= 5; my /= 5; qx /= 5; qy /= 5;
    double a = 0, b = 0, var = 0;
    for (int j = 0; j < 5; ++j) {
      const double ux = px[j] - mx, uy = py[j] - my;
      const double vx = static_cast<double>(txf[j]) - qx;
      const double vy = static_cast<double>(tyf[j]) - qy;
      a += ux * vx + uy * vy;
      b += ux * vy - uy * vx;
      var += ux * ux + uy * uy;
    }
    double* o = coef + 6 * f;
    image_index[f] = n;
    const double nrm = sqrt(a * a + b * b);
    if (!(var > 0.0) || !(nrm > 0.0)) {           // degenerate landmarks: skimage returns NaNs
      for (int j = 0; j < 6; ++j) o[j] = nan("");
      continue;
    }
    const double s = nrm / var, c = a / nrm, sn = b / nrm;
    // forward: q = s R p + t,  R = [[c, -sn], [sn, c]],  t = mu_q - s R mu_p
    const double t0 = qx - s * (c * mx - sn * my), t1 = qy - s * (sn * mx + c * my);
    // inverse: p = R^T (q - t) / s
    const double is = 1.0 / s;
    o[0] = c * is;  o[1] = sn * is; o[2] = -(c * t0 + sn * t1) * is;
    o[3] = -sn * is; o[4] = c * is; o[5] = -(-sn * t0 + c * t1) * is;
  }
}

void face_similarity_launch(const float* det, const int* count, int N, int max_det, float scale,
                            int cap, double* coef, int* image_index, int* total, cudaStream_t s) {
  if (N == 0) return;
  face_similarity_kernel<<<N, 64, 0, s>>>(det, count, N, max_det, scale, cap, coef, image_index, total);
  TR_CUDA(cudaGetLastError());
}

void face_align_launch(const uint8_t* frames, int H, int W, const double* coef,
                       const int* image_index, int F, uint8_t* out, int S, cudaStream_t s) {
  if (F == 0) return;
  TR_CHECK(S <= 128, "crop side");
  face_align_kernel<<<dim3(S, F), 128, 0, s>>>(frames, H, W, coef, image_index, out, S);
  TR_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// Faces given WITHOUT landmarks (arcface/wrapper.py:75-99): the image is resized so that its
// longer side is 112 with PIL's default Image.resize — for RGB images Pillow's antialiased
// BICUBIC resampler — and centred on a zero 112x112 canvas, channels BGR.  Pillow's 8-bit
// resampler (Resample.c) is integer arithmetic on precomputed fixed-point tables, so the device
// version is bit-exact with it:
//   * per axis and output sample: the window [first, first + count) of input samples under the
//     filter (support 2 * max(1, in/out)), bicubic weights (a = -0.5) evaluated in double,
//     normalised by their sum, rounded half away from zero to 22 fractional bits;
//   * horizontal pass over all rows, result rounded to uint8 ((acc + 2^21) >> 22, clamped),
//     then the vertical pass over that intermediate image, rounded the same way.
// The tables depend only on (input size, output size): the host builds them in double exactly
// as Pillow does and ships them with the per-image descriptors in one small upload.
// ---------------------------------------------------------------------------
namespace {

constexpr int kResampleBits = 32 - 8 - 2;

inline double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct LetterDesc {
  long long pix_off, tmp_off;          // bytes into the pixel blob / into the intermediate area
  int h, w, oh, ow, x_min, y_min;
  int kh_off, bh_off, ksize_h;         // int32 offsets into the table area
  int kv_off, bv_off, ksize_v;
};

__device__ __forceinline__ uint8_t resample_round(int acc) {
  return static_cast<uint8_t>(min(max(acc >> kResampleBits, 0), 255));
}

__global__ void __launch_bounds__(256)
letterbox_rows_kernel(const uint8_t* __restrict__ pixels, const LetterDesc* __restrict__ descs,
                      const int* __restrict__ tables, uint8_t* __restrict__ tmp) {
  const LetterDesc d = descs[blockIdx.y];
  const long idx = static_cast<long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long>(d.h) * d.ow) return;
  const int row = static_cast<int>(idx / d.ow), x = static_cast<int>(idx % d.ow);
  const int first = tables[d.bh_off + 2 * x], count = tables[d.bh_off + 2 * x + 1];
  const int* k = tables + d.kh_off + x * d.ksize_h;
  const uint8_t* src = pixels + d.pix_off + (static_cast<long>(row) * d.w + first) * 3;
  int a0 = 1 << (kResampleBits - 1), a1 = a0, a2 = a0;
  for (int i = 0; i < count; ++i) {
    const int c = k[i];
    a0 += src[3 * i] * c;
    a1 += src[3 * i + 1] * c;
    a2 += src[3 * i + 2] * c;
  }
  uint8_t* o = tmp + d.tmp_off + idx * 3;
  o[0] = resample_round(a0);
  o[1] = resample_round(a1);
  o[2] = resample_round(a2);
}

__global__ void __launch_bounds__(128)
letterbox_cols_kernel(const LetterDesc* __restrict__ descs, const int* __restrict__ tables,
                      const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out, int S) {
  const LetterDesc d = descs[blockIdx.y];
  const int y = blockIdx.x, x = threadIdx.x;
  if (x >= S) return;
  const int ox = x - d.x_min, oy = y - d.y_min;
  uint8_t px[3] = {0, 0, 0};
  if (ox >= 0 && ox < d.ow && oy >= 0 && oy < d.oh) {
    const int first = tables[d.bv_off + 2 * oy], count = tables[d.bv_off + 2 * oy + 1];
    const int* k = tables + d.kv_off + oy * d.ksize_v;
    const uint8_t* src = tmp + d.tmp_off + (static_cast<long>(first) * d.ow + ox) * 3;
    int a0 = 1 << (kResampleBits - 1), a1 = a0, a2 = a0;
    for (int i = 0; i < count; ++i) {
      const int c = k[i];
      const uint8_t* p = src + static_cast<long>(i) * d.ow * 3;
      a0 += p[0] * c;
      a1 += p[1] * c;
      a2 += p[2] * c;
    }
    px[0] = resample_round(a0);
    px[1] = resample_round(a1);
    px[2] = resample_round(a2);
  }
  uint8_t* o = out + static_cast<long>(blockIdx.y) * 3 * S * S + static_cast<long>(y) * S + x;
  o[0] = px[2];
  o[static_cast<long>(S) * S] = px[1];
  o[2L * S * S] = px[0];
}

int resample_ksize(int in_size, int out_size) {
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return static_cast<int>(std::ceil(2.0 * filterscale)) * 2 + 1;
}

struct LetterGeometry { int oh, ow, x_min, y_min; };

LetterGeometry letter_geometry(int h, int w, int S) {
  // (reference :77-83) scale on the longer side, int() truncation, centred
  const double scale = static_cast<double>(S) / (w > h ? w : h);
  LetterGeometry g;
  g.ow = static_cast<int>(w * scale);
  g.oh = static_cast<int>(h * scale);
  g.x_min = static_cast<int>((S - g.ow) / 2.0);
  g.y_min = static_cast<int>((S - g.oh) / 2.0);
  return g;
}

size_t align16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }

}  // namespace

int resample_table_host(int in_size, int out_size, int* bounds, int* coeffs) {
  TR_CHECK(in_size > 0 && out_size > 0, "resample sizes");
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  if (!bounds || !coeffs) return ksize;
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = pil_bicubic((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int* k = coeffs + static_cast<long>(xx) * ksize;
    for (int x = 0; x < ksize; ++x) {
      if (x >= xmax) { k[x] = 0; continue; }
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0 ? static_cast<int>(-0.5 + v * (1 << kResampleBits))
                   : static_cast<int>(0.5 + v * (1 << kResampleBits));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

namespace {

// Sizes of the three workspace areas for a list of images (sizes: n x (height, width)).
struct LetterLayout { size_t desc_bytes, table_ints, tmp_bytes; };

LetterLayout letter_layout(const int* sizes, int n, int S) {
  LetterLayout L{align16(sizeof(LetterDesc) * static_cast<size_t>(n)), 0, 0};
  for (int i = 0; i < n; ++i) {
    const int h = sizes[2 * i], w = sizes[2 * i + 1];
    TR_CHECK(h > 0 && w > 0, "image " + std::to_string(i) + " is empty");
    const LetterGeometry g = letter_geometry(h, w, S);
    TR_CHECK(g.ow > 0 && g.oh > 0, "height and width must be > 0 (image " + std::to_string(i) + " is too thin)");
    L.table_ints += static_cast<size_t>(g.ow) * (resample_ksize(w, g.ow) + 2)
                  + static_cast<size_t>(g.oh) * (resample_ksize(h, g.oh) + 2);
    L.tmp_bytes += align16(static_cast<size_t>(h) * g.ow * 3);
  }
  return L;
}

}  // namespace

size_t face_letterbox_workspace_bytes(const int* sizes, int n, int S) {
  const LetterLayout L = letter_layout(sizes, n, S);
  return L.desc_bytes + align16(L.table_ints * 4) + L.tmp_bytes;
}

void face_letterbox_launch(const uint8_t* pixels, const long long* offsets, const int* sizes, int n,
                           int S, void* workspace, uint8_t* out, cudaStream_t s) {
  if (n == 0) return;
  TR_CHECK(S <= 128, "crop side");
  const LetterLayout L = letter_layout(sizes, n, S);
  const size_t head_bytes = L.desc_bytes + align16(L.table_ints * 4);
  std::vector<char> head(head_bytes, 0);
  LetterDesc* descs = reinterpret_cast<LetterDesc*>(head.data());
  int* tables = reinterpret_cast<int*>(head.data() + L.desc_bytes);
  size_t t = 0, tmp_off = 0;
  long max_elems = 0;
  for (int i = 0; i < n; ++i) {
    const int h = sizes[2 * i], w = sizes[2 * i + 1];
    const LetterGeometry g = letter_geometry(h, w, S);
    LetterDesc& d = descs[i];
    d.pix_off = offsets[i];
    d.tmp_off = static_cast<long long>(tmp_off);
    d.h = h; d.w = w; d.oh = g.oh; d.ow = g.ow; d.x_min = g.x_min; d.y_min = g.y_min;
    d.ksize_h = resample_ksize(w, g.ow);
    d.bh_off = static_cast<int>(t); t += 2 * static_cast<size_t>(g.ow);
    d.kh_off = static_cast<int>(t); t += static_cast<size_t>(g.ow) * d.ksize_h;
    resample_table_host(w, g.ow, tables + d.bh_off, tables + d.kh_off);
    d.ksize_v = resample_ksize(h, g.oh);
    d.bv_off = static_cast<int>(t); t += 2 * static_cast<size_t>(g.oh);
    d.kv_off = static_cast<int>(t); t += static_cast<size_t>(g.oh) * d.ksize_v;
    resample_table_host(h, g.oh, tables + d.bv_off, tables + d.kv_off);
    tmp_off += align16(static_cast<size_t>(h) * g.ow * 3);
    max_elems = std::max(max_elems, static_cast<long>(h) * g.ow);
  }
  TR_CHECK(t == L.table_ints, "table layout");
  char* ws = static_cast<char*>(workspace);
  // pageable source: the runtime stages it before returning, so `head` may go out of scope
  TR_CUDA(cudaMemcpyAsync(ws, head.data(), head_bytes, cudaMemcpyHostToDevice, s));
  const LetterDesc* ddesc = reinterpret_cast<const LetterDesc*>(ws);
  const int* dtab = reinterpret_cast<const int*>(ws + L.desc_bytes);
  uint8_t* dtmp = reinterpret_cast<uint8_t*>(ws + head_bytes);
  letterbox_rows_kernel<<<dim3(static_cast<unsigned>((max_elems + 255) / 256), n), 256, 0, s>>>(
      pixels, ddesc, dtab, dtmp);
  TR_CUDA(cudaGetLastError());
  letterbox_cols_kernel<<<dim3(S, n), 128, 0, s>>>(ddesc, dtab, dtmp, out, S);
  TR_CUDA(cudaGetLastError());
}

}  // namespace trb
