// Shared declarations for the terran_b200 native library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <stdexcept>
#include <string>

namespace trb {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const std::string& msg) { throw Error(msg); }

#define TR_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess)                                                         \
      ::trb::fail(std::string(#expr) + " -> " + cudaGetErrorString(_e) + " at " +  \
                  __FILE__ + ":" + std::to_string(__LINE__));                      \
  } while (0)

#define TR_CHECK(cond, msg)                                                        \
  do {                                                                             \
    if (!(cond)) ::trb::fail(std::string("check failed: ") + #cond + ": " + (msg)); \
  } while (0)

// Per-device caches (function attributes, constant tables, SM counts) are indexed by the
// CUDA device ordinal: every model class accepts an arbitrary device=.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) fail(std::string("cudaGetDevice -> ") + cudaGetErrorString(e));
  if (dev < 0 || dev >= kMaxDevices) fail("device ordinal out of range: " + std::to_string(dev));
  return dev;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_PRELU = 2 };

// One NHWC fp16 activation buffer view: element (n,h,w,c) lives at
// ptr[((n*H + h)*W + w)*cs + coff + c].
struct View {
  __half* ptr = nullptr;
  int N = 0, H = 0, W = 0;
  int cs = 0;    // channel stride (total channels of the underlying buffer)
  int coff = 0;  // channel offset of this view
  int C = 0;     // channels of this view (padded to a multiple of 8)
};

// Everything a convolution launch needs, shared by the tcgen05 and the direct
// kernels so that the two can be cross-checked on identical arguments.
struct ConvArgs {
  View in;            // input view (C = cin_pad)
  View out;           // primary output (fp16) — unused when out_f32 != nullptr
  View out2;          // optional second output: out2 = out * scale2 + shift2
  View res;           // optional residual added after the activation
  int res_up2 = 0;    // residual is read at (h>>1, w>>1) (nearest x2 upsample)
  float* out_f32 = nullptr;   // optional fp32 NHWC output with stride out.cs/coff
  const __half* w = nullptr;  // [cout_pad][kh][kw][cin_pad]
  const float* scale = nullptr;   // [cout_pad]
  const float* shift = nullptr;   // [cout_pad]
  const float* slope = nullptr;   // [cout_pad] (PReLU) or null
  const float* scale2 = nullptr;  // [cout_pad] or null
  const float* shift2 = nullptr;
  // optional [9][cout_pad] border-class shifts replacing `shift` (3x3, pad 1, stride 1):
  // class = 3 * (first / inner / last output row) + (first / inner / last output column)
  const float* shift9 = nullptr;
  // grouped convolution (conv_patch only; the executor splits it for the other kernels): the
  // cout_pad filter rows are `groups` equal blocks, block g reads channels in.coff + g * cin_pad
  int groups = 1;
  int patch_rows = 0;  // conv_patch: force R rows of 8-pixel groups per tile (0 = choose)
  int pool2 = 0;       // conv_patch: fuse the 2x2 / stride-2 max-pool that follows; `out` is the POOLED view
  View pool_out;       // conv_tc: OPTIONAL pooled output view — the plan fuses the pool if its tile layout allows
                       // (conv_tc_plan_pooled tells), and then writes this tensor INSTEAD of `out`
  int cout_pad = 0;   // multiple of 16
  int cout_store = 0; // channels actually written (multiple of 8, <= cout_pad)
  int cin_pad = 0;    // multiple of 16
  int kh = 1, kw = 1, stride = 1, pad = 0;
  int act = ACT_NONE;
  int H_out = 0, W_out = 0;
  // Stream-K scratch shared by the plans of one stream (conv_tc_sk_scratch_bytes() bytes,
  // zero-initialised); null: the plan allocates its own.
  void* sk_scratch = nullptr;
};

// conv_tc.cu
size_t conv_tc_sk_scratch_bytes();
bool conv_tc_eligible(const ConvArgs& a);
struct ConvTcPlan;   // opaque: tensor maps + launch geometry
ConvTcPlan* conv_tc_plan_create(const ConvArgs& a);
void conv_tc_plan_destroy(ConvTcPlan* p);
void conv_tc_launch(const ConvTcPlan* p, cudaStream_t s);
double conv_tc_plan_flops(const ConvTcPlan* p);
int conv_tc_last_timeout();   // pipeline wait that timed out before a trap (0 = none)
int* conv_tc_error_flag();    // device-visible word a trapping kernel writes its timeout code to

// conv_patch.cu — tcgen05 kernel with filters on M and a resident input patch on N
bool conv_patch_eligible(const ConvArgs& a);
size_t conv_patch_scratch_bytes();
struct ConvPatchPlan;
ConvPatchPlan* conv_patch_plan_create(const ConvArgs& a, int* err_flag);
bool conv_tc_plan_pooled(const ConvTcPlan* p);   // did the plan take ConvArgs::pool_out?
void conv_patch_plan_destroy(ConvPatchPlan* p);
void conv_patch_launch(const ConvPatchPlan* p, cudaStream_t s);
void conv_patch_plan_describe(const ConvPatchPlan* p, int* axis, int* R, int* tiles, int* stages);

// conv_direct.cu
void conv_direct_launch(const ConvArgs& a, cudaStream_t s);

// conv_mma.cu — warp-level mma.sync kernels for small-channel layers (RetinaFace)
bool conv_mma_eligible(const ConvArgs& a);
void conv_mma_launch(const ConvArgs& a, cudaStream_t s);
// Fused depthwise 3x3 (pad 1, stride 1|2) + BN + ReLU -> 1x1 conv + scale/shift (+ReLU).
struct SepArgs {
  View in, out;
  const float* dw_w;      // fp32 [3][3][cin_pad] (unfused cross-check path)
  const __half* dw_w16;   // the same filter in fp16 (fused kernel)
  const float* dw_scale; const float* dw_shift;
  int stride;             // of the depthwise stage
  const __half* w;        // fp16 [cout_pad][cin_pad]
  const float* scale; const float* shift;
  int cin_pad, cout_pad, cout_store, act;
};
bool sep_mma_eligible(const SepArgs& a);
void sep_mma_launch(const SepArgs& a, cudaStream_t s);

struct StemArgs {
  const uint8_t* in;      // u8 image, arbitrary element strides
  long sn, sh, sw, sc;    // element strides of (n, h, w, c)
  int N, H, W;            // input dims (3 channels)
  float in_scale, in_shift;   // x' = x * in_scale + in_shift on in-bounds taps
  View out, out2;
  const float* w;         // fp32 [cout][3][3][3] (kh, kw, c)
  const float* scale; const float* shift; const float* slope;
  const float* scale2; const float* shift2;
  int cout, stride, act;
  int use_mma;            // 1: tensor-core kernel (fp16 operands), 0: fp32 CUDA-core kernel
};
void stem_launch(const StemArgs& a, cudaStream_t s);

struct DwArgs {
  View in, out;
  const float* w;         // fp32 [3][3][C]
  const float* scale; const float* shift;
  int stride;             // pad 1, 3x3, ReLU
};
void dwconv_launch(const DwArgs& a, cudaStream_t s);
void maxpool2_launch(const View& in, const View& out, cudaStream_t s);
void copy_slice_launch(const View& in, const View& out, cudaStream_t s);
// NHWC fp16 slice -> NCHW fp32 (C real channels)
void export_nchw_launch(const View& in, int C, float* out, cudaStream_t s);
void export_nchw_f32_launch(const float* in, int N, int H, int W, int cs, int coff, int C,
                            float* out, int softmax_pairs, cudaStream_t s);
void l2_normalize_launch(const float* in, float* out, int N, int D, cudaStream_t s);
void resize_bilinear_u8_launch(const uint8_t* src, int N, int H, int W, uint8_t* dst, int h,
                               int w, cudaStream_t s);

}  // namespace trb
