// RetinaFace / OpenPose post-processing entry points (host side).
#pragma once
#include "common.cuh"

namespace trb {

// Head tensors of the three pyramid levels, index 0 = stride 32, 1 = 16, 2 = 8.
//  fused == 0: reference layout, NCHW fp32: cls (N,4,h,w) already soft-maxed,
//              bbox (N,8,h,w), lmk (N,20,h,w).
//  fused == 1: cls[l] points at an NHWC fp32 tensor of 32 channels
//              [4 class logits | 8 bbox | 20 landmark]; soft-max done in-kernel.
struct DetHeads {
  const float* cls[3];
  const float* bbox[3];
  const float* lmk[3];
  int fused;
  float anchor_lo[3][2];
  float anchor_hi[3][2];
};

size_t detect_workspace_bytes(int N, int H, int W);
void detect_post_launch(const DetHeads& heads, int N, int H, int W, float thr, double nms_thr,
                        int max_det, void* workspace, int* out_count, int* out_cand,
                        float* out_det, cudaStream_t s);

// ---- OpenPose parse
constexpr int kPeakCap = 512;     // peaks per (frame, part)
constexpr int kCandCap = 4096;    // accepted pairs per (frame, limb)
constexpr int kHumanCap = 128;    // humans per frame

struct PoseOut {
  int* count;        // [N] humans after filtering
  int* keypoints;    // [N][kHumanCap][18][3] int32 (x, y, present)
  double* score;     // [N][kHumanCap]
  int* status;       // [N] overflow bits: 1 peaks, 2 candidates, 4 humans
};

size_t pose_workspace_bytes(int N);
void pose_parse_launch(const float* paf, const float* heat, int N, int h, int w, double scale,
                       void* workspace, const PoseOut& out, cudaStream_t s);
void bicubic_table_host(float out[32]);

// ---- face alignment (PIL-exact inverse affine bilinear warp to (F,3,S,S) BGR)
void face_similarity_launch(const float* det, const int* count, int N, int max_det, float scale,
                            int cap, double* coef, int* image_index, int* total, cudaStream_t s);
void face_align_launch(const uint8_t* frames, int H, int W, const double* coef,
                       const int* image_index, int F, uint8_t* out, int S, cudaStream_t s);

// ---- faces without landmarks (PIL-exact antialiased bicubic resize to max side S, centred on a
//      zero canvas, (n,3,S,S) BGR).  sizes: n x (height, width) and offsets: n byte offsets of the
//      RGB HWC images inside `pixels`, both on the HOST; pixels / workspace / out on the device.
int resample_table_host(int in_size, int out_size, int* bounds, int* coeffs);
size_t face_letterbox_workspace_bytes(const int* sizes, int n, int S);
void face_letterbox_launch(const uint8_t* pixels, const long long* offsets, const int* sizes, int n,
                           int S, void* workspace, uint8_t* out, cudaStream_t s);

}  // namespace trb
