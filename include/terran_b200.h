/* terran_b200 — C ABI of the B200-native Terran hot path.
 *
 * The reference (terran-project/terran) is pure Python/PyTorch and has no FFI;
 * each entry point below names the reference code it replaces.  All pointers
 * marked "device" are CUDA device pointers on the current device (the Python
 * side passes torch tensor .data_ptr()); `stream` is a cudaStream_t (may be 0).
 * Every function returns 0 on success and a non-zero code on failure, in which
 * case tr_last_error() describes the failure (the Python binding raises).
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef TERRAN_B200_H
#define TERRAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tr_net tr_net;

/* ---- library ----------------------------------------------------------- */
int tr_version(void);
const char* tr_last_error(void);
/* cudaSetDevice + capability check (sm_100 required). Replaces the implicit
 * `default_device` selection of terran/defaults.py:3-5. */
int tr_init(int device);

/* ---- layer program ("net") ----------------------------------------------
 * A net is a list of fused layer ops over NHWC fp16 activation buffers plus a
 * packed weight blob (BatchNorm folded, fp16 [cout][kh][kw][cin] filters).
 * terran_b200/weights.py builds it from the reference state_dict layout
 * (SURVEY.md Appendix C); shapes are inferred per call from (N,H,W).
 * Replaces nn.Module construction + load_state_dict
 * (retinaface/wrapper.py:16-22, arcface/wrapper.py:13-19, openpose/wrapper.py:26-34). */
enum { TR_OP_STEM = 0, TR_OP_CONV = 1, TR_OP_DWCONV = 2, TR_OP_MAXPOOL = 3, TR_OP_COPY = 4,
       TR_OP_VIEW = 5,
       /* depthwise 3x3 (+BN+ReLU) fused with the 1x1 conv (+BN+ReLU) that follows it: the
        * sep_block of one ConvSepBlock and the conv_block of the next (retinaface/model.py:6-50).
        * k/stride/pad describe the depthwise stage, w/scale/shift the 1x1, dw_* the depthwise. */
       TR_OP_SEPCONV = 6 };
/* tr_op_desc.engine */
enum { TR_ENGINE_AUTO = 0,   /* tcgen05 implicit GEMM where eligible, else the direct kernel */
       TR_ENGINE_MMA = 1 };  /* warp-level mma.sync kernel for small-channel layers where eligible */
enum { TR_ACT_NONE = 0, TR_ACT_RELU = 1, TR_ACT_PRELU = 2 };
enum { TR_SYNC_FORK = 1, TR_SYNC_JOIN = 2 };

typedef struct tr_buffer_desc {
  int32_t channels;   /* total channels (multiple of 8) */
  int32_t is_f32;     /* 1: fp32 NHWC output buffer, 0: fp16 */
} tr_buffer_desc;

typedef struct tr_op_desc {
  int32_t type;
  int32_t in, in_coff, in_c;        /* input buffer (-1 = the u8 image), channel offset, channels */
  int32_t out, out_coff, out_c;     /* output buffer, channel offset, channels written */
  int32_t out2, out2_coff;          /* optional second output (out*scale2+shift2) or -1 */
  int32_t res, res_coff, res_up2;   /* optional residual (-1 = none); res_up2: read at (h/2,w/2) */
  int32_t k, stride, pad, act;
  int32_t cout_pad;                 /* filter rows in the blob (multiple of 16) */
  int32_t cin_real, cout_real;      /* un-padded channel counts (algorithmic flop accounting) */
  int32_t force_direct;             /* 1: never use the tcgen05 kernel for this op */
  int32_t lane;                     /* 0: the caller's stream; 1: the net's side stream (independent branch) */
  int32_t sync;                     /* TR_SYNC_FORK: ops of lane 1 issued later start after this op;
                                       TR_SYNC_JOIN: this op starts after everything issued on lane 1 */
  int64_t w_off, scale_off, shift_off, slope_off, scale2_off, shift2_off; /* blob byte offsets, -1 = none */
  float in_scale, in_shift;         /* stem only: x' = x*in_scale + in_shift on in-bounds taps */
  int64_t dw_w_off, dw_scale_off, dw_shift_off;   /* TR_OP_SEPCONV: depthwise filter [3][3][C] fp32 + BN */
  int64_t dw_w16_off;               /* the same filter rounded to fp16 (fused kernel operand) */
  int32_t engine;                   /* TR_OP_CONV: TR_ENGINE_* */
  /* TR_OP_CONV: grouped convolution.  groups = G > 1: the filter rows and the output channels
   * are G equal blocks; block g reads input channels [in_coff + g * in_c, +in_c) — how the two
   * OpenPose branches (same shapes, different inputs) run as ONE launch.  0 / 1 = dense. */
  int32_t groups;
  /* TR_OP_CONV, 3x3 pad 1 stride 1: [9][cout_pad] fp32 border-class shifts replacing the
   * shift vector, class = 3 * (first / inner / last output row) + (first / inner / last
   * column).  How a BatchNorm that PRECEDES a zero-padded conv is folded into it exactly
   * (arcface/model.py:11-14): its scale goes into the filters, its shift becomes a term that
   * depends on which taps are in bounds.  -1 = none. */
  int64_t shift9_off;
} tr_op_desc;

int tr_net_create(const tr_buffer_desc* buffers, int n_buffers, const tr_op_desc* ops, int n_ops,
                  const void* weights_host, size_t weight_bytes, tr_net** out);
void tr_net_destroy(tr_net* net);
/* 0 = auto (tcgen05 where eligible), 1 = force the CUDA-core direct kernels (cross-check). */
int tr_net_set_mode(tr_net* net, int force_direct);
/* Run the program on a u8 image batch with arbitrary element strides (so both
 * NHWC RGB frames and the reference's NCHW BGR crops are accepted as they are). */
int tr_net_run(tr_net* net, const uint8_t* image_dev, int N, int H, int W, int64_t stride_n,
               int64_t stride_h, int64_t stride_w, int64_t stride_c, void* stream);
/* Buffer of the most recent run: device pointer and NHWC dims. */
int tr_net_buffer(tr_net* net, int buffer, void** ptr, int* N, int* H, int* W, int* channels);
/* Export C channels starting at coff of an fp16 buffer to NCHW fp32 (device). */
int tr_net_export_nchw(tr_net* net, int buffer, int coff, int C, float* out_dev, void* stream);
/* Same for an fp32 buffer; softmax_pairs != 0 soft-maxes channel a against a^2
 * (the class-pair softmax of retinaface/model.py:283-295). */
int tr_net_export_nchw_f32(tr_net* net, int buffer, int coff, int C, float* out_dev,
                           int softmax_pairs, void* stream);
/* Algorithmic (un-padded) flops of the tensor-core ops and launch counts of the
 * most recent run (for bench.py). */
int tr_net_stats(tr_net* net, double* tc_flops, int* tc_launches, int* total_launches);
/* Per-op device timing: when enabled every op of a run is bracketed by CUDA
 * events on the launch stream; tr_net_profile synchronises and returns, per op,
 * the milliseconds, whether it ran on the tcgen05 kernel and its algorithmic
 * flops.  Returns the number of ops through *n_ops (at most cap are written). */
int tr_net_set_profile(tr_net* net, int enable);
int tr_net_profile(tr_net* net, float* ms, int32_t* is_tc, double* flops, int cap, int* n_ops);

/* ---- models from a checkpoint, without Python --------------------------------
 * The layer programs of the three reference models are built by the library itself from the
 * reference's state_dict, passed as a flat "TRSD" blob of named tensors:
 *   "TRSD", u32 version (1), u32 count, then per tensor: u16 name length, name, u8 dtype
 *   (0 = float32, 1 = int64), u8 ndim, i64 dims[ndim], u64 byte count, zero padding to an 8-byte
 *   boundary, data.   (terran_b200.weights.pack_state_dict writes it from a torch state_dict.)
 * tr_program_* are host-only (no CUDA call): they replace nn.Module construction +
 * load_state_dict + BatchNorm folding (retinaface/wrapper.py:16-22, arcface/wrapper.py:13-19,
 * openpose/wrapper.py:26-34). */
typedef struct tr_program tr_program;
/* model: "retinaface" | "arcface" | "openpose"; flags bit 0 (retinaface only): one op per
 * reference layer instead of the fused depthwise+1x1 program (cross-check). */
int tr_program_build(const char* model, const void* state_dict_blob, size_t bytes, int flags,
                     tr_program** out);
void tr_program_destroy(tr_program* program);
/* roles8: retinaface = head buffers of stride 32, 16, 8; arcface = embedding buffer;
 * openpose = maps buffer, PAF channel offset, heat-map channel offset; unused = -1. */
int tr_program_info(const tr_program* program, int* n_buffers, int* n_ops, size_t* blob_bytes,
                    int32_t* roles8);
int tr_program_copy(const tr_program* program, tr_buffer_desc* buffers, tr_op_desc* ops, void* blob);
int tr_net_create_from_program(const tr_program* program, tr_net** out);

/* Single-call models: tr_*_create builds program + net from the checkpoint blob; tr_*_forward
 * runs one batch on device-resident inputs and leaves the results on the device.  Scratch
 * buffers live in the handle (grow-only).
 *
 * RetinaFace.call (retinaface/wrapper.py:133-238) on frames already resized by the caller:
 * frames (N,H,W,3) uint8 RGB; count (N,) int32; det (N,max_det,16) float32 rows
 * [score, x1,y1,x2,y2, 10 landmark coords, anchor index bits] in score order after NMS. */
typedef struct tr_model tr_model;
int tr_retinaface_create(const void* state_dict_blob, size_t bytes, tr_model** out);
int tr_retinaface_forward(tr_model* model, const uint8_t* frames_dev, int N, int H, int W,
                          float threshold, double nms_threshold, int max_det, int32_t* count_dev,
                          float* det_dev, void* stream);
/* FaceResNet100.forward + L2 normalise (arcface/model.py:87-97, wrapper.py:174-176): crops
 * (N,112,112,3) uint8 RGB (layout 0) or (N,3,112,112) uint8 BGR as the reference feeds its model
 * (layout 1); emb (N,512) float32. */
int tr_arcface_create(const void* state_dict_blob, size_t bytes, tr_model** out);
int tr_arcface_forward(tr_model* model, const uint8_t* crops_dev, int N, int layout, int normalise,
                       float* emb_dev, void* stream);
/* OpenPose.call after the resize (openpose/wrapper.py:207-485): frames (N,H,W,3) uint8 RGB at
 * network resolution, scale = short_side / min(original H, W); outputs as tr_openpose_parse. */
int tr_openpose_create(const void* state_dict_blob, size_t bytes, tr_model** out);
int tr_openpose_forward(tr_model* model, const uint8_t* frames_dev, int N, int H, int W, double scale,
                        int32_t* count_dev, int32_t* keypoints_dev, double* score_dev,
                        int32_t* status_dev, void* stream);
tr_net* tr_model_net(tr_model* model);
void tr_model_destroy(tr_model* model);

/* ---- single ops (parity tests, roofline micro-benchmarks) ---------------- */
/* NHWC fp16 convolution with fused scale/shift/activation/residual epilogue.
 * use_tc: 1 = tcgen05 implicit GEMM, 0 = CUDA-core direct kernel, 2 = warp-level mma.sync
 * kernel (small-channel 1x1 / 3x3 stride-1 layers only). */
int tr_conv2d(const void* in_dev, int N, int H, int W, int in_cs, int in_coff, int cin_pad,
              const void* w_dev, const float* scale_dev, const float* shift_dev,
              const float* slope_dev, int cout_pad, int cout_store, int k, int stride, int pad,
              int act, const void* res_dev, int res_cs, int res_up2, void* out_dev, int out_cs,
              int out_coff, int out_is_f32, int use_tc, int repeat, float* ms, void* stream);

/* Fused depthwise 3x3 + BN + ReLU -> 1x1 conv + scale/shift (+act): the single-op entry of
 * TR_OP_SEPCONV.  fused = 1: one mma.sync kernel; fused = 0: the depthwise kernel into
 * tmp_dev ((N,Ho,Wo,cin_pad) fp16) followed by the direct 1x1 kernel (cross-check). */
int tr_sepconv2d(const void* in_dev, int N, int H, int W, int in_cs, int in_coff, int cin_pad,
                 const float* dw_w_dev, const void* dw_w16_dev, const float* dw_scale_dev,
                 const float* dw_shift_dev, int stride, const void* w_dev, const float* scale_dev, const float* shift_dev,
                 int cout_pad, int cout_store, int act, void* out_dev, int out_cs, int out_coff,
                 void* tmp_dev, int fused, int repeat, float* ms, void* stream);

/* ---- RetinaFace post-processing ------------------------------------------
 * Replaces anchors_plane / decode_bboxes / decode_landmarks / threshold /
 * argsort / torchvision.ops.nms of retinaface/wrapper.py:154-236.
 * heads (fused == 0): the 9 reference outputs, order s32,s16,s8 x (prob NCHW (N,4,h,w)
 * soft-maxed, bbox (N,8,h,w), landmark (N,20,h,w)), fp32 device pointers.
 * Output rows: [score, x1,y1,x2,y2, 10 landmark coords, anchor index (int bits)]. */
size_t tr_detect_workspace_bytes(int N, int H, int W);
int tr_retinaface_decode_nms(const float* const* heads9_dev, int N, int H, int W, float threshold,
                             double nms_threshold, int max_det, void* workspace_dev,
                             int32_t* out_count_dev, int32_t* out_candidates_dev,
                             float* out_det_dev, void* stream);
/* Same on the fused fp32 NHWC head buffers (buffer ids for stride 32,16,8) of a net run. */
int tr_retinaface_detect(tr_net* net, const int* head_buffers3, float threshold,
                         double nms_threshold, int max_det, void* workspace_dev,
                         int32_t* out_count_dev, int32_t* out_candidates_dev, float* out_det_dev,
                         void* stream);

/* ---- ArcFace ---------------------------------------------------------------
 * Row-wise x / ||x||_2 with zero rows divided by 1: replaces
 * sklearn.preprocessing.normalize(axis=1) of arcface/wrapper.py:174-176. */
int tr_l2_normalize(const float* in_dev, float* out_dev, int N, int D, void* stream);

/* Inverse-affine bilinear warp of F faces to (F,3,side,side) uint8 BGR crops, bit-exact
 * with PIL Image.transform(AFFINE, BILINEAR, fillcolor=0): replaces the per-face host warp of
 * arcface/wrapper.py:22-72.  frames: (N,H,W,3) uint8 RGB; coef: F x 6 doubles (the PIL AFFINE
 * data = first two rows of the inverse similarity); image_index: frame of each face. */
int tr_face_align(const uint8_t* frames_dev, int H, int W, const double* coef_dev,
                  const int32_t* image_index_dev, int F, uint8_t* out_dev, int side, void* stream);

/* Faces given WITHOUT landmarks: replaces preprocess_face_no_landmarks of arcface/wrapper.py:75-99
 * (PIL Image.resize to longer side `side` — Pillow's antialiased BICUBIC for RGB — centred on a
 * zero canvas, channels BGR), bit-exact with Pillow's 8-bit resampler.  n RGB HWC uint8 images of
 * any size lie back to back in pixels_dev; offsets_host[i] is the byte offset of image i and
 * sizes_host[2i], sizes_host[2i+1] its height and width (both arrays on the HOST: the fixed-point
 * filter tables are built on the host from the sizes).  out_dev: (n,3,side,side) uint8.
 * workspace_dev: tr_face_letterbox_workspace_bytes(sizes_host, n, side) bytes (0 = bad sizes;
 * tr_last_error says why).  An image so thin that its resized shorter side is 0 pixels is an
 * error, as in Pillow ("height and width must be > 0"). */
size_t tr_face_letterbox_workspace_bytes(const int32_t* sizes_host, int n, int side);
int tr_face_letterbox(const uint8_t* pixels_dev, const int64_t* offsets_host, const int32_t* sizes_host,
                      int n, int side, void* workspace_dev, uint8_t* out_dev, void* stream);
/* Host-only helper (no GPU needed): Pillow's bicubic filter table of one axis, as the kernels
 * use it.  bounds_host: out_size x (first input sample, count); coeffs_host: out_size x ksize
 * int32 weights with 22 fractional bits.  Returns ksize (also when both pointers are NULL, to
 * size the buffers), -1 on error. */
int tr_resample_table(int in_size, int out_size, int32_t* bounds_host, int32_t* coeffs_host);

/* The 2x3 inverse similarity (PIL AFFINE data) of every detected face, computed on the device
 * from the detection rows of tr_retinaface_detect (so detect -> align -> embed needs no host
 * round trip for landmarks): landmarks are mapped back to frame pixels as Detection.resize_out
 * does (value / scale, round half to even), then fitted to the 5-point template in closed form
 * — replaces skimage SimilarityTransform.estimate + np.linalg.inv of arcface/wrapper.py:47-61.
 * det: (N, max_det, 16) rows, count: (N,) survivors per frame.  Faces are numbered frame by
 * frame in row order; coef: cap x 6 doubles, image_index: cap, *total: number of faces. */
int tr_face_similarity(const float* det_dev, const int32_t* count_dev, int N, int max_det, float scale,
                       int cap, double* coef_dev, int32_t* image_index_dev, int32_t* total_dev,
                       void* stream);

/* ---- OpenPose parse ---------------------------------------------------------
 * Replaces openpose/wrapper.py:214-483: x8 bicubic up-sampling (fused, never
 * materialised), peak extraction, PAF line integrals, greedy limb matching,
 * human assembly, keypoint rescale.  paf (N,38,h,w), heat (N,19,h,w) fp32 NCHW.
 * keypoints: [N][TR_HUMAN_CAP][18][3] int32 (x, y, present); score [N][TR_HUMAN_CAP] f64;
 * status bits per frame: 1 peak cap, 2 candidate cap, 4 human cap exceeded. */
enum { TR_PEAK_CAP = 512, TR_CAND_CAP = 4096, TR_HUMAN_CAP = 128 };
size_t tr_pose_workspace_bytes(int N);
int tr_openpose_parse(const float* paf_dev, const float* heat_dev, int N, int h, int w, double scale,
                      void* workspace_dev, int32_t* out_count_dev, int32_t* out_keypoints_dev,
                      double* out_score_dev, int32_t* out_status_dev, void* stream);
void tr_bicubic_table(float out32[32]);

/* ---- frame pre-processing ---------------------------------------------------
 * cv2.resize(INTER_LINEAR) on uint8 NHWC, bit-exact: replaces the host resize of
 * face/detection/__init__.py:15-57 and pose/openpose/wrapper.py:93-113. */
int tr_resize_bilinear_u8(const uint8_t* src_dev, int N, int H, int W, uint8_t* dst_dev, int h,
                          int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif
