// Layer-program executor and the C ABI (include/terran_b200.h).
//
// A net is built once from a Python-side description (terran_b200/weights.py);
// per input shape the executor infers every buffer's dims, allocates the
// activation buffers, builds the TMA tensor maps / launch geometry of each op
// (a "plan", cached by (N,H,W)) and then a run is a plain sequence of kernel
// launches on the caller's stream.
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/terran_b200.h"
#include "detect_post.cuh"
#include "program.h"

namespace trb {

namespace {

thread_local std::string g_last_error;

struct Buf {
  void* ptr = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  bool f32 = false;
  bool alias = false;
  bool sized = false;
};

struct PreparedOp {
  tr_op_desc d;
  double flops = 0;          // algorithmic (un-padded) flops of this op
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ConvArgs conv;
  ConvTcPlan* tc = nullptr;
  ConvPatchPlan* pt = nullptr;   // TR_OP_CONV on the resident-patch tcgen05 kernel
  // grouped conv on the other kernels: one launch per group
  std::vector<ConvArgs> gconv;
  std::vector<ConvTcPlan*> gtc;
  StemArgs stem;
  DwArgs dw;
  SepArgs sep;
  bool mma = false;          // TR_OP_CONV on the warp-level mma.sync kernel
  void* sep_tmp = nullptr;   // force_direct: depthwise output between the two unfused kernels
  View vin, vout;
  bool skip = false;        // fused into the previous op (2x2 max-pool in the conv epilogue): no launch
};

constexpr size_t kMaxPlans = 8;      // cached (N, H, W) plans per net

struct Plan {
  std::vector<Buf> bufs;
  std::vector<PreparedOp> ops;
  double tc_flops = 0;
  int tc_launches = 0, launches = 0;
  int N = 0, H = 0, W = 0;
  unsigned long long last_use = 0;   // tr_net::use_clock of the last run (LRU eviction)
  ~Plan() {
    for (auto& o : ops) {
      if (o.tc) conv_tc_plan_destroy(o.tc);
      if (o.pt) conv_patch_plan_destroy(o.pt);
      for (auto* g : o.gtc)
        if (g) conv_tc_plan_destroy(g);
      if (o.sep_tmp) cudaFree(o.sep_tmp);
      if (o.e0) cudaEventDestroy(o.e0);
      if (o.e1) cudaEventDestroy(o.e1);
    }
    for (auto& b : bufs)
      if (b.ptr && !b.alias) cudaFree(b.ptr);
  }
};

}  // namespace

}  // namespace trb

using namespace trb;

struct tr_net {
  std::vector<tr_buffer_desc> buffers;
  std::vector<tr_op_desc> ops;
  uint8_t* weights = nullptr;   // device copy of the blob
  size_t weight_bytes = 0;
  int force_direct = 0;
  int profile = 0;
  int lanes = 1;                // 0: issue every op on the caller's stream
  // Side stream for ops of lane 1 (independent branches that fill each other's scheduling
  // tails) and the two events that order it against the caller's stream.
  void* sk_scratch[2] = {nullptr, nullptr};   // stream-K scratch, one per lane
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::map<std::tuple<int, int, int, int>, std::unique_ptr<Plan>> plans;
  unsigned long long use_clock = 0;
  Plan* last = nullptr;
  ~tr_net() {
    plans.clear();
    if (weights) cudaFree(weights);
    for (void* s : sk_scratch)
      if (s) cudaFree(s);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
  }
};

namespace trb {
namespace {

template <class T>
const T* blob_ptr(const tr_net* net, int64_t off) {
  if (off < 0) return nullptr;
  TR_CHECK(size_t(off) < net->weight_bytes && off % 16 == 0, "bad blob offset");
  return reinterpret_cast<const T*>(net->weights + off);
}

// TRB_MMA=0 keeps TR_ENGINE_MMA / TR_OP_SEPCONV ops off the mma.sync kernels (A/B measurements).
bool mma_enabled() {
  static const bool on = [] { const char* e = getenv("TRB_MMA"); return !e || atoi(e) != 0; }();
  return on;
}

// TRB_PATCH: 0 = never, 1 = auto (layers where the resident-patch kernel measured faster,
// profiles/r02_patch_vs_plain.txt), 2 = every eligible layer.
int patch_mode() {
  static const int mode = [] { const char* e = getenv("TRB_PATCH"); return e ? atoi(e) : 1; }();
  return mode;
}

bool patch_wanted(const ConvArgs& a) {
  const int mode = patch_mode();
  if (!mode || !conv_patch_eligible(a)) return false;
  if (mode >= 2) return true;
  // (64-filter layers run, but half of every 128-row MMA is padding: measured 25-60 % slower
  // than conv_tc_kernel's pixels-on-M tiling for them)
  if (a.cout_pad % 128) return false;
  // Measured faster than conv_tc_kernel on every eligible OpenPose layer (1x1, 3x3 and 7x7,
  // 64..512 input channels) once the map is large enough for the 8 x R tiles to fill their
  // MMA columns; small maps (ArcFace 14x14 / 7x7) waste too many of them.
  const int H = a.H_out, W = a.W_out;
  auto fill8 = [](int n) { return double(n) / (8.0 * ((n + 7) / 8)); };
  auto fill_r = [](int n) {            // best fill with R in [20, 32] rows per tile (N = 160..256)
    double best = 0;
    for (int r = 20; r <= 32; r += 2) best = std::max(best, double(n) / (double(r) * ((n + r - 1) / r)));
    return best;
  };
  const double best = std::max(fill8(W) * fill_r(H), fill8(H) * fill_r(W));
  return best >= 0.85;
}

// Stream-K scratch of one lane (stream), shared by all its conv plans: sized for either kernel.
void* lane_scratch(tr_net* net, bool side) {
  void*& scratch = net->sk_scratch[side];
  if (!scratch) {
    const size_t bytes = std::max(conv_tc_sk_scratch_bytes(), conv_patch_scratch_bytes());
    TR_CUDA(cudaMalloc(&scratch, bytes));
    TR_CUDA(cudaMemset(scratch, 0, bytes));
  }
  return scratch;
}

View make_view(const Buf& b, int coff, int C) {
  View v;
  v.ptr = static_cast<__half*>(b.ptr);
  v.N = b.N; v.H = b.H; v.W = b.W; v.cs = b.C; v.coff = coff; v.C = C;
  return v;
}

void set_dims(Buf& b, int N, int H, int W, const char* what) {
  if (b.sized) {
    TR_CHECK(b.N == N && b.H == H && b.W == W,
             std::string("buffer written with inconsistent dims by ") + what);
    return;
  }
  b.N = N; b.H = H; b.W = W; b.sized = true;
}

Plan* build_plan(tr_net* net, int N, int H, int W) {
  auto plan = std::make_unique<Plan>();
  plan->N = N; plan->H = H; plan->W = W;
  plan->bufs.resize(net->buffers.size());
  for (size_t i = 0; i < net->buffers.size(); ++i) {
    plan->bufs[i].C = net->buffers[i].channels;
    plan->bufs[i].f32 = net->buffers[i].is_f32 != 0;
    TR_CHECK(plan->bufs[i].C % 8 == 0, "buffer channels must be a multiple of 8");
  }
  auto& B = plan->bufs;
  // ---- pass 1: shape inference
  for (const tr_op_desc& d : net->ops) {
    int iH, iW, iN;
    if (d.in < 0) { iN = N; iH = H; iW = W; }
    else {
      TR_CHECK(B[d.in].sized, "op reads a buffer no earlier op wrote");
      iN = B[d.in].N; iH = B[d.in].H; iW = B[d.in].W;
    }
    switch (d.type) {
      case TR_OP_STEM:
      case TR_OP_CONV:
      case TR_OP_SEPCONV:
      case TR_OP_DWCONV: {
        const int oH = (iH + 2 * d.pad - d.k) / d.stride + 1, oW = (iW + 2 * d.pad - d.k) / d.stride + 1;
        TR_CHECK(oH > 0 && oW > 0, "image too small for the network");
        set_dims(B[d.out], iN, oH, oW, "conv");
        if (d.out2 >= 0) set_dims(B[d.out2], iN, oH, oW, "conv out2");
        break;
      }
      case TR_OP_MAXPOOL:
        TR_CHECK(iH >= 2 && iW >= 2, "image too small for the network");
        set_dims(B[d.out], iN, iH / 2, iW / 2, "maxpool");
        break;
      case TR_OP_COPY:
        set_dims(B[d.out], iN, iH, iW, "copy");
        break;
      case TR_OP_VIEW: {
        TR_CHECK(B[d.out].C == iH * iW * B[d.in].C, "flatten view channel mismatch");
        set_dims(B[d.out], iN, 1, 1, "view");
        B[d.out].alias = true;
        break;
      }
      default: fail("unknown op type");
    }
  }
  // ---- pass 2: allocate
  for (auto& b : B) {
    if (!b.sized || b.alias) continue;
    const size_t bytes = size_t(b.N) * b.H * b.W * b.C * (b.f32 ? 4 : 2);
    TR_CUDA(cudaMalloc(&b.ptr, bytes + 256));
    // Padding channels (and slices no op writes) must read as zero.
    TR_CUDA(cudaMemset(b.ptr, 0, bytes + 256));
  }
  for (const tr_op_desc& d : net->ops)
    if (d.type == TR_OP_VIEW) B[d.out].ptr = B[d.in].ptr;
  // ---- pass 3: prepare launches
  // A 2x2 max-pool whose input is written by the op right before it and read by nothing else
  // (VGG: conv -> ReLU -> pool) is fused into that conv's epilogue when the resident-patch
  // kernel runs it (TRB_POOL_FUSE=0 keeps the separate kernel).
  static const bool pool_fuse = [] {
    const char* e = getenv("TRB_POOL_FUSE");
    return (!e || atoi(e) != 0) && !getenv("TRB_PT_DEBUG") && !getenv("TRB_PT_GENERIC") && !getenv("TRB_PT_TMA_STORE");
  }();
  auto sole_consumer_is_next_pool = [&](size_t i) {
    if (i + 1 >= net->ops.size()) return false;
    const tr_op_desc &c = net->ops[i], &m = net->ops[i + 1];
    if (m.type != TR_OP_MAXPOOL || m.in != c.out || m.in_coff != c.out_coff || m.in_c != c.out_c) return false;
    for (size_t j = 0; j < net->ops.size(); ++j) {
      if (j == i || j == i + 1) continue;
      const tr_op_desc& o = net->ops[j];
      if (o.in == c.out || o.res == c.out || o.out == c.out || o.out2 == c.out) return false;
    }
    return true;
  };
  bool skip_next = false;
  for (size_t op_i = 0; op_i < net->ops.size(); ++op_i) {
    const tr_op_desc& d = net->ops[op_i];
    PreparedOp po{};
    po.d = d;
    po.skip = skip_next;
    skip_next = false;
    switch (d.type) {
      case TR_OP_STEM: {
        StemArgs& a = po.stem;
        a = StemArgs{};
        a.N = N; a.H = H; a.W = W;
        a.in_scale = d.in_scale; a.in_shift = d.in_shift;
        a.out = make_view(B[d.out], d.out_coff, d.out_c);
        if (d.out2 >= 0) a.out2 = make_view(B[d.out2], d.out2_coff, d.out_c);
        a.w = blob_ptr<float>(net, d.w_off);
        a.scale = blob_ptr<float>(net, d.scale_off); a.shift = blob_ptr<float>(net, d.shift_off);
        a.slope = blob_ptr<float>(net, d.slope_off);
        a.scale2 = blob_ptr<float>(net, d.scale2_off); a.shift2 = blob_ptr<float>(net, d.shift2_off);
        a.cout = d.out_c; a.stride = d.stride; a.act = d.act;
        a.use_mma = !net->force_direct && !d.force_direct;
        TR_CHECK(d.k == 3 && d.pad == 1, "stem is 3x3 pad 1");
        TR_CHECK(!B[d.out].f32, "stem output is fp16");
        break;
      }
      case TR_OP_CONV: {
        ConvArgs& a = po.conv;
        a = ConvArgs{};
        a.in = make_view(B[d.in], d.in_coff, d.in_c);
        a.out = make_view(B[d.out], d.out_coff, d.out_c);
        if (B[d.out].f32) a.out_f32 = static_cast<float*>(B[d.out].ptr);
        if (d.out2 >= 0) a.out2 = make_view(B[d.out2], d.out2_coff, d.out_c);
        if (d.res >= 0) {
          a.res = make_view(B[d.res], d.res_coff, d.out_c);
          a.res_up2 = d.res_up2;
          if (d.res_up2)
            TR_CHECK((B[d.out].H + 1) / 2 <= B[d.res].H && (B[d.out].W + 1) / 2 <= B[d.res].W,
                     "upsampled residual too small");
          else
            TR_CHECK(B[d.res].H == B[d.out].H && B[d.res].W == B[d.out].W, "residual dims");
        }
        a.w = blob_ptr<__half>(net, d.w_off);
        a.scale = blob_ptr<float>(net, d.scale_off); a.shift = blob_ptr<float>(net, d.shift_off);
        a.slope = blob_ptr<float>(net, d.slope_off);
        a.scale2 = blob_ptr<float>(net, d.scale2_off); a.shift2 = blob_ptr<float>(net, d.shift2_off);
        a.shift9 = blob_ptr<float>(net, d.shift9_off);
        TR_CHECK(!a.shift9 || (d.k == 3 && d.pad == 1 && d.stride == 1), "border-class shifts are for 3x3 pad-1 stride-1 convs");
        a.cout_pad = d.cout_pad; a.cout_store = d.out_c; a.cin_pad = d.in_c;
        a.kh = a.kw = d.k; a.stride = d.stride; a.pad = d.pad; a.act = d.act;
        a.H_out = B[d.out].H; a.W_out = B[d.out].W;
        TR_CHECK(d.in_coff + d.in_c <= B[d.in].C && d.out_coff + d.out_c <= B[d.out].C,
                 "channel slice out of range");
        po.flops = 2.0 * B[d.out].N * a.H_out * a.W_out * double(d.cout_real) * d.k * d.k * d.cin_real;
        const int G = d.groups > 1 ? d.groups : 1;
        a.groups = G;
        if (G > 1) {
          TR_CHECK(d.cout_pad % G == 0 && d.out_c % G == 0 && d.res < 0 && d.out2 < 0 && !B[d.out].f32,
                   "grouped conv: equal filter / output blocks, plain fp16 epilogue");
          TR_CHECK(d.in_coff + G * d.in_c <= B[d.in].C, "grouped conv: input slice out of range");
        }
        po.mma = G == 1 && !net->force_direct && !d.force_direct && d.engine == TR_ENGINE_MMA && mma_enabled() &&
                 !a.shift9 && conv_mma_eligible(a);
        bool use_tc = !po.mma && !net->force_direct && !d.force_direct;
        if (use_tc && G == 1) use_tc = conv_tc_eligible(a);
        // A fully-connected layer (1x1 conv on 1x1 maps: ArcFace's 25088 -> 512) has one pixel
        // per image — nothing for an 8 x R pixel tile.  The batch IS a pixel axis, though: NHWC
        // rows of a (N,1,1,C) tensor are the pixels of a (1, N/8, 8, C) map, bit for bit, so the
        // resident-patch kernel takes it as one image and splits the long K over all SMs
        // (stream-K): 8 tiles x 392 channel chunks instead of 8 CTAs grinding through K alone.
        ConvArgs fc = a;
        const bool as_fc = use_tc && G == 1 && d.k == 1 && a.in.H == 1 && a.in.W == 1 && a.in.N % 8 == 0 &&
                           a.in.N >= 64 && a.cin_pad >= 1024 && !a.res.ptr && patch_mode() != 0;
        if (as_fc) {
          fc.in.H = fc.out.H = a.in.N / 8; fc.in.W = fc.out.W = 8; fc.in.N = fc.out.N = 1;
          fc.H_out = fc.in.H; fc.W_out = 8;
          fc.patch_rows = 8;        // small tiles: more of them, so each is split over fewer CTAs
        }
        if (as_fc && conv_patch_eligible(fc)) {
          fc.sk_scratch = lane_scratch(net, d.lane == 1);
          po.pt = conv_patch_plan_create(fc, conv_tc_error_flag());
          plan->tc_flops += po.flops;
          plan->tc_launches++;
        } else if (use_tc && patch_wanted(a)) {
          if (pool_fuse && G == 1 && d.act == TR_ACT_RELU && d.res < 0 && d.out2 < 0 && !a.shift9 && !a.out_f32 &&
              d.out_c % 128 == 0 && sole_consumer_is_next_pool(op_i)) {
            const tr_op_desc& m = net->ops[op_i + 1];
            a.pool2 = 1;
            a.out = make_view(B[m.out], m.out_coff, m.out_c);      // the POOLED tensor
            skip_next = true;
          }
          a.sk_scratch = lane_scratch(net, d.lane == 1);
          po.pt = conv_patch_plan_create(a, conv_tc_error_flag());
          plan->tc_flops += po.flops;
          plan->tc_launches++;
        } else if (G > 1) {
          // the other kernels take one group at a time: block g of the filters / outputs,
          // input channels in_coff + g * in_c
          for (int g = 0; g < G; ++g) {
            ConvArgs c = a;
            c.groups = 1;
            c.in.coff = d.in_coff + g * d.in_c;
            c.out.coff = d.out_coff + g * (d.out_c / G); c.out.C = d.out_c / G;
            c.cout_pad = d.cout_pad / G; c.cout_store = d.out_c / G;
            c.w = a.w + size_t(g) * c.cout_pad * d.k * d.k * d.in_c;
            c.scale = a.scale + g * c.cout_pad; c.shift = a.shift + g * c.cout_pad;
            if (a.slope) c.slope = a.slope + g * c.cout_pad;
            ConvTcPlan* tcp = nullptr;
            if (use_tc && conv_tc_eligible(c)) {
              c.sk_scratch = lane_scratch(net, d.lane == 1);
              tcp = conv_tc_plan_create(c);
              plan->tc_launches++;
            }
            po.gconv.push_back(c);
            po.gtc.push_back(tcp);
          }
          if (use_tc) plan->tc_flops += po.flops;
        } else if (use_tc) {
          a.sk_scratch = lane_scratch(net, d.lane == 1);
          if (pool_fuse && d.act == TR_ACT_RELU && d.res < 0 && d.out2 < 0 && !a.shift9 && !a.out_f32 &&
              sole_consumer_is_next_pool(op_i)) {
            const tr_op_desc& m = net->ops[op_i + 1];
            a.pool_out = make_view(B[m.out], m.out_coff, m.out_c);
          }
          po.tc = conv_tc_plan_create(a);
          if (a.pool_out.ptr && conv_tc_plan_pooled(po.tc)) skip_next = true;
          plan->tc_flops += po.flops;
          plan->tc_launches++;
        }
        break;
      }
      case TR_OP_SEPCONV: {
        SepArgs& a = po.sep;
        a = SepArgs{};
        a.in = make_view(B[d.in], d.in_coff, d.in_c);
        a.out = make_view(B[d.out], d.out_coff, d.out_c);
        a.dw_w = blob_ptr<float>(net, d.dw_w_off);
        a.dw_w16 = blob_ptr<__half>(net, d.dw_w16_off);
        a.dw_scale = blob_ptr<float>(net, d.dw_scale_off); a.dw_shift = blob_ptr<float>(net, d.dw_shift_off);
        a.stride = d.stride;
        a.w = blob_ptr<__half>(net, d.w_off);
        a.scale = blob_ptr<float>(net, d.scale_off); a.shift = blob_ptr<float>(net, d.shift_off);
        a.cin_pad = d.in_c; a.cout_pad = d.cout_pad; a.cout_store = d.out_c; a.act = d.act;
        TR_CHECK(d.k == 3 && d.pad == 1, "fused depthwise stage is 3x3 pad 1");
        TR_CHECK(!B[d.out].f32, "sepconv output is fp16");
        TR_CHECK(d.in_coff + d.in_c <= B[d.in].C && d.out_coff + d.out_c <= B[d.out].C,
                 "channel slice out of range");
        po.flops = 2.0 * B[d.out].N * B[d.out].H * B[d.out].W * (double(d.cout_real) * d.cin_real + 9.0 * d.cin_real);
        po.mma = !net->force_direct && !d.force_direct && mma_enabled();
        if (po.mma) TR_CHECK(sep_mma_eligible(a), "sepconv: unsupported (channels, stride) combination");
        if (!po.mma) {
          // Cross-check path: the depthwise kernel into a scratch tensor, then the direct 1x1.
          const size_t bytes = size_t(B[d.out].N) * B[d.out].H * B[d.out].W * d.in_c * 2;
          TR_CUDA(cudaMalloc(&po.sep_tmp, bytes + 256));
          DwArgs& w = po.dw;
          w.in = a.in;
          w.out = View{static_cast<__half*>(po.sep_tmp), B[d.out].N, B[d.out].H, B[d.out].W, d.in_c, 0, d.in_c};
          w.w = a.dw_w; w.scale = a.dw_scale; w.shift = a.dw_shift; w.stride = d.stride;
          ConvArgs& c = po.conv;
          c = ConvArgs{};
          c.in = w.out; c.out = a.out;
          c.w = a.w; c.scale = a.scale; c.shift = a.shift;
          c.cout_pad = d.cout_pad; c.cout_store = d.out_c; c.cin_pad = d.in_c;
          c.act = d.act; c.H_out = B[d.out].H; c.W_out = B[d.out].W;
        }
        break;
      }
      case TR_OP_DWCONV: {
        DwArgs& a = po.dw;
        a.in = make_view(B[d.in], d.in_coff, d.in_c);
        a.out = make_view(B[d.out], d.out_coff, d.out_c);
        a.w = blob_ptr<float>(net, d.w_off);
        a.scale = blob_ptr<float>(net, d.scale_off); a.shift = blob_ptr<float>(net, d.shift_off);
        a.stride = d.stride;
        TR_CHECK(d.k == 3 && d.pad == 1 && d.act == TR_ACT_RELU, "depthwise op is 3x3 pad 1 + ReLU");
        break;
      }
      case TR_OP_MAXPOOL:
      case TR_OP_COPY:
        po.vin = make_view(B[d.in], d.in_coff, d.in_c);
        po.vout = make_view(B[d.out], d.out_coff, d.out_c);
        break;
      case TR_OP_VIEW:
        break;
    }
    if (d.type != TR_OP_VIEW && !po.skip)
      plan->launches += (d.type == TR_OP_SEPCONV && !po.mma) ? 2 : (po.gconv.empty() ? 1 : int(po.gconv.size()));
    plan->ops.push_back(po);
  }
  Plan* raw = plan.get();
  // Bounded cache, least-recently-used eviction ONE plan at a time (lists of differently sized
  // images produce a new shape per call: clearing everything would re-allocate every buffer
  // of every cached shape).  The plan that ran last (net->last) is never the victim.
  raw->last_use = ++net->use_clock;
  while (net->plans.size() >= kMaxPlans) {
    auto victim = net->plans.end();
    for (auto it = net->plans.begin(); it != net->plans.end(); ++it)
      if (it->second.get() != net->last && (victim == net->plans.end() || it->second->last_use < victim->second->last_use))
        victim = it;
    if (victim == net->plans.end()) break;
    net->plans.erase(victim);
  }
  net->plans[std::make_tuple(N, H, W, net->force_direct)] = std::move(plan);
  return raw;
}

void run_plan(tr_net* net, Plan* plan, const uint8_t* image, int64_t sn, int64_t sh, int64_t sw,
              int64_t sc, cudaStream_t main_stream, bool profile) {
  // Lanes: ops of lane 1 go to the net's side stream.  A FORK op marks the point of the main
  // stream the side stream has to reach first; a JOIN op (and the end of the program) waits
  // for the side stream.  Per-op profiling keeps everything on one stream.
  // NVTX: one range per net run, one per op (visible in nsys / ncu --nvtx; a no-op otherwise)
  struct Range {
    explicit Range(const char* n) { nvtxRangePushA(n); }
    ~Range() { nvtxRangePop(); }
  };
  static const char* kOpNames[] = {"tr:stem", "tr:conv", "tr:dwconv", "tr:maxpool", "tr:copy", "tr:view", "tr:sepconv"};
  Range run_range("tr_net_run");
  const bool lanes = net->lanes && !profile;
  bool fork_pending = false, side_dirty = false;
  auto ensure_side = [&] {
    if (net->side) return;
    TR_CUDA(cudaStreamCreateWithFlags(&net->side, cudaStreamNonBlocking));
    TR_CUDA(cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming));
    TR_CUDA(cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming));
  };
  auto join = [&] {
    if (!side_dirty) return;
    TR_CUDA(cudaEventRecord(net->ev_join, net->side));
    TR_CUDA(cudaStreamWaitEvent(main_stream, net->ev_join, 0));
    side_dirty = false;
  };
  for (auto& po : plan->ops) {
    cudaStream_t s = main_stream;
    if (lanes) {
      if (po.d.sync & TR_SYNC_JOIN) join();
      if (po.d.lane == 1) {
        ensure_side();
        if (!fork_pending && !side_dirty) {      // no explicit fork point: everything issued so far
          TR_CUDA(cudaEventRecord(net->ev_fork, main_stream));
          fork_pending = true;
        }
        if (fork_pending) {
          TR_CUDA(cudaStreamWaitEvent(net->side, net->ev_fork, 0));
          fork_pending = false;
        }
        s = net->side;
        side_dirty = true;
      }
    }
    Range op_range(po.pt ? "tr:conv(patch)" : (po.tc ? "tr:conv(tcgen05)" : kOpNames[po.d.type]));
    if (profile && po.d.type != TR_OP_VIEW) {
      if (!po.e0) { TR_CUDA(cudaEventCreate(&po.e0)); TR_CUDA(cudaEventCreate(&po.e1)); }
      TR_CUDA(cudaEventRecord(po.e0, s));
    }
    switch (po.d.type) {
      case TR_OP_STEM: {
        StemArgs a = po.stem;
        a.in = image; a.sn = sn; a.sh = sh; a.sw = sw; a.sc = sc;
        stem_launch(a, s);
        break;
      }
      case TR_OP_CONV:
        if (po.pt) conv_patch_launch(po.pt, s);
        else if (!po.gconv.empty()) {
          for (size_t g = 0; g < po.gconv.size(); ++g) {
            if (po.gtc[g]) conv_tc_launch(po.gtc[g], s);
            else conv_direct_launch(po.gconv[g], s);
          }
        }
        else if (po.tc) conv_tc_launch(po.tc, s);
        else if (po.mma) conv_mma_launch(po.conv, s);
        else conv_direct_launch(po.conv, s);
        break;
      case TR_OP_SEPCONV:
        if (po.mma) sep_mma_launch(po.sep, s);
        else { dwconv_launch(po.dw, s); conv_direct_launch(po.conv, s); }
        break;
      case TR_OP_DWCONV: dwconv_launch(po.dw, s); break;
      case TR_OP_MAXPOOL: if (!po.skip) maxpool2_launch(po.vin, po.vout, s); break;
      case TR_OP_COPY: copy_slice_launch(po.vin, po.vout, s); break;
      default: break;
    }
    if (profile && po.d.type != TR_OP_VIEW) TR_CUDA(cudaEventRecord(po.e1, s));
    if (lanes && (po.d.sync & TR_SYNC_FORK)) {
      ensure_side();
      TR_CUDA(cudaEventRecord(net->ev_fork, main_stream));
      fork_pending = true;
    }
  }
  join();
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  }
}

const float kAnchorLo[3][2] = {{-248.f, -120.f}, {-56.f, -24.f}, {-8.f, 0.f}};
const float kAnchorHi[3][2] = {{263.f, 135.f}, {71.f, 39.f}, {23.f, 15.f}};

void fill_anchor_refs(DetHeads& h) {
  for (int l = 0; l < 3; ++l)
    for (int a = 0; a < 2; ++a) { h.anchor_lo[l][a] = kAnchorLo[l][a]; h.anchor_hi[l][a] = kAnchorHi[l][a]; }
}

}  // namespace
}  // namespace trb

extern "C" {

int tr_version(void) { return 100; }

const char* tr_last_error(void) { return g_last_error.c_str(); }

int tr_init(int device) {
  return guarded([&] {
    TR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TR_CUDA(cudaGetDeviceProperties(&prop, device));
    TR_CHECK(prop.major == 10, "terran_b200 needs an sm_100 (Blackwell B200) device, found sm_" +
                                   std::to_string(prop.major) + std::to_string(prop.minor));
  });
}

int tr_net_create(const tr_buffer_desc* buffers, int n_buffers, const tr_op_desc* ops, int n_ops,
                  const void* weights_host, size_t weight_bytes, tr_net** out) {
  return guarded([&] {
    auto net = std::make_unique<tr_net>();
    net->buffers.assign(buffers, buffers + n_buffers);
    net->ops.assign(ops, ops + n_ops);
    for (const auto& d : net->ops) {
      TR_CHECK(d.out >= 0 && d.out < n_buffers && d.in < n_buffers, "op buffer id out of range");
      TR_CHECK(d.out2 < n_buffers && d.res < n_buffers, "op buffer id out of range");
    }
    net->weight_bytes = weight_bytes;
    TR_CUDA(cudaMalloc(&net->weights, weight_bytes + 256));
    TR_CUDA(cudaMemcpy(net->weights, weights_host, weight_bytes, cudaMemcpyHostToDevice));
    if (const char* e = getenv("TRB_LANES")) net->lanes = atoi(e) != 0;
    *out = net.release();
  });
}

void tr_net_destroy(tr_net* net) { delete net; }

int tr_net_set_mode(tr_net* net, int force_direct) {
  return guarded([&] { net->force_direct = force_direct ? 1 : 0; });
}

int tr_net_run(tr_net* net, const uint8_t* image_dev, int N, int H, int W, int64_t stride_n,
               int64_t stride_h, int64_t stride_w, int64_t stride_c, void* stream) {
  return guarded([&] {
    TR_CHECK(N > 0 && H > 0 && W > 0, "empty batch");
    auto it = net->plans.find(std::make_tuple(N, H, W, net->force_direct));
    Plan* plan = it != net->plans.end() ? it->second.get() : build_plan(net, N, H, W);
    plan->last_use = ++net->use_clock;
    net->last = plan;
    run_plan(net, plan, image_dev, stride_n, stride_h, stride_w, stride_c,
             static_cast<cudaStream_t>(stream), net->profile != 0);
  });
}

int tr_net_buffer(tr_net* net, int buffer, void** ptr, int* N, int* H, int* W, int* channels) {
  return guarded([&] {
    TR_CHECK(net->last, "no run yet");
    TR_CHECK(buffer >= 0 && size_t(buffer) < net->last->bufs.size(), "buffer id");
    const Buf& b = net->last->bufs[buffer];
    *ptr = b.ptr; *N = b.N; *H = b.H; *W = b.W; *channels = b.C;
  });
}

int tr_net_export_nchw(tr_net* net, int buffer, int coff, int C, float* out_dev, void* stream) {
  return guarded([&] {
    TR_CHECK(net->last, "no run yet");
    const Buf& b = net->last->bufs.at(buffer);
    TR_CHECK(!b.f32 && coff + C <= b.C, "export slice");
    export_nchw_launch(make_view(b, coff, C), C, out_dev, static_cast<cudaStream_t>(stream));
  });
}

int tr_net_export_nchw_f32(tr_net* net, int buffer, int coff, int C, float* out_dev,
                           int softmax_pairs, void* stream) {
  return guarded([&] {
    TR_CHECK(net->last, "no run yet");
    const Buf& b = net->last->bufs.at(buffer);
    TR_CHECK(b.f32 && coff + C <= b.C, "export slice");
    TR_CHECK(!softmax_pairs || C == 4, "pair softmax is over the 4 class channels");
    export_nchw_f32_launch(static_cast<const float*>(b.ptr), b.N, b.H, b.W, b.C, coff, C, out_dev,
                           softmax_pairs, static_cast<cudaStream_t>(stream));
  });
}

int tr_net_stats(tr_net* net, double* tc_flops, int* tc_launches, int* total_launches) {
  return guarded([&] {
    TR_CHECK(net->last, "no run yet");
    *tc_flops = net->last->tc_flops;
    *tc_launches = net->last->tc_launches;
    *total_launches = net->last->launches;
  });
}

int tr_net_set_profile(tr_net* net, int enable) {
  return guarded([&] { net->profile = enable ? 1 : 0; });
}

int tr_net_profile(tr_net* net, float* ms, int32_t* is_tc, double* flops, int cap, int* n_ops) {
  return guarded([&] {
    TR_CHECK(net->last, "no run yet");
    int n = 0;
    for (auto& po : net->last->ops) {
      if (po.d.type == TR_OP_VIEW) continue;
      TR_CHECK(po.e0 && po.e1, "profiling was not enabled for the last run");
      TR_CUDA(cudaEventSynchronize(po.e1));
      if (n < cap) {
        TR_CUDA(cudaEventElapsedTime(ms + n, po.e0, po.e1));
        is_tc[n] = (po.tc || po.pt || (!po.gtc.empty() && po.gtc[0])) ? 1 : 0;
        flops[n] = po.flops;
      }
      ++n;
    }
    *n_ops = n;
  });
}

int tr_conv2d(const void* in_dev, int N, int H, int W, int in_cs, int in_coff, int cin_pad,
              const void* w_dev, const float* scale_dev, const float* shift_dev,
              const float* slope_dev, int cout_pad, int cout_store, int k, int stride, int pad,
              int act, const void* res_dev, int res_cs, int res_up2, void* out_dev, int out_cs,
              int out_coff, int out_is_f32, int use_tc, int repeat, float* ms, void* stream) {
  return guarded([&] {
    ConvArgs a{};
    a.in.ptr = static_cast<__half*>(const_cast<void*>(in_dev));
    a.in.N = N; a.in.H = H; a.in.W = W; a.in.cs = in_cs; a.in.coff = in_coff; a.in.C = cin_pad;
    a.H_out = (H + 2 * pad - k) / stride + 1;
    a.W_out = (W + 2 * pad - k) / stride + 1;
    a.out.ptr = static_cast<__half*>(out_dev);
    a.out.N = N; a.out.H = a.H_out; a.out.W = a.W_out; a.out.cs = out_cs; a.out.coff = out_coff;
    a.out.C = cout_store;
    if (out_is_f32) a.out_f32 = static_cast<float*>(out_dev);
    if (res_dev) {
      a.res.ptr = static_cast<__half*>(const_cast<void*>(res_dev));
      a.res.N = N; a.res.cs = res_cs; a.res.coff = 0; a.res.C = cout_store;
      a.res.H = res_up2 ? (a.H_out + 1) / 2 : a.H_out;
      a.res.W = res_up2 ? (a.W_out + 1) / 2 : a.W_out;
      a.res_up2 = res_up2;
    }
    a.w = static_cast<const __half*>(w_dev);
    a.scale = scale_dev; a.shift = shift_dev; a.slope = slope_dev;
    a.cout_pad = cout_pad; a.cout_store = cout_store; a.cin_pad = cin_pad;
    a.kh = a.kw = k; a.stride = stride; a.pad = pad; a.act = act;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ConvTcPlan* plan = nullptr;
    if (use_tc == 1) plan = conv_tc_plan_create(a);
    ConvPatchPlan* pplan = use_tc == 3 ? conv_patch_plan_create(a, conv_tc_error_flag()) : nullptr;
    if (use_tc == 2) TR_CHECK(conv_mma_eligible(a), "layer not supported by the mma.sync kernel");
    cudaEvent_t e0, e1;
    TR_CUDA(cudaEventCreate(&e0));
    TR_CUDA(cudaEventCreate(&e1));
    auto once = [&] {
      if (plan) conv_tc_launch(plan, s);
      else if (pplan) conv_patch_launch(pplan, s);
      else if (use_tc == 2) conv_mma_launch(a, s);
      else conv_direct_launch(a, s);
    };
    once();
    TR_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < repeat; ++i) once();
    TR_CUDA(cudaEventRecord(e1, s));
    cudaError_t err = cudaStreamSynchronize(s);
    if (plan) conv_tc_plan_destroy(plan);
    if (pplan) conv_patch_plan_destroy(pplan);
    if (err != cudaSuccess)
      fail(std::string("conv launch failed: ") + cudaGetErrorString(err) + " (pipeline timeout code " +
           std::to_string(conv_tc_last_timeout()) + ")");
    float t = 0.f;
    TR_CUDA(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = repeat > 0 ? t / repeat : 0.f;
  });
}

int tr_sepconv2d(const void* in_dev, int N, int H, int W, int in_cs, int in_coff, int cin_pad,
                 const float* dw_w_dev, const void* dw_w16_dev, const float* dw_scale_dev,
                 const float* dw_shift_dev, int stride, const void* w_dev, const float* scale_dev, const float* shift_dev,
                 int cout_pad, int cout_store, int act, void* out_dev, int out_cs, int out_coff,
                 void* tmp_dev, int fused, int repeat, float* ms, void* stream) {
  return guarded([&] {
    TR_CHECK(stride == 1 || stride == 2, "depthwise stride");
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    SepArgs a{};
    a.in = View{static_cast<__half*>(const_cast<void*>(in_dev)), N, H, W, in_cs, in_coff, cin_pad};
    a.out = View{static_cast<__half*>(out_dev), N, Ho, Wo, out_cs, out_coff, cout_store};
    a.dw_w = dw_w_dev; a.dw_w16 = static_cast<const __half*>(dw_w16_dev); a.dw_scale = dw_scale_dev; a.dw_shift = dw_shift_dev; a.stride = stride;
    a.w = static_cast<const __half*>(w_dev); a.scale = scale_dev; a.shift = shift_dev;
    a.cin_pad = cin_pad; a.cout_pad = cout_pad; a.cout_store = cout_store; a.act = act;
    DwArgs d{};
    ConvArgs c{};
    if (!fused) {
      TR_CHECK(tmp_dev, "unfused sepconv needs the scratch tensor");
      d.in = a.in;
      d.out = View{static_cast<__half*>(tmp_dev), N, Ho, Wo, cin_pad, 0, cin_pad};
      d.w = dw_w_dev; d.scale = dw_scale_dev; d.shift = dw_shift_dev; d.stride = stride;
      c.in = d.out; c.out = a.out; c.w = a.w; c.scale = scale_dev; c.shift = shift_dev;
      c.cout_pad = cout_pad; c.cout_store = cout_store; c.cin_pad = cin_pad; c.act = act;
      c.H_out = Ho; c.W_out = Wo;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto once = [&] {
      if (fused) sep_mma_launch(a, s);
      else { dwconv_launch(d, s); conv_direct_launch(c, s); }
    };
    cudaEvent_t e0, e1;
    TR_CUDA(cudaEventCreate(&e0));
    TR_CUDA(cudaEventCreate(&e1));
    once();
    TR_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < repeat; ++i) once();
    TR_CUDA(cudaEventRecord(e1, s));
    TR_CUDA(cudaStreamSynchronize(s));
    float t = 0.f;
    TR_CUDA(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = repeat > 0 ? t / repeat : 0.f;
  });
}

size_t tr_detect_workspace_bytes(int N, int H, int W) { return detect_workspace_bytes(N, H, W); }

int tr_retinaface_decode_nms(const float* const* heads9_dev, int N, int H, int W, float threshold,
                             double nms_threshold, int max_det, void* workspace_dev,
                             int32_t* out_count_dev, int32_t* out_candidates_dev,
                             float* out_det_dev, void* stream) {
  return guarded([&] {
    DetHeads h{};
    for (int l = 0; l < 3; ++l) {
      h.cls[l] = heads9_dev[3 * l]; h.bbox[l] = heads9_dev[3 * l + 1]; h.lmk[l] = heads9_dev[3 * l + 2];
    }
    h.fused = 0;
    fill_anchor_refs(h);
    detect_post_launch(h, N, H, W, threshold, nms_threshold, max_det, workspace_dev, out_count_dev,
                       out_candidates_dev, out_det_dev, static_cast<cudaStream_t>(stream));
  });
}

int tr_retinaface_detect(tr_net* net, const int* head_buffers3, float threshold,
                         double nms_threshold, int max_det, void* workspace_dev,
                         int32_t* out_count_dev, int32_t* out_candidates_dev, float* out_det_dev,
                         void* stream) {
  return guarded([&] {
    TR_CHECK(net->last, "no run yet");
    DetHeads h{};
    int N = 0;
    for (int l = 0; l < 3; ++l) {
      const Buf& b = net->last->bufs.at(head_buffers3[l]);
      TR_CHECK(b.f32 && b.C == 32, "fused head buffers are fp32 with 32 channels");
      h.cls[l] = static_cast<const float*>(b.ptr);
      N = b.N;
    }
    h.fused = 1;
    fill_anchor_refs(h);
    const int H = net->last->H, W = net->last->W;   // image dims of the last run
    for (int l = 0; l < 3; ++l) {
      const int st = l == 0 ? 32 : (l == 1 ? 16 : 8);
      const Buf& b = net->last->bufs.at(head_buffers3[l]);
      TR_CHECK(b.H == ceil_div(H, st) && b.W == ceil_div(W, st),
               "head dims do not match ceil(H/stride): " + std::to_string(b.H) + "x" + std::to_string(b.W));
    }
    detect_post_launch(h, N, H, W, threshold, nms_threshold, max_det, workspace_dev, out_count_dev,
                       out_candidates_dev, out_det_dev, static_cast<cudaStream_t>(stream));
  });
}

int tr_l2_normalize(const float* in_dev, float* out_dev, int N, int D, void* stream) {
  return guarded([&] { l2_normalize_launch(in_dev, out_dev, N, D, static_cast<cudaStream_t>(stream)); });
}

int tr_face_align(const uint8_t* frames_dev, int H, int W, const double* coef_dev,
                  const int32_t* image_index_dev, int F, uint8_t* out_dev, int side, void* stream) {
  return guarded([&] {
    face_align_launch(frames_dev, H, W, coef_dev, image_index_dev, F, out_dev, side,
                      static_cast<cudaStream_t>(stream));
  });
}

int tr_face_similarity(const float* det_dev, const int32_t* count_dev, int N, int max_det, float scale,
                       int cap, double* coef_dev, int32_t* image_index_dev, int32_t* total_dev,
                       void* stream) {
  return guarded([&] {
    face_similarity_launch(det_dev, count_dev, N, max_det, scale, cap, coef_dev, image_index_dev,
                           total_dev, static_cast<cudaStream_t>(stream));
  });
}

int tr_resample_table(int in_size, int out_size, int32_t* bounds_host, int32_t* coeffs_host) {
  int ksize = -1;
  const int rc = guarded([&] { ksize = resample_table_host(in_size, out_size, bounds_host, coeffs_host); });
  return rc ? -1 : ksize;
}

size_t tr_face_letterbox_workspace_bytes(const int32_t* sizes_host, int n, int side) {
  size_t bytes = 0;
  guarded([&] { bytes = face_letterbox_workspace_bytes(sizes_host, n, side); });
  return bytes;
}

int tr_face_letterbox(const uint8_t* pixels_dev, const int64_t* offsets_host, const int32_t* sizes_host,
                      int n, int side, void* workspace_dev, uint8_t* out_dev, void* stream) {
  return guarded([&] {
    face_letterbox_launch(pixels_dev, reinterpret_cast<const long long*>(offsets_host), sizes_host, n,
                          side, workspace_dev, out_dev, static_cast<cudaStream_t>(stream));
  });
}

size_t tr_pose_workspace_bytes(int N) { return pose_workspace_bytes(N); }

int tr_openpose_parse(const float* paf_dev, const float* heat_dev, int N, int h, int w, double scale,
                      void* workspace_dev, int32_t* out_count_dev, int32_t* out_keypoints_dev,
                      double* out_score_dev, int32_t* out_status_dev, void* stream) {
  return guarded([&] {
    PoseOut o{out_count_dev, out_keypoints_dev, out_score_dev, out_status_dev};
    pose_parse_launch(paf_dev, heat_dev, N, h, w, scale, workspace_dev, o,
                      static_cast<cudaStream_t>(stream));
  });
}

void tr_bicubic_table(float out32[32]) { bicubic_table_host(out32); }


/* ---- programs and single-call models ------------------------------------- */
struct tr_program { trb::Program p; };

struct tr_model {
  int kind = 0;                 // 0 retinaface, 1 arcface, 2 openpose
  tr_net* net = nullptr;
  int roles[8];
  void* ws = nullptr; size_t ws_bytes = 0;      // post-processing workspace
  float* f0 = nullptr; size_t f0_bytes = 0;     // PAF / raw embedding
  float* f1 = nullptr; size_t f1_bytes = 0;     // heat maps
  int32_t* cand = nullptr; size_t cand_bytes = 0;
  uint8_t* u8 = nullptr; size_t u8_bytes = 0;   // padded crop batch
  ~tr_model() {
    if (net) tr_net_destroy(net);
    for (void* p : {ws, static_cast<void*>(f0), static_cast<void*>(f1), static_cast<void*>(cand), static_cast<void*>(u8)})
      if (p) cudaFree(p);
  }
};

extern "C++" {
namespace {
template <class T>
void grow(T*& ptr, size_t& have, size_t need) {
  if (have >= need) return;
  if (ptr) TR_CUDA(cudaFree(ptr));
  ptr = nullptr; have = 0;
  TR_CUDA(cudaMalloc(reinterpret_cast<void**>(&ptr), need));
  have = need;
}

int model_create(const char* name, int kind, const void* blob, size_t bytes, tr_model** out) {
  return guarded([&] {
    auto prog = std::make_unique<tr_program>();
    trb::program_build(name, blob, bytes, 0, prog->p);
    auto m = std::make_unique<tr_model>();
    m->kind = kind;
    for (int i = 0; i < 8; ++i) m->roles[i] = prog->p.roles[i];
    if (tr_net_create_from_program(prog.get(), &m->net) != 0) fail(g_last_error);
    *out = m.release();
  });
}
}  // namespace
}  // extern "C++"

int tr_program_build(const char* model, const void* state_dict_blob, size_t bytes, int flags,
                     tr_program** out) {
  return guarded([&] {
    auto prog = std::make_unique<tr_program>();
    trb::program_build(model, state_dict_blob, bytes, flags, prog->p);
    *out = prog.release();
  });
}

void tr_program_destroy(tr_program* program) { delete program; }

int tr_program_info(const tr_program* program, int* n_buffers, int* n_ops, size_t* blob_bytes,
                    int32_t* roles8) {
  return guarded([&] {
    *n_buffers = int(program->p.buffers.size());
    *n_ops = int(program->p.ops.size());
    *blob_bytes = program->p.blob.size();
    for (int i = 0; i < 8; ++i) roles8[i] = program->p.roles[i];
  });
}

int tr_program_copy(const tr_program* program, tr_buffer_desc* buffers, tr_op_desc* ops, void* blob) {
  return guarded([&] {
    std::copy(program->p.buffers.begin(), program->p.buffers.end(), buffers);
    std::copy(program->p.ops.begin(), program->p.ops.end(), ops);
    memcpy(blob, program->p.blob.data(), program->p.blob.size());
  });
}

int tr_net_create_from_program(const tr_program* program, tr_net** out) {
  const trb::Program& p = program->p;
  return tr_net_create(p.buffers.data(), int(p.buffers.size()), p.ops.data(), int(p.ops.size()),
                       p.blob.data(), p.blob.size(), out);
}

int tr_retinaface_create(const void* blob, size_t bytes, tr_model** out) {
  return model_create("retinaface", 0, blob, bytes, out);
}
int tr_arcface_create(const void* blob, size_t bytes, tr_model** out) {
  return model_create("arcface", 1, blob, bytes, out);
}
int tr_openpose_create(const void* blob, size_t bytes, tr_model** out) {
  return model_create("openpose", 2, blob, bytes, out);
}
tr_net* tr_model_net(tr_model* model) { return model->net; }
void tr_model_destroy(tr_model* model) { delete model; }

int tr_retinaface_forward(tr_model* m, const uint8_t* frames_dev, int N, int H, int W, float threshold,
                          double nms_threshold, int max_det, int32_t* count_dev, float* det_dev,
                          void* stream) {
  return guarded([&] {
    TR_CHECK(m->kind == 0, "not a RetinaFace model");
    // model channel order is BGR: start at channel 2 and walk backwards (retinaface/wrapper.py:144-146)
    if (tr_net_run(m->net, frames_dev + 2, N, H, W, int64_t(H) * W * 3, int64_t(W) * 3, 3, -1, stream)) fail(g_last_error);
    grow(m->ws, m->ws_bytes, tr_detect_workspace_bytes(N, H, W));
    grow(m->cand, m->cand_bytes, size_t(N) * 4);
    if (tr_retinaface_detect(m->net, m->roles, threshold, nms_threshold, max_det, m->ws, count_dev, m->cand,
                             det_dev, stream)) fail(g_last_error);
  });
}

int tr_arcface_forward(tr_model* m, const uint8_t* crops_dev, int N, int layout, int normalise, float* emb_dev,
                       void* stream) {
  return guarded([&] {
    TR_CHECK(m->kind == 1, "not an ArcFace model");
    TR_CHECK(N > 0, "empty batch");
    const int S = 112;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Batches are padded to a multiple of 8 crops (zeros): the final FC then runs as a
    // (1, N/8, 8, C) map on the resident-patch kernel with its K split over all SMs.
    const int Np = round_up(N, 8);
    const size_t crop_bytes = size_t(3) * S * S;
    if (Np != N) {
      grow(m->u8, m->u8_bytes, Np * crop_bytes);
      TR_CUDA(cudaMemcpyAsync(m->u8, crops_dev, N * crop_bytes, cudaMemcpyDeviceToDevice, st));
      TR_CUDA(cudaMemsetAsync(m->u8 + N * crop_bytes, 0, (Np - N) * crop_bytes, st));
      crops_dev = m->u8;
    }
    int rc = layout == 0
                 ? tr_net_run(m->net, crops_dev + 2, Np, S, S, int64_t(S) * S * 3, int64_t(S) * 3, 3, -1, stream)
                 : tr_net_run(m->net, crops_dev, Np, S, S, int64_t(3) * S * S, S, 1, int64_t(S) * S, stream);
    if (rc) fail(g_last_error);
    grow(m->f0, m->f0_bytes, size_t(Np) * 512 * 4);
    if (tr_net_export_nchw_f32(m->net, m->roles[0], 0, 512, m->f0, 0, stream)) fail(g_last_error);
    if (normalise) {
      if (tr_l2_normalize(m->f0, emb_dev, N, 512, stream)) fail(g_last_error);
    } else {
      TR_CUDA(cudaMemcpyAsync(emb_dev, m->f0, size_t(N) * 512 * 4, cudaMemcpyDeviceToDevice, st));
    }
  });
}

int tr_openpose_forward(tr_model* m, const uint8_t* frames_dev, int N, int H, int W, double scale,
                        int32_t* count_dev, int32_t* keypoints_dev, double* score_dev, int32_t* status_dev,
                        void* stream) {
  return guarded([&] {
    TR_CHECK(m->kind == 2, "not an OpenPose model");
    if (tr_net_run(m->net, frames_dev, N, H, W, int64_t(H) * W * 3, int64_t(W) * 3, 3, 1, stream)) fail(g_last_error);
    void* ptr; int n, h, w, c;
    if (tr_net_buffer(m->net, m->roles[0], &ptr, &n, &h, &w, &c)) fail(g_last_error);
    grow(m->f0, m->f0_bytes, size_t(N) * 38 * h * w * 4);
    grow(m->f1, m->f1_bytes, size_t(N) * 19 * h * w * 4);
    if (tr_net_export_nchw(m->net, m->roles[0], m->roles[1], 38, m->f0, stream)) fail(g_last_error);
    if (tr_net_export_nchw(m->net, m->roles[0], m->roles[2], 19, m->f1, stream)) fail(g_last_error);
    grow(m->ws, m->ws_bytes, tr_pose_workspace_bytes(N));
    if (tr_openpose_parse(m->f0, m->f1, N, h, w, scale, m->ws, count_dev, keypoints_dev, score_dev, status_dev,
                          stream)) fail(g_last_error);
  });
}

int tr_resize_bilinear_u8(const uint8_t* src_dev, int N, int H, int W, uint8_t* dst_dev, int h,
                          int w, void* stream) {
  return guarded([&] {
    resize_bilinear_u8_launch(src_dev, N, H, W, dst_dev, h, w, static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"
