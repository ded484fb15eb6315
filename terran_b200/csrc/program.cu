// Weight ingestion behind the C ABI: reference state_dict -> layer program + packed blob.
//
// Reads the checkpoint key layout of the reference models (SURVEY.md Appendix C;
// retinaface/model.py, arcface/model.py, openpose/model.py) from a flat "TRSD" blob of named
// tensors, folds every BatchNorm into fp32 per-channel scale/shift vectors, repacks filters to
// the [cout][kh][kw][cin] fp16 layout the tcgen05 kernels' TMA descriptors expect and emits the
// op list net.cu executes.  Host-only code (no CUDA call): a non-Python host builds a model with
// tr_retinaface_create / tr_arcface_create / tr_openpose_create from the same blob.
//
// TRSD blob: "TRSD", u32 version (1), u32 count, then per tensor
//   u16 name length, name bytes, u8 dtype (0 = f32, 1 = i64), u8 ndim, i64 dims[ndim],
//   u64 byte count, zero padding to an 8-byte boundary, data.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/terran_b200.h"
#include "common.cuh"
#include "program.h"

namespace trb {

namespace {

int r_up(int x, int m) { return (x + m - 1) / m * m; }

struct Tensor {
  std::vector<int64_t> dims;
  const float* f = nullptr;     // null for non-f32 tensors
  int64_t numel = 0;
};

struct StateDict {
  std::map<std::string, Tensor> t;
  bool has(const std::string& k) const { return t.count(k) != 0; }
  const Tensor& at(const std::string& k) const {
    auto it = t.find(k);
    if (it == t.end()) fail("checkpoint has no tensor '" + k + "'");
    if (!it->second.f) fail("checkpoint tensor '" + k + "' is not float32");
    return it->second;
  }
};

StateDict parse_state_dict(const void* blob, size_t bytes) {
  const uint8_t* p = static_cast<const uint8_t*>(blob);
  const uint8_t* end = p + bytes;
  auto need = [&](size_t n) { if (size_t(end - p) < n) fail("truncated state_dict blob"); };
  need(12);
  if (memcmp(p, "TRSD", 4) != 0) fail("state_dict blob: bad magic (expected TRSD)");
  uint32_t version, count;
  memcpy(&version, p + 4, 4); memcpy(&count, p + 8, 4);
  if (version != 1) fail("state_dict blob: unsupported version");
  p += 12;
  StateDict sd;
  for (uint32_t i = 0; i < count; ++i) {
    need(2);
    uint16_t nl; memcpy(&nl, p, 2); p += 2;
    need(nl + 2);
    std::string name(reinterpret_cast<const char*>(p), nl); p += nl;
    const uint8_t dtype = p[0], ndim = p[1]; p += 2;
    need(size_t(ndim) * 8 + 8);
    Tensor t;
    t.numel = 1;
    for (int d = 0; d < ndim; ++d) { int64_t v; memcpy(&v, p, 8); p += 8; t.dims.push_back(v); t.numel *= v; }
    uint64_t nb; memcpy(&nb, p, 8); p += 8;
    p += (8 - (reinterpret_cast<uintptr_t>(p) - reinterpret_cast<uintptr_t>(blob)) % 8) % 8;
    need(nb);
    if (dtype == 0) {
      if (nb != uint64_t(t.numel) * 4) fail("state_dict blob: size mismatch for " + name);
      if (reinterpret_cast<uintptr_t>(p) % 4) fail("state_dict blob must be 8-byte aligned");
      t.f = reinterpret_cast<const float*>(p);
    }
    p += nb;
    sd.t[name] = t;
  }
  return sd;
}

using Vec = std::vector<float>;

struct ConvOpts {
  int in_coff = 0, out_coff = 0, cin_pad = 0, stride = 1, pad = -1, act = TR_ACT_NONE;
  const std::vector<int>* in_map = nullptr;
  const Vec* slope = nullptr;
  int res = -1, res_coff = 0, res_up2 = 0;
  int lane = 0, sync = 0, engine = TR_ENGINE_AUTO;
  const std::vector<Vec>* shift9 = nullptr;     // 9 vectors of cout floats
  int groups = 0;                                // > 1: w is (cout, cin per group, k, k)
};

struct Builder {
  Program& P;
  explicit Builder(Program& p) : P(p) {}

  int64_t add(const void* data, size_t bytes) {
    P.blob.resize(r_up(int(P.blob.size()), 16), 0);
    const int64_t off = int64_t(P.blob.size());
    const uint8_t* b = static_cast<const uint8_t*>(data);
    P.blob.insert(P.blob.end(), b, b + bytes);
    return off;
  }
  int64_t add_vec(const Vec* v, int n_pad) {
    if (!v) return -1;
    Vec tmp(n_pad, 0.f);
    for (size_t i = 0; i < v->size() && int(i) < n_pad; ++i) tmp[i] = (*v)[i];
    return add(tmp.data(), tmp.size() * 4);
  }
  int buffer(int channels, bool f32 = false) {
    P.buffers.push_back(tr_buffer_desc{channels, f32 ? 1 : 0});
    return int(P.buffers.size()) - 1;
  }
  static tr_op_desc blank() {
    tr_op_desc d{};
    d.in = -1; d.out = -1; d.out2 = -1; d.res = -1;
    d.k = 1; d.stride = 1;
    d.w_off = d.scale_off = d.shift_off = d.slope_off = d.scale2_off = d.shift2_off = -1;
    d.in_scale = 1.f; d.in_shift = 0.f;
    d.dw_w_off = d.dw_scale_off = d.dw_shift_off = d.dw_w16_off = -1;
    d.shift9_off = -1;
    return d;
  }

  // w: (cout, cin, k, k) fp32
  void conv(const float* w, int cout, int cin, int k, const Vec& scale, const Vec& shift, int in, int out,
            const ConvOpts& o) {
    const int pad = o.pad < 0 ? k / 2 : o.pad;
    const int cin_pad = o.cin_pad ? o.cin_pad : (cin <= 8 ? r_up(cin, 8) : r_up(cin, 16));
    const int cout_pad = r_up(cout, 16);
    std::vector<__half> packed(size_t(cout_pad) * k * k * cin_pad, __float2half_rn(0.f));
    for (int oc = 0; oc < cout; ++oc)
      for (int c = 0; c < cin; ++c) {
        const int pos = o.in_map ? (*o.in_map)[c] : c;
        for (int r = 0; r < k; ++r)
          for (int s = 0; s < k; ++s)
            packed[((size_t(oc) * k + r) * k + s) * cin_pad + pos] =
                __float2half_rn(w[((size_t(oc) * cin + c) * k + r) * k + s]);
      }
    tr_op_desc d = blank();
    d.type = TR_OP_CONV;
    d.in = in; d.in_coff = o.in_coff; d.in_c = cin_pad;
    d.out = out; d.out_coff = o.out_coff; d.out_c = r_up(cout, 8);
    d.res = o.res; d.res_coff = o.res_coff; d.res_up2 = o.res_up2;
    d.k = k; d.stride = o.stride; d.pad = pad; d.act = o.act; d.cout_pad = cout_pad;
    d.cin_real = cin; d.cout_real = cout; d.lane = o.lane; d.sync = o.sync; d.engine = o.engine;
    d.groups = o.groups;
    d.w_off = add(packed.data(), packed.size() * 2);
    d.scale_off = add_vec(&scale, cout_pad);
    d.shift_off = add_vec(&shift, cout_pad);
    d.slope_off = add_vec(o.slope, cout_pad);
    if (o.shift9) {
      Vec m(size_t(9) * cout_pad, 0.f);
      for (int c = 0; c < 9; ++c)
        for (int i = 0; i < cout; ++i) m[size_t(c) * cout_pad + i] = (*o.shift9)[c][i];
      d.shift9_off = add(m.data(), m.size() * 4);
    }
    P.ops.push_back(d);
  }

  void stem(const float* w, int cout, const Vec& scale, const Vec& shift, int out, int stride, int act,
            const Vec* slope, float in_scale, float in_shift) {
    Vec wt(size_t(cout) * 27);                         // (cout, 3, 3, 3) -> [cout][kh][kw][c]
    for (int oc = 0; oc < cout; ++oc)
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s) wt[((size_t(oc) * 3 + r) * 3 + s) * 3 + c] = w[((size_t(oc) * 3 + c) * 3 + r) * 3 + s];
    tr_op_desc d = blank();
    d.type = TR_OP_STEM;
    d.out = out; d.out_c = cout; d.k = 3; d.stride = stride; d.pad = 1; d.act = act; d.cout_pad = cout;
    d.in_scale = in_scale; d.in_shift = in_shift;
    d.w_off = add(wt.data(), wt.size() * 4);
    d.scale_off = add_vec(&scale, cout);
    d.shift_off = add_vec(&shift, cout);
    d.slope_off = add_vec(slope, cout);
    P.ops.push_back(d);
  }

  // dw_w: (C, 1, 3, 3)
  void dwconv(const float* dw_w, int c, const Vec& scale, const Vec& shift, int in, int out, int stride) {
    Vec wt(size_t(9) * c);
    for (int ch = 0; ch < c; ++ch)
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) wt[(size_t(r) * 3 + s) * c + ch] = dw_w[(size_t(ch) * 3 + r) * 3 + s];
    tr_op_desc d = blank();
    d.type = TR_OP_DWCONV;
    d.in = in; d.in_c = c; d.out = out; d.out_c = c; d.k = 3; d.stride = stride; d.pad = 1;
    d.act = TR_ACT_RELU; d.cout_pad = c;
    d.w_off = add(wt.data(), wt.size() * 4);
    d.scale_off = add_vec(&scale, c);
    d.shift_off = add_vec(&shift, c);
    P.ops.push_back(d);
  }

  // dw_w: (C, 1, 3, 3); w: (cout, C, 1, 1)
  void sepconv(const float* dw_w, const Vec& dw_scale, const Vec& dw_shift, const float* w, int cout, int cin,
               const Vec& scale, const Vec& shift, int in, int out, int stride) {
    const int cin_pad = cin <= 8 ? r_up(cin, 8) : r_up(cin, 16), cout_pad = r_up(cout, 16);
    std::vector<__half> packed(size_t(cout_pad) * cin_pad, __float2half_rn(0.f));
    for (int oc = 0; oc < cout; ++oc)
      for (int c = 0; c < cin; ++c) packed[size_t(oc) * cin_pad + c] = __float2half_rn(w[size_t(oc) * cin + c]);
    Vec dwp(size_t(9) * cin_pad, 0.f);
    for (int c = 0; c < cin; ++c)
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) dwp[(size_t(r) * 3 + s) * cin_pad + c] = dw_w[(size_t(c) * 3 + r) * 3 + s];
    std::vector<__half> dwp16(dwp.size());
    for (size_t i = 0; i < dwp.size(); ++i) dwp16[i] = __float2half_rn(dwp[i]);
    tr_op_desc d = blank();
    d.type = TR_OP_SEPCONV;
    d.in = in; d.in_c = cin_pad; d.out = out; d.out_c = r_up(cout, 8); d.k = 3; d.stride = stride; d.pad = 1;
    d.act = TR_ACT_RELU; d.cout_pad = cout_pad; d.cin_real = cin; d.cout_real = cout;
    d.w_off = add(packed.data(), packed.size() * 2);
    d.scale_off = add_vec(&scale, cout_pad);
    d.shift_off = add_vec(&shift, cout_pad);
    d.dw_w_off = add(dwp.data(), dwp.size() * 4);
    d.dw_w16_off = add(dwp16.data(), dwp16.size() * 2);
    d.dw_scale_off = add_vec(&dw_scale, cin_pad);
    d.dw_shift_off = add_vec(&dw_shift, cin_pad);
    P.ops.push_back(d);
  }

  void simple(int type, int in, int in_coff, int out, int out_coff, int channels) {
    tr_op_desc d = blank();
    d.type = type; d.in = in; d.in_coff = in_coff; d.in_c = channels; d.out = out; d.out_coff = out_coff;
    d.out_c = channels;
    if (type == TR_OP_MAXPOOL) { d.k = 2; d.stride = 2; }
    if (type == TR_OP_VIEW) { d.in_c = 0; d.out_c = 0; }
    P.ops.push_back(d);
  }
};

// (scale, shift) of BN(conv + bias) as an affine of the conv output (eval mode).
void bn_fold(const StateDict& sd, const std::string& prefix, double eps, const Tensor* conv_bias, Vec& scale,
             Vec& shift) {
  const Tensor &g = sd.at(prefix + ".weight"), &b = sd.at(prefix + ".bias"),
               &m = sd.at(prefix + ".running_mean"), &v = sd.at(prefix + ".running_var");
  const int n = int(g.numel);
  scale.resize(n); shift.resize(n);
  for (int i = 0; i < n; ++i) {
    const double s = double(g.f[i]) / std::sqrt(double(v.f[i]) + eps);
    const double bias = conv_bias ? double(conv_bias->f[i]) : 0.0;
    scale[i] = float(s);
    shift[i] = float(double(b.f[i]) + (bias - double(m.f[i])) * s);
  }
}

Vec ones(int n) { return Vec(size_t(n), 1.f); }
Vec to_vec(const Tensor& t) { return Vec(t.f, t.f + t.numel); }

// ------------------------------------------------------------------ RetinaFace
struct SepBlock { int cin, cout, stride; };
const SepBlock kScales0[] = {{8, 16, 2}, {16, 32, 1}, {32, 32, 2}, {32, 64, 1}, {64, 64, 2}};
const SepBlock kScales1[] = {{64, 128, 1}, {128, 128, 1}, {128, 128, 1}, {128, 128, 1}, {128, 128, 1}, {128, 128, 2}};

void build_retinaface(const StateDict& sd, bool fused, Program& P) {
  Builder B(P);
  const int engine = fused ? TR_ENGINE_MMA : TR_ENGINE_AUTO;
  Vec s, t, ds, dt;
  auto cbr = [&](const std::string& pc, const std::string& pb, int in, int out, double eps, ConvOpts o) {
    const Tensor& w = sd.at(pc + ".weight");
    bn_fold(sd, pb, eps, sd.has(pc + ".bias") ? &sd.at(pc + ".bias") : nullptr, s, t);
    const int cout = int(w.dims[0]), cin = int(w.dims[1]), k = int(w.dims[2]);
    o.act = TR_ACT_RELU;
    // the warp-level kernel wins where either channel count is <= 16 (profiles/r01_retinaface_mma.txt)
    // and on the 1x1 FPN laterals (a tcgen05 tile per 128 pixels with K = 64..256 is all fixed cost)
    static const bool lateral_mma = [] { const char* e = getenv("TRB_FPN_MMA"); return e && atoi(e) != 0; }();
    o.engine = (std::min(cout, cin) <= 16 || (k == 1 && lateral_mma)) ? engine : TR_ENGINE_AUTO;
    B.conv(w.f, cout, cin, k, s, t, in, out, o);
  };
  const int b = B.buffer(8);
  bn_fold(sd, "base.first_conv_block.1", 1e-5, nullptr, s, t);
  B.stem(sd.at("base.first_conv_block.0.weight").f, 8, s, t, b, 2, TR_ACT_RELU, nullptr, 1.f, 0.f);

  // stem -> dw -> [1x1 -> dw]* -> 1x1: pair every depthwise with the 1x1 that FOLLOWS it
  struct Blk { std::string prefix; int cout, stride; };
  std::vector<Blk> blocks;
  for (int i = 0; i < 5; ++i) blocks.push_back({"base.scales.0." + std::to_string(i), kScales0[i].cout, kScales0[i].stride});
  for (int i = 0; i < 6; ++i) blocks.push_back({"base.scales.1." + std::to_string(i), kScales1[i].cout, kScales1[i].stride});
  blocks.push_back({"base.final_conv.0", 256, 1});
  struct Pw { std::string conv, bn; int cout; };
  struct Dw { std::string conv, bn; int stride; bool valid; };
  std::vector<Pw> pointwise;
  std::vector<Dw> next_dw;
  for (auto& bl : blocks) {
    pointwise.push_back({bl.prefix + ".conv_block.0", bl.prefix + ".conv_block.1", bl.cout});
    next_dw.push_back({bl.prefix + ".sep_block.0", bl.prefix + ".sep_block.1", bl.stride, true});
  }
  pointwise.push_back({"base.final_conv.1", "base.final_conv.2", 256});
  next_dw.push_back({"", "", 1, false});
  Dw pending{"base.first_conv_block.3", "base.first_conv_block.4", 1, true};
  int x = b, ch = 8;
  std::vector<int> taps;
  for (size_t i = 0; i < pointwise.size(); ++i) {
    const Pw& pw = pointwise[i];
    bn_fold(sd, pending.bn, 1e-5, nullptr, ds, dt);
    bn_fold(sd, pw.bn, 1e-5, nullptr, s, t);
    const int y = B.buffer(pw.cout);
    const Tensor& w = sd.at(pw.conv + ".weight");
    if (fused) {
      B.sepconv(sd.at(pending.conv + ".weight").f, ds, dt, w.f, pw.cout, int(w.dims[1]), s, t, x, y, pending.stride);
    } else {
      const int d = B.buffer(ch);
      B.dwconv(sd.at(pending.conv + ".weight").f, ch, ds, dt, x, d, pending.stride);
      ConvOpts o; o.act = TR_ACT_RELU;
      B.conv(w.f, pw.cout, int(w.dims[1]), 1, s, t, d, y, o);
    }
    if (i == 4 || i == 10) taps.push_back(y);
    x = y; ch = pw.cout; pending = next_dw[i];
  }
  const int c32 = x, c8 = taps[0], c16 = taps[1];
  const double e = 2e-5;
  const int p32 = B.buffer(64);
  cbr("refiner.conv_stride32.0", "refiner.conv_stride32.1", c32, p32, e, ConvOpts{});
  const int p16 = B.buffer(64);
  { ConvOpts o; o.res = p32; o.res_up2 = 1; cbr("refiner.conv_stride16.0", "refiner.conv_stride16.1", c16, p16, e, o); }
  const int a16 = B.buffer(64);
  cbr("refiner.aggr_stride16.0", "refiner.aggr_stride16.1", p16, a16, e, ConvOpts{});
  const int p8 = B.buffer(64);
  { ConvOpts o; o.res = a16; o.res_up2 = 1; cbr("refiner.conv_stride8.0", "refiner.conv_stride8.1", c8, p8, e, o); }
  const int a8 = B.buffer(64);
  // The three pyramid levels' context modules and heads are independent chains of small,
  // latency-bound kernels: stride 8 stays on the caller's stream, strides 16 and 32 run on the
  // net's side stream (forked here, joined at the end of the program).
  { ConvOpts o; o.sync = TR_SYNC_FORK; cbr("refiner.aggr_stride8.0", "refiner.aggr_stride8.1", p8, a8, e, o); }

  const int strides[3] = {8, 16, 32}, feats[3] = {a8, a16, p32};
  int heads[3] = {-1, -1, -1};
  for (int si = 0; si < 3; ++si) {
    const std::string st = std::to_string(strides[si]), p = "refiner.context_stride" + st;
    const int ctx = B.buffer(64), red = B.buffer(16), tmp = B.buffer(16);
    const int lane = si == 0 ? 0 : 1;
    { ConvOpts o; o.lane = lane; cbr(p + ".context_3x3.0", p + ".context_3x3.1", feats[si], ctx, e, o); }
    { ConvOpts o; o.lane = lane; cbr(p + ".dimension_reducer.0", p + ".dimension_reducer.1", feats[si], red, e, o); }
    { ConvOpts o; o.lane = lane; o.out_coff = 32; cbr(p + ".context_5x5.0", p + ".context_5x5.1", red, ctx, e, o); }
    { ConvOpts o; o.lane = lane; cbr(p + ".context_7x7.0", p + ".context_7x7.1", red, tmp, e, o); }
    { ConvOpts o; o.lane = lane; o.out_coff = 48; cbr(p + ".context_7x7.3", p + ".context_7x7.4", tmp, ctx, e, o); }
    // fused head: [4 class logits | 8 bbox | 20 landmark] -> fp32
    Vec w(size_t(32) * 64), bias(32);
    const char* names[3] = {"cls", "bbox", "landmark"};
    const int counts[3] = {4, 8, 20};
    int row = 0;
    for (int h = 0; h < 3; ++h) {
      const Tensor& hw = sd.at(std::string("outputs.") + names[h] + "_stride" + st + ".weight");
      const Tensor& hb = sd.at(std::string("outputs.") + names[h] + "_stride" + st + ".bias");
      for (int r = 0; r < counts[h]; ++r, ++row) {
        memcpy(&w[size_t(row) * 64], hw.f + size_t(r) * 64, 64 * 4);
        bias[row] = hb.f[r];
      }
    }
    const int head = B.buffer(32, true);
    ConvOpts o; o.engine = engine; o.lane = lane;
    B.conv(w.data(), 32, 64, 1, ones(32), bias, ctx, head, o);
    heads[si] = head;
  }
  P.roles[0] = heads[2]; P.roles[1] = heads[1]; P.roles[2] = heads[0];   // stride 32, 16, 8
}

// --------------------------------------------------------------------- ArcFace
const int kArcChannels[5] = {64, 64, 128, 256, 512};

// conv_W(pad(x*s0 + t0)) * s1 + t1 = conv_{W*s0}(pad(x)) * s1 + shift9[class]  (see weights.py)
void pre_bn_fold(const Tensor& w, const Vec& s0, const Vec& t0, const Vec& s1, const Vec& t1, Vec& wf,
                 std::vector<Vec>& shift9) {
  const int cout = int(w.dims[0]), cin = int(w.dims[1]);
  wf.resize(size_t(w.numel));
  std::vector<double> per_tap(size_t(cout) * 9, 0.0);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int k = 0; k < 9; ++k) {
        const size_t i = (size_t(o) * cin + c) * 9 + k;
        per_tap[size_t(o) * 9 + k] += double(w.f[i]) * double(t0[c]);
        wf[i] = float(double(w.f[i]) * double(s0[c]));
      }
  shift9.assign(9, Vec(size_t(cout)));
  const int lo[3] = {1, 0, 0}, hi[3] = {2, 2, 1};       // filter rows / columns that stay in bounds
  for (int rc = 0; rc < 3; ++rc)
    for (int cc = 0; cc < 3; ++cc)
      for (int o = 0; o < cout; ++o) {
        double term = 0;
        for (int r = lo[rc]; r <= hi[rc]; ++r)
          for (int s = lo[cc]; s <= hi[cc]; ++s) term += per_tap[size_t(o) * 9 + r * 3 + s];
        shift9[rc * 3 + cc][o] = float(term * double(s1[o]) + double(t1[o]));
      }
}

void build_arcface(const StateDict& sd, const int* units, int n_stages, Program& P) {
  Builder B(P);
  const double e = 2e-5;
  Vec s, t, s0, t0, s1, t1, wf;
  std::vector<Vec> shift9;
  int x = B.buffer(kArcChannels[0]);
  bn_fold(sd, "initial_layer.1", e, nullptr, s, t);
  const Vec slope0 = to_vec(sd.at("initial_layer.2.weight"));
  B.stem(sd.at("initial_layer.0.weight").f, 64, s, t, x, 1, TR_ACT_PRELU, &slope0, 0.0078125f,
         -127.5f * 0.0078125f);
  for (int si = 0; si < n_stages; ++si) {
    const int cout = kArcChannels[si + 1];
    const int y_full = B.buffer(cout), y = B.buffer(cout), sc = B.buffer(cout);
    const int xs[2] = {B.buffer(cout), B.buffer(cout)};
    for (int u = 0; u < units[si]; ++u) {
      const std::string p = "stages." + std::to_string(si) + "." + std::to_string(u);
      const int stride = u == 0 ? 2 : 1;
      bn_fold(sd, p + ".body.0", e, nullptr, s0, t0);
      bn_fold(sd, p + ".body.2", e, nullptr, s1, t1);
      const Tensor& w1 = sd.at(p + ".body.1.weight");
      pre_bn_fold(w1, s0, t0, s1, t1, wf, shift9);
      const int yb = u == 0 ? y_full : y;
      const Vec slope = to_vec(sd.at(p + ".body.3.weight"));
      { ConvOpts o; o.act = TR_ACT_PRELU; o.slope = &slope; o.shift9 = &shift9;
        B.conv(wf.data(), int(w1.dims[0]), int(w1.dims[1]), 3, s1, shift9[4], x, yb, o); }
      int res = x;
      if (u == 0) {
        bn_fold(sd, p + ".shortcut.1", e, nullptr, s, t);
        const Tensor& ws = sd.at(p + ".shortcut.0.weight");
        ConvOpts o; o.stride = 2;
        B.conv(ws.f, int(ws.dims[0]), int(ws.dims[1]), 1, s, t, x, sc, o);
        res = sc;
      }
      const int x_new = xs[u & 1];
      bn_fold(sd, p + ".body.5", e, nullptr, s, t);
      const Tensor& w2 = sd.at(p + ".body.4.weight");
      { ConvOpts o; o.stride = stride; o.res = res;
        B.conv(w2.f, int(w2.dims[0]), int(w2.dims[1]), 3, s, t, yb, x_new, o); }
      x = x_new;
    }
  }
  // final_layer: BN2d (no padding follows: folds into the FC exactly), Flatten in (C,H,W) order
  // -> permuted to (H,W,C), Linear, BN1d.
  const int Cl = kArcChannels[n_stages];
  bn_fold(sd, "final_layer.0", e, nullptr, s0, t0);
  const Tensor& W = sd.at("final_layer.3.weight");
  const Tensor& bfc = sd.at("final_layer.3.bias");
  const int K = int(W.dims[1]), hw = K / Cl, nout = int(W.dims[0]);
  Vec Wf(size_t(nout) * K), scale(nout), shift(nout);
  const Tensor &g = sd.at("final_layer.4.weight"), &b = sd.at("final_layer.4.bias"),
               &m = sd.at("final_layer.4.running_mean"), &v = sd.at("final_layer.4.running_var");
  for (int o = 0; o < nout; ++o) {
    double bias = double(bfc.f[o]);
    for (int c = 0; c < Cl; ++c)
      for (int px = 0; px < hw; ++px) {
        const double w = double(W.f[size_t(o) * K + size_t(c) * hw + px]);
        bias += w * double(t0[c]);
        Wf[size_t(o) * K + size_t(px) * Cl + c] = float(w * double(s0[c]));
      }
    const double sc1 = double(g.f[o]) / std::sqrt(double(v.f[o]) + e);
    scale[o] = float(sc1);
    shift[o] = float(double(b.f[o]) + (bias - double(m.f[o])) * sc1);
  }
  const int flat = B.buffer(hw * Cl);
  B.simple(TR_OP_VIEW, x, 0, flat, 0, 0);
  const int emb = B.buffer(nout, true);
  { ConvOpts o; o.pad = 0; B.conv(Wf.data(), nout, K, 1, scale, shift, flat, emb, o); }
  P.roles[0] = emb;
}

// -------------------------------------------------------------------- OpenPose
struct TrunkItem { const char* name; int cin, cout; };     // name == nullptr: 2x2 max-pool
const TrunkItem kTrunk[] = {
    {"conv1_1", 3, 64}, {"conv1_2", 64, 64}, {nullptr, 0, 0}, {"conv2_1", 64, 128}, {"conv2_2", 128, 128},
    {nullptr, 0, 0}, {"conv3_1", 128, 256}, {"conv3_2", 256, 256}, {"conv3_3", 256, 256}, {"conv3_4", 256, 256},
    {nullptr, 0, 0}, {"conv4_1", 256, 512}, {"conv4_2", 512, 512}, {"conv4_3_CPM", 512, 256},
    {"conv4_4_CPM", 256, 128}};

struct StageLayer { std::string name; int cin, cout, k; bool relu; };
std::vector<StageLayer> stage_layers(int stage, int branch) {
  const int cout = branch == 1 ? 38 : 19;
  const std::string L = "L" + std::to_string(branch), S = std::to_string(stage);
  if (stage == 1)
    return {{"conv5_1_CPM_" + L, 128, 128, 3, true}, {"conv5_2_CPM_" + L, 128, 128, 3, true},
            {"conv5_3_CPM_" + L, 128, 128, 3, true}, {"conv5_4_CPM_" + L, 128, 512, 1, true},
            {"conv5_5_CPM_" + L, 512, cout, 1, false}};
  // the reference's no_relu_layers typo keeps the ReLU on Mconv7_stage6_L2 (openpose/model.py:32-39)
  const bool last_relu = stage == 6 && branch == 2;
  std::vector<StageLayer> v;
  v.push_back({"Mconv1_stage" + S + "_" + L, 185, 128, 7, true});
  for (int i = 2; i <= 5; ++i) v.push_back({"Mconv" + std::to_string(i) + "_stage" + S + "_" + L, 128, 128, 7, true});
  v.push_back({"Mconv6_stage" + S + "_" + L, 128, 128, 1, true});
  v.push_back({"Mconv7_stage" + S + "_" + L, 128, cout, 1, last_relu});
  return v;
}

void build_openpose(const StateDict& sd, Program& P) {
  Builder B(P);
  // cat[PAF 38, heat 19, trunk 128] lives in a padded 192-channel buffer
  // [PAF 0..37 | pad | heat 40..58 | pad | trunk 64..191]
  std::vector<int> cat_map;
  for (int i = 0; i < 38; ++i) cat_map.push_back(i);
  for (int i = 0; i < 19; ++i) cat_map.push_back(40 + i);
  for (int i = 0; i < 128; ++i) cat_map.push_back(64 + i);
  const int cat[2] = {B.buffer(192), B.buffer(192)};
  int x = -1, ch = 3;
  const int n_items = int(sizeof(kTrunk) / sizeof(kTrunk[0]));
  for (int i = 0; i < n_items; ++i) {
    const TrunkItem& it = kTrunk[i];
    if (!it.name) {
      const int y = B.buffer(ch);
      B.simple(TR_OP_MAXPOOL, x, 0, y, 0, ch);
      x = y;
      continue;
    }
    const Tensor& w = sd.at(std::string("model0.") + it.name + ".weight");
    const Vec bias = to_vec(sd.at(std::string("model0.") + it.name + ".bias"));
    int y = -1;
    if (it.cin == 3) {
      y = B.buffer(it.cout);
      B.stem(w.f, it.cout, ones(it.cout), bias, y, 1, TR_ACT_RELU, nullptr, 1.0f / 255.0f, -0.5f);
    } else if (i == n_items - 1) {
      ConvOpts o; o.out_coff = 64; o.act = TR_ACT_RELU;
      B.conv(w.f, it.cout, it.cin, 3, ones(it.cout), bias, x, cat[0], o);
      B.simple(TR_OP_COPY, cat[0], 64, cat[1], 64, 128);
    } else {
      y = B.buffer(it.cout);
      ConvOpts o; o.act = TR_ACT_RELU;
      B.conv(w.f, it.cout, it.cin, 3, ones(it.cout), bias, x, y, o);
    }
    x = y; ch = it.cout;
  }
  for (int stage = 1; stage <= 6; ++stage) {
    const int src = stage == 1 ? cat[0] : cat[stage % 2];
    const int dst = stage == 1 ? cat[0] : cat[(stage + 1) % 2];
    const std::vector<StageLayer> specs[2] = {stage_layers(stage, 1), stage_layers(stage, 2)};
    const std::string m[2] = {"model" + std::to_string(stage) + "_1.", "model" + std::to_string(stage) + "_2."};
    // Both branches have the same layer shapes up to their last layer.  Layer 0 reads the same
    // tensor in both (one conv, filter banks stacked); every further layer but the last is ONE
    // grouped conv (groups = 2) over a [branch 1 | branch 2] channel pair — half the launches,
    // and twice the tiles per launch to spread over the SMs.  The last layers (38 / 19 filters,
    // written into the concat buffer) are one conv over the pair as well (below).
    auto stacked = [&](size_t li, Vec& w, Vec& bias) {
      const StageLayer &l1 = specs[0][li], &l2 = specs[1][li];
      const Tensor &w1 = sd.at(m[0] + l1.name + ".weight"), &w2 = sd.at(m[1] + l2.name + ".weight");
      w.resize(size_t(w1.numel + w2.numel));
      memcpy(w.data(), w1.f, size_t(w1.numel) * 4);
      memcpy(w.data() + w1.numel, w2.f, size_t(w2.numel) * 4);
      bias.clear();
      for (const Tensor* tb : {&sd.at(m[0] + l1.name + ".bias"), &sd.at(m[1] + l2.name + ".bias")})
        bias.insert(bias.end(), tb->f, tb->f + tb->numel);
    };
    Vec w, bias;
    const size_t n_layers = specs[0].size();
    int xb = -1, ch = 0;                      // current [branch 1 | branch 2] buffer, channels per branch
    const int tmp[2] = {B.buffer(256), B.buffer(256)};
    for (size_t li = 0; li + 1 < n_layers; ++li) {
      const StageLayer& l = specs[0][li];
      stacked(li, w, bias);
      ConvOpts o; o.act = TR_ACT_RELU;
      if (li == 0) {
        if (stage == 1) o.in_coff = 64;
        else { o.in_map = &cat_map; o.cin_pad = 192; }
      } else {
        o.groups = 2;
      }
      const int y = 2 * l.cout != 256 ? B.buffer(2 * l.cout) : tmp[li & 1];
      B.conv(w.data(), 2 * l.cout, l.cin, l.k, ones(2 * l.cout), bias, li == 0 ? src : xb, y, o);
      xb = y; ch = l.cout;
    }
    {
      // Last layers (1x1, 38 PAF / 19 heat-map filters): ONE conv over the [branch 1 | branch 2]
      // pair whose 64 output rows are the concat buffer's channels 0..63 — rows 0..37 carry the
      // PAF filters on the first `ch` input channels, rows 40..58 the heat-map filters on the
      // second `ch`, everything else is zero (the pad channels 38, 39, 59..63 are written as 0,
      // which is what they hold anyway).  Adding exact zeros changes no fp32 sum, so the result
      // is bit-identical to two per-branch launches; one launch instead of two per stage.
      const StageLayer &l1 = specs[0][n_layers - 1], &l2 = specs[1][n_layers - 1];
      const Tensor &w1 = sd.at(m[0] + l1.name + ".weight"), &w2 = sd.at(m[1] + l2.name + ".weight");
      const Vec b1 = to_vec(sd.at(m[0] + l1.name + ".bias")), b2 = to_vec(sd.at(m[1] + l2.name + ".bias"));
      const int cin2 = 2 * ch;
      Vec wm(size_t(64) * cin2, 0.f), bm(64, 0.f), slope(64, 1.f);
      for (int oc = 0; oc < l1.cout; ++oc) {
        for (int c = 0; c < ch; ++c) wm[size_t(oc) * cin2 + c] = w1.f[size_t(oc) * ch + c];
        bm[oc] = b1[oc];
        if (l1.relu) slope[oc] = 0.f;
      }
      for (int oc = 0; oc < l2.cout; ++oc) {
        for (int c = 0; c < ch; ++c) wm[size_t(40 + oc) * cin2 + ch + c] = w2.f[size_t(oc) * ch + c];
        bm[40 + oc] = b2[oc];
        if (l2.relu) slope[40 + oc] = 0.f;
      }
      // a ReLU on one branch only (the reference keeps it on Mconv7_stage6_L2) is a per-channel
      // PReLU with slope 0 (ReLU) or 1 (identity): max(y, 0) + slope * min(y, 0), exact
      ConvOpts o;
      if (l1.relu || l2.relu) { o.act = TR_ACT_PRELU; o.slope = &slope; }
      B.conv(wm.data(), 64, cin2, 1, ones(64), bm, xb, dst, o);
      P.ops.back().cin_real = ch;                       // algorithmic work: (38 + 19) x ch MACs per pixel
      P.ops.back().cout_real = l1.cout + l2.cout;
    }
  }
  P.roles[0] = cat[(6 + 1) % 2]; P.roles[1] = 0; P.roles[2] = 40;
}

}  // namespace

void program_build(const char* model, const void* blob, size_t bytes, int flags, Program& P) {
  const StateDict sd = parse_state_dict(blob, bytes);
  const std::string m = model ? model : "";
  for (int& r : P.roles) r = -1;
  if (m == "retinaface") {
    build_retinaface(sd, (flags & 1) == 0, P);
  } else if (m == "arcface") {
    // depth read from the checkpoint: units per stage = number of stages.{s}.{u} blocks
    int units[4] = {0, 0, 0, 0};
    for (int s = 0; s < 4; ++s)
      while (sd.has("stages." + std::to_string(s) + "." + std::to_string(units[s]) + ".body.1.weight")) ++units[s];
    int n_stages = 0;
    while (n_stages < 4 && units[n_stages] > 0) ++n_stages;
    TR_CHECK(n_stages > 0, "arcface checkpoint has no stages");
    build_arcface(sd, units, n_stages, P);
  } else if (m == "openpose") {
    build_openpose(sd, P);
  } else {
    fail("unknown model '" + m + "' (retinaface | arcface | openpose)");
  }
}

}  // namespace trb
