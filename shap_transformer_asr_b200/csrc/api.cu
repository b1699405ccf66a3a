// C ABI (include/w2s.h): handle, weight re-layout, workspace, per-batch launch plans and the
// orchestration of one masked-coalition forward.  Everything below the boundary is CUDA for sm_100a;
// there is no CPU path.
#include "../../include/w2s.h"
#include "gemm.cuh"
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace w2s;
typedef __nv_bfloat16 bf16;

namespace {

thread_local std::string g_create_error;

struct LayerW {
  // wav2vec2 encoder layer (post-LN / stable-LN)
  bf16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  // conformer extras
  bf16 *f2w1 = nullptr, *f2w2 = nullptr, *wpos = nullptr, *pw1 = nullptr, *pw2 = nullptr;
  float *f2b1 = nullptr, *f2b2 = nullptr;
  float *lnf1_g = nullptr, *lnf1_b = nullptr, *lnf2_g = nullptr, *lnf2_b = nullptr;
  float *lnc_g = nullptr, *lnc_b = nullptr, *lnfin_g = nullptr, *lnfin_b = nullptr;
  float *dw_w = nullptr, *dw_scale = nullptr, *dw_shift = nullptr;  // depthwise taps [H][k], folded BatchNorm
  bf16* pos_proj = nullptr;  // [2T-1, H], built per clip length
};

struct Step {
  std::string name;
  std::function<std::string(cudaStream_t)> run;
  double flops = 0.0;  // algorithmic FLOPs (2*MAC) of this launch, 0 for non-contraction kernels
  double bytes = 0.0;  // algorithmic HBM bytes of this launch (HBM-bound kernels)
};

struct ProfRec {
  std::string name;
  cudaEvent_t e0, e1;
  double flops, bytes;
};

struct Plan {
  int n = 0;
  std::vector<Step> steps;
  std::vector<GemmLaunch*> gemms;
  std::vector<AttnRelPlan*> attn_rel;
  std::vector<PosConvPlan*> posconv;
  std::vector<AttnFaPlan*> attn_fa;
  // the whole tile forward as one CUDA graph (captured on first use; per-call arguments travel through DynArgs)
  cudaGraphExec_t exec = nullptr;
  bool graph_failed = false;
  ~Plan() {
    if (exec) cudaGraphExecDestroy(exec);
    for (auto* f : attn_fa) attention_fa_free(f);
    for (auto* pc : posconv) posconv_free(pc);
    for (auto* g : gemms) delete g;
    for (auto* a : attn_rel) attention_rel_free(a);
  }
};

__global__ void set_dyn_kernel(DynArgs* dst, const DynArgs v) { *dst = v; }

// expected-gradients path: transposed dense weights of one encoder layer (dX = dY W runs as a contraction with W^T)
struct GradW {
  bf16 *wqkvT = nullptr, *woT = nullptr, *w1T = nullptr, *w2T = nullptr;
  // conformer: second macaron feed-forward, pointwise convs, rotary q|k and v projections, reversed depthwise taps
  bf16 *f2w1T = nullptr, *f2w2T = nullptr, *pw1T = nullptr, *pw2T = nullptr, *wqkT = nullptr, *wvT = nullptr;
  float* dw_flip = nullptr;
  // relative positions, per clip length: linear_pos(pe) [2T'-1, H] and its per-head transpose [heads][64][Rp]
  bf16 *pos_proj = nullptr, *pos_projT = nullptr;
};
struct GradPlan;

}  // namespace

struct w2s_handle {
  w2s_config cfg{};
  int device = 0;
  int num_sms = 148;
  bool auto_batch = false;   // max_batch == 0 at create: pick the batch tile per clip length
  std::string err;
  std::vector<void*> allocs;      // weights: live until destroy
  std::vector<void*> ws_allocs;   // workspace: re-made when the clip length changes

  // weights
  float *conv0_w = nullptr, *conv0_b = nullptr, *norm0_g = nullptr, *norm0_b = nullptr;
  float *ln0_wbar = nullptr, *ln0_gram = nullptr, *ln0_wb = nullptr;
  bf16* ln0_wb48 = nullptr;
  float ln0_bmean = 0.f, ln0_b2mean = 0.f;
  bf16* conv_w[W2S_MAX_CONV_LAYERS] = {};
  float* conv_b[W2S_MAX_CONV_LAYERS] = {};
  float* conv_ln_g[W2S_MAX_CONV_LAYERS] = {};
  float* conv_ln_b[W2S_MAX_CONV_LAYERS] = {};
  float *fp_ln_g = nullptr, *fp_ln_b = nullptr, *fp_b = nullptr;
  bf16* fp_w = nullptr;
  bf16* pos_w = nullptr;
  float* pos_b = nullptr;
  float *enc_ln_g = nullptr, *enc_ln_b = nullptr;
  std::vector<LayerW> layers;
  bf16* head_w = nullptr;
  float* head_b = nullptr;

  // clip state
  long long L = 0;
  int M = 0, zwords = 0;
  float baseline = 0.f;
  float* clip = nullptr;
  uint16_t* seg_id = nullptr;
  long long clip_cap = 0;

  // targets
  int mode = W2S_OUT_MAX, D = 0, max_frame = 0;
  int *frames = nullptr, *tokens = nullptr;
  int targets_cap = 0;
  std::vector<int32_t> targets_host;   // staging of the last w2s_set_targets arrays (source of the async upload)

  // workspace (sized for max_batch rows of the current clip length)
  long long ws_L = -1;
  int T = 0, Tp = 0;
  std::vector<int> Tl;  // conv output lengths
  float *gn_a = nullptr, *gn_b = nullptr;
  bf16* gn_wb = nullptr;
  bf16 *bufA = nullptr, *bufB = nullptr;
  bf16 *fpn = nullptr, *h0 = nullptr, *hp = nullptr, *hb = nullptr, *h1 = nullptr, *qkv = nullptr, *ctx = nullptr,
       *ffn = nullptr;
  float* pre = nullptr;
  float* logits = nullptr;
  bf16* hrot = nullptr;
  double* wls_work = nullptr;
  long long wls_cap = 0;
  std::map<int, std::unique_ptr<Plan>> plans;

  // optional per-launch CUDA-event profile (bench.py roofline leg)
  bool profiling = false;
  std::vector<ProfRec> prof;

  // expected-gradients path (grad_plan.cuh): backward weights, plans per tile size, per-call pointers
  bool grad_ready = false;
  std::vector<GradW> gradw;
  bf16* conv_wT[W2S_MAX_CONV_LAYERS] = {};
  bf16* fp_wT = nullptr;
  bf16* pos_w_bwd = nullptr;          // flipped / transposed positional-conv filters (built at create)
  float *grad_ones = nullptr, *grad_zeros = nullptr;   // [H]: identity scale / shift of the depthwise backward
  std::vector<void*> grad_pos_allocs;                  // relative-position tables of the current grad_L
  std::map<int, std::shared_ptr<GradPlan>> grad_plans;
  long long grad_L = -1;
  int grad_tile = 32;
  bool grad_debug = false, grad_debug_built = false;
  bool grad_attn_simt = false;        // cross-check: attention backward on the CUDA-core kernels (w2s_grad_debug bit 1)
  bool grad_attn_unfused = false;     // cross-check: attention backward as batched contractions + row kernels (bit 2)
  int grad_rules = 0, grad_rules_built = 0;   // w2s_grad_rules: DeepLIFT handler rules on paired [explained | reference] rows
  // target frames of a tile travel through a ring of pinned slots (a pageable source would make every call wait for the
  // stream); a slot is reused only after the copy that read it has completed
  static constexpr int kFrameSlots = 8;
  int32_t* grad_frames_pinned = nullptr;          // [kFrameSlots][grad_tile]
  cudaEvent_t grad_frames_done[kFrameSlots] = {};
  unsigned grad_frames_next = 0;
  float *grad_out = nullptr, *grad_out_val = nullptr;
  const float* grad_gout = nullptr;

  // per-call arguments of the plan's kernels (kernels.cuh: DynArgs), rewritten before every tile
  DynArgs* dyn_dev = nullptr;
  int head_ldl = 0;             // logits row stride: vocab rounded up to a multiple of 32
  cudaStream_t cap_stream = nullptr;   // graphs are captured here (the caller's stream may be the legacy stream)
  bool use_graphs = true;
  long long launches = 0;       // kernel launches issued by w2s_eval / w2s_eval_waveforms since create

  ~w2s_handle() {
    for (auto& r : prof) {
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    plans.clear();
    grad_plans.clear();
    for (void* p : grad_pos_allocs) cudaFree(p);
    if (grad_frames_pinned) cudaFreeHost(grad_frames_pinned);
    for (auto& e : grad_frames_done)
      if (e) cudaEventDestroy(e);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (dyn_dev) cudaFree(dyn_dev);
    if (clip) cudaFree(clip);
    if (seg_id) cudaFree(seg_id);
    if (frames) cudaFree(frames);
    if (tokens) cudaFree(tokens);
    if (wls_work) cudaFree(wls_work);
    for (void* p : ws_allocs) cudaFree(p);
    for (void* p : allocs) cudaFree(p);
  }
};

namespace {

template <typename T>
std::string dalloc(std::vector<void*>& pool, T** out, size_t count, bool zero = false) {
  void* p = nullptr;
  const size_t bytes = count * sizeof(T) + 65536;  // slack: strided-conv views may touch a few rows past the end
  W2S_CUDA_OK(cudaMalloc(&p, bytes));
  if (zero) W2S_CUDA_OK(cudaMemset(p, 0, bytes));
  pool.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return "";
}

struct WeightTable {
  std::map<std::string, std::pair<const float*, int64_t>> t;
  std::string prefix;
  std::string get(const std::string& name, int64_t expect, const float** out) const {
    auto it = t.find(name);
    if (it == t.end()) return "missing weight '" + name + "'";
    if (expect >= 0 && it->second.second != expect)
      return "weight '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " +
             std::to_string(expect);
    *out = it->second.first;
    return "";
  }
};

std::string copy_f32(w2s_handle* h, const WeightTable& wt, const std::string& name, int64_t n, float** dst) {
  const float* src = nullptr;
  W2S_TRY(wt.get(name, n, &src));
  W2S_TRY(dalloc(h->allocs, dst, (size_t)n));
  W2S_CUDA_OK(cudaMemcpy(*dst, src, sizeof(float) * n, cudaMemcpyDeviceToDevice));
  return "";
}
std::string copy_bf16(w2s_handle* h, const WeightTable& wt, const std::string& name, int64_t n, bf16* dst) {
  const float* src = nullptr;
  W2S_TRY(wt.get(name, n, &src));
  return launch_cast_bf16(src, dst, n, 0);
}

std::string load_weights(w2s_handle* h, const WeightTable& wt) {
  const w2s_config& c = h->cfg;
  const std::string P = wt.prefix;
  const int H = c.hidden_size, I = c.intermediate_size, V = c.vocab_size;
  const bool layer = c.feat_extract_norm == 1;
  // ---- feature encoder ----
  for (int l = 0; l < c.num_conv_layers; ++l) {
    const std::string cp = P + "feature_extractor.conv_layers." + std::to_string(l) + ".";
    const int Cin = l == 0 ? 1 : c.conv_dim[l - 1], Cout = c.conv_dim[l], kw = c.conv_kernel[l];
    if (l == 0) {
      W2S_TRY(copy_f32(h, wt, cp + "conv.weight", (int64_t)Cout * kw, &h->conv0_w));
      if (c.conv_bias) W2S_TRY(copy_f32(h, wt, cp + "conv.bias", Cout, &h->conv0_b));
      W2S_TRY(copy_f32(h, wt, cp + "layer_norm.weight", Cout, &h->norm0_g));
      W2S_TRY(copy_f32(h, wt, cp + "layer_norm.bias", Cout, &h->norm0_b));
      if (layer) {
        W2S_TRY(dalloc(h->allocs, &h->ln0_wbar, 16));
        W2S_TRY(dalloc(h->allocs, &h->ln0_gram, 256));
        W2S_TRY(dalloc(h->allocs, &h->ln0_wb, 16));
        float* sc = nullptr;
        W2S_TRY(dalloc(h->allocs, &sc, 2));
        W2S_TRY(launch_conv0_ln_prep(h->conv0_w, h->conv0_b, Cout, kw, h->ln0_wbar, h->ln0_gram, h->ln0_wb, sc, 0));
        float hs[2];
        W2S_CUDA_OK(cudaMemcpy(hs, sc, sizeof(hs), cudaMemcpyDeviceToHost));
        h->ln0_bmean = hs[0];
        h->ln0_b2mean = hs[1];
        if (kw == 10 && Cout % 64 == 0 && Cout <= 512) {
          W2S_TRY(dalloc(h->allocs, &h->ln0_wb48, (size_t)Cout * 48));
          W2S_TRY(launch_conv0_ln_b(h->conv0_w, h->conv0_b, h->norm0_g, h->norm0_b, Cout, kw, h->ln0_wb48, 0));
        }
      }
    } else {
      const float* src = nullptr;
      W2S_TRY(wt.get(cp + "conv.weight", (int64_t)Cout * Cin * kw, &src));
      W2S_TRY(dalloc(h->allocs, &h->conv_w[l], (size_t)Cout * Cin * kw));
      W2S_TRY(launch_repack_conv(src, h->conv_w[l], Cout, Cin, kw, 0));
      if (c.conv_bias) W2S_TRY(copy_f32(h, wt, cp + "conv.bias", Cout, &h->conv_b[l]));
      if (layer) {
        W2S_TRY(copy_f32(h, wt, cp + "layer_norm.weight", Cout, &h->conv_ln_g[l]));
        W2S_TRY(copy_f32(h, wt, cp + "layer_norm.bias", Cout, &h->conv_ln_b[l]));
      }
    }
  }
  // ---- feature projection ----
  const int Cl = c.conv_dim[c.num_conv_layers - 1];
  W2S_TRY(copy_f32(h, wt, P + "feature_projection.layer_norm.weight", Cl, &h->fp_ln_g));
  W2S_TRY(copy_f32(h, wt, P + "feature_projection.layer_norm.bias", Cl, &h->fp_ln_b));
  W2S_TRY(dalloc(h->allocs, &h->fp_w, (size_t)H * Cl));
  W2S_TRY(copy_bf16(h, wt, P + "feature_projection.projection.weight", (int64_t)H * Cl, h->fp_w));
  W2S_TRY(copy_f32(h, wt, P + "feature_projection.projection.bias", H, &h->fp_b));
  // ---- encoder ----
  W2S_TRY(copy_f32(h, wt, P + "encoder.layer_norm.weight", H, &h->enc_ln_g));
  W2S_TRY(copy_f32(h, wt, P + "encoder.layer_norm.bias", H, &h->enc_ln_b));
  h->layers.resize(c.num_hidden_layers);
  if (c.kind == 0) {
    const int G = c.num_conv_pos_embedding_groups, kp = c.num_conv_pos_embeddings, cpg = H / G;
    const float* src = nullptr;
    W2S_TRY(wt.get(P + "encoder.pos_conv_embed.conv.weight", (int64_t)H * cpg * kp, &src));
    W2S_TRY(dalloc(h->allocs, &h->pos_w, (size_t)H * kp * 64));
    W2S_TRY(launch_repack_posconv(src, h->pos_w, H, G, kp, 0));
    W2S_TRY(dalloc(h->allocs, &h->pos_w_bwd, (size_t)H * kp * 64));
    W2S_TRY(launch_repack_posconv_bwd(src, h->pos_w_bwd, H, G, kp, 0));
    W2S_TRY(copy_f32(h, wt, P + "encoder.pos_conv_embed.conv.bias", H, &h->pos_b));
    for (int l = 0; l < c.num_hidden_layers; ++l) {
      LayerW& w = h->layers[l];
      const std::string lp = P + "encoder.layers." + std::to_string(l) + ".";
      W2S_TRY(dalloc(h->allocs, &w.wqkv, (size_t)3 * H * H));
      W2S_TRY(copy_bf16(h, wt, lp + "attention.q_proj.weight", (int64_t)H * H, w.wqkv));
      W2S_TRY(copy_bf16(h, wt, lp + "attention.k_proj.weight", (int64_t)H * H, w.wqkv + (size_t)H * H));
      W2S_TRY(copy_bf16(h, wt, lp + "attention.v_proj.weight", (int64_t)H * H, w.wqkv + (size_t)2 * H * H));
      W2S_TRY(dalloc(h->allocs, &w.bqkv, (size_t)3 * H));
      const float* b = nullptr;
      const char* nm[3] = {"attention.q_proj.bias", "attention.k_proj.bias", "attention.v_proj.bias"};
      for (int j = 0; j < 3; ++j) {
        W2S_TRY(wt.get(lp + nm[j], H, &b));
        W2S_CUDA_OK(cudaMemcpy(w.bqkv + (size_t)j * H, b, sizeof(float) * H, cudaMemcpyDeviceToDevice));
      }
      W2S_TRY(dalloc(h->allocs, &w.wo, (size_t)H * H));
      W2S_TRY(copy_bf16(h, wt, lp + "attention.out_proj.weight", (int64_t)H * H, w.wo));
      W2S_TRY(copy_f32(h, wt, lp + "attention.out_proj.bias", H, &w.bo));
      W2S_TRY(copy_f32(h, wt, lp + "layer_norm.weight", H, &w.ln1_g));
      W2S_TRY(copy_f32(h, wt, lp + "layer_norm.bias", H, &w.ln1_b));
      W2S_TRY(dalloc(h->allocs, &w.w1, (size_t)I * H));
      W2S_TRY(copy_bf16(h, wt, lp + "feed_forward.intermediate_dense.weight", (int64_t)I * H, w.w1));
      W2S_TRY(copy_f32(h, wt, lp + "feed_forward.intermediate_dense.bias", I, &w.b1));
      W2S_TRY(dalloc(h->allocs, &w.w2, (size_t)H * I));
      W2S_TRY(copy_bf16(h, wt, lp + "feed_forward.output_dense.weight", (int64_t)H * I, w.w2));
      W2S_TRY(copy_f32(h, wt, lp + "feed_forward.output_dense.bias", H, &w.b2));
      W2S_TRY(copy_f32(h, wt, lp + "final_layer_norm.weight", H, &w.ln2_g));
      W2S_TRY(copy_f32(h, wt, lp + "final_layer_norm.bias", H, &w.ln2_b));
    }
  } else {
    const int kd = c.conv_depthwise_kernel_size;
    for (int l = 0; l < c.num_hidden_layers; ++l) {
      LayerW& w = h->layers[l];
      const std::string lp = P + "encoder.layers." + std::to_string(l) + ".";
      auto lin = [&](const std::string& nm, int64_t out, int64_t in, bf16** wdst, float** bdst) -> std::string {
        W2S_TRY(dalloc(h->allocs, wdst, (size_t)out * in));
        W2S_TRY(copy_bf16(h, wt, lp + nm + ".weight", out * in, *wdst));
        if (bdst) W2S_TRY(copy_f32(h, wt, lp + nm + ".bias", out, bdst));
        return "";
      };
      W2S_TRY(copy_f32(h, wt, lp + "ffn1_layer_norm.weight", H, &w.lnf1_g));
      W2S_TRY(copy_f32(h, wt, lp + "ffn1_layer_norm.bias", H, &w.lnf1_b));
      W2S_TRY(lin("ffn1.intermediate_dense", I, H, &w.w1, &w.b1));
      W2S_TRY(lin("ffn1.output_dense", H, I, &w.w2, &w.b2));
      W2S_TRY(copy_f32(h, wt, lp + "self_attn_layer_norm.weight", H, &w.ln1_g));
      W2S_TRY(copy_f32(h, wt, lp + "self_attn_layer_norm.bias", H, &w.ln1_b));
      // relative positions: the query projection is emitted twice, (q + u | q + v | k | v), with pos_bias_u / pos_bias_v
      // folded into the fp32 bias of the respective copy (HF :524-527 adds them to q before the two score matmuls)
      const bool relp = c.position_embeddings_type == 1;
      const int nq = relp ? 2 : 1;
      W2S_TRY(dalloc(h->allocs, &w.wqkv, (size_t)(nq + 2) * H * H));
      W2S_TRY(dalloc(h->allocs, &w.bqkv, (size_t)(nq + 2) * H));
      for (int j = 0; j < nq; ++j)
        W2S_TRY(copy_bf16(h, wt, lp + "self_attn.linear_q.weight", (int64_t)H * H, w.wqkv + (size_t)j * H * H));
      W2S_TRY(copy_bf16(h, wt, lp + "self_attn.linear_k.weight", (int64_t)H * H, w.wqkv + (size_t)nq * H * H));
      W2S_TRY(copy_bf16(h, wt, lp + "self_attn.linear_v.weight", (int64_t)H * H, w.wqkv + (size_t)(nq + 1) * H * H));
      {
        const float *bq = nullptr, *bk = nullptr, *bv = nullptr;
        W2S_TRY(wt.get(lp + "self_attn.linear_q.bias", H, &bq));
        W2S_TRY(wt.get(lp + "self_attn.linear_k.bias", H, &bk));
        W2S_TRY(wt.get(lp + "self_attn.linear_v.bias", H, &bv));
        for (int j = 0; j < nq; ++j)
          W2S_CUDA_OK(cudaMemcpy(w.bqkv + (size_t)j * H, bq, sizeof(float) * H, cudaMemcpyDeviceToDevice));
        W2S_CUDA_OK(cudaMemcpy(w.bqkv + (size_t)nq * H, bk, sizeof(float) * H, cudaMemcpyDeviceToDevice));
        W2S_CUDA_OK(cudaMemcpy(w.bqkv + (size_t)(nq + 1) * H, bv, sizeof(float) * H, cudaMemcpyDeviceToDevice));
      }
      W2S_TRY(lin("self_attn.linear_out", H, H, &w.wo, &w.bo));
      if (relp) {
        W2S_TRY(lin("self_attn.linear_pos", H, H, &w.wpos, nullptr));
        const float *u = nullptr, *v = nullptr;
        W2S_TRY(wt.get(lp + "self_attn.pos_bias_u", H, &u));
        W2S_TRY(wt.get(lp + "self_attn.pos_bias_v", H, &v));
        W2S_TRY(launch_axpy(u, w.bqkv, H, 0));
        W2S_TRY(launch_axpy(v, w.bqkv + H, H, 0));
      }
      W2S_TRY(copy_f32(h, wt, lp + "conv_module.layer_norm.weight", H, &w.lnc_g));
      W2S_TRY(copy_f32(h, wt, lp + "conv_module.layer_norm.bias", H, &w.lnc_b));
      {
        const float* src = nullptr;
        W2S_TRY(wt.get(lp + "conv_module.pointwise_conv1.weight", (int64_t)2 * H * H, &src));
        W2S_TRY(dalloc(h->allocs, &w.pw1, (size_t)2 * H * H));
        W2S_TRY(launch_repack_glu(src, w.pw1, H, H, 0));
      }
      {
        const float* src = nullptr;
        W2S_TRY(wt.get(lp + "conv_module.depthwise_conv.weight", (int64_t)H * kd, &src));
        W2S_TRY(dalloc(h->allocs, &w.dw_w, (size_t)H * kd));
        W2S_TRY(launch_transpose_f32(src, w.dw_w, H, kd, 0));   // [H][k] -> [k][H]: lanes read consecutive channels
      }
      {
        const float *g = nullptr, *b = nullptr, *mu = nullptr, *var = nullptr;
        W2S_TRY(wt.get(lp + "conv_module.batch_norm.weight", H, &g));
        W2S_TRY(wt.get(lp + "conv_module.batch_norm.bias", H, &b));
        W2S_TRY(wt.get(lp + "conv_module.batch_norm.running_mean", H, &mu));
        W2S_TRY(wt.get(lp + "conv_module.batch_norm.running_var", H, &var));
        W2S_TRY(dalloc(h->allocs, &w.dw_scale, (size_t)H));
        W2S_TRY(dalloc(h->allocs, &w.dw_shift, (size_t)H));
        W2S_TRY(launch_bn_fold(g, b, mu, var, H, 1e-5f, w.dw_scale, w.dw_shift, 0));
      }
      W2S_TRY(lin("conv_module.pointwise_conv2", H, H, &w.pw2, nullptr));
      W2S_TRY(copy_f32(h, wt, lp + "ffn2_layer_norm.weight", H, &w.lnf2_g));
      W2S_TRY(copy_f32(h, wt, lp + "ffn2_layer_norm.bias", H, &w.lnf2_b));
      W2S_TRY(lin("ffn2.intermediate_dense", I, H, &w.f2w1, &w.f2b1));
      W2S_TRY(lin("ffn2.output_dense", H, I, &w.f2w2, &w.f2b2));
      W2S_TRY(copy_f32(h, wt, lp + "final_layer_norm.weight", H, &w.lnfin_g));
      W2S_TRY(copy_f32(h, wt, lp + "final_layer_norm.bias", H, &w.lnfin_b));
    }
  }
  // lm_head rows are padded with zeros to a multiple of 32 so that any vocabulary size runs on the contraction kernel
  // (the reduction kernel only reads the first V entries of a logits row)
  {
    const int Vp = (V + 31) / 32 * 32;
    h->head_ldl = Vp;
    W2S_TRY(dalloc(h->allocs, &h->head_w, (size_t)Vp * H, true));
    W2S_TRY(copy_bf16(h, wt, "lm_head.weight", (int64_t)V * H, h->head_w));
    const float* b = nullptr;
    W2S_TRY(wt.get("lm_head.bias", V, &b));
    W2S_TRY(dalloc(h->allocs, &h->head_b, (size_t)Vp, true));
    W2S_CUDA_OK(cudaMemcpy(h->head_b, b, sizeof(float) * V, cudaMemcpyDeviceToDevice));
  }
  W2S_CUDA_OK(cudaDeviceSynchronize());
  return "";
}

GemmProblem plain_problem(const bf16* a, long long rows, int K, const bf16* w, int N) {
  GemmProblem p;
  p.a = a; p.a_cols = K; p.a_rows = rows; p.a_batches = 1; p.a_row_stride = K; p.a_batch_stride = rows * (long long)K;
  p.a_kb_per_row = K / 64; p.a_g_col = 0;
  p.w = w; p.M = (int)rows; p.N = N; p.K = K; p.Bz = 1; p.G = 1;
  p.epi.ldg = 0; p.epi.ldb = 0; p.epi.ldm = N;
  return p;
}

int64_t num_frames(const w2s_config& c, int64_t L, std::vector<int>* lens) {
  int64_t n = L;
  for (int l = 0; l < c.num_conv_layers; ++l) {
    if (n < c.conv_kernel[l]) return 0;
    n = (n - c.conv_kernel[l]) / c.conv_stride[l] + 1;
    if (lens) lens->push_back((int)n);
  }
  return n;
}

void free_workspace(w2s_handle* h) {
  h->plans.clear();
  for (void* p : h->ws_allocs) cudaFree(p);
  h->ws_allocs.clear();
  h->ws_L = -1;
}

std::string ensure_workspace(w2s_handle* h, long long L) {
  if (h->ws_L == L) return "";
  cudaDeviceSynchronize();
  free_workspace(h);
  const w2s_config& c = h->cfg;
  h->Tl.clear();
  const int64_t T = num_frames(c, L, &h->Tl);
  if (T <= 0) return "clip of " + std::to_string(L) + " samples is shorter than the conv receptive field";
  h->T = (int)T;
  h->Tp = (int)((T + 63) / 64 * 64);
  if (h->auto_batch) {
    // rows = tile * T' should fill whole waves of 128-row tiles on all SMs (tile = floor(SMs * 128 * k / T')), with at
    // least 128 coalitions per tile to amortise per-launch prologues, and the conv0 output (the largest buffer) <= 8 GB
    long long k = 1;
    while ((long long)h->num_sms * 128 * k / T < 128) ++k;
    long long tile = (long long)h->num_sms * 128 * k / T;
    const long long conv0_bytes = (long long)h->Tl[0] * c.conv_dim[0] * 2;
    while (tile > 8 && tile * conv0_bytes > (8LL << 30)) tile /= 2;
    h->cfg.max_batch = (int)tile;
  }
  const size_t nb = (size_t)c.max_batch;
  const int H = c.hidden_size, I = c.intermediate_size;
  const int Cl = c.conv_dim[c.num_conv_layers - 1];
  auto& pool = h->ws_allocs;
  W2S_TRY(dalloc(pool, &h->gn_a, nb * c.conv_dim[0]));
  W2S_TRY(dalloc(pool, &h->gn_b, nb * c.conv_dim[0]));
  W2S_TRY(dalloc(pool, &h->gn_wb, nb * c.conv_dim[0] * 32));
  size_t szA = 0, szB = 0;
  for (int l = 0; l < c.num_conv_layers; ++l) {
    const size_t s = nb * (size_t)h->Tl[l] * c.conv_dim[l];
    if (l % 2 == 0) szA = s > szA ? s : szA;
    else szB = s > szB ? s : szB;
  }
  W2S_TRY(dalloc(pool, &h->bufA, szA));
  W2S_TRY(dalloc(pool, &h->bufB, szB ? szB : 1));
  const size_t rows = nb * (size_t)T;
  W2S_TRY(dalloc(pool, &h->fpn, rows * Cl));
  W2S_TRY(dalloc(pool, &h->h0, rows * H));
  W2S_TRY(dalloc(pool, &h->hb, rows * H));
  W2S_TRY(dalloc(pool, &h->h1, rows * H));
  W2S_TRY(dalloc(pool, &h->pre, rows * H));
  W2S_TRY(dalloc(pool, &h->qkv, rows * 4 * H));
  W2S_TRY(dalloc(pool, &h->ctx, rows * H));
  W2S_TRY(dalloc(pool, &h->ffn, rows * I));
  W2S_TRY(dalloc(pool, &h->logits, rows * (size_t)h->head_ldl));
  if (c.kind == 0) {
    W2S_TRY(dalloc(pool, &h->hp,
                   nb * (size_t)(T + c.num_conv_pos_embeddings) * c.num_conv_pos_embedding_groups * 64));
  } else {
    if (c.position_embeddings_type == 2) W2S_TRY(dalloc(pool, &h->hrot, rows * H));
    if (c.position_embeddings_type == 1) {
      // relative position table and its per-layer projection linear_pos(pe): input independent, built once per
      // clip length (HF modeling_wav2vec2_conformer.py:159-205, :509-518)
      const long long R = 2 * T - 1;
      bf16* pe = nullptr;
      W2S_TRY(dalloc(pool, &pe, (size_t)R * H));
      W2S_TRY(launch_relpos((int)T, H, pe, 0));
      for (int l = 0; l < c.num_hidden_layers; ++l) {
        LayerW& w = h->layers[l];
        W2S_TRY(dalloc(pool, &w.pos_proj, (size_t)R * H));
        GemmProblem p = plain_problem(pe, R, H, w.wpos, H);
        p.epi.out = w.pos_proj;
        GemmLaunch gl;
        W2S_TRY(gemm_prepare(p, h->num_sms, &gl));
        W2S_TRY((h->cfg.flags & W2S_FLAG_VALIDATE_GEMM) ? gemm_launch_simt(gl, 0) : gemm_launch_tc(gl, 0));
      }
      W2S_CUDA_OK(cudaDeviceSynchronize());
    }
  }
  h->ws_L = L;
  return "";
}

// ------------------------------------------------------------------------------------------------
// plan construction: every launch of one batch-tile forward, with tensor maps encoded once
// ------------------------------------------------------------------------------------------------
struct PlanBuilder {
  w2s_handle* h;
  Plan* plan;
  int n;
  bool simt_gemm, simt_attn;

  std::string add_gemm(const std::string& name, const GemmProblem& p) {
    GemmLaunch* gl = new GemmLaunch();
    plan->gemms.push_back(gl);
    W2S_TRY(gemm_prepare(p, h->num_sms, gl));
    const bool simt = simt_gemm;
    Step st{name, [gl, simt](cudaStream_t s) { return simt ? gemm_launch_simt(*gl, s) : gemm_launch_tc(*gl, s); }};
    st.flops = 2.0 * p.M * (double)p.N * p.K * p.Bz * p.G;
    plan->steps.push_back(std::move(st));
    return "";
  }
  void add(const std::string& name, std::function<std::string(cudaStream_t)> f, double flops = 0.0, double bytes = 0.0) {
    Step st{name, std::move(f)};
    st.flops = flops;
    st.bytes = bytes;
    plan->steps.push_back(std::move(st));
  }
  std::string add_ln(const std::string& name, const void* in, int in_fp32, long long rows, int H, const float* g,
                     const float* b, float eps, int act, bf16* out, float* out_f32, const bf16* residual = nullptr) {
    // algorithmic HBM bytes: the row read once (+ the bf16 residual) and every output written once
    const double bytes = (double)rows * H * ((in_fp32 ? 4 : 2) + (residual ? 2 : 0) + (out ? 2 : 0) + (out_f32 ? 4 : 0));
    add(name, [=](cudaStream_t s) { return launch_layernorm(in, in_fp32, rows, H, g, b, eps, act, out, out_f32, s, residual); },
        0.0, bytes);
    return "";
  }
  static GemmProblem plain(const bf16* a, long long rows, int K, const bf16* w, int N) {
    return plain_problem(a, rows, K, w, N);
  }

  std::string build_conformer();

  // K9: lm_head on the contraction kernel (N = vocab padded to a multiple of 32) into an fp32 logits buffer, then the
  // warp-level reduction, which reads mode / targets / destination from the per-call argument block.  The fused
  // CUDA-core head kernel exists for W2S_FLAG_VALIDATE_GEMM only.
  std::string add_head() {
    const w2s_config& c = h->cfg;
    const int T = h->T, H = c.hidden_size, V = c.vocab_size;
    HeadParams hp{};
    hp.h = h->hb; hp.w = h->head_w; hp.bias = h->head_b;
    hp.n = n; hp.T = T; hp.H = H; hp.V = V; hp.ldl = h->head_ldl; hp.dyn = h->dyn_dev;
    if (simt_gemm) {
      add("head", [=](cudaStream_t s) { return launch_head(hp, s); });
      return "";
    }
    GemmProblem p = plain(h->hb, (long long)n * T, H, h->head_w, h->head_ldl);
    p.epi.bias = h->head_b;
    p.epi.out = h->logits;
    p.epi.out_fp32 = 1;
    W2S_TRY(add_gemm("lm_head", p));
    const float* logits = h->logits;
    add("head_reduce", [=](cudaStream_t s) { return launch_head_reduce(logits, hp, s); }, 0.0,
        (double)n * T * V * 4.0);
    return "";
  }

  std::string build() {
    const w2s_config& c = h->cfg;
    const int T = h->T, H = c.hidden_size, I = c.intermediate_size;
    const bool layer = c.feat_extract_norm == 1;
    const long long rows = (long long)n * T;
    w2s_handle* hh = h;
    const int nn = n;

    // ---- K1: conv0 + norm + GELU -------------------------------------------------------------------
    {
      Conv0Params cp{};
      cp.dyn = h->dyn_dev;
      cp.n = n; cp.L = (int)h->ws_L; cp.T0 = h->Tl[0]; cp.C = c.conv_dim[0]; cp.kw = c.conv_kernel[0];
      cp.stride = c.conv_stride[0];
      cp.w = h->conv0_w; cp.bias = h->conv0_b; cp.gamma = h->norm0_g; cp.beta = h->norm0_b;
      cp.gn_a = h->gn_a; cp.gn_b = h->gn_b; cp.gn_wb = h->gn_wb;
      cp.ln_wbar = h->ln0_wbar; cp.ln_gram = h->ln0_gram; cp.ln_wb = h->ln0_wb;
      cp.ln_bmean = h->ln0_bmean; cp.ln_b2mean = h->ln0_b2mean; cp.ln_wb48 = h->ln0_wb48;
      cp.out = h->bufA;
      if (!layer)
        add("conv0_stats", [=](cudaStream_t s) { return launch_conv0_stats(cp, s); }, 0.0,
            (4.0 * (double)h->ws_L + (double)cp.C * (8.0 + 64.0)) * n);
      add("conv0", [=](cudaStream_t s) { return launch_conv0(cp, layer, s); },
          2.0 * cp.T0 * (double)cp.C * cp.kw * n, ((double)cp.T0 * cp.C * 2.0 + 4.0 * (double)h->ws_L) * n);
    }
    // ---- K2: conv1..6 as implicit GEMM over the stride-row view of the previous layer ---------------------
    bf16* cur = h->bufA;
    for (int l = 1; l < c.num_conv_layers; ++l) {
      bf16* nxt = (l % 2 == 1) ? h->bufB : h->bufA;
      const int Cin = c.conv_dim[l - 1], Cout = c.conv_dim[l], kw = c.conv_kernel[l], st = c.conv_stride[l];
      const int Tin = h->Tl[l - 1], Tout = h->Tl[l];
      if ((st * Cin) % 64) return "conv layer " + std::to_string(l) + ": stride * in_channels must be a multiple of 64";
      GemmProblem p;
      p.a = cur; p.a_cols = (long long)st * Cin; p.a_rows = (Tin + st - 1) / st; p.a_batches = n;
      p.a_row_stride = (long long)st * Cin; p.a_batch_stride = (long long)Tin * Cin;
      p.a_kb_per_row = st * Cin / 64; p.a_g_col = 0;
      p.w = h->conv_w[l]; p.M = Tout; p.N = Cout; p.K = kw * Cin; p.Bz = n; p.G = 1;
      p.epi.bias = h->conv_b[l];
      p.epi.act = layer ? ACT_NONE : ACT_GELU;
      p.epi.out = nxt; p.epi.ldb = (long long)Tout * Cout; p.epi.ldm = Cout;
      W2S_TRY(add_gemm("conv" + std::to_string(l), p));
      if (layer)
        add_ln("conv" + std::to_string(l) + "_ln_gelu", nxt, 0, (long long)n * Tout, Cout, h->conv_ln_g[l],
               h->conv_ln_b[l], 1e-5f, ACT_GELU, nxt, nullptr);
      cur = nxt;
    }
    // ---- K3: feature projection ---------------------------------------------------------------------------
    const int Cl = c.conv_dim[c.num_conv_layers - 1];
    add_ln("featproj_ln", cur, 0, rows, Cl, h->fp_ln_g, h->fp_ln_b, c.layer_norm_eps, ACT_NONE, h->fpn, nullptr);
    {
      GemmProblem p = plain(h->fpn, rows, Cl, h->fp_w, H);
      p.epi.bias = h->fp_b;
      if (c.kind == 0) {
        p.epi.out = h->h0;
      } else {  // conformer: the un-normalised residual stream lives in fp32
        p.epi.out = h->pre;
        p.epi.out_fp32 = 1;
      }
      W2S_TRY(add_gemm("featproj", p));
    }
    if (c.kind != 0) return build_conformer();
    const bool stable = c.do_stable_layer_norm != 0;
    // ---- K4: positional conv (grouped, k=128) + GELU + residual (+ LayerNorm) ----------------------------
    {
      const int G = c.num_conv_pos_embedding_groups, kp = c.num_conv_pos_embeddings, cpg = H / G;
      add("pos_pad", [=](cudaStream_t s) { return launch_pos_pad(hh->h0, nn, T, H, G, kp, hh->hp, s); }, 0.0,
          ((double)T * H + (double)(T + kp) * G * 64) * 2.0 * n);
      GemmProblem p;
      p.a = h->hp; p.a_cols = (long long)G * 64; p.a_rows = T + kp; p.a_batches = n;
      p.a_row_stride = (long long)G * 64; p.a_batch_stride = (long long)(T + kp) * G * 64;
      p.a_kb_per_row = 1; p.a_g_col = 64;
      p.w = h->pos_w; p.M = T; p.N = cpg; p.K = kp * 64; p.Bz = n; p.G = G;
      p.epi.bias = h->pos_b; p.epi.act = ACT_GELU;
      p.epi.residual = h->h0; p.epi.res_fp32 = 0;
      p.epi.out = h->pre; p.epi.out_fp32 = 1;
      p.epi.ldg = cpg; p.epi.ldb = (long long)T * H; p.epi.ldm = H;
      if (!simt_gemm && posconv_supported(H, G, kp)) {
        PosConvPlan* pc = nullptr;
        W2S_TRY(posconv_prepare(h->hp, h->pos_w, n, T, H, G, kp, p.epi, h->num_sms, &pc));
        plan->posconv.push_back(pc);
        add("pos_conv", [=](cudaStream_t s) { return posconv_launch(pc, s); }, 2.0 * n * T * (double)H * cpg * kp);
      } else {
        W2S_TRY(add_gemm("pos_conv", p));
      }
      if (!stable)
        add_ln("encoder_ln", h->pre, 1, rows, H, h->enc_ln_g, h->enc_ln_b, c.layer_norm_eps, ACT_NONE, h->hb, nullptr);
    }
    // ---- K5-K8: transformer layers ------------------------------------------------------------------------
    AttnParams ap{};
    ap.qkv = h->qkv; ap.ctx = h->ctx; ap.B = n; ap.T = T; ap.Tp = h->Tp; ap.H = H;
    ap.heads = c.num_attention_heads; ap.hd = H / c.num_attention_heads;
    ap.ld = 3 * H; ap.q_off = 0; ap.qv_off = 0; ap.k_off = H; ap.v_off = 2 * H;
    ap.scale = 1.0f / sqrtf((float)ap.hd);
    // attention: the persistent tcgen05 kernel, or -- W2S_FLAG_VALIDATE_ATTN only -- the CUDA-core cross-check.
    // A shape the tensor-core kernel cannot take is an error, never a silent change of code path.
    AttnFaPlan* afl = nullptr;
    if (!simt_attn) {
      if (!attention_fa_supported(ap))
        return "attention: head_dim " + std::to_string(ap.hd) + " is not supported (the tcgen05 kernel needs head_dim 64)";
      W2S_TRY(attention_fa_prepare(ap, h->num_sms, &afl));
      plan->attn_fa.push_back(afl);
    }
    const double attn_flops = 4.0 * T * (double)T * H * n;
    const int act = c.hidden_act == 1 ? ACT_SWISH : ACT_GELU;
    const bool ln_res = !stable && (H == 128 || H == 256 || H == 512 || H == 768 || H == 1024);
    // Post-LN models write the pre-LayerNorm tensor of out_proj / ffn2 as bf16: it halves the traffic of the one
    // HBM-bound contraction and of both LayerNorms (163.5 vs 170.4 ms per C2 step) at the price of one more bf16
    // rounding per sub-layer, inside the parity tolerances (tests/test_gpu_parity.py).  W2S_FLAG_FP32_PRELN restores fp32.
    const bool bf16_preln = (c.flags & W2S_FLAG_FP32_PRELN) == 0;
    const int pre32 = (bf16_preln && ln_res) ? 0 : 1;
    for (int l = 0; l < c.num_hidden_layers; ++l) {
      const LayerW& w = h->layers[l];
      const std::string ls = "L" + std::to_string(l) + ".";
      if (stable) add_ln(ls + "ln1", h->pre, 1, rows, H, w.ln1_g, w.ln1_b, c.layer_norm_eps, ACT_NONE, h->hb, nullptr);
      {
        GemmProblem p = plain(h->hb, rows, H, w.wqkv, 3 * H);
        p.epi.bias = w.bqkv;
        p.epi.out = h->qkv;
        W2S_TRY(add_gemm(ls + "qkv", p));
      }
      if (afl) add(ls + "attention", [=](cudaStream_t s) { return attention_fa_launch(afl, s); }, attn_flops);
      else add(ls + "attention", [=](cudaStream_t s) { return launch_attention_simt(ap, s); }, attn_flops);
      {
        GemmProblem p = plain(h->ctx, rows, H, w.wo, H);
        p.epi.bias = w.bo;
        // post-LN: the bf16 residual is added by the LayerNorm kernel (coalesced), not by the GEMM epilogue
        if (stable || !ln_res) {
          p.epi.residual = stable ? (const void*)h->pre : (const void*)h->hb;
          p.epi.res_fp32 = stable ? 1 : 0;
        }
        p.epi.out = h->pre; p.epi.out_fp32 = pre32;
        W2S_TRY(add_gemm(ls + "out_proj", p));
      }
      if (stable) add_ln(ls + "ln2", h->pre, 1, rows, H, w.ln2_g, w.ln2_b, c.layer_norm_eps, ACT_NONE, h->h1, nullptr);
      else add_ln(ls + "ln1", h->pre, pre32, rows, H, w.ln1_g, w.ln1_b, c.layer_norm_eps, ACT_NONE, h->h1, nullptr,
                  ln_res ? h->hb : nullptr);
      {
        GemmProblem p = plain(h->h1, rows, H, w.w1, I);
        p.epi.bias = w.b1; p.epi.act = act;
        p.epi.out = h->ffn;
        W2S_TRY(add_gemm(ls + "ffn1", p));
      }
      {
        GemmProblem p = plain(h->ffn, rows, I, w.w2, H);
        p.epi.bias = w.b2;
        if (stable || !ln_res) {
          p.epi.residual = stable ? (const void*)h->pre : (const void*)h->h1;
          p.epi.res_fp32 = stable ? 1 : 0;
        }
        p.epi.out = h->pre; p.epi.out_fp32 = pre32;
        W2S_TRY(add_gemm(ls + "ffn2", p));
      }
      if (!stable) add_ln(ls + "ln2", h->pre, pre32, rows, H, w.ln2_g, w.ln2_b, c.layer_norm_eps, ACT_NONE, h->hb, nullptr,
                          ln_res ? h->h1 : nullptr);
    }
    if (stable)
      add_ln("encoder_ln", h->pre, 1, rows, H, h->enc_ln_g, h->enc_ln_b, c.layer_norm_eps, ACT_NONE, h->hb, nullptr);
    // ---- K9: lm_head + reduction ---------------------------------------------------------------------------
    W2S_TRY(add_head());
    return "";
  }
};

std::string PlanBuilder::build_conformer() {
  // HF wav2vec2_conformer/modeling_wav2vec2_conformer.py:633-717 (encoder), :568-630 (layer).  The residual
  // stream `pre` is fp32; every sub-block reads a LayerNorm'd bf16 copy and adds its result back in the
  // GEMM epilogue (alpha = 0.5 for the two macaron feed-forward blocks).
  const w2s_config& c = h->cfg;
  const int T = h->T, H = c.hidden_size, I = c.intermediate_size;
  const long long rows = (long long)n * T;
  const int act = c.hidden_act == 1 ? ACT_SWISH : ACT_GELU;
  w2s_handle* hh = h;
  const int nn = n;
  AttnParams ap{};
  const bool rel = c.position_embeddings_type == 1, rotary = c.position_embeddings_type == 2;
  ap.qkv = h->qkv; ap.ctx = h->ctx; ap.B = n; ap.T = T; ap.Tp = h->Tp; ap.H = H;
  ap.heads = c.num_attention_heads; ap.hd = H / c.num_attention_heads;
  ap.scale = 1.0f / sqrtf((float)ap.hd);
  const int nq = rel ? 2 : 1;
  ap.ld = (nq + 2) * H; ap.q_off = 0; ap.qv_off = rel ? H : 0; ap.k_off = nq * H; ap.v_off = (nq + 1) * H;
  AttnFaPlan* afl = nullptr;
  if (!simt_attn && !rel) {
    if (!attention_fa_supported(ap))
      return "attention: head_dim " + std::to_string(ap.hd) + " is not supported (the tcgen05 kernel needs head_dim 64)";
    W2S_TRY(attention_fa_prepare(ap, h->num_sms, &afl));
    plan->attn_fa.push_back(afl);
  }
  double attn_flops = 4.0 * T * (double)T * H * n;
  if (rel) attn_flops += 2.0 * T * (2.0 * T - 1) * H * n;   // (q + v) . linear_pos(pe) over the 2T'-1 relative positions
  auto ffn = [&](const std::string& ls, const float* lg, const float* lb, const bf16* w1, const float* b1,
                 const bf16* w2, const float* b2) -> std::string {
    add_ln(ls + "ln", h->pre, 1, rows, H, lg, lb, 1e-5f, ACT_NONE, h->hb, nullptr);
    GemmProblem p = plain(h->hb, rows, H, w1, I);
    p.epi.bias = b1; p.epi.act = act; p.epi.out = h->ffn;
    W2S_TRY(add_gemm(ls + "ffn1", p));
    GemmProblem q = plain(h->ffn, rows, I, w2, H);
    q.epi.bias = b2; q.epi.alpha = 0.5f;
    q.epi.residual = h->pre; q.epi.res_fp32 = 1; q.epi.out = h->pre; q.epi.out_fp32 = 1;
    W2S_TRY(add_gemm(ls + "ffn2", q));
    return "";
  };
  for (int l = 0; l < c.num_hidden_layers; ++l) {
    const LayerW& w = h->layers[l];
    const std::string ls = "L" + std::to_string(l) + ".";
    W2S_TRY(ffn(ls + "mac1_", w.lnf1_g, w.lnf1_b, w.w1, w.b1, w.w2, w.b2));
    // ---- self-attention ----
    add_ln(ls + "attn_ln", h->pre, 1, rows, H, w.ln1_g, w.ln1_b, 1e-5f, ACT_NONE, h->hb, nullptr);
    if (rotary) {
      const int base = c.rotary_embedding_base, hd = ap.hd;
      add(ls + "rotary", [=](cudaStream_t s) { return launch_rotary(hh->hb, rows, T, H, hd, base, hh->hrot, s); }, 0.0,
          (double)rows * H * 4.0);
      GemmProblem p = plain(h->hrot, rows, H, w.wqkv, 2 * H);
      p.epi.bias = w.bqkv; p.epi.out = h->qkv; p.epi.ldm = 3 * H;
      W2S_TRY(add_gemm(ls + "qk", p));
      GemmProblem v = plain(h->hb, rows, H, w.wqkv + (size_t)2 * H * H, H);
      v.epi.bias = w.bqkv + 2 * H; v.epi.out = h->qkv + 2 * H; v.epi.ldm = 3 * H;
      W2S_TRY(add_gemm(ls + "v", v));
    } else {
      GemmProblem p = plain(h->hb, rows, H, w.wqkv, (nq + 2) * H);
      p.epi.bias = w.bqkv; p.epi.out = h->qkv;
      W2S_TRY(add_gemm(ls + "qkv", p));
    }
    if (afl) {
      add(ls + "attention", [=](cudaStream_t s) { return attention_fa_launch(afl, s); }, attn_flops);
    } else {
      AttnParams lp = ap;
      if (rel) lp.pos_proj = w.pos_proj;
      if (rel && !simt_attn) {
        if (!attention_rel_supported(lp))
          return "attention (relative positions): head_dim " + std::to_string(lp.hd) + " is not supported (needs 64)";
        AttnRelPlan* rp = nullptr;
        W2S_TRY(attention_rel_prepare(lp, &rp));
        plan->attn_rel.push_back(rp);
        add(ls + "attention", [=](cudaStream_t s) { return attention_rel_launch(rp, s); }, attn_flops);
      } else {
        add(ls + "attention", [=](cudaStream_t s) { return launch_attention_simt(lp, s); }, attn_flops);
      }
    }
    {
      GemmProblem p = plain(h->ctx, rows, H, w.wo, H);
      p.epi.bias = w.bo; p.epi.residual = h->pre; p.epi.res_fp32 = 1; p.epi.out = h->pre; p.epi.out_fp32 = 1;
      W2S_TRY(add_gemm(ls + "out_proj", p));
    }
    // ---- convolution module: LN -> pointwise (GLU epilogue) -> depthwise + BatchNorm + act -> pointwise ----
    add_ln(ls + "conv_ln", h->pre, 1, rows, H, w.lnc_g, w.lnc_b, 1e-5f, ACT_NONE, h->hb, nullptr);
    {
      GemmProblem p = plain(h->hb, rows, H, w.pw1, 2 * H);
      p.epi.glu = 1; p.epi.out = h->h1; p.epi.ldm = H;
      W2S_TRY(add_gemm(ls + "pw1_glu", p));
    }
    {
      const int kd = c.conv_depthwise_kernel_size;
      const float *dw = w.dw_w, *sc = w.dw_scale, *sh = w.dw_shift;
      add(ls + "depthwise", [=](cudaStream_t s) { return launch_depthwise(hh->h1, nn, T, H, kd, dw, sc, sh, act, hh->ctx, s); },
          2.0 * rows * (double)H * kd, (double)rows * H * 4.0);
    }
    {
      GemmProblem p = plain(h->ctx, rows, H, w.pw2, H);
      p.epi.residual = h->pre; p.epi.res_fp32 = 1; p.epi.out = h->pre; p.epi.out_fp32 = 1;
      W2S_TRY(add_gemm(ls + "pw2", p));
    }
    W2S_TRY(ffn(ls + "mac2_", w.lnf2_g, w.lnf2_b, w.f2w1, w.f2b1, w.f2w2, w.f2b2));
    add_ln(ls + "final_ln", h->pre, 1, rows, H, w.lnfin_g, w.lnfin_b, 1e-5f, ACT_NONE, nullptr, h->pre);
  }
  add_ln("encoder_ln", h->pre, 1, rows, H, h->enc_ln_g, h->enc_ln_b, c.layer_norm_eps, ACT_NONE, h->hb, nullptr);
  W2S_TRY(add_head());
  return "";
}

std::string get_plan(w2s_handle* h, int n, Plan** out) {
  auto it = h->plans.find(n);
  if (it != h->plans.end()) {
    *out = it->second.get();
    return "";
  }
  std::unique_ptr<Plan> pl(new Plan());
  pl->n = n;
  PlanBuilder b{h, pl.get(), n, (h->cfg.flags & W2S_FLAG_VALIDATE_GEMM) != 0, (h->cfg.flags & W2S_FLAG_VALIDATE_ATTN) != 0};
  W2S_TRY(b.build());
  *out = pl.get();
  h->plans[n] = std::move(pl);
  return "";
}

int64_t out_width(const w2s_handle* h, int64_t L) {
  const int64_t T = num_frames(h->cfg, L, nullptr);
  switch (h->mode) {
    case W2S_OUT_MAX: return T;
    case W2S_OUT_MEAN: return 1;
    case W2S_OUT_LOGITS: return T * h->cfg.vocab_size;
    default: return h->D;
  }
}

// Capture every launch of a tile plan into one CUDA graph (on the handle's own stream: the caller's may be the legacy
// default stream, which cannot be captured).  A capture that fails for any reason leaves the plan on the eager path.
void capture_plan(w2s_handle* h, Plan* pl) {
  pl->graph_failed = true;
  if (cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  bool ok = true;
  for (const Step& st : pl->steps)
    if (!st.run(h->cap_stream).empty()) {
      ok = false;
      break;
    }
  cudaGraph_t g = nullptr;
  if (cudaStreamEndCapture(h->cap_stream, &g) != cudaSuccess || !g) ok = false;
  if (ok && cudaGraphInstantiate(&pl->exec, g, 0) != cudaSuccess) {
    pl->exec = nullptr;
    ok = false;
  }
  if (g) cudaGraphDestroy(g);
  cudaGetLastError();
  pl->graph_failed = !ok;
}

// One evaluation call: rows in tiles of max_batch (the last tile ragged), each tile = one DynArgs update + the plan.
std::string run_batches(w2s_handle* h, const uint32_t* zbits, const float* x, long long ld, int64_t K, float* out,
                        cudaStream_t s) {
  const int64_t width = out_width(h, h->ws_L);
  if (width <= 0) return "no outputs selected (w2s_set_targets)";
  const bool graphs = h->use_graphs && !h->profiling;
  for (int64_t k0 = 0; k0 < K; k0 += h->cfg.max_batch) {
    const int n = (int)((K - k0) < h->cfg.max_batch ? (K - k0) : h->cfg.max_batch);
    Plan* pl = nullptr;
    W2S_TRY(get_plan(h, n, &pl));
    DynArgs d{};
    if (zbits) {
      d.clip = h->clip; d.seg_id = h->seg_id; d.zbits = zbits + k0 * h->zwords; d.zwords = h->zwords;
      d.baseline = h->baseline;
    } else {
      d.x = x + k0 * ld; d.ld = ld;
    }
    d.out = out + k0 * width;
    d.mode = h->mode; d.D = h->D; d.frames = h->frames; d.tokens = h->tokens;
    set_dyn_kernel<<<1, 1, 0, s>>>(h->dyn_dev, d);
    W2S_CUDA_OK(cudaGetLastError());
    h->launches += 1 + (long long)pl->steps.size();
    if (graphs && !pl->exec && !pl->graph_failed) capture_plan(h, pl);
    if (graphs && pl->exec) {
      W2S_CUDA_OK(cudaGraphLaunch(pl->exec, s));
      continue;
    }
    for (const Step& st : pl->steps) {
      ProfRec rec;
      if (h->profiling) {
        rec.name = st.name;
        rec.flops = st.flops;
        rec.bytes = st.bytes;
        cudaEventCreate(&rec.e0);
        cudaEventCreate(&rec.e1);
        cudaEventRecord(rec.e0, s);
      }
      std::string e = st.run(s);
      if (h->profiling) {
        cudaEventRecord(rec.e1, s);
        h->prof.push_back(rec);
      }
      if (!e.empty()) return st.name + ": " + e;
    }
  }
  return "";
}

std::string check_targets(const w2s_handle* h) {
  if (h->mode == W2S_OUT_LOGIT || h->mode == W2S_OUT_LOGPROB) {
    if (h->D <= 0 || !h->frames || !h->tokens) return "targets not set (w2s_set_targets)";
    if (h->max_frame >= h->T)
      return "target frame " + std::to_string(h->max_frame) + " beyond the clip's " + std::to_string(h->T) + " frames";
  }
  return "";
}

#include "grad_plan.cuh"

int fail(w2s_handle* h, const std::string& e) {
  h->err = e;
  return 1;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int w2s_create(const w2s_config* cfg, const char* const* names, const float* const* ptrs, const int64_t* numels,
               int n_weights, int device, w2s_handle** out) {
  g_create_error.clear();
  if (!cfg || !out) {
    g_create_error = "null argument";
    return 1;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device: this library has no CPU path";
    return 1;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    g_create_error = "cudaSetDevice failed";
    return 1;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + "; this library is built for sm_100a only";
    return 1;
  }
  std::unique_ptr<w2s_handle> h(new w2s_handle());
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->auto_batch = h->cfg.max_batch <= 0;
  if (h->auto_batch) h->cfg.max_batch = 64;
  // the validation modes launch eagerly
  h->use_graphs = (cfg->flags & (W2S_FLAG_NO_GRAPH | W2S_FLAG_VALIDATE_GEMM | W2S_FLAG_VALIDATE_ATTN)) == 0;
  if (cudaMalloc((void**)&h->dyn_dev, sizeof(DynArgs)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "out of device memory";
    return 1;
  }
  if (cfg->num_conv_layers < 1 || cfg->num_conv_layers > W2S_MAX_CONV_LAYERS) {
    g_create_error = "num_conv_layers out of range";
    return 1;
  }
  if (cfg->hidden_size % cfg->num_attention_heads) {
    g_create_error = "hidden_size must be divisible by num_attention_heads";
    return 1;
  }
  std::string e = gemm_init();
  if (e.empty()) e = attention_rel_init();
  if (e.empty()) e = posconv_init();
  if (e.empty()) e = attention_fa_init();
  if (e.empty()) e = attention_bwd_init();
  if (e.empty()) {
    WeightTable wt;
    wt.prefix = cfg->kind == 1 ? "wav2vec2_conformer." : "wav2vec2.";
    for (int i = 0; i < n_weights; ++i) wt.t[names[i]] = {ptrs[i], numels[i]};
    e = load_weights(h.get(), wt);
  }
  if (!e.empty()) {
    g_create_error = e;
    return 1;
  }
  *out = h.release();
  return 0;
}

void w2s_destroy(w2s_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  delete h;
}

const char* w2s_last_error(const w2s_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t w2s_num_frames(const w2s_handle* h, int64_t num_samples) { return num_frames(h->cfg, num_samples, nullptr); }

int w2s_set_clip(w2s_handle* h, const float* x_dev, int64_t L, const int32_t* seg_bounds_host, int M, float baseline,
                 void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (L <= 0 || M <= 0 || M > 2048) return fail(h, "set_clip: need L > 0 and 1 <= M <= 2048");
  if (seg_bounds_host[0] != 0 || seg_bounds_host[M] != L) return fail(h, "set_clip: seg_bounds must span [0, L]");
  std::vector<uint16_t> seg((size_t)L);
  for (int m = 0; m < M; ++m) {
    if (seg_bounds_host[m + 1] < seg_bounds_host[m]) return fail(h, "set_clip: seg_bounds must be ascending");
    for (int64_t i = seg_bounds_host[m]; i < seg_bounds_host[m + 1]; ++i) seg[(size_t)i] = (uint16_t)m;
  }
  std::string e = ensure_workspace(h, L);
  if (!e.empty()) return fail(h, e);
  if (h->clip_cap < L) {
    if (h->clip) cudaFree(h->clip);
    if (h->seg_id) cudaFree(h->seg_id);
    if (cudaMalloc((void**)&h->clip, sizeof(float) * (L + 16)) != cudaSuccess ||
        cudaMalloc((void**)&h->seg_id, sizeof(uint16_t) * (L + 16)) != cudaSuccess)
      return fail(h, "set_clip: out of device memory");
    h->clip_cap = L;
  }
  if (cudaMemcpyAsync(h->clip, x_dev, sizeof(float) * L, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
      cudaMemcpyAsync(h->seg_id, seg.data(), sizeof(uint16_t) * L, cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return fail(h, std::string("set_clip: copy failed: ") + cudaGetErrorString(cudaGetLastError()));
  h->L = L;
  h->M = M;
  h->zwords = (M + 31) / 32;
  h->baseline = baseline;
  return 0;
}

int w2s_set_targets(w2s_handle* h, const int32_t* frame_idx_host, const int32_t* token_idx_host, int D, int mode,
                    void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (mode < W2S_OUT_MAX || mode > W2S_OUT_LOGITS) return fail(h, "set_targets: unknown mode");
  if (mode == W2S_OUT_LOGIT || mode == W2S_OUT_LOGPROB) {
    if (D <= 0 || !frame_idx_host || !token_idx_host) return fail(h, "set_targets: need D > 0 (frame, token) pairs");
    for (int d = 0; d < D; ++d)
      if (token_idx_host[d] < 0 || token_idx_host[d] >= h->cfg.vocab_size || frame_idx_host[d] < 0)
        return fail(h, "set_targets: target out of range");
    if (h->targets_cap < D) {
      // evaluations in flight on the stream still read the old arrays
      if (cudaStreamSynchronize(s) != cudaSuccess) return fail(h, "set_targets: stream error");
      if (h->frames) cudaFree(h->frames);
      if (h->tokens) cudaFree(h->tokens);
      h->frames = h->tokens = nullptr;
      h->targets_cap = 0;
      const int cap = D < 1024 ? 1024 : D;
      if (cudaMalloc((void**)&h->frames, sizeof(int) * cap) != cudaSuccess ||
          cudaMalloc((void**)&h->tokens, sizeof(int) * cap) != cudaSuccess)
        return fail(h, "set_targets: out of device memory");
      h->targets_cap = cap;
    }
    // ordered on the caller's stream behind any evaluation still running with the previous targets; the host arrays
    // are staged in the handle so the caller's buffers may be released when this returns
    h->targets_host.assign(frame_idx_host, frame_idx_host + D);
    h->targets_host.insert(h->targets_host.end(), token_idx_host, token_idx_host + D);
    if (cudaMemcpyAsync(h->frames, h->targets_host.data(), sizeof(int) * D, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(h->tokens, h->targets_host.data() + D, sizeof(int) * D, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return fail(h, "set_targets: copy failed");
    h->D = D;
    h->max_frame = 0;
    for (int d = 0; d < D; ++d) h->max_frame = frame_idx_host[d] > h->max_frame ? frame_idx_host[d] : h->max_frame;
  } else {
    h->D = 0;
  }
  h->mode = mode;
  return 0;
}

int64_t w2s_out_width(const w2s_handle* h, int64_t num_samples) { return out_width(h, num_samples); }

int w2s_eval(w2s_handle* h, const uint32_t* z_bits_dev, int64_t K, float* out_dev, void* stream) {
  if (h->L <= 0) return fail(h, "eval: no clip set (w2s_set_clip)");
  if (K == 0) return 0;   // empty coalition matrix: nothing to evaluate
  if (K < 0 || !z_bits_dev || !out_dev) return fail(h, "eval: null buffer");
  // the workspace (and with it T') follows the clip of THIS call: w2s_eval_waveforms may have left it at another length
  std::string e = ensure_workspace(h, h->L);
  if (e.empty()) e = check_targets(h);
  if (e.empty()) e = run_batches(h, z_bits_dev, nullptr, 0, K, out_dev, (cudaStream_t)stream);
  return e.empty() ? 0 : fail(h, "eval: " + e);
}

int w2s_eval_waveforms(w2s_handle* h, const float* x_dev, int64_t n, int64_t L, int64_t ld, float* out_dev,
                       void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !x_dev || !out_dev) return fail(h, "eval_waveforms: null buffer");
  if (ld < L) return fail(h, "eval_waveforms: row stride smaller than the row length");
  std::string e = ensure_workspace(h, L);
  if (e.empty()) e = check_targets(h);
  if (e.empty()) e = run_batches(h, nullptr, x_dev, ld, n, out_dev, (cudaStream_t)stream);
  return e.empty() ? 0 : fail(h, "eval_waveforms: " + e);
}

int w2s_grad_waveforms(w2s_handle* h, const float* x_dev, int64_t n, int64_t L, int64_t ld, const int32_t* frames_host,
                       float* grad_dev, float* out_dev, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !x_dev || !frames_host || !grad_dev) return fail(h, "grad_waveforms: null buffer");
  if (ld < L) return fail(h, "grad_waveforms: row stride smaller than the row length");
  std::string e = run_grad(h, x_dev, ld, L, n, frames_host, nullptr, grad_dev, out_dev, (cudaStream_t)stream);
  return e.empty() ? 0 : fail(h, "grad_waveforms: " + e);
}

int w2s_vjp_waveforms(w2s_handle* h, const float* x_dev, int64_t n, int64_t L, int64_t ld, const float* gout_dev,
                      float* grad_dev, float* out_dev, void* stream) {
  if (n == 0) return 0;
  if (n < 0 || !x_dev || !gout_dev || !grad_dev) return fail(h, "vjp_waveforms: null buffer");
  if (ld < L) return fail(h, "vjp_waveforms: row stride smaller than the row length");
  std::string e = run_grad(h, x_dev, ld, L, n, nullptr, gout_dev, grad_dev, out_dev, (cudaStream_t)stream);
  return e.empty() ? 0 : fail(h, "vjp_waveforms: " + e);
}

int w2s_grad_debug(w2s_handle* h, int on) {
  const bool snaps = (on & 1) != 0, simt = (on & 2) != 0, unfused = (on & 4) != 0;
  if (simt != h->grad_attn_simt || unfused != h->grad_attn_unfused) h->grad_L = -1;   // rebuild the plans
  h->grad_debug = snaps;
  h->grad_attn_simt = simt;
  h->grad_attn_unfused = unfused;
  return 0;
}

int w2s_grad_rules(w2s_handle* h, int rules) {
  if (!h) return 1;
  if (rules & ~3) return fail(h, "grad_rules: unknown rule bits");
  h->grad_rules = rules;
  return 0;
}

int64_t w2s_grad_peek(w2s_handle* h, const char* name, void* dst_dev, int64_t max_bytes, void* stream) {
  for (auto& kv : h->grad_plans) {
    auto it = kv.second->peek.find(name);
    if (it == kv.second->peek.end()) continue;
    const int64_t bytes = (int64_t)it->second.second;
    if (bytes > max_bytes) return -bytes;
    if (cudaMemcpyAsync(dst_dev, it->second.first, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess)
      return 0;
    return bytes;
  }
  return 0;
}

int w2s_mask(w2s_handle* h, const uint32_t* z_bits_dev, int64_t K, float* out_dev, void* stream) {
  if (h->L <= 0) return fail(h, "mask: no clip set (w2s_set_clip)");
  std::string e = launch_mask(h->clip, h->seg_id, z_bits_dev, h->zwords, K, h->L, h->baseline, out_dev, h->L,
                              (cudaStream_t)stream);
  return e.empty() ? 0 : fail(h, e);
}

int w2s_wls(w2s_handle* h, const uint32_t* z_bits_dev, const double* w_dev, const float* y_dev, int64_t K, int M,
            int D, const double* fx_dev, const double* fnull_dev, double* phi_dev, int32_t* status_dev, void* stream) {
  const long long need = 2 * ((long long)(M - 1) * (M - 1) + (long long)(M - 1) * D);
  if (h->wls_cap < need) {
    cudaStreamSynchronize((cudaStream_t)stream);
    if (h->wls_work) cudaFree(h->wls_work);
    if (cudaMalloc((void**)&h->wls_work, sizeof(double) * need) != cudaSuccess) return fail(h, "wls: out of device memory");
    h->wls_cap = need;
  }
  std::string e = launch_wls(z_bits_dev, (M + 31) / 32, w_dev, y_dev, K, M, D, fx_dev, fnull_dev, phi_dev, status_dev,
                             h->wls_work, (cudaStream_t)stream);
  return e.empty() ? 0 : fail(h, e);
}

int w2s_debug_gemm(int use_tcgen05, const void* a_bf16, const void* w_bf16, const float* bias, const void* residual,
                   int res_fp32, float alpha, void* out, int M, int N, int K, int act, int out_fp32, void* stream) {
  g_create_error.clear();
  std::string e = gemm_init();
  if (e.empty()) {
    GemmProblem p = PlanBuilder::plain((const bf16*)a_bf16, M, K, (const bf16*)w_bf16, N);
    p.epi.bias = bias; p.epi.act = act; p.epi.out = out; p.epi.out_fp32 = out_fp32;
    p.epi.residual = residual; p.epi.res_fp32 = res_fp32; p.epi.alpha = alpha;
    GemmLaunch gl;
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    e = gemm_prepare(p, sms, &gl);
    if (e.empty()) e = use_tcgen05 ? gemm_launch_tc(gl, (cudaStream_t)stream) : gemm_launch_simt(gl, (cudaStream_t)stream);
  }
  if (!e.empty()) {
    g_create_error = e;
    return 1;
  }
  return 0;
}

int w2s_profile_enable(w2s_handle* h, int on) {
  cudaDeviceSynchronize();
  for (auto& r : h->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  h->prof.clear();
  h->profiling = on != 0;
  return 0;
}

int64_t w2s_profile_read(w2s_handle* h, char* names, int64_t names_cap, double* ms, double* flops, double* bytes,
                         int64_t* counts, int64_t max_entries) {
  cudaDeviceSynchronize();
  // aggregate by launch class: strip the "L<i>." layer prefix
  std::vector<std::string> keys;
  std::vector<double> tms, tfl, tby;
  std::vector<int64_t> cnt;
  for (auto& r : h->prof) {
    std::string k = r.name;
    std::string pre;
    if (k.compare(0, 5, "grad.") == 0) {   // gradient path: "grad.L3.qkv" / "grad.B3.qkv_bwd" -> "grad.qkv" / "grad.qkv_bwd"
      pre = "grad.";
      k = k.substr(5);
    }
    if (k.size() > 1 && (k[0] == 'L' || k[0] == 'B') && isdigit((unsigned char)k[1])) k = k.substr(k.find('.') + 1);
    k = pre + k;
    size_t i = 0;
    for (; i < keys.size(); ++i)
      if (keys[i] == k) break;
    if (i == keys.size()) {
      keys.push_back(k);
      tms.push_back(0.0); tfl.push_back(0.0); tby.push_back(0.0); cnt.push_back(0);
    }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    tms[i] += t; tfl[i] += r.flops; tby[i] += r.bytes; cnt[i] += 1;
  }
  std::string joined;
  int64_t n = 0;
  for (size_t i = 0; i < keys.size() && n < max_entries; ++i, ++n) {
    joined += keys[i] + "\n";
    ms[n] = tms[i]; flops[n] = tfl[i]; bytes[n] = tby[i]; counts[n] = cnt[i];
  }
  if ((int64_t)joined.size() + 1 > names_cap) return -1;
  memcpy(names, joined.c_str(), joined.size() + 1);
  return n;
}

int w2s_kernel_count(const w2s_handle* h, int64_t* launches_per_batch, int64_t* batch_tile) {
  if (batch_tile) *batch_tile = h->cfg.max_batch;
  int64_t n = 0;
  for (auto& kv : h->plans)
    if ((int64_t)kv.second->steps.size() > n) n = (int64_t)kv.second->steps.size();
  if (launches_per_batch) *launches_per_batch = n + 1;  // + the per-call argument kernel
  return 0;
}

int64_t w2s_launch_count(const w2s_handle* h) { return h->launches; }

double w2s_flops_per_forward(const w2s_handle* h, int64_t L) {
  const w2s_config& c = h->cfg;
  std::vector<int> Tl;
  const double T = (double)num_frames(c, L, &Tl);
  if (T <= 0) return 0.0;
  double f = 0.0;
  for (int l = 0; l < c.num_conv_layers; ++l)
    f += 2.0 * Tl[l] * c.conv_dim[l] * (l == 0 ? 1 : c.conv_dim[l - 1]) * c.conv_kernel[l];
  const double H = c.hidden_size, I = c.intermediate_size;
  f += 2.0 * T * c.conv_dim[c.num_conv_layers - 1] * H;
  if (c.kind == 0) f += 2.0 * T * H * (H / c.num_conv_pos_embedding_groups) * c.num_conv_pos_embeddings;
  double per_layer = 2.0 * T * H * 3 * H + 4.0 * T * T * H + 2.0 * T * H * H + 4.0 * T * H * I;
  if (c.kind == 1) {
    per_layer += 4.0 * T * H * I;                                   // second macaron FFN
    per_layer += 2.0 * T * H * 2 * H + 2.0 * T * H * c.conv_depthwise_kernel_size + 2.0 * T * H * H;  // conv module
    if (c.position_embeddings_type == 1) per_layer += 2.0 * (2 * T - 1) * H * H + 2.0 * T * (2 * T - 1) * H;
  }
  f += per_layer * c.num_hidden_layers;
  f += 2.0 * T * H * c.vocab_size;
  return f;
}

}  // extern "C"
