// Expected-gradients path (the reference's production explainer: shap.GradientExplainer over ModelWrapper,
// shap_calculation.py:125-162): forward pass with saved activations + backward pass to the waveform, for one batch tile.
// Included by api.cu inside its anonymous namespace (it needs w2s_handle / Step / PlanBuilder::plain).
//
// Scope of this slice: Wav2Vec2ForCTC -- both front ends (feat_extract_norm = "group" / "layer") and both encoder orders
// (post-LN: facebook/wav2vec2-base-960h, the model the reference runs, and wav2vec2-large-960h; stable-LN: the
// wav2vec2-large-lv60 family), GELU.  The conformer encoder is not built.  Only d(output)/d(input) is
// computed: no weight gradients.  Every dense backward contraction dX = dY W runs on the tcgen05 contraction kernels of
// the forward pass with pre-transposed weights; attention backward and the normalisation / activation / conv-gather
// steps are CUDA-core kernels (grad.cu).
#pragma once

struct GradLayerBuf {
  bf16* qkv = nullptr;    // [rows, 3H]   saved q | k | v
  float* s1 = nullptr;    // [rows, H]    h + attention(h): input of layer_norm
  bf16* u = nullptr;      // [rows, I]    pre-activation of the feed-forward
  float* s2 = nullptr;    // [rows, H]    h1 + ffn(h1): input of final_layer_norm
};

struct GradPlan {
  int n = 0;
  std::vector<Step> steps;
  std::vector<GemmLaunch*> gemms;
  std::vector<PosConvPlan*> posconv;
  std::vector<AttnFaPlan*> attn_fa;
  std::vector<void*> allocs;
  // buffers the entry point reads / snapshots
  float* logits = nullptr;
  float* dA = nullptr;
  bf16* D[2] = {nullptr, nullptr};
  int* frames = nullptr;
  std::map<std::string, std::pair<const void*, size_t>> peek;   // debug: name -> (device buffer, bytes) snapshots
  ~GradPlan() {
    for (auto* f : attn_fa) attention_fa_free(f);
    for (auto* pc : posconv) posconv_free(pc);
    for (auto* g : gemms) delete g;
    for (void* p : allocs) cudaFree(p);
  }
};

std::string grad_supported(const w2s_handle* h) {
  const w2s_config& c = h->cfg;
  if (c.kind != 0) return "gradient path: only Wav2Vec2ForCTC is built (conformer: not yet)";
  if (c.hidden_act != 0) return "gradient path: only GELU is built";
  if (c.hidden_size != c.num_attention_heads * 64) return "gradient path: head_dim must be 64";
  if (c.num_conv_pos_embeddings % 2) return "gradient path: odd positional-conv kernels are not built";
  if (c.conv_kernel[0] != 10 || c.conv_dim[0] % 64 || c.conv_dim[0] > 512)
    return "gradient path: conv0 must be k = 10 with 64..512 channels (a multiple of 64)";
  return "";
}

// transposed copies of the dense weights for dX = dY W (once per handle)
std::string grad_prepare_weights(w2s_handle* h) {
  if (h->grad_ready) return "";
  const w2s_config& c = h->cfg;
  const int H = c.hidden_size, I = c.intermediate_size;
  auto tr = [&](const bf16* src, int R, int C, bf16** dst) -> std::string {
    W2S_TRY(dalloc(h->allocs, dst, (size_t)R * C));
    return launch_transpose_bf16(src, *dst, R, C, 0);
  };
  h->gradw.resize(c.num_hidden_layers);
  for (int l = 0; l < c.num_hidden_layers; ++l) {
    const LayerW& w = h->layers[l];
    GradW& g = h->gradw[l];
    W2S_TRY(tr(w.wqkv, 3 * H, H, &g.wqkvT));   // [3H][H] -> [H][3H]
    W2S_TRY(tr(w.wo, H, H, &g.woT));
    W2S_TRY(tr(w.w1, I, H, &g.w1T));           // [I][H] -> [H][I]
    W2S_TRY(tr(w.w2, H, I, &g.w2T));           // [H][I] -> [I][H]
  }
  for (int l = 1; l < c.num_conv_layers; ++l)
    W2S_TRY(tr(h->conv_w[l], c.conv_dim[l], c.conv_kernel[l] * c.conv_dim[l - 1], &h->conv_wT[l]));   // [O][kw C] -> [kw C][O]
  const int Cl = c.conv_dim[c.num_conv_layers - 1];
  W2S_TRY(tr(h->fp_w, H, Cl, &h->fp_wT));      // [H][Cl] -> [Cl][H]
  W2S_CUDA_OK(cudaDeviceSynchronize());
  h->grad_ready = true;
  return "";
}

struct GradBuilder {
  w2s_handle* h;
  GradPlan* plan;
  int n;
  long long L;
  bool debug = false;

  template <typename TT>
  std::string alloc(TT** out, size_t count, bool zero = false) { return dalloc(plan->allocs, out, count, zero); }
  void add(const std::string& name, std::function<std::string(cudaStream_t)> f) {
    plan->steps.push_back(Step{name, std::move(f)});
  }
  std::string add_gemm(const std::string& name, const GemmProblem& p) {
    GemmLaunch* gl = new GemmLaunch();
    plan->gemms.push_back(gl);
    W2S_TRY(gemm_prepare(p, h->num_sms, gl));
    add(name, [gl](cudaStream_t s) { return gemm_launch_tc(*gl, s); });
    plan->steps.back().flops = 2.0 * p.M * (double)p.N * p.K * p.Bz * p.G;
    return "";
  }
  // debug snapshot of a gradient buffer right after the step that produced it
  std::string snap(const std::string& name, const void* src, size_t bytes) {
    if (!debug) return "";
    uint8_t* dst = nullptr;
    W2S_TRY(alloc(&dst, bytes));
    add("snap." + name, [=](cudaStream_t s) -> std::string {
      W2S_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
      return "";
    });
    plan->peek[name] = {dst, bytes};
    return "";
  }

  std::string build() {
    const w2s_config& c = h->cfg;
    const int H = c.hidden_size, I = c.intermediate_size, NL = c.num_hidden_layers, NC = c.num_conv_layers;
    const int C0 = c.conv_dim[0];
    // NB: every value a step lambda uses must be a LOCAL of this function (captured by value): the builder itself is gone
    // by the time the steps run
    std::vector<int> Tl;
    const int T = (int)num_frames(c, L, &Tl);
    if (T <= 0) return "clip shorter than the conv receptive field";
    const long long rows = (long long)n * T;
    const int nn = n;
    w2s_handle* hh = h;

    // ---- buffers ------------------------------------------------------------------------------------------------
    std::vector<bf16*> u(NC), y(NC);
    for (int l = 0; l < NC; ++l) {
      W2S_TRY(alloc(&u[l], (size_t)n * Tl[l] * c.conv_dim[l]));
      W2S_TRY(alloc(&y[l], (size_t)n * Tl[l] * c.conv_dim[l]));
    }
    const bool layer = c.feat_extract_norm == 1, stable = c.do_stable_layer_norm != 0;
    std::vector<bf16*> cpre(NC, nullptr);   // layer-norm front end: conv + bias, the input of the per-frame LayerNorm
    float* ln_rstd0 = nullptr;              // and conv0's per-frame rstd
    if (layer) {
      for (int l = 1; l < NC; ++l) W2S_TRY(alloc(&cpre[l], (size_t)n * Tl[l] * c.conv_dim[l]));
      W2S_TRY(alloc(&ln_rstd0, (size_t)n * Tl[0]));
    }
    float *gn_a = nullptr, *gn_b = nullptr;
    bf16* gn_wb = nullptr;
    W2S_TRY(alloc(&gn_a, (size_t)n * C0));
    W2S_TRY(alloc(&gn_b, (size_t)n * C0));
    W2S_TRY(alloc(&gn_wb, (size_t)n * C0 * 32));
    const int Cl = c.conv_dim[NC - 1];
    bf16 *fpn = nullptr, *h0 = nullptr, *hp = nullptr, *upos = nullptr, *hb = nullptr, *h1 = nullptr, *ctx = nullptr, *ffn = nullptr;
    float* pre0 = nullptr;
    const int G = c.num_conv_pos_embedding_groups, kp = c.num_conv_pos_embeddings, cpg = H / G;
    W2S_TRY(alloc(&fpn, (size_t)rows * Cl));
    W2S_TRY(alloc(&h0, (size_t)rows * H));
    W2S_TRY(alloc(&hp, (size_t)n * (T + kp) * G * 64));
    W2S_TRY(alloc(&upos, (size_t)rows * H));
    W2S_TRY(alloc(&pre0, (size_t)rows * H));
    W2S_TRY(alloc(&hb, (size_t)rows * H));
    W2S_TRY(alloc(&h1, (size_t)rows * H));
    W2S_TRY(alloc(&ctx, (size_t)rows * H));
    W2S_TRY(alloc(&ffn, (size_t)rows * I));
    W2S_TRY(alloc(&plan->logits, (size_t)rows * h->head_ldl));
    std::vector<GradLayerBuf> lb(NL);
    for (int l = 0; l < NL; ++l) {
      W2S_TRY(alloc(&lb[l].qkv, (size_t)rows * 3 * H));
      W2S_TRY(alloc(&lb[l].s1, (size_t)rows * H));
      W2S_TRY(alloc(&lb[l].u, (size_t)rows * I));
      W2S_TRY(alloc(&lb[l].s2, (size_t)rows * H));
    }
    // backward
    float *dA = nullptr, *dS = nullptr, *dT = nullptr, *attn_stats = nullptr, *dFp = nullptr, *m12 = nullptr, *gtap = nullptr;
    bf16 *dS16 = nullptr, *dF = nullptr, *dC = nullptr, *dQKV = nullptr, *dcol = nullptr;
    W2S_TRY(alloc(&dA, (size_t)rows * H));
    W2S_TRY(alloc(&dS, (size_t)rows * H));
    if (stable) W2S_TRY(alloc(&dT, (size_t)rows * H));
    W2S_TRY(alloc(&dS16, (size_t)rows * H));
    W2S_TRY(alloc(&dF, (size_t)rows * I));
    W2S_TRY(alloc(&dC, (size_t)rows * H));
    W2S_TRY(alloc(&dQKV, (size_t)rows * 3 * H));
    W2S_TRY(alloc(&attn_stats, (size_t)rows * c.num_attention_heads * 3));
    // attention backward on the tensor cores: [n, heads, T, Tp] score-shaped buffers and per-head transposes, zero padding
    const int heads = c.num_attention_heads;
    const int Tp = (T + 63) / 64 * 64;
    const size_t nsc = (size_t)n * heads * T * Tp, ntr = (size_t)n * heads * 64 * Tp;
    float *aS = nullptr, *adP = nullptr;
    bf16 *aP = nullptr, *aPT = nullptr, *adS = nullptr, *adST = nullptr, *aQT = nullptr, *aKT = nullptr, *adOT = nullptr;
    const bool attn_tc = !h->grad_attn_simt;
    if (attn_tc) {
      W2S_TRY(alloc(&aS, nsc));
      W2S_TRY(alloc(&adP, nsc));
      W2S_TRY(alloc(&aP, nsc, true));
      W2S_TRY(alloc(&aPT, nsc, true));
      W2S_TRY(alloc(&adS, nsc, true));
      W2S_TRY(alloc(&adST, nsc, true));
      W2S_TRY(alloc(&aQT, ntr, true));
      W2S_TRY(alloc(&aKT, ntr, true));
      W2S_TRY(alloc(&adOT, ntr, true));
    }
    W2S_TRY(alloc(&dFp, (size_t)rows * Cl));
    size_t dmax = 0, colmax = 0;
    for (int l = 0; l < NC; ++l) {
      dmax = std::max(dmax, (size_t)n * Tl[l] * c.conv_dim[l]);
      if (l >= 1) colmax = std::max(colmax, (size_t)n * Tl[l] * c.conv_kernel[l] * c.conv_dim[l - 1]);
    }
    W2S_TRY(alloc(&plan->D[0], dmax));
    W2S_TRY(alloc(&plan->D[1], dmax));
    W2S_TRY(alloc(&dcol, colmax));
    W2S_TRY(alloc(&m12, (size_t)n * C0 * 2));
    W2S_TRY(alloc(&gtap, (size_t)n * Tl[0] * c.conv_kernel[0]));
    W2S_TRY(alloc(&plan->frames, (size_t)n));
    plan->dA = dA;

    // ================================ forward, activations saved =================================================
    {
      Conv0Params cp{};
      cp.dyn = h->dyn_dev;
      cp.n = n; cp.L = (int)L; cp.T0 = Tl[0]; cp.C = C0; cp.kw = c.conv_kernel[0]; cp.stride = c.conv_stride[0];
      cp.w = h->conv0_w; cp.bias = layer ? h->conv0_b : nullptr; cp.gamma = h->norm0_g; cp.beta = h->norm0_b;
      cp.gn_a = gn_a; cp.gn_b = gn_b; cp.gn_wb = gn_wb;
      cp.ln_wbar = h->ln0_wbar; cp.ln_gram = h->ln0_gram; cp.ln_wb = h->ln0_wb;
      cp.ln_bmean = h->ln0_bmean; cp.ln_b2mean = h->ln0_b2mean; cp.ln_wb48 = h->ln0_wb48; cp.ln_rstd_out = ln_rstd0;
      cp.out = u[0];
      cp.pre_act = 1;   // store the normalised pre-activation; GELU follows as its own step
      if (layer && !h->ln0_wb48) return "gradient path: the layer-norm conv0 needs the tensor-core form (64..512 channels)";
      if (!layer) add("conv0_stats", [=](cudaStream_t s) { return launch_conv0_stats(cp, s); });
      add("conv0", [=](cudaStream_t s) { return launch_conv0(cp, layer, s); });
      const long long ne = (long long)n * Tl[0] * C0;
      bf16 *uu = u[0], *yy = y[0];
      add("conv0_gelu", [=](cudaStream_t s) { return launch_gelu_fwd(uu, yy, ne, s); });
      W2S_TRY(snap("f.convu0", u[0], sizeof(bf16) * (size_t)ne));
    }
    for (int l = 1; l < NC; ++l) {
      const int Cin = c.conv_dim[l - 1], Cout = c.conv_dim[l], kw = c.conv_kernel[l], st = c.conv_stride[l];
      const int Tin = Tl[l - 1], Tout = Tl[l];
      if ((st * Cin) % 64) return "conv layer " + std::to_string(l) + ": stride * in_channels must be a multiple of 64";
      GemmProblem p;
      p.a = y[l - 1]; p.a_cols = (long long)st * Cin; p.a_rows = (Tin + st - 1) / st; p.a_batches = n;
      p.a_row_stride = (long long)st * Cin; p.a_batch_stride = (long long)Tin * Cin;
      p.a_kb_per_row = st * Cin / 64; p.a_g_col = 0;
      p.w = h->conv_w[l]; p.M = Tout; p.N = Cout; p.K = kw * Cin; p.Bz = n; p.G = 1;
      p.epi.bias = h->conv_b[l];
      p.epi.act = ACT_NONE;
      p.epi.out = layer ? cpre[l] : u[l]; p.epi.ldb = (long long)Tout * Cout; p.epi.ldm = Cout;
      W2S_TRY(add_gemm("conv" + std::to_string(l), p));
      const long long ne = (long long)n * Tout * Cout;
      bf16 *uu = u[l], *yy = y[l];
      if (layer) {
        const bf16* cc = cpre[l];
        const float *lg = h->conv_ln_g[l], *lb2 = h->conv_ln_b[l];
        const long long lrows = (long long)n * Tout;
        add("conv" + std::to_string(l) + "_ln",
            [=](cudaStream_t s) { return launch_layernorm(cc, 0, lrows, Cout, lg, lb2, 1e-5f, ACT_NONE, uu, nullptr, s); });
      }
      add("conv" + std::to_string(l) + "_gelu", [=](cudaStream_t s) { return launch_gelu_fwd(uu, yy, ne, s); });
      W2S_TRY(snap("f.convu" + std::to_string(l), u[l], sizeof(bf16) * (size_t)ne));
    }
    {
      bf16* y6 = y[NC - 1];
      const float *g = h->fp_ln_g, *b = h->fp_ln_b;
      const float eps = c.layer_norm_eps;
      add("featproj_ln", [=](cudaStream_t s) { return launch_layernorm(y6, 0, rows, Cl, g, b, eps, ACT_NONE, fpn, nullptr, s); });
      GemmProblem p = PlanBuilder::plain(fpn, rows, Cl, h->fp_w, H);
      p.epi.bias = h->fp_b; p.epi.out = h0;
      W2S_TRY(add_gemm("featproj", p));
      W2S_TRY(snap("f.conv" + std::to_string(NC - 1), y6, sizeof(bf16) * (size_t)rows * Cl));
      W2S_TRY(snap("f.h0", h0, sizeof(bf16) * (size_t)rows * H));
    }
    {
      add("pos_pad", [=](cudaStream_t s) { return launch_pos_pad(h0, nn, T, H, G, kp, hp, s, kp / 2); });
      EpiParams e;
      e.bias = h->pos_b; e.act = ACT_NONE; e.out = upos; e.out_fp32 = 0;
      e.ldg = cpg; e.ldb = (long long)T * H; e.ldm = H;
      if (!posconv_supported(H, G, kp)) return "gradient path: positional conv shape not supported by the tcgen05 kernel";
      PosConvPlan* pc = nullptr;
      W2S_TRY(posconv_prepare(hp, h->pos_w, n, T, H, G, kp, e, h->num_sms, &pc));
      plan->posconv.push_back(pc);
      add("pos_conv", [=](cudaStream_t s) { return posconv_launch(pc, s); });
      add("pos_add", [=](cudaStream_t s) { return launch_add_gelu(h0, upos, pre0, rows * H, s); });
      const float *g = h->enc_ln_g, *b = h->enc_ln_b;
      const float eps = c.layer_norm_eps;
      if (!stable) {
        add("encoder_ln", [=](cudaStream_t s) { return launch_layernorm(pre0, 1, rows, H, g, b, eps, ACT_NONE, hb, nullptr, s); });
        W2S_TRY(snap("f.layer0", hb, sizeof(bf16) * (size_t)rows * H));
      }
    }
    AttnParams ap{};
    ap.ctx = ctx; ap.B = n; ap.T = T; ap.Tp = (T + 63) / 64 * 64; ap.H = H;
    ap.heads = c.num_attention_heads; ap.hd = 64;
    ap.ld = 3 * H; ap.q_off = 0; ap.qv_off = 0; ap.k_off = H; ap.v_off = 2 * H;
    ap.scale = 0.125f;
    const float eps = c.layer_norm_eps;
    for (int l = 0; l < NL; ++l) {
      const LayerW& w = h->layers[l];
      const GradLayerBuf B = lb[l];
      const std::string ls = "L" + std::to_string(l) + ".";
      // stable-LN (pre-LN) layers: r1 = residual stream entering the layer (fp32), s1 = r1 + attention(LN1(r1)),
      // s2 = s1 + ffn(LN2(s1)); post-LN layers: s1 = hb + attention(hb), h1 = LN1(s1), s2 = h1 + ffn(h1), hb' = LN2(s2)
      const float* r1 = l == 0 ? pre0 : lb[l - 1].s2;
      if (stable) {
        const float *g = w.ln1_g, *b = w.ln1_b;
        add(ls + "ln1", [=](cudaStream_t s) { return launch_layernorm(r1, 1, rows, H, g, b, eps, ACT_NONE, hb, nullptr, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(hb, rows, H, w.wqkv, 3 * H);
        p.epi.bias = w.bqkv; p.epi.out = B.qkv;
        W2S_TRY(add_gemm(ls + "qkv", p));
      }
      {
        AttnParams lp = ap;
        lp.qkv = B.qkv;
        AttnFaPlan* afl = nullptr;
        W2S_TRY(attention_fa_prepare(lp, h->num_sms, &afl));
        plan->attn_fa.push_back(afl);
        add(ls + "attention", [=](cudaStream_t s) { return attention_fa_launch(afl, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(ctx, rows, H, w.wo, H);
        p.epi.bias = w.bo;
        if (stable) {
          p.epi.residual = r1; p.epi.res_fp32 = 1;
        } else {
          p.epi.residual = hb; p.epi.res_fp32 = 0;
        }
        p.epi.out = B.s1; p.epi.out_fp32 = 1;
        W2S_TRY(add_gemm(ls + "out_proj", p));
      }
      {
        const float *g = stable ? w.ln2_g : w.ln1_g, *b = stable ? w.ln2_b : w.ln1_b;
        const float* in = B.s1;
        add(ls + (stable ? "ln2" : "ln1"),
            [=](cudaStream_t s) { return launch_layernorm(in, 1, rows, H, g, b, eps, ACT_NONE, h1, nullptr, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(h1, rows, H, w.w1, I);
        p.epi.bias = w.b1; p.epi.act = ACT_NONE; p.epi.out = B.u;
        W2S_TRY(add_gemm(ls + "ffn1", p));
        bf16* uu = B.u;
        add(ls + "ffn_gelu", [=](cudaStream_t s) { return launch_gelu_fwd(uu, ffn, rows * I, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(ffn, rows, I, w.w2, H);
        p.epi.bias = w.b2;
        if (stable) {
          p.epi.residual = B.s1; p.epi.res_fp32 = 1;
        } else {
          p.epi.residual = h1; p.epi.res_fp32 = 0;
        }
        p.epi.out = B.s2; p.epi.out_fp32 = 1;
        W2S_TRY(add_gemm(ls + "ffn2", p));
      }
      if (!stable) {
        const float *g = w.ln2_g, *b = w.ln2_b;
        const float* in = B.s2;
        add(ls + "ln2", [=](cudaStream_t s) { return launch_layernorm(in, 1, rows, H, g, b, eps, ACT_NONE, hb, nullptr, s); });
        W2S_TRY(snap("f.layer" + std::to_string(l + 1), hb, sizeof(bf16) * (size_t)rows * H));
      }
    }
    if (stable) {   // the encoder's LayerNorm comes last (HF wav2vec2/modeling_wav2vec2.py: Wav2Vec2EncoderStableLayerNorm)
      const float *g = h->enc_ln_g, *b = h->enc_ln_b;
      const float* in = lb[NL - 1].s2;
      add("encoder_ln", [=](cudaStream_t s) { return launch_layernorm(in, 1, rows, H, g, b, eps, ACT_NONE, hb, nullptr, s); });
    }
    {
      GemmProblem p = PlanBuilder::plain(hb, rows, H, h->head_w, h->head_ldl);
      p.epi.bias = h->head_b; p.epi.out = plan->logits; p.epi.out_fp32 = 1;
      W2S_TRY(add_gemm("lm_head", p));
      W2S_TRY(snap("f.logits", plan->logits, sizeof(float) * (size_t)rows * h->head_ldl));
    }

    // ================================ backward to the waveform ===================================================
    // attention backward of one layer: (saved q | k | v, dC = d ctx) -> dQKV; shared by both encoder orders
    auto add_attention_bwd = [&](const std::string& ls, const GradLayerBuf& B) -> std::string {
      if (!attn_tc) {
        const bf16* q = B.qkv;
        add(ls + "attention_bwd", [=](cudaStream_t s) { return launch_attn_bwd(q, dC, nn, T, H, heads, 0.125f, dQKV, attn_stats, s); });
      } else {
        const bf16* qkv = B.qkv;
        // operands per (coalition b, head g): columns g*64 of a [rows, ld] buffer, or row block g*T of a [n, heads, T, Tp] one
        auto from_rows = [&](const bf16* base, int ld) {   // A = 64 columns of head g, K = 64
          GemmProblem p;
          p.a = base; p.a_cols = ld; p.a_rows = T; p.a_batches = n; p.a_row_stride = ld; p.a_batch_stride = (long long)T * ld;
          p.a_kb_per_row = 1; p.a_g_col = 64;
          p.M = T; p.K = 64; p.Bz = n; p.G = heads;
          return p;
        };
        auto from_scores = [&](const bf16* base) {        // A = [T, Tp] block of (b, g), K = Tp
          GemmProblem p;
          p.a = base; p.a_cols = Tp; p.a_rows = (long long)heads * T; p.a_batches = n; p.a_row_stride = Tp;
          p.a_batch_stride = (long long)heads * T * Tp; p.a_kb_per_row = Tp / 64; p.a_g_col = 0; p.a_g_row = T;
          p.M = T; p.K = Tp; p.Bz = n; p.G = heads;
          return p;
        };
        auto w_rows_of = [&](GemmProblem& p, const bf16* base, int ld) {   // W = the T rows of head g in a [rows, ld] buffer
          p.w = base; p.w_row_stride = ld; p.w_g_stride = 64; p.w_batch_stride = (long long)T * ld; p.w_rows = T; p.N = Tp;
        };
        auto w_transposed = [&](GemmProblem& p, const bf16* base) {        // W = [64, Tp] block of (b, g)
          p.w = base; p.w_row_stride = Tp; p.w_g_stride = 64LL * Tp; p.w_batch_stride = (long long)heads * 64 * Tp; p.N = 64;
        };
        auto score_out = [&](GemmProblem& p, float* out, float alpha) {
          p.epi.out = out; p.epi.out_fp32 = 1; p.epi.alpha = alpha;
          p.epi.ldg = (long long)T * Tp; p.epi.ldb = (long long)heads * T * Tp; p.epi.ldm = Tp;
        };
        auto qkv_out = [&](GemmProblem& p, bf16* out, float alpha) {
          p.epi.out = out; p.epi.out_fp32 = 0; p.epi.alpha = alpha;
          p.epi.ldg = 64; p.epi.ldb = (long long)T * 3 * H; p.epi.ldm = 3 * H;
        };
        {
          GemmProblem p = from_rows(qkv, 3 * H);                 // S = scale Q K^T
          w_rows_of(p, qkv + H, 3 * H);
          score_out(p, aS, 0.125f);
          W2S_TRY(add_gemm(ls + "attn_scores", p));
        }
        add(ls + "attn_softmax", [=](cudaStream_t s) { return launch_attn_softmax_t(aS, nn * heads, T, Tp, aP, aPT, s); });
        {
          GemmProblem p = from_rows(dC, H);                      // dP = dO V^T
          w_rows_of(p, qkv + 2 * H, 3 * H);
          score_out(p, adP, 1.0f);
          W2S_TRY(add_gemm(ls + "attn_dp", p));
        }
        add(ls + "attn_ds", [=](cudaStream_t s) { return launch_attn_ds_t(aP, adP, nn * heads, T, Tp, adS, adST, s); });
        add(ls + "attn_transposes", [=](cudaStream_t s) -> std::string {
          W2S_TRY(launch_head_transpose(qkv, 3 * H, 0, nn, T, Tp, heads, aQT, s));
          W2S_TRY(launch_head_transpose(qkv, 3 * H, H, nn, T, Tp, heads, aKT, s));
          return launch_head_transpose(dC, H, 0, nn, T, Tp, heads, adOT, s);
        });
        {
          GemmProblem p = from_scores(adS);                      // dQ = scale dS K
          w_transposed(p, aKT);
          qkv_out(p, dQKV, 0.125f);
          W2S_TRY(add_gemm(ls + "attn_dq", p));
        }
        {
          GemmProblem p = from_scores(adST);                     // dK = scale dS^T Q
          w_transposed(p, aQT);
          qkv_out(p, dQKV + H, 0.125f);
          W2S_TRY(add_gemm(ls + "attn_dk", p));
        }
        {
          GemmProblem p = from_scores(aPT);                      // dV = P^T dO
          w_transposed(p, adOT);
          qkv_out(p, dQKV + 2 * H, 1.0f);
          W2S_TRY(add_gemm(ls + "attn_dv", p));
        }
      }
      return "";
    };
    {
      const float* lg = plan->logits;
      const int ldl = h->head_ldl, V = c.vocab_size;
      const bf16* hw = h->head_w;
      const int* fr = plan->frames;
      float* seed = stable ? dT : dA;
      // per-row target frame (w2s_grad_waveforms), or an upstream gradient over all frames (w2s_vjp_waveforms)
      add("head_bwd", [=](cudaStream_t s) {
        if (hh->grad_gout) return launch_head_vjp(lg, ldl, V, hw, nn, T, H, hh->grad_gout, seed, hh->grad_out_val, s);
        return launch_head_bwd(lg, ldl, V, hw, nn, T, H, fr, seed, hh->grad_out_val, s);
      });
      if (stable) {   // final encoder LayerNorm: dA = d (residual stream after the last layer)
        const float* g = h->enc_ln_g;
        const float* x = lb[NL - 1].s2;
        add("encoder_ln_bwd", [=](cudaStream_t s) { return launch_ln_bwd(dT, x, 1, rows, H, g, eps, nullptr, dA, nullptr, s); });
      }
      W2S_TRY(snap("layer" + std::to_string(NL), dA, sizeof(float) * rows * H));
    }
    for (int l = NL - 1; l >= 0 && stable; --l) {
      // invariant: dA = d s2_l (gradient of the residual stream leaving layer l)
      const LayerW& w = h->layers[l];
      const GradW& gw = h->gradw[l];
      const GradLayerBuf B = lb[l];
      const std::string ls = "B" + std::to_string(l) + ".";
      const float* r1 = l == 0 ? pre0 : lb[l - 1].s2;
      add(ls + "cast", [=](cudaStream_t s) { return launch_grad_cast(dA, nullptr, dS16, rows * H, s); });
      {
        GemmProblem p = PlanBuilder::plain(dS16, rows, H, gw.w2T, I);
        p.epi.out = dF;
        W2S_TRY(add_gemm(ls + "ffn2_bwd", p));
        const bf16* uu = B.u;
        add(ls + "gelu_bwd", [=](cudaStream_t s) { return launch_gelu_bwd(uu, dF, rows * I, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(dF, rows, I, gw.w1T, H);
        p.epi.out = dT; p.epi.out_fp32 = 1;
        W2S_TRY(add_gemm(ls + "ffn1_bwd", p));
      }
      {   // d s1 = d s2 + LN2^T (d h1)
        const float* g = w.ln2_g;
        const float* x = B.s1;
        add(ls + "ln2_bwd", [=](cudaStream_t s) { return launch_ln_bwd(dT, x, 1, rows, H, g, eps, dA, dS, dS16, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(dS16, rows, H, gw.woT, H);
        p.epi.out = dC;
        W2S_TRY(add_gemm(ls + "out_proj_bwd", p));
      }
      W2S_TRY(add_attention_bwd(ls, B));
      {
        GemmProblem p = PlanBuilder::plain(dQKV, rows, 3 * H, gw.wqkvT, H);
        p.epi.out = dT; p.epi.out_fp32 = 1;
        W2S_TRY(add_gemm(ls + "qkv_bwd", p));
      }
      {   // d r1 = d s1 + LN1^T (d hb)
        const float* g = w.ln1_g;
        add(ls + "ln1_bwd", [=](cudaStream_t s) { return launch_ln_bwd(dT, r1, 1, rows, H, g, eps, dS, dA, nullptr, s); });
      }
      W2S_TRY(snap("layer" + std::to_string(l), dA, sizeof(float) * rows * H));
    }
    for (int l = NL - 1; l >= 0 && !stable; --l) {
      const LayerW& w = h->layers[l];
      const GradW& gw = h->gradw[l];
      const GradLayerBuf B = lb[l];
      const std::string ls = "B" + std::to_string(l) + ".";
      {
        const float* g = w.ln2_g;
        const float* x = B.s2;
        add(ls + "ln2_bwd", [=](cudaStream_t s) { return launch_ln_bwd(dA, x, 1, rows, H, g, eps, nullptr, dS, dS16, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(dS16, rows, H, gw.w2T, I);
        p.epi.out = dF;
        W2S_TRY(add_gemm(ls + "ffn2_bwd", p));
        const bf16* uu = B.u;
        add(ls + "gelu_bwd", [=](cudaStream_t s) { return launch_gelu_bwd(uu, dF, rows * I, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(dF, rows, I, gw.w1T, H);
        p.epi.residual = dS; p.epi.res_fp32 = 1; p.epi.out = dA; p.epi.out_fp32 = 1;
        W2S_TRY(add_gemm(ls + "ffn1_bwd", p));
      }
      {
        const float* g = w.ln1_g;
        const float* x = B.s1;
        add(ls + "ln1_bwd", [=](cudaStream_t s) { return launch_ln_bwd(dA, x, 1, rows, H, g, eps, nullptr, dS, dS16, s); });
      }
      {
        GemmProblem p = PlanBuilder::plain(dS16, rows, H, gw.woT, H);
        p.epi.out = dC;
        W2S_TRY(add_gemm(ls + "out_proj_bwd", p));
      }
      W2S_TRY(add_attention_bwd(ls, B));
      {
        GemmProblem p = PlanBuilder::plain(dQKV, rows, 3 * H, gw.wqkvT, H);
        p.epi.residual = dS; p.epi.res_fp32 = 1; p.epi.out = dA; p.epi.out_fp32 = 1;
        W2S_TRY(add_gemm(ls + "qkv_bwd", p));
      }
      W2S_TRY(snap("layer" + std::to_string(l), dA, sizeof(float) * rows * H));
    }
    {
      // encoder input: hb0 = LN(pre0), pre0 = h0 + gelu(pos_conv(h0))
      const float* g = h->enc_ln_g;
      if (stable)   // the residual stream enters the first layer un-normalised: d pre0 = dA
        add("pre0_copy", [=](cudaStream_t s) -> std::string {
          W2S_CUDA_OK(cudaMemcpyAsync(dS, dA, sizeof(float) * rows * H, cudaMemcpyDeviceToDevice, s));
          return "";
        });
      else
        add("encoder_ln_bwd", [=](cudaStream_t s) { return launch_ln_bwd(dA, pre0, 1, rows, H, g, eps, nullptr, dS, nullptr, s); });
      add("pos_gelu_bwd", [=](cudaStream_t s) { return launch_grad_cast(dS, upos, dS16, rows * H, s); });
      add("pos_pad_bwd", [=](cudaStream_t s) { return launch_pos_pad(dS16, nn, T, H, G, kp, hp, s, kp / 2 - 1); });
      EpiParams e;
      e.act = ACT_NONE; e.residual = dS; e.res_fp32 = 1; e.out = dA; e.out_fp32 = 1;
      e.ldg = cpg; e.ldb = (long long)T * H; e.ldm = H;
      PosConvPlan* pc = nullptr;
      W2S_TRY(posconv_prepare(hp, h->pos_w_bwd, n, T, H, G, kp, e, h->num_sms, &pc));
      plan->posconv.push_back(pc);
      add("pos_conv_bwd", [=](cudaStream_t s) { return posconv_launch(pc, s); });
      W2S_TRY(snap("h0", dA, sizeof(float) * rows * H));
      add("h0_cast", [=](cudaStream_t s) { return launch_grad_cast(dA, nullptr, dS16, rows * H, s); });
      GemmProblem p = PlanBuilder::plain(dS16, rows, H, h->fp_wT, Cl);
      p.epi.out = dFp; p.epi.out_fp32 = 1;
      W2S_TRY(add_gemm("featproj_bwd", p));
      const float* gl = h->fp_ln_g;
      const bf16* y6 = y[NC - 1];
      bf16* d6 = plan->D[(NC - 1) & 1];
      add("featproj_ln_bwd", [=](cudaStream_t s) { return launch_ln_bwd(dFp, y6, 0, rows, Cl, gl, eps, nullptr, nullptr, d6, s); });
      W2S_TRY(snap("conv" + std::to_string(NC - 1), d6, sizeof(bf16) * rows * Cl));
      const bf16* u6 = u[NC - 1];
      add("conv" + std::to_string(NC - 1) + "_gelu_bwd", [=](cudaStream_t s) { return launch_gelu_bwd(u6, d6, rows * Cl, s); });
    }
    for (int l = NC - 1; l >= 1; --l) {
      // D[l & 1] holds d u_l [n, T_l, C_l]; contraction with the transposed filters, then gather to d u_(l-1)
      const int Cin = c.conv_dim[l - 1], Cout = c.conv_dim[l], kw = c.conv_kernel[l], st = c.conv_stride[l];
      const int Tin = Tl[l - 1], Tout = Tl[l];
      bf16* dul = plan->D[l & 1];
      bf16* dprev = plan->D[(l - 1) & 1];
      if (layer) {   // through the per-frame LayerNorm: d (conv + bias) from d u_l, in place
        const bf16* cc = cpre[l];
        const float* lg = h->conv_ln_g[l];
        const long long lrows = (long long)n * Tout;
        add("conv" + std::to_string(l) + "_ln_bwd",
            [=](cudaStream_t s) { return launch_ln_bwd(dul, cc, 0, lrows, Cout, lg, 1e-5f, nullptr, nullptr, dul, s, 0); });
      }
      GemmProblem p = PlanBuilder::plain(dul, (long long)n * Tout, Cout, h->conv_wT[l], kw * Cin);
      p.epi.out = dcol;
      W2S_TRY(add_gemm("conv" + std::to_string(l) + "_bwd", p));
      const bf16* up = u[l - 1];
      add("conv" + std::to_string(l) + "_gather",
          [=](cudaStream_t s) { return launch_conv_gather(dcol, nn, Tin, Tout, Cin, kw, st, up, dprev, s); });
      W2S_TRY(snap("convu" + std::to_string(l - 1), dprev, sizeof(bf16) * (size_t)n * Tin * Cin));
    }
    {
      const bf16* du0 = plan->D[0];
      const bf16* u0 = u[0];
      const int T0 = Tl[0], kw = c.conv_kernel[0], st = c.conv_stride[0];
      const float *w0 = h->conv0_w, *gam = h->norm0_g, *bet = h->norm0_b;
      const long long LL = L;
      add("conv0_bwd", [=](cudaStream_t s) {
        if (layer) return launch_conv0_ln_bwd(du0, u0, ln_rstd0, nn, LL, T0, C0, kw, st, w0, gam, bet, gtap, hh->grad_out, LL, s);
        return launch_conv0_bwd(du0, u0, nn, LL, T0, C0, kw, st, w0, gn_a, gam, bet, m12, gtap, hh->grad_out, LL, s);
      });
    }
    return "";
  }
};

std::string get_grad_plan(w2s_handle* h, int n, long long L, GradPlan** out) {
  if (h->grad_L != L || h->grad_debug_built != h->grad_debug) {
    cudaDeviceSynchronize();
    h->grad_plans.clear();
    h->grad_L = L;
    h->grad_debug_built = h->grad_debug;
  }
  auto it = h->grad_plans.find(n);
  if (it != h->grad_plans.end()) {
    *out = it->second.get();
    return "";
  }
  std::shared_ptr<GradPlan> pl(new GradPlan());
  pl->n = n;
  GradBuilder b{h, pl.get(), n, L};
  b.debug = h->grad_debug;
  W2S_TRY(b.build());
  *out = pl.get();
  h->grad_plans[n] = pl;
  return "";
}

// rows in tiles of `tile`; per tile: argument block, target frames, forward + backward
std::string run_grad(w2s_handle* h, const float* x, long long ld, long long L, int64_t n, const int32_t* frames_host,
                     const float* gout, float* grad, float* out_val, cudaStream_t s) {
  W2S_TRY(grad_supported(h));
  W2S_TRY(grad_prepare_weights(h));
  const int64_t T = num_frames(h->cfg, L, nullptr);
  if (T <= 0) return "clip shorter than the conv receptive field";
  if (frames_host) {
    for (int64_t i = 0; i < n; ++i)
      if (frames_host[i] < 0 || frames_host[i] >= T)
        return "target frame " + std::to_string(frames_host[i]) + " outside the clip's " + std::to_string(T) + " frames";
    h->grad_frames_host.assign(frames_host, frames_host + n);
  }
  const int tile = h->grad_tile;
  for (int64_t k0 = 0; k0 < n; k0 += tile) {
    const int nt = (int)((n - k0) < tile ? (n - k0) : tile);
    GradPlan* pl = nullptr;
    W2S_TRY(get_grad_plan(h, nt, L, &pl));
    DynArgs d{};
    d.x = x + k0 * ld; d.ld = ld;
    set_dyn_kernel<<<1, 1, 0, s>>>(h->dyn_dev, d);
    if (frames_host)
      W2S_CUDA_OK(cudaMemcpyAsync(pl->frames, h->grad_frames_host.data() + k0, sizeof(int) * nt, cudaMemcpyHostToDevice, s));
    h->grad_out = grad + k0 * L;
    h->grad_gout = gout ? gout + k0 * T : nullptr;
    h->grad_out_val = out_val ? out_val + k0 * (gout ? T : 1) : nullptr;
    for (const Step& st : pl->steps) {
      ProfRec rec;
      if (h->profiling) {
        rec.name = "grad." + st.name;
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
    h->launches += (long long)pl->steps.size() + 1;
  }
  return "";
}
