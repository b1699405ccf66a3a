/*
 * w2s.h -- C ABI of the B200-native masked-coalition evaluation path ("wave-to-scores").
 *
 * Drop-in boundary for ONE path of HagenMarin/SHAP-Transformer-ASR: the model-evaluation
 * callback an explainer invokes on masked audio coalitions.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repo; HF = the
 * third-party `transformers` package the reference imports).
 *
 * Conventions: extern "C", plain pointers and sizes, `int` status (0 = ok, non-zero =
 * error; text via w2s_last_error) -- no C++ exceptions and no torch types cross this
 * boundary.  The caller owns every buffer it passes; the handle owns weights, workspace
 * and plans.  Calls are asynchronous on the given CUDA stream (cudaStream_t passed as
 * void*), with no hidden device synchronisation unless stated.  One handle per
 * (device, stream); a handle is not re-entrant.  There is no CPU fallback: every entry
 * point fails with an error when no sm_100 device is present.
 */
#ifndef W2S_H_
#define W2S_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct w2s_handle w2s_handle;

#define W2S_MAX_CONV_LAYERS 8

/* Output reductions (w2s_set_targets `mode`). */
enum {
  W2S_OUT_MAX     = 0, /* max logit per frame           -> [n, T']  shap_calculation.py:50            */
  W2S_OUT_LOGIT   = 1, /* logits[:, t_d, tok_d]          -> [n, D]   feasability_tests/w2v2conformer.py:40-42 */
  W2S_OUT_LOGPROB = 2, /* log_softmax(logits)[t_d,tok_d] -> [n, D]   north-star per-character CTC output      */
  W2S_OUT_MEAN    = 3, /* mean over vocab and time       -> [n, 1]   lime_shap_wav2vec2_comparison.py:68-70   */
  W2S_OUT_LOGITS  = 4  /* raw logits                     -> [n, T', V]  (target selection, parity tests)      */
};

/* w2s_config.flags */
enum {
  W2S_FLAG_VALIDATE_GEMM = 1,  /* run every contraction on the CUDA-core validation kernels (same bf16 data) */
  W2S_FLAG_VALIDATE_ATTN = 2,  /* run attention on the CUDA-core validation kernel                            */
  W2S_FLAG_FP32_PRELN    = 4,  /* post-LN models: fp32 (instead of bf16) pre-LayerNorm tensors -- A/B measurement */
  W2S_FLAG_NO_GRAPH      = 16  /* launch the kernels of a tile one by one instead of replaying a CUDA graph        */
};

/* Mirrors transformers.Wav2Vec2Config / Wav2Vec2ConformerConfig (the objects the reference
 * obtains from from_pretrained at shap_calculation.py:218-219, w2v2conformer.py:58-59). */
typedef struct {
  int32_t kind;                 /* 0 = Wav2Vec2ForCTC, 1 = Wav2Vec2ConformerForCTC */
  int32_t num_conv_layers;
  int32_t conv_dim[W2S_MAX_CONV_LAYERS];
  int32_t conv_kernel[W2S_MAX_CONV_LAYERS];
  int32_t conv_stride[W2S_MAX_CONV_LAYERS];
  int32_t conv_bias;            /* 0 / 1 */
  int32_t feat_extract_norm;    /* 0 = "group", 1 = "layer" */
  int32_t hidden_size;
  int32_t num_hidden_layers;
  int32_t num_attention_heads;
  int32_t intermediate_size;
  int32_t num_conv_pos_embeddings;
  int32_t num_conv_pos_embedding_groups;
  int32_t vocab_size;
  float   layer_norm_eps;
  int32_t do_stable_layer_norm; /* 0 / 1 */
  int32_t position_embeddings_type; /* conformer: 0 none, 1 "relative", 2 "rotary" */
  int32_t conv_depthwise_kernel_size;
  int32_t hidden_act;           /* 0 = "gelu" (erf form), 1 = "swish" */
  int32_t rotary_embedding_base;
  int32_t max_batch;            /* coalitions evaluated per internal batch tile (0 = choose) */
  int32_t flags;                /* W2S_FLAG_* */
} w2s_config;

/* Replaces model construction + .to(device) (shap_calculation.py:218-219,258).
 * `names[i]` are HF state_dict keys ("wav2vec2.feature_extractor.conv_layers.0.conv.weight", ...,
 * "lm_head.bias"; the weight-normed pos-conv must be passed folded as
 * "<prefix>encoder.pos_conv_embed.conv.weight"), `ptrs[i]` fp32 DEVICE pointers in the HF layout with
 * `numels[i]` elements.  Weights are converted/re-laid-out into handle-owned storage; the caller may
 * free its copies after the call returns (the call synchronises the device once). */
int w2s_create(const w2s_config* cfg, const char* const* names, const float* const* ptrs,
               const int64_t* numels, int n_weights, int device, w2s_handle** out);

void w2s_destroy(w2s_handle* h);

/* NULL handle -> message of the last failed w2s_create on this thread. */
const char* w2s_last_error(const w2s_handle* h);

/* L -> T' (HF wav2vec2/modeling_wav2vec2.py:1005-1024). */
int64_t w2s_num_frames(const w2s_handle* h, int64_t num_samples);

/* Replaces the masker's captured state (feasability_tests/conformer_test.ipynb:138-141): the already
 * normalised clip x[L] (fp32, device; copied), the segment partition seg_bounds[M+1] (host, ascending,
 * seg_bounds[0]=0, seg_bounds[M]=L) and the baseline fill value (0.0 in the reference). */
int w2s_set_clip(w2s_handle* h, const float* x_dev, int64_t L, const int32_t* seg_bounds_host, int M,
                 float baseline, void* stream);

/* Replaces timestep_to_explain / token_id_to_explain (w2v2conformer.py:93-110) and the character-frame
 * list of visualization.py:319-327: D (frame, token) pairs (host arrays, copied before the call returns; ignored
 * for MAX/MEAN/LOGITS).  The upload is ordered on `stream` behind evaluations already queued there. */
int w2s_set_targets(w2s_handle* h, const int32_t* frame_idx_host, const int32_t* token_idx_host, int D,
                    int mode, void* stream);

/* Number of fp32 values one evaluated row produces under the current mode for clips of L samples. */
int64_t w2s_out_width(const w2s_handle* h, int64_t num_samples);

/* The fused masker + model callable: coalition bit matrix z_bits[K, ceil(M/32)] (device; bit m of row k
 * set = segment m KEPT, clear = replaced by the baseline) -> out[K, width] fp32 (device). */
int w2s_eval(w2s_handle* h, const uint32_t* z_bits_dev, int64_t K, float* out_dev, void* stream);

/* Replaces ModelWrapper.forward (shap_calculation.py:31-50), predict_function
 * (w2v2conformer.py:116-131) and lime_predict_fn (lime_shap_wav2vec2_comparison.py:60-71) for explicit
 * waveform rows x[n, L] (fp32, device, row stride `ld` elements) -> out[n, width] fp32 (device). */
int w2s_eval_waveforms(w2s_handle* h, const float* x_dev, int64_t n, int64_t L, int64_t ld, float* out_dev,
                       void* stream);

/* Expected-gradients path -- the reference's production explainer differentiates ModelWrapper's output
 * (shap.GradientExplainer(wrapped_model, ...).shap_values, shap_calculation.py:133,162; per output index
 * `outputs = self.model(*X); outputs[:, idx]` then autograd, conformer_test.ipynb:95).  For explicit waveform rows
 * x[n, L] (fp32, device, row stride `ld`) and one output frame per row (host array): out[r] = max_v logits[r,
 * frames[r], v] (shap_calculation.py:50) and grad[r, :] = d out[r] / d x[r, :] (fp32 [n, L], device).  Input
 * gradients only -- no weight gradients.  First slice: Wav2Vec2ForCTC with the group-norm front end and a post-LN
 * encoder (wav2vec2-base / -large); other configurations return an error. */
int w2s_grad_waveforms(w2s_handle* h, const float* x_dev, int64_t n, int64_t L, int64_t ld,
                       const int32_t* frames_host, float* grad_dev, float* out_dev, void* stream);
/* The same backward pass seeded with an upstream gradient over ALL output frames -- the vector-Jacobian product autograd
 * asks of ModelWrapper.forward (B1 in SURVEY.md 8b: "autograd graph required"): gout[n, T'] (fp32, device) ->
 * grad[r, :] = sum_t gout[r, t] d out[r, t] / d x[r, :], out[n, T'] = the max logits (may be NULL).  With it the
 * reference's shap.GradientExplainer(wrapped_model, ...) runs unmodified on callbacks.ModelWrapper. */
int w2s_vjp_waveforms(w2s_handle* h, const float* x_dev, int64_t n, int64_t L, int64_t ld, const float* gout_dev,
                      float* grad_dev, float* out_dev, void* stream);
/* Test hooks of the gradient path.  `on` bit 0: keep snapshots of the gradient after every encoder layer / conv layer
 * ("layer<l>", "h0", "conv<l>", "convu<l>", forward activations "f.*"); bit 1: run attention backward on the CUDA-core
 * cross-check kernels, bit 2: as batched tensor-core contractions + row kernels, instead of the fused tcgen05 kernel.  w2s_grad_peek copies a snapshot out (returns bytes copied, 0 = unknown name, < 0 = buffer
 * too small by that many bytes). */
int w2s_grad_debug(w2s_handle* h, int on);
/* DeepLIFT handler rules of feasability_tests/custom_shap_handlers.py:35-80 for the backward pass of w2s_grad_waveforms.
 * With rules != 0 every call carries PAIRED rows: rows [0, n/2) are explained inputs, row n/2 + r is the reference
 * (background) row of row r; n even and at most 32.  Reference rows receive no output seed: their gradient rows are zero.
 *   bit 0: SiLU activations use the rescale multiplier (act(x) - act(ref)) / (x - ref) (shap's nonlinear_1d; the ordinary
 *          derivative where |x - ref| < 1e-6); LayerNorm / GroupNorm are linear_1d, i.e. the ordinary gradient, like every
 *          module shap has no handler for (GELUActivation, attention);
 *   bit 1: the reference's GLU handler as written: GLU inputs that differ from the reference by >= 1e-6 receive
 *          grad_output * 5e-6 (a placeholder its author left in; off by default).
 * rules = 0 restores the plain gradient. */
int w2s_grad_rules(w2s_handle* h, int rules);
int64_t w2s_grad_peek(w2s_handle* h, const char* name, void* dst_dev, int64_t max_bytes, void* stream);

/* Standalone masker (conformer_test.ipynb:138-141 semantics with keep-bits): materialise
 * out[K, L] = keep ? x : baseline for the clip set by w2s_set_clip. */
int w2s_mask(w2s_handle* h, const uint32_t* z_bits_dev, int64_t K, float* out_dev, void* stream);

/* KernelSHAP constrained weighted least squares (shap KernelExplainer.solve with l1_reg=False; SURVEY.md
 * Appendix A step 6): z_bits[K, ceil(M/32)], kernel weights w[K] (fp64), y[K, D] fp32, fx[D], fnull[D]
 * (fp64) -> phi[M, D] fp64, all device pointers.  `status_dev` (int32, device, required) receives 0 when the
 * normal matrix was factored by Cholesky; 2 when the design is rank deficient (fewer distinct coalitions than
 * features, or a feature that never varies) and the minimum-norm least-squares solution was computed instead
 * -- what shap's `solve` returns through its numpy.linalg.lstsq fallback -- by conjugate gradients on the device;
 * 1 if that iteration did not converge (phi is then not to be used). */
int w2s_wls(w2s_handle* h, const uint32_t* z_bits_dev, const double* w_dev, const float* y_dev, int64_t K,
            int M, int D, const double* fx_dev, const double* fnull_dev, double* phi_dev, int32_t* status_dev,
            void* stream);

/* The random half of shap.KernelExplainer's coalition sampler (third-party, never called by the reference: SURVEY.md
 * Appendix A step 4), on the HOST: for every entry of `ind_set` (the subset sizes drawn by np.random.choice) one
 * legacy np.random.permutation(M) on the MT19937 state `mt_key[624]` / `*mt_pos` taken from np.random.get_state()
 * (updated in place, to be handed back with np.random.set_state), de-duplicated rows appended to the bit-packed
 * matrix z_words[cap_rows, ceil(M/32)] from row `added` on, with their hit counts in weights[] and the paired
 * complement rows where shap adds them.  Returns the new row count, or -1 on a bad argument / overflow.
 * No device is touched. */
int64_t w2s_sample_rows(int M, int n_full, int n_paired, const int64_t* ind_set, int64_t n_ind,
                        int64_t samples_left, int64_t added, uint32_t* mt_key, int32_t* mt_pos,
                        uint32_t* z_words, double* weights, int64_t cap_rows, int64_t* ind_used);

/* Test / profiling hooks for the individual kernels (used by tests/ and bench.py only). */
/* out[M, N] = act(a[M, K] w[N, K]^T + bias) * alpha + residual (residual == out with fp32 data: in-place
 * accumulation, which the CTA-pair kernel performs with TMA reduce-add). */
int w2s_debug_gemm(int use_tcgen05, const void* a_bf16, const void* w_bf16, const float* bias,
                   const void* residual, int res_fp32, float alpha, void* out, int M, int N, int K, int act,
                   int out_fp32, void* stream);
/* Per-launch CUDA-event profile on the launching stream (bench.py roofline leg): enable, run evaluations,
 * then read per launch class (newline-separated names) total milliseconds, algorithmic FLOPs / bytes and
 * launch counts.  Returns the number of classes, or -1 if `names_cap` is too small. */
int w2s_profile_enable(w2s_handle* h, int on);
int64_t w2s_profile_read(w2s_handle* h, char* names, int64_t names_cap, double* ms, double* flops, double* bytes,
                         int64_t* counts, int64_t max_entries);
int w2s_kernel_count(const w2s_handle* h, int64_t* launches_per_batch, int64_t* batch_tile);
/* kernel launches issued by w2s_eval / w2s_eval_waveforms on this handle since w2s_create (counted at launch,
 * graph-replayed kernels included) */
int64_t w2s_launch_count(const w2s_handle* h);
/* algorithmic FLOPs (2*MAC) of one coalition forward for clips of L samples */
double w2s_flops_per_forward(const w2s_handle* h, int64_t num_samples);

#ifdef __cplusplus
}
#endif
#endif /* W2S_H_ */
