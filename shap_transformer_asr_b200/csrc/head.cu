// K9: lm_head (H -> V) fused with the output reduction of the callback.
//   max logit per frame            shap_calculation.py:50
//   logit at (frame, token)        feasability_tests/w2v2conformer.py:40-42
//   log-softmax at (frame, token)  north-star per-character CTC log-probability
//   mean over vocab and time       feasability_tests/lime_shap_wav2vec2_comparison.py:68-70
// One warp per (row, frame) work item: lane = vocabulary entry, so log-softmax and gather are warp
// shuffles (HF wav2vec2/modeling_wav2vec2.py:1705-1708 for the linear layer).
#include "kernels.cuh"
#include "../../include/w2s.h"

namespace w2s {

constexpr int HEAD_MAXJ = 4;  // V <= 128

__global__ void __launch_bounds__(256) head_kernel(const HeadParams p, long long items) {
  extern __shared__ uint8_t smem[];
  const int ldw = p.H + 2;  // bf16 elements; (H+2)/2 odd -> conflict-free row stride
  __nv_bfloat16* ws = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* hs = ws + (size_t)p.V * ldw;  // 8 warps x H
  for (int i = threadIdx.x; i < p.V * p.H; i += blockDim.x) {
    const int v = i / p.H, k = i - v * p.H;
    ws[v * ldw + k] = p.w[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* hrow = hs + (size_t)warp * p.H;
  const bool targeted = (p.mode == W2S_OUT_LOGIT || p.mode == W2S_OUT_LOGPROB);
  const int per = targeted ? p.D : p.T;
  for (long long it = (long long)blockIdx.x * 8 + warp; it < items; it += (long long)gridDim.x * 8) {
    const int b = (int)(it / per);
    const int d = (int)(it - (long long)b * per);
    const int t = targeted ? p.frames[d] : d;
    const __nv_bfloat16* src = p.h + ((long long)b * p.T + t) * p.H;
    __syncwarp();
    for (int k = lane * 2; k < p.H; k += 64)
      *reinterpret_cast<uint32_t*>(hrow + k) = *reinterpret_cast<const uint32_t*>(src + k);
    __syncwarp();
    float logit[HEAD_MAXJ];
#pragma unroll
    for (int j = 0; j < HEAD_MAXJ; ++j) {
      const int v = lane + 32 * j;
      float acc = 0.f;
      if (v < p.V) {
        const __nv_bfloat16* wr = ws + (size_t)v * ldw;
        for (int k = 0; k < p.H; k += 2) {
          const uint32_t hh = *reinterpret_cast<const uint32_t*>(hrow + k);
          const uint32_t ww = *reinterpret_cast<const uint32_t*>(wr + k);
          acc = fmaf(bf16_lo(hh), bf16_lo(ww), acc);
          acc = fmaf(bf16_hi(hh), bf16_hi(ww), acc);
        }
        acc += p.bias[v];
      }
      logit[j] = acc;
    }
    if (p.mode == W2S_OUT_LOGITS) {
#pragma unroll
      for (int j = 0; j < HEAD_MAXJ; ++j) {
        const int v = lane + 32 * j;
        if (v < p.V) p.out[((long long)b * p.T + t) * p.V + v] = logit[j];
      }
      continue;
    }
    float mx = -INFINITY, sm = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_MAXJ; ++j)
      if (lane + 32 * j < p.V) {
        mx = fmaxf(mx, logit[j]);
        sm += logit[j];
      }
    mx = warp_max(mx);
    if (p.mode == W2S_OUT_MAX) {
      if (lane == 0) p.out[(long long)b * p.T + t] = mx;
    } else if (p.mode == W2S_OUT_MEAN) {
      sm = warp_sum(sm);
      if (lane == 0) atomicAdd(p.out + b, sm / ((float)p.V * (float)p.T));
    } else {
      const int tok = p.tokens[d];
      float sel = 0.f;
#pragma unroll
      for (int j = 0; j < HEAD_MAXJ; ++j) {
        const float cand = __shfl_sync(0xffffffffu, logit[j], tok & 31);
        if ((tok >> 5) == j) sel = cand;
      }
      if (p.mode == W2S_OUT_LOGPROB) {
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < HEAD_MAXJ; ++j)
          if (lane + 32 * j < p.V) se += __expf(logit[j] - mx);
        se = warp_sum(se);
        sel -= mx + __logf(se);
      }
      if (lane == 0) p.out[(long long)b * p.D + d] = sel;
    }
  }
}

// second stage of the tensor-core head: logits[rows, V] (fp32, bias included) -> reduction; one warp per work item
__global__ void __launch_bounds__(256) head_reduce_kernel(const float* __restrict__ logits, const HeadParams p,
                                                           long long items) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool targeted = (p.mode == W2S_OUT_LOGIT || p.mode == W2S_OUT_LOGPROB);
  const int per = targeted ? p.D : p.T;
  for (long long it = (long long)blockIdx.x * 8 + warp; it < items; it += (long long)gridDim.x * 8) {
    const int b = (int)(it / per);
    const int d = (int)(it - (long long)b * per);
    const int t = targeted ? p.frames[d] : d;
    const float* row = logits + ((long long)b * p.T + t) * p.V;
    float logit[HEAD_MAXJ];
    float mx = -INFINITY, sm = 0.f;
#pragma unroll
    for (int j = 0; j < HEAD_MAXJ; ++j) {
      const int v = lane + 32 * j;
      logit[j] = v < p.V ? row[v] : -INFINITY;
      if (v < p.V) {
        mx = fmaxf(mx, logit[j]);
        sm += logit[j];
      }
    }
    mx = warp_max(mx);
    if (p.mode == W2S_OUT_MAX) {
      if (lane == 0) p.out[(long long)b * p.T + t] = mx;
    } else if (p.mode == W2S_OUT_MEAN) {
      sm = warp_sum(sm);
      if (lane == 0) atomicAdd(p.out + b, sm / ((float)p.V * (float)p.T));
    } else {
      const int tok = p.tokens[d];
      float sel = 0.f;
#pragma unroll
      for (int j = 0; j < HEAD_MAXJ; ++j) {
        const float cand = __shfl_sync(0xffffffffu, logit[j], tok & 31);
        if ((tok >> 5) == j) sel = cand;
      }
      if (p.mode == W2S_OUT_LOGPROB) {
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < HEAD_MAXJ; ++j)
          if (lane + 32 * j < p.V) se += __expf(logit[j] - mx);
        se = warp_sum(se);
        sel -= mx + __logf(se);
      }
      if (lane == 0) p.out[(long long)b * p.D + d] = sel;
    }
  }
}

std::string launch_head_reduce(const float* logits, const HeadParams& p, cudaStream_t s) {
  if (p.V > 32 * HEAD_MAXJ) return "head: vocab_size > 128 not supported";
  const bool targeted = (p.mode == W2S_OUT_LOGIT || p.mode == W2S_OUT_LOGPROB);
  if (targeted && (p.D <= 0 || !p.frames || !p.tokens)) return "head: targets not set (w2s_set_targets)";
  const long long items = (long long)p.n * (targeted ? p.D : p.T);
  if (items == 0) return "";
  if (p.mode == W2S_OUT_MEAN) W2S_CUDA_OK(cudaMemsetAsync(p.out, 0, sizeof(float) * p.n, s));
  long long blocks = (items + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  W2S_CUDA_OK(launch_pdl(head_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, s, 1, logits, p, items));
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

static bool g_head_attr = false;

std::string launch_head(const HeadParams& p, cudaStream_t s) {
  if (p.V > 32 * HEAD_MAXJ) return "head: vocab_size > 128 not supported";
  if (p.H % 2) return "head: hidden size must be even";
  const bool targeted = (p.mode == W2S_OUT_LOGIT || p.mode == W2S_OUT_LOGPROB);
  if (targeted && (p.D <= 0 || !p.frames || !p.tokens)) return "head: targets not set (w2s_set_targets)";
  const long long items = (long long)p.n * (targeted ? p.D : p.T);
  if (items == 0) return "";
  const size_t smem = ((size_t)p.V * (p.H + 2) + 8 * (size_t)p.H) * sizeof(__nv_bfloat16);
  if (!g_head_attr) {
    W2S_CUDA_OK(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    g_head_attr = true;
  }
  if (smem > 200 * 1024) return "head: vocab * hidden too large for shared memory";
  if (p.mode == W2S_OUT_MEAN) W2S_CUDA_OK(cudaMemsetAsync(p.out, 0, sizeof(float) * p.n, s));
  long long blocks = (items + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  head_kernel<<<(unsigned)blocks, 256, smem, s>>>(p, items);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

}  // namespace w2s
