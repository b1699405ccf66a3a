// Conformer-only CUDA-core kernels (HF wav2vec2_conformer/modeling_wav2vec2_conformer.py):
//   depthwise conv k=31 + eval-mode BatchNorm + activation   (:360-417, ConvolutionModule)
//   rotary position rotation of the query/key input            (:489-507)
//   relative position table (sin/cos)                          (:159-205)
#include "kernels.cuh"

namespace w2s {

// scale = gamma / sqrt(var + eps), shift = beta - mean * scale   (BatchNorm1d in eval mode, running statistics)
__global__ void bn_fold_kernel(const float* g, const float* b, const float* mean, const float* var, int n, float eps,
                               float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = g[i] * rsqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = b[i] - mean[i] * s;
  }
}
std::string launch_bn_fold(const float* g, const float* b, const float* mean, const float* var, int n, float eps,
                           float* scale, float* shift, cudaStream_t s) {
  bn_fold_kernel<<<(n + 255) / 256, 256, 0, s>>>(g, b, mean, var, n, eps, scale, shift);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

// out[b, t, c] = act((sum_j w[c, j] in[b, t + j - pad, c]) * scale[c] + shift[c]), zero padding in time.
// One warp = 16 consecutive frames x 64 channels (lane = channel pair on the packed fp32 pipe).  Output-stationary:
// each input row is loaded once (128 B per warp, coalesced) and scattered into the <= 16 accumulators it touches;
// all tap indices are compile-time after unrolling, so filters and accumulators stay in registers.
template <int K>
__global__ void __launch_bounds__(256, 2) depthwise_kernel(const __nv_bfloat16* __restrict__ in, int T, int H,
                                                         const float* __restrict__ w, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int act,
                                                         __nv_bfloat16* __restrict__ out) {
  constexpr int TT = 16, PAD = (K - 1) / 2, ROWS = 8 * TT + K - 1;
  __shared__ uint4 xs[ROWS * 8];   // [ROWS][64 channels] bf16, staged by the whole CTA with 16-byte loads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  const int c_blk = blockIdx.y * 64;
  {
    const int tb = blockIdx.x * 8 * TT - PAD;
    const bool vec_ok = (H % 8 == 0) && (c_blk + 64 <= H);
    for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) {
      const int r = i >> 3, q = i & 7;
      const int ti = tb + r;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (ti >= 0 && ti < T) {
        const __nv_bfloat16* src = in + ((long long)b * T + ti) * H + c_blk + q * 8;
        if (vec_ok) {
          v = *reinterpret_cast<const uint4*>(src);
        } else {
          __nv_bfloat16 tmp[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) tmp[e] = (c_blk + q * 8 + e < H) ? src[e] : __float2bfloat16_rn(0.f);
          v = *reinterpret_cast<uint4*>(tmp);
        }
      }
      xs[i] = v;
    }
  }
  __syncthreads();
  const int t0 = (blockIdx.x * 8 + warp) * TT;
  const int c = c_blk + lane * 2;
  if (t0 >= T || c >= H) return;
  const uint32_t* xw = reinterpret_cast<const uint32_t*>(xs) + warp * TT * 32 + lane;
  float2 wr[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wr[j] = __ldg(reinterpret_cast<const float2*>(w + (long long)j * H + c));   // taps stored [K][H]
  float2 acc[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) acc[t] = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < TT + K - 1; ++r) {
    const uint32_t u = xw[r * 32];   // zero outside [0, T): staged that way
    const float2 x = make_float2(bf16_lo(u), bf16_hi(u));
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const int j = r - t;  // tap index, static after unrolling
      if (j >= 0 && j < K) acc[t] = __ffma2_rn(wr[j], x, acc[t]);
    }
  }
  const float2 sc = make_float2(scale[c], scale[c + 1]), sh = make_float2(shift[c], shift[c + 1]);
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    if (t0 + t < T) {
      const float2 y = __ffma2_rn(acc[t], sc, sh);
      *reinterpret_cast<uint32_t*>(out + ((long long)b * T + t0 + t) * H + c) =
          pack_bf16x2(apply_act(y.x, act), apply_act(y.y, act));
    }
  }
}
std::string launch_depthwise(const __nv_bfloat16* in, int B, int T, int H, int k, const float* w, const float* scale,
                             const float* shift, int act, __nv_bfloat16* out, cudaStream_t s) {
  if (H % 2) return "depthwise conv: channel count must be even";
  if (B == 0) return "";
  dim3 grid((T + 127) / 128, (H + 63) / 64, B);
  switch (k) {
    case 31: W2S_CUDA_OK(launch_pdl(depthwise_kernel<31>, grid, dim3(256), 0, s, 1, in, T, H, w, scale, shift, act, out)); break;
    case 15: W2S_CUDA_OK(launch_pdl(depthwise_kernel<15>, grid, dim3(256), 0, s, 1, in, T, H, w, scale, shift, act, out)); break;
    case 7: W2S_CUDA_OK(launch_pdl(depthwise_kernel<7>, grid, dim3(256), 0, s, 1, in, T, H, w, scale, shift, act, out)); break;
    case 3: W2S_CUDA_OK(launch_pdl(depthwise_kernel<3>, grid, dim3(256), 0, s, 1, in, T, H, w, scale, shift, act, out)); break;
    default: return "depthwise conv: kernel size " + std::to_string(k) + " not instantiated (31, 15, 7, 3)";
  }
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

__global__ void transpose_f32_kernel(const float* src, float* dst, int R, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * C) {
    const int r = i / C, cc = i - r * C;
    dst[(long long)cc * R + r] = src[i];
  }
}
// dst[C][R] = src[R][C]^T (one-off weight re-layout)
std::string launch_transpose_f32(const float* src, float* dst, int R, int C, cudaStream_t s) {
  transpose_f32_kernel<<<(R * C + 255) / 256, 256, 0, s>>>(src, dst, R, C);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

// rotary: out[b, t, h, :] = x * cos(t) + rotate_half(x) * sin(t), angle[t, i] = t * base^(-2 (i mod hd/2) / hd)
__global__ void __launch_bounds__(256) rotary_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int T, int H,
                                                      int hd, float log2_base, __nv_bfloat16* __restrict__ out,
                                                      float sign) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const long long r = i / H;
  const int c = (int)(i - r * H);
  const int t = (int)(r % T);
  const int d = c % hd, half = hd / 2;
  const int fi = d % half;
  const float inv_freq = exp2f(-log2_base * (2.0f * fi) / hd);
  float sn, cs;
  sincosf((float)t * inv_freq, &sn, &cs);
  const float xv = __bfloat162float(x[i]);
  const float other = d < half ? -__bfloat162float(x[i + half]) : __bfloat162float(x[i - half]);
  out[i] = __float2bfloat16_rn(xv * cs + other * (sign * sn));
}
std::string launch_rotary(const __nv_bfloat16* x, long long rows, int T, int H, int hd, int base, __nv_bfloat16* out,
                          cudaStream_t s, int inverse) {
  if (rows == 0) return "";
  const long long n = rows * H;
  rotary_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, rows, T, H, hd, log2f((float)base), out,
                                                                        inverse ? -1.0f : 1.0f);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

// relative position table: row r (0 .. 2T-2) encodes relative position T-1-r; even columns sin, odd columns cos
__global__ void relpos_kernel(int T, int H, __nv_bfloat16* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = (2 * T - 1) * H;
  if (i >= n) return;
  const int r = i / H, c = i - r * H;
  const float pos = (float)(T - 1 - r);
  const float div = expf((float)(c & ~1) * -(logf(10000.0f) / (float)H));
  out[i] = __float2bfloat16_rn((c & 1) ? cosf(pos * div) : sinf(pos * div));
}
std::string launch_relpos(int T, int H, __nv_bfloat16* out, cudaStream_t s) {
  const int n = (2 * T - 1) * H;
  relpos_kernel<<<(n + 255) / 256, 256, 0, s>>>(T, H, out);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

}  // namespace w2s
