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
// CTA = (64-frame tile, 64-channel block, row); the (64 + k - 1) x 64 input window is staged in shared memory.
template <int KMAX>
__global__ void __launch_bounds__(256) depthwise_kernel(const __nv_bfloat16* __restrict__ in, int T, int H, int k,
                                                         const float* __restrict__ w, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int act,
                                                         __nv_bfloat16* __restrict__ out) {
  __shared__ float xs[64 + KMAX][64 + 1];
  const int t0 = blockIdx.x * 64, c0 = blockIdx.y * 64, b = blockIdx.z;
  const int pad = (k - 1) / 2;
  const int nrow = 64 + k - 1;
  for (int i = threadIdx.x; i < nrow * 64; i += blockDim.x) {
    const int r = i >> 6, c = i & 63;
    const int t = t0 + r - pad;
    float v = 0.f;
    if (t >= 0 && t < T && c0 + c < H) v = __bfloat162float(in[((long long)b * T + t) * H + c0 + c]);
    xs[r][c] = v;
  }
  __syncthreads();
  const int c = threadIdx.x & 63;
  if (c0 + c >= H) return;
  float wr[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) wr[j] = j < k ? __ldg(w + (long long)(c0 + c) * k + j) : 0.f;
  const float sc = scale[c0 + c], sh = shift[c0 + c];
  for (int r = threadIdx.x >> 6; r < 64; r += 4) {
    const int t = t0 + r;
    if (t >= T) break;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j < k) acc = fmaf(wr[j], xs[r + j][c], acc);
    out[((long long)b * T + t) * H + c0 + c] = __float2bfloat16_rn(apply_act(fmaf(acc, sc, sh), act));
  }
}
std::string launch_depthwise(const __nv_bfloat16* in, int B, int T, int H, int k, const float* w, const float* scale,
                             const float* shift, int act, __nv_bfloat16* out, cudaStream_t s) {
  if (k > 32 || (k & 1) == 0) return "depthwise conv: kernel size must be odd and <= 31";
  if (B == 0) return "";
  dim3 grid((T + 63) / 64, (H + 63) / 64, B);
  depthwise_kernel<32><<<grid, 256, 0, s>>>(in, T, H, k, w, scale, shift, act, out);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

// rotary: out[b, t, h, :] = x * cos(t) + rotate_half(x) * sin(t), angle[t, i] = t * base^(-2 (i mod hd/2) / hd)
__global__ void __launch_bounds__(256) rotary_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int T, int H,
                                                      int hd, float log2_base, __nv_bfloat16* __restrict__ out) {
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
  out[i] = __float2bfloat16_rn(xv * cs + other * sn);
}
std::string launch_rotary(const __nv_bfloat16* x, long long rows, int T, int H, int hd, int base, __nv_bfloat16* out,
                          cudaStream_t s) {
  if (rows == 0) return "";
  const long long n = rows * H;
  rotary_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, rows, T, H, hd, log2f((float)base), out);
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
