// K12: KernelSHAP constrained weighted least squares on the device, fp64.
// Restates shap KernelExplainer.solve with l1_reg=False: the last feature is
// eliminated with the efficiency constraint,
//   X[k, i] = z[k, i] - z[k, M-1]                         (entries in {-1, 0, 1})
//   r[k, d] = y[k, d] - fnull[d] - z[k, M-1] (fx[d] - fnull[d])
//   (X^T W X) w = X^T W r,   phi[:M-1] = w,   phi[M-1] = fx - fnull - sum(w),   |phi| < 1e-10 -> 0
// The normal matrix is SPD for any sampled design with positive kernel weights, so it is factored by
// Cholesky (shap calls numpy.linalg.solve; same solution up to fp64 round-off).
#include "kernels.cuh"

namespace w2s {

__device__ __forceinline__ int zbit(const uint32_t* z, int i) { return (z[i >> 5] >> (i & 31)) & 1; }

// A[i, j] = sum_k w_k X[k,i] X[k,j]   (n = M-1), 16x16 tile per CTA, K walked in chunks staged in smem
__global__ void __launch_bounds__(256) wls_gram_kernel(const uint32_t* __restrict__ zbits, int zwords,
                                                        const double* __restrict__ w, long long K, int M,
                                                        double* __restrict__ A) {
  const int n = M - 1;
  const int i = blockIdx.y * 16 + (threadIdx.x >> 4);
  const int j = blockIdx.x * 16 + (threadIdx.x & 15);
  __shared__ signed char xi[256][16];
  __shared__ signed char xj[256][16];
  __shared__ double ws[256];
  double acc = 0.0;
  for (long long k0 = 0; k0 < K; k0 += 256) {
    const long long k = k0 + threadIdx.x;
    if (k < K) {
      const uint32_t* z = zbits + k * zwords;
      const int last = zbit(z, M - 1);
      for (int c = 0; c < 16; ++c) {
        const int ii = blockIdx.y * 16 + c, jj = blockIdx.x * 16 + c;
        xi[threadIdx.x][c] = ii < n ? (signed char)(zbit(z, ii) - last) : 0;
        xj[threadIdx.x][c] = jj < n ? (signed char)(zbit(z, jj) - last) : 0;
      }
      ws[threadIdx.x] = w[k];
    } else {
      for (int c = 0; c < 16; ++c) xi[threadIdx.x][c] = xj[threadIdx.x][c] = 0;
      ws[threadIdx.x] = 0.0;
    }
    __syncthreads();
    const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
#pragma unroll 8
    for (int kk = 0; kk < 256; ++kk) acc += ws[kk] * (double)(xi[kk][ti] * xj[kk][tj]);
    __syncthreads();
  }
  if (i < n && j < n) A[(long long)i * n + j] = acc;
}

// Bm[i, d] = sum_k w_k X[k,i] r[k,d]
__global__ void __launch_bounds__(256) wls_rhs_kernel(const uint32_t* __restrict__ zbits, int zwords,
                                                       const double* __restrict__ w, const float* __restrict__ y,
                                                       long long K, int M, int D, const double* __restrict__ fx,
                                                       const double* __restrict__ fnull, double* __restrict__ Bm) {
  const int n = M - 1;
  const int d = blockIdx.x * 16 + (threadIdx.x & 15);
  const int i = blockIdx.y * 16 + (threadIdx.x >> 4);
  __shared__ signed char xi[256][16];
  __shared__ signed char lastb[256];
  __shared__ double ws[256];
  const double fn = d < D ? fnull[d] : 0.0;
  const double delta = d < D ? fx[d] - fn : 0.0;
  double acc = 0.0;
  for (long long k0 = 0; k0 < K; k0 += 256) {
    const long long k = k0 + threadIdx.x;
    if (k < K) {
      const uint32_t* z = zbits + k * zwords;
      const int last = zbit(z, M - 1);
      for (int c = 0; c < 16; ++c) {
        const int ii = blockIdx.y * 16 + c;
        xi[threadIdx.x][c] = ii < n ? (signed char)(zbit(z, ii) - last) : 0;
      }
      lastb[threadIdx.x] = (signed char)last;
      ws[threadIdx.x] = w[k];
    } else {
      for (int c = 0; c < 16; ++c) xi[threadIdx.x][c] = 0;
      lastb[threadIdx.x] = 0;
      ws[threadIdx.x] = 0.0;
    }
    __syncthreads();
    const int ti = threadIdx.x >> 4;
    const long long kmax = (K - k0) < 256 ? (K - k0) : 256;
    if (d < D) {
      for (int kk = 0; kk < kmax; ++kk) {
        const int x = xi[kk][ti];
        if (x != 0) {
          const double r = (double)y[(k0 + kk) * D + d] - fn - (double)lastb[kk] * delta;
          acc += ws[kk] * (double)x * r;
        }
      }
    }
    __syncthreads();
  }
  if (i < n && d < D) Bm[(long long)i * D + d] = acc;
}

// In-place Cholesky A = L L^T (lower), one CTA; then forward/back substitution for all D right-hand sides
// (thread = output dimension) and the phi assembly.
__global__ void __launch_bounds__(1024) wls_solve_kernel(double* __restrict__ A, double* __restrict__ Bm, int M,
                                                          int D, const double* __restrict__ fx,
                                                          const double* __restrict__ fnull, double* __restrict__ phi,
                                                          int32_t* __restrict__ status) {
  const int n = M - 1;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      const double djj = A[(long long)j * n + j];
      if (!(djj > 0.0)) bad = 1;
      A[(long long)j * n + j] = sqrt(djj > 0.0 ? djj : 1.0);
    }
    __syncthreads();
    const double ljj = A[(long long)j * n + j];
    for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) A[(long long)i * n + j] /= ljj;
    __syncthreads();
    // trailing update of the lower triangle: A[i, k] -= L[i, j] L[k, j], j < k <= i
    const int rem = n - j - 1;
    for (int idx = threadIdx.x; idx < rem * rem; idx += blockDim.x) {
      const int i = j + 1 + idx / rem, k = j + 1 + idx % rem;
      if (k <= i) A[(long long)i * n + k] -= A[(long long)i * n + j] * A[(long long)k * n + j];
    }
    __syncthreads();
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    // L u = b
    for (int i = 0; i < n; ++i) {
      double s = Bm[(long long)i * D + d];
      for (int k = 0; k < i; ++k) s -= A[(long long)i * n + k] * Bm[(long long)k * D + d];
      Bm[(long long)i * D + d] = s / A[(long long)i * n + i];
    }
    // L^T w = u
    for (int i = n - 1; i >= 0; --i) {
      double s = Bm[(long long)i * D + d];
      for (int k = i + 1; k < n; ++k) s -= A[(long long)k * n + i] * Bm[(long long)k * D + d];
      Bm[(long long)i * D + d] = s / A[(long long)i * n + i];
    }
    double tot = 0.0;
    for (int i = 0; i < n; ++i) {
      const double v = Bm[(long long)i * D + d];
      tot += v;
      phi[(long long)i * D + d] = fabs(v) < 1e-10 ? 0.0 : v;
    }
    const double last = (fx[d] - fnull[d]) - tot;
    phi[(long long)n * D + d] = fabs(last) < 1e-10 ? 0.0 : last;
  }
  __syncthreads();
  if (threadIdx.x == 0 && status) *status = bad;
}

std::string launch_wls(const uint32_t* zbits, int zwords, const double* w, const float* y, long long K, int M,
                       int D, const double* fx, const double* fnull, double* phi, int32_t* status, double* work,
                       cudaStream_t s) {
  if (M < 2) return "wls: need at least 2 features";
  if (K <= 0 || D <= 0) return "wls: empty problem";
  const int n = M - 1;
  double* A = work;
  double* Bm = work + (long long)n * n;
  dim3 g1((n + 15) / 16, (n + 15) / 16);
  wls_gram_kernel<<<g1, 256, 0, s>>>(zbits, zwords, w, K, M, A);
  dim3 g2((D + 15) / 16, (n + 15) / 16);
  wls_rhs_kernel<<<g2, 256, 0, s>>>(zbits, zwords, w, y, K, M, D, fx, fnull, Bm);
  wls_solve_kernel<<<1, 1024, 0, s>>>(A, Bm, M, D, fx, fnull, phi, status);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

}  // namespace w2s
