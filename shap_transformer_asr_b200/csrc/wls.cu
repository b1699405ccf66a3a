// K12: KernelSHAP constrained weighted least squares on the device, fp64.
// Restates shap KernelExplainer.solve with l1_reg=False: the last feature is
// eliminated with the efficiency constraint,
//   X[k, i] = z[k, i] - z[k, M-1]                         (entries in {-1, 0, 1})
//   r[k, d] = y[k, d] - fnull[d] - z[k, M-1] (fx[d] - fnull[d])
//   (X^T W X) w = X^T W r,   phi[:M-1] = w,   phi[M-1] = fx - fnull - sum(w),   |phi| < 1e-10 -> 0
// The normal matrix is symmetric positive SEMI-definite; it is definite when the rows span the M-1 differenced
// columns (always at the BASELINE configs: every single-feature coalition is enumerated) and is then factored by
// Cholesky (shap calls numpy.linalg.solve; same solution up to fp64 round-off).  When a pivot collapses -- fewer
// distinct coalitions than features, or a feature that never varies -- shap falls back to numpy.linalg.lstsq on the
// sqrt-weighted system, i.e. the minimum-norm least-squares solution.  The device path does the same without a host
// round trip: conjugate gradients on the (consistent) normal equations started from zero stay in range(X^T W X) and
// converge to exactly that minimum-norm solution; the kernel is always launched and returns at once when the
// factorisation succeeded.  status: 0 = Cholesky, 2 = singular design solved by CG, 1 = CG did not converge.
#include "kernels.cuh"

namespace w2s {

__device__ __forceinline__ int zbit(const uint32_t* z, int i) { return (z[i >> 5] >> (i & 31)) & 1; }
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

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
                                                          const double* __restrict__ A0, int32_t* __restrict__ status) {
  const int n = M - 1;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      const double djj = A[(long long)j * n + j];
      // a pivot that lost 11 digits against the original diagonal is a rank deficiency, not a number
      const bool ok = djj > 1e-11 * A0[(long long)j * n + j] && djj > 0.0;
      if (!ok) bad = 2;
      A[(long long)j * n + j] = sqrt(ok ? djj : 1.0);
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
  if (threadIdx.x == 0) *status = bad;
}

// Minimum-norm solution of A w = b for a singular PSD A (see the header): one CTA per right-hand side, vectors in
// shared memory, one warp per matrix row in the product.  Leaves at once unless the factorisation flagged the design.
__global__ void __launch_bounds__(256) wls_cg_kernel(const double* __restrict__ A, const double* __restrict__ Bm, int M,
                                                      int D, const double* __restrict__ fx,
                                                      const double* __restrict__ fnull, double* __restrict__ phi,
                                                      int32_t* __restrict__ status) {
  if (*status == 0) return;
  extern __shared__ double cg[];
  const int n = M - 1;
  double *x = cg, *r = cg + n, *pv = cg + 2 * n, *q = cg + 3 * n;
  __shared__ double s_red[8];
  __shared__ double s_val;
  const int d = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto block_sum = [&](double v) {
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_red[w];
      s_val = t;
    }
    __syncthreads();
    return s_val;
  };
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double b = Bm[(long long)i * D + d];
    x[i] = 0.0; r[i] = b; pv[i] = b;
    part += b * b;
  }
  const double bb = block_sum(part);
  double rr = bb;
  int it = 0;
  const int maxit = 4 * n + 50;
  const double tol = 1e-26 * bb;   // |r| <= 1e-13 |b|
  while (rr > tol && it < maxit) {
    for (int i = warp; i < n; i += 8) {
      const double* row = A + (long long)i * n;
      double acc = 0.0;
      for (int k = lane; k < n; k += 32) acc += row[k] * pv[k];
      acc = warp_sum_d(acc);
      if (lane == 0) q[i] = acc;
    }
    __syncthreads();
    part = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) part += pv[i] * q[i];
    const double pq = block_sum(part);
    if (!(pq > 0.0)) break;   // direction fell into the null space: converged as far as fp64 goes
    const double alpha = rr / pq;
    part = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      x[i] += alpha * pv[i];
      const double ri = r[i] - alpha * q[i];
      r[i] = ri;
      part += ri * ri;
    }
    const double rr_new = block_sum(part);
    const double beta = rr_new / rr;
    for (int i = threadIdx.x; i < n; i += blockDim.x) pv[i] = r[i] + beta * pv[i];
    __syncthreads();
    rr = rr_new;
    ++it;
  }
  part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i];
    part += v;
    phi[(long long)i * D + d] = fabs(v) < 1e-10 ? 0.0 : v;
  }
  const double tot = block_sum(part);
  if (threadIdx.x == 0) {
    const double last = (fx[d] - fnull[d]) - tot;
    phi[(long long)n * D + d] = fabs(last) < 1e-10 ? 0.0 : last;
    if (rr > 1e-16 * bb && bb > 0.0) atomicExch(status, 1);   // |r| > 1e-8 |b|: did not converge
  }
}

std::string launch_wls(const uint32_t* zbits, int zwords, const double* w, const float* y, long long K, int M,
                       int D, const double* fx, const double* fnull, double* phi, int32_t* status, double* work,
                       cudaStream_t s) {
  if (M < 2) return "wls: need at least 2 features";
  if (K <= 0 || D <= 0) return "wls: empty problem";
  if (!status) return "wls: status pointer is required";
  const int n = M - 1;
  const long long nn = (long long)n * n, nd = (long long)n * D;
  double *A = work, *Bm = work + nn, *A0 = work + nn + nd, *B0 = work + 2 * nn + nd;
  dim3 g1((n + 15) / 16, (n + 15) / 16);
  wls_gram_kernel<<<g1, 256, 0, s>>>(zbits, zwords, w, K, M, A0);
  dim3 g2((D + 15) / 16, (n + 15) / 16);
  wls_rhs_kernel<<<g2, 256, 0, s>>>(zbits, zwords, w, y, K, M, D, fx, fnull, B0);
  // the factorisation and the substitutions run in place on copies; the originals feed the CG fallback
  W2S_CUDA_OK(cudaMemcpyAsync(A, A0, sizeof(double) * nn, cudaMemcpyDeviceToDevice, s));
  W2S_CUDA_OK(cudaMemcpyAsync(Bm, B0, sizeof(double) * nd, cudaMemcpyDeviceToDevice, s));
  wls_solve_kernel<<<1, 1024, 0, s>>>(A, Bm, M, D, fx, fnull, phi, A0, status);
  const size_t cg_smem = sizeof(double) * 4 * (size_t)n;
  static bool attr_set = false;
  if (!attr_set) {
    W2S_CUDA_OK(cudaFuncSetAttribute(wls_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4 * 2048));
    attr_set = true;
  }
  wls_cg_kernel<<<D, 256, cg_smem, s>>>(A0, B0, M, D, fx, fnull, phi, status);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

}  // namespace w2s
