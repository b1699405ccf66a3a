// Epilogue shared by every contraction kernel: CH consecutive columns of one output row.
#pragma once
#include "gemm.cuh"

namespace w2s {

// ------------------------------------------------------------------------------------------------
// epilogue shared by both kernels: CH consecutive columns of one output row
// ------------------------------------------------------------------------------------------------
template <int CH>
__device__ __forceinline__ void epi_store(const EpiParams& e, int N, int g, int b, int m, int ncol0, float* v) {
  // N is a multiple of the column chunk in both kernels (N % 4 == 0; tcgen05 tiles divide N exactly)
  if (ncol0 >= N) return;
  if (e.bias) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + (long long)g * N + ncol0);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) {
      const float4 t = __ldg(bp + j);
      v[4 * j] += t.x;
      v[4 * j + 1] += t.y;
      v[4 * j + 2] += t.z;
      v[4 * j + 3] += t.w;
    }
  }
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < CH; j += 2) {
      const float2 r = gelu_erf2(make_float2(v[j], v[j + 1]));
      v[j] = r.x;
      v[j + 1] = r.y;
    }
  } else if (e.act == ACT_SWISH) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = swish(v[j]);
  }
  int nout0 = ncol0, nvals = CH, Nout = N;
  if (e.glu) {
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) v[j] = v[2 * j] * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v[2 * j + 1]));
    nout0 = ncol0 >> 1;
    nvals = CH / 2;
    Nout = N >> 1;
  }
  const long long off = (long long)g * e.ldg + (long long)b * e.ldb + (long long)m * e.ldm + nout0;
  if constexpr (CH < 8) {
    // scalar path (validation kernel)
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (j < nvals && nout0 + j < Nout) {
        float r = 0.f;
        if (e.residual)
          r = e.res_fp32 ? reinterpret_cast<const float*>(e.residual)[off + j]
                         : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.residual)[off + j]);
        const float o = fmaf(v[j], e.alpha, r);
        if (e.out_fp32) reinterpret_cast<float*>(e.out)[off + j] = o;
        else reinterpret_cast<__nv_bfloat16*>(e.out)[off + j] = __float2bfloat16_rn(o);
      }
    }
  } else {
    if (e.residual) {
      if (e.res_fp32) {
        const float* rp = reinterpret_cast<const float*>(e.residual) + off;
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          if (j < nvals && nout0 + j < Nout) {
            float4 r = *reinterpret_cast<const float4*>(rp + j);
            v[j] = fmaf(v[j], e.alpha, r.x);
            v[j + 1] = fmaf(v[j + 1], e.alpha, r.y);
            v[j + 2] = fmaf(v[j + 2], e.alpha, r.z);
            v[j + 3] = fmaf(v[j + 3], e.alpha, r.w);
          }
        }
      } else {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(e.residual) + off;
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          if (j < nvals && nout0 + j < Nout) {
            uint2 r = *reinterpret_cast<const uint2*>(rp + j);
            v[j] = fmaf(v[j], e.alpha, bf16_lo(r.x));
            v[j + 1] = fmaf(v[j + 1], e.alpha, bf16_hi(r.x));
            v[j + 2] = fmaf(v[j + 2], e.alpha, bf16_lo(r.y));
            v[j + 3] = fmaf(v[j + 3], e.alpha, bf16_hi(r.y));
          }
        }
      }
    } else if (e.alpha != 1.0f) {
#pragma unroll
      for (int j = 0; j < CH; ++j) v[j] *= e.alpha;
    }
    if (e.out_fp32) {
      float* op = reinterpret_cast<float*>(e.out) + off;
#pragma unroll
      for (int j = 0; j < CH; j += 4)
        if (j < nvals && nout0 + j < Nout)
          *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out) + off;
#pragma unroll
      for (int j = 0; j < CH; j += 8) {
        if (j < nvals && nout0 + j < Nout) {
          uint4 u;
          u.x = pack_bf16x2(v[j], v[j + 1]);
          u.y = pack_bf16x2(v[j + 2], v[j + 3]);
          u.z = pack_bf16x2(v[j + 4], v[j + 5]);
          u.w = pack_bf16x2(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(op + j) = u;
        }
      }
    }
  }
}


// Pair-kernel epilogue, first half: bias (from the warp's shared-memory strip) + activation + alpha / residual on
// 32 consecutive columns of one row; the caller stores the result (TMA staging).
__device__ __forceinline__ void epi_math32(const EpiParams& e, const float* sbias, int N, int g, int b, int m,
                                           int ncol0, bool row_ok, float* v, bool skip_residual = false) {
  if (sbias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = *reinterpret_cast<const float4*>(sbias + 4 * j);
      v[4 * j] += t.x;
      v[4 * j + 1] += t.y;
      v[4 * j + 2] += t.z;
      v[4 * j + 3] += t.w;
    }
  }
  if (e.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 r = gelu_erf2(make_float2(v[j], v[j + 1]));
      v[j] = r.x;
      v[j + 1] = r.y;
    }
  } else if (e.act == ACT_SWISH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = swish(v[j]);
  }
  if (e.residual && !skip_residual) {
    if (row_ok) {
      const long long off = (long long)g * e.ldg + (long long)b * e.ldb + (long long)m * e.ldm + ncol0;
      if (e.res_fp32) {
        const float* rp = reinterpret_cast<const float*>(e.residual) + off;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 r = *reinterpret_cast<const float4*>(rp + j);
          v[j] = fmaf(v[j], e.alpha, r.x);
          v[j + 1] = fmaf(v[j + 1], e.alpha, r.y);
          v[j + 2] = fmaf(v[j + 2], e.alpha, r.z);
          v[j + 3] = fmaf(v[j + 3], e.alpha, r.w);
        }
      } else {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(e.residual) + off;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 r = *reinterpret_cast<const uint4*>(rp + j);
          v[j] = fmaf(v[j], e.alpha, bf16_lo(r.x));
          v[j + 1] = fmaf(v[j + 1], e.alpha, bf16_hi(r.x));
          v[j + 2] = fmaf(v[j + 2], e.alpha, bf16_lo(r.y));
          v[j + 3] = fmaf(v[j + 3], e.alpha, bf16_hi(r.y));
          v[j + 4] = fmaf(v[j + 4], e.alpha, bf16_lo(r.z));
          v[j + 5] = fmaf(v[j + 5], e.alpha, bf16_hi(r.z));
          v[j + 6] = fmaf(v[j + 6], e.alpha, bf16_lo(r.w));
          v[j + 7] = fmaf(v[j + 7], e.alpha, bf16_hi(r.w));
        }
      }
    }
  } else if (e.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
  }
}

}  // namespace w2s
