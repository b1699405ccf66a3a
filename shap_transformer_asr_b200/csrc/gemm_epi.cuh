// Epilogue shared by every contraction kernel: CH consecutive columns of one output row.
#pragma once
#include "gemm.cuh"

namespace w2s {

// Partial LayerNorm statistics (sum, sum of squares) of 32 consecutive stored values.  The summation order is part of the
// contract -- two interleaved chains (even / odd columns, i.e. one packed FADD2 / FFMA2 chain), then their sum -- so that
// every contraction kernel produces the same bits for the same row.
__device__ __forceinline__ float2 ln_partial32(const float* v) {
  float2 s = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float2 t = make_float2(v[j], v[j + 1]);
    s = __fadd2_rn(s, t);
    q = __ffma2_rn(t, t, q);
  }
  return make_float2(s.x + s.y, q.x + q.y);
}

// ------------------------------------------------------------------------------------------------
// epilogue shared by both kernels: CH consecutive columns of one output row
// ------------------------------------------------------------------------------------------------
template <int CH>
__device__ __forceinline__ void epi_store(const EpiParams& e, int N, int g, int b, int m, int ncol0, float* v) {
  // N is a multiple of the column chunk in both kernels (N % 4 == 0; tcgen05 tiles divide N exactly)
  if (ncol0 >= N) return;
  if (e.ln_in) {
    // consumer form of a carried LayerNorm: rstd (acc - mean c1[n]) + c0[n] = fma(a, acc, fma(b, c1[n], c0[n])) with
    // a = rstd, b = -rstd mean (c0 arrives as the bias).  Every kernel uses exactly this operation sequence.
    const float2 st = __ldg(e.ln_in + m);
    const float a = st.y, bb = -st.x * st.y;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      v[j] = fmaf(a, v[j], fmaf(bb, __ldg(e.ln_c1 + ncol0 + j), __ldg(e.bias + ncol0 + j)));
  } else if (e.bias) {
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
        if (e.res_ln) {
          const float2 st = __ldg(e.res_ln + m);
          r = fmaf(fmaf(r, st.y, -st.x * st.y), __ldg(e.res_g + nout0 + j), __ldg(e.res_b + nout0 + j));
        }
        const float o = fmaf(v[j], e.alpha, r);
        v[j] = o;
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
        float ra = 1.f, rs = 0.f;   // residual LayerNorm on the fly: (r - mean) rstd = r ra + rs
        if (e.res_ln) {
          const float2 st = __ldg(e.res_ln + m);
          ra = st.y;
          rs = -st.x * st.y;
        }
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          if (j < nvals && nout0 + j < Nout) {
            uint2 r = *reinterpret_cast<const uint2*>(rp + j);
            float r4[4] = {bf16_lo(r.x), bf16_hi(r.x), bf16_lo(r.y), bf16_hi(r.y)};
            if (e.res_ln) {
              const float4 gg = __ldg(reinterpret_cast<const float4*>(e.res_g + nout0 + j));
              const float4 bb = __ldg(reinterpret_cast<const float4*>(e.res_b + nout0 + j));
              r4[0] = fmaf(fmaf(r4[0], ra, rs), gg.x, bb.x);
              r4[1] = fmaf(fmaf(r4[1], ra, rs), gg.y, bb.y);
              r4[2] = fmaf(fmaf(r4[2], ra, rs), gg.z, bb.z);
              r4[3] = fmaf(fmaf(r4[3], ra, rs), gg.w, bb.w);
            }
            v[j] = fmaf(v[j], e.alpha, r4[0]);
            v[j + 1] = fmaf(v[j + 1], e.alpha, r4[1]);
            v[j + 2] = fmaf(v[j + 2], e.alpha, r4[2]);
            v[j + 3] = fmaf(v[j + 3], e.alpha, r4[3]);
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
  if (e.stats_out) {
    // partial LayerNorm statistics of the stored row, one (sum, sum of squares) per 32-column block, summed in column
    // order so that every contraction kernel produces the same bits
    if constexpr (CH == 32) {
      e.stats_out[(long long)m * (N >> 5) + (ncol0 >> 5)] = ln_partial32(v);
    } else if constexpr (CH == 4) {
      // validation kernel: 8 neighbouring threads hold the 32 columns of a block (tree order: not bit-identical to the
      // sequential sum of the tensor-core kernels, same value to fp32 round-off)
      float s1 = (v[0] + v[1]) + (v[2] + v[3]);
      float s2 = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], v[3] * v[3])));
      const unsigned msk = __activemask();
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s1 += __shfl_xor_sync(msk, s1, o);
        s2 += __shfl_xor_sync(msk, s2, o);
      }
      if (((ncol0 >> 2) & 7) == 0) e.stats_out[(long long)m * (N >> 5) + (ncol0 >> 5)] = make_float2(s1, s2);
    }
  }
}


// Pair-kernel epilogue, first half: bias (from the warp's shared-memory strip) + activation + alpha / residual on
// 32 consecutive columns of one row; the caller stores the result (TMA staging).
// `sc1` / `sg` / `sb`: the warp's shared-memory strips of ln_c1 / res_g / res_b for these 32 columns; `st_in` / `st_res`:
// the row's (mean, rstd) for the consumer form / the residual LayerNorm; `rpre`: the 32 bf16 residual values of this
// sub-block already in registers (the caller issued the loads one sub-block ahead, so their latency is not exposed).
__device__ __forceinline__ void epi_math32(const EpiParams& e, const float* sbias, int N, int g, int b, int m,
                                           int ncol0, bool row_ok, float* v, bool skip_residual = false,
                                           const float* sc1 = nullptr, float2 st_in = make_float2(0.f, 1.f),
                                           const float* sg = nullptr, const float* sb = nullptr,
                                           float2 st_res = make_float2(0.f, 1.f), const uint4* rpre = nullptr) {
  if (sc1) {   // carried LayerNorm, consumer form (same operation sequence as epi_store), on the packed fp32 pipe
    const float2 a2 = make_float2(st_in.y, st_in.y);
    const float bb = -st_in.x * st_in.y;
    const float2 b2 = make_float2(bb, bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 c = *reinterpret_cast<const float4*>(sc1 + 4 * j);
      const float4 t = *reinterpret_cast<const float4*>(sbias + 4 * j);
      const float2 lo = __ffma2_rn(a2, make_float2(v[4 * j], v[4 * j + 1]), __ffma2_rn(b2, make_float2(c.x, c.y), make_float2(t.x, t.y)));
      const float2 hi = __ffma2_rn(a2, make_float2(v[4 * j + 2], v[4 * j + 3]), __ffma2_rn(b2, make_float2(c.z, c.w), make_float2(t.z, t.w)));
      v[4 * j] = lo.x; v[4 * j + 1] = lo.y; v[4 * j + 2] = hi.x; v[4 * j + 3] = hi.y;
    }
  } else if (sbias) {
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
        const float rsv = -st_res.x * st_res.y;
        const float2 ra2 = make_float2(st_res.y, st_res.y), rs2 = make_float2(rsv, rsv);
        const float2 al2 = make_float2(e.alpha, e.alpha);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 r = rpre ? rpre[j >> 3] : *reinterpret_cast<const uint4*>(rp + j);
          float2 r2[4] = {make_float2(bf16_lo(r.x), bf16_hi(r.x)), make_float2(bf16_lo(r.y), bf16_hi(r.y)),
                          make_float2(bf16_lo(r.z), bf16_hi(r.z)), make_float2(bf16_lo(r.w), bf16_hi(r.w))};
          if (sg) {   // residual LayerNorm on the fly: fma(fma(r, rstd, -mean rstd), gamma, beta), as in epi_store
            const float4 g0 = *reinterpret_cast<const float4*>(sg + j), g1 = *reinterpret_cast<const float4*>(sg + j + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(sb + j), b1 = *reinterpret_cast<const float4*>(sb + j + 4);
            r2[0] = __ffma2_rn(__ffma2_rn(r2[0], ra2, rs2), make_float2(g0.x, g0.y), make_float2(b0.x, b0.y));
            r2[1] = __ffma2_rn(__ffma2_rn(r2[1], ra2, rs2), make_float2(g0.z, g0.w), make_float2(b0.z, b0.w));
            r2[2] = __ffma2_rn(__ffma2_rn(r2[2], ra2, rs2), make_float2(g1.x, g1.y), make_float2(b1.x, b1.y));
            r2[3] = __ffma2_rn(__ffma2_rn(r2[3], ra2, rs2), make_float2(g1.z, g1.w), make_float2(b1.z, b1.w));
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 o = __ffma2_rn(make_float2(v[j + 2 * t], v[j + 2 * t + 1]), al2, r2[t]);
            v[j + 2 * t] = o.x;
            v[j + 2 * t + 1] = o.y;
          }
        }
      }
    }
  } else if (e.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
  }
}

}  // namespace w2s
