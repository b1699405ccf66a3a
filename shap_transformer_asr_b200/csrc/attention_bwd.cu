// Self-attention backward (no mask, head_dim 64) as ONE persistent, warp-specialised tcgen05 kernel -- the expected-gradients
// path's largest cost (shap_calculation.py:125-162 needs d out / d waveform).  The unfused form (grad_plan.cuh: five batched
// contractions + two row kernels per layer) writes and re-reads four [n, heads, T', T'] score-shaped tensors, ~4.8 GB per
// layer at 32 rows of an 11.5 s clip; here scores never leave the SM.
//
// Work item = (128-key block j, head, coalition).  K_j and V_j are loaded once; the item walks the query blocks i:
//   warp 13 (TMA)   Q_i and dO_i through a two-stage ring
//   warp 12 (MMA)   S = Q_i K_j^T and dP = dO_i V_j^T into TMEM; after the softmax warps have produced P and dS:
//                   dV_j += P^T dO_i,  dK_j += dS^T Q_i  (accumulating over i in TMEM; P^T / dS^T are the SAME shared-memory
//                   tiles read through an MN-major descriptor, Q_i / dO_i / K_j as MN-major B operands),  dQ_i = dS K_j
//   warps 0-7       P = exp2(S c - LSE_i),  dS = P (dP - D_i) scale  (a thread owns a query row and 64 of the 128 keys);
//                   LSE_i comes from the forward kernel (attention_fa.cu), D_i = dO_i . O_i from attn_delta_kernel
//   warps 8-11      drain dQ_i with vector reductions into the fp32 accumulator (a query block receives one partial per key
//                   block), and at the end of the item write dV_j / dK_j as bf16 rows of the d(q | k | v) buffer
// HF wav2vec2/modeling_wav2vec2.py:438-463 (softmax(Q K^T / sqrt d) V, eval mode), differentiated.
#include "kernels.cuh"
#include "gemm.cuh"

namespace w2s {

struct AttnBwdDev {
  const float* lse;     // [B, heads, T] log2-domain log-sum-exp of the scaled scores (forward kernel)
  const float* delta;   // [B, heads, T] D_i = dO_i . O_i
  float* dq;            // [B*T, H] fp32 accumulator (zeroed by the caller)
  __nv_bfloat16* dqkv;  // [B*T, ld]: dK and dV rows are written here
  int B, T, H, heads, ld, k_off, v_off, nblk, num_items;
  float scale, scale_log2e;
};
struct AttnBwdPlan {
  CUtensorMap mapQ, mapK, mapV, mapDO;
  AttnBwdDev dev;
  int grid;
};

constexpr int AB_SK = 0, AB_SV = 16384, AB_SQ = 2 * 16384, AB_SDO = AB_SQ + 2 * 16384, AB_SP = AB_SDO + 2 * 16384,
              AB_SDS = AB_SP + 32768, AB_BAR = AB_SDS + 32768;
constexpr size_t AB_SMEM = AB_BAR + 256 + 1024;
constexpr int AB_THREADS = 448;   // 8 softmax + 4 drain warps + MMA + TMA
// TMEM columns
constexpr uint32_t AB_TS = 0, AB_TDP = 128, AB_TDV = 256, AB_TDK = 320, AB_TDQ = 384;

// MN-major operand, 128-byte swizzle: rows are the K index (128 B each, 8-row groups 1024 B apart), 64-element chunks of the
// MN index are `lbo` bytes apart
__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                     const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapDO, const AttnBwdDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + AB_BAR;
  const uint32_t kv_full = bars, kv_empty = bars + 8, s_full = bars + 16, s_empty = bars + 24, p_full = bars + 32,
                 p_empty = bars + 40, dq_full = bars + 48, dq_empty = bars + 56, dkv_full = bars + 64, dkv_empty = bars + 72;
  auto qd_full = [&](int s) { return bars + 80 + 8u * s; };
  auto qd_empty = [&](int s) { return bars + 96 + 8u * s; };
  const uint32_t tmem_slot = bars + 112;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + AB_BAR + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NQ = p.nblk;   // query blocks per item (= key blocks per clip)
  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    tma_prefetch_desc(&mapDO);
  }
  if (warp == 12 && lane == 0) {
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_empty, 4);
    for (int s = 0; s < 2; ++s) {
      mbar_init(qd_full(s), 1);
      mbar_init(qd_empty(s), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 12) {
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 13) {
    if (elect_one()) {
      uint32_t qc = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int j = item % p.nblk, h = (item / p.nblk) % p.heads, b = item / (p.nblk * p.heads);
        mbar_wait(kv_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_expect_tx(kv_full, 32768);
        tma_load_4d(base + AB_SK, &mapK, kv_full, 0, j * 128, h, b);
        tma_load_4d(base + AB_SV, &mapV, kv_full, 0, j * 128, h, b);
#pragma unroll 1
        for (int i = 0; i < NQ; ++i, ++qc) {
          const int st = qc & 1;
          mbar_wait(qd_empty(st), ((qc >> 1) & 1u) ^ 1u);
          mbar_expect_tx(qd_full(st), 32768);
          tma_load_4d(base + AB_SQ + st * 16384, &mapQ, qd_full(st), 0, i * 128, h, b);
          tma_load_4d(base + AB_SDO + st * 16384, &mapDO, qd_full(st), 0, i * 128, h, b);
        }
      }
    }
  } else if (warp == 12) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);                                  // A, B K-major
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);         // A, B MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64) | (1u << 16);                      // A K-major, B MN-major
      uint32_t qc = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        mbar_wait(kv_full, (uint32_t)it & 1u);
#pragma unroll 1
        for (int i = 0; i < NQ; ++i, ++qc) {
          const int st = qc & 1;
          const uint32_t sq = base + AB_SQ + st * 16384, sdo = base + AB_SDO + st * 16384;
          mbar_wait(qd_full(st), (qc >> 1) & 1u);
          mbar_wait(s_empty, (qc & 1u) ^ 1u);
          tc_fence_after();
          {
            const uint64_t dq_ = umma_desc_sw128(sq), dk_ = umma_desc_sw128(base + AB_SK);
            const uint64_t ddo = umma_desc_sw128(sdo), dv_ = umma_desc_sw128(base + AB_SV);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + AB_TS, dq_ + 2u * k, dk_ + 2u * k, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + AB_TDP, ddo + 2u * k, dv_ + 2u * k, idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(s_full);
          mbar_wait(p_full, qc & 1u);
          if (i == 0) mbar_wait(dkv_empty, ((uint32_t)it & 1u) ^ 1u);
          mbar_wait(dq_empty, (qc & 1u) ^ 1u);
          tc_fence_after();
          // dV_j += P^T dO_i,  dK_j += dS^T Q_i: M = keys (two 64-key halves 16 KB apart), K = the 128 query rows
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t a = ab_desc_mn(base + AB_SP + k * 2048, 16384), bb = ab_desc_mn(sdo + k * 2048, 16384);
            umma_bf16(tmem + AB_TDV, a, bb, idesc_t, (i | k) != 0 ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t a = ab_desc_mn(base + AB_SDS + k * 2048, 16384), bb = ab_desc_mn(sq + k * 2048, 16384);
            umma_bf16(tmem + AB_TDK, a, bb, idesc_t, (i | k) != 0 ? 1u : 0u);
          }
          // dQ_i = dS K_j: M = queries, K = keys (two halves), B = K_j read MN-major (N = head dim, K = key rows)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t a = umma_desc_sw128(base + AB_SDS + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem + AB_TDQ, a + 2u * k, ab_desc_mn(base + AB_SK + kb * 8192 + k * 2048, 16384), idesc_q,
                        (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(p_empty);
          umma_commit(dq_full);
          umma_commit(qd_empty(st));
        }
        umma_commit(dkv_full);
        umma_commit(kv_empty);
      }
    }
  } else if (warp < 8) {
    // softmax / dS warps: row = 32 (warp & 3) + lane, keys 64 (warp >> 2) .. + 63 of the block
    const int qd = warp & 3, half = warp >> 2;
    const int row = qd * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(qd * 32) << 16);
    const uint32_t sw = (uint32_t)row & 7u;
    const uint32_t p_row = base + AB_SP + half * 16384 + row * 128, ds_row = base + AB_SDS + half * 16384 + row * 128;
    uint32_t qc = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int j = item % p.nblk, h = (item / p.nblk) % p.heads, b = item / (p.nblk * p.heads);
      const int nvalid = p.T - j * 128;   // keys of this block that exist
      const float* lse_bh = p.lse + ((long long)b * p.heads + h) * p.T;
      const float* del_bh = p.delta + ((long long)b * p.heads + h) * p.T;
#pragma unroll 1
      for (int i = 0; i < NQ; ++i, ++qc) {
        const int qi = i * 128 + row;
        const bool valid = qi < p.T;
        const float nlse = valid ? -lse_bh[qi] : -1.0e30f;   // rows past the clip: P = 0
        const float dlt = valid ? del_bh[qi] : 0.f;
        mbar_wait(s_full, qc & 1u);
        tc_fence_after();
        mbar_wait(p_empty, (qc & 1u) ^ 1u);   // the MMAs of the previous query block have read P / dS
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float s[32], dp[32];
          tmem_ld_32x32_issue(trow + AB_TS + half * 64 + c * 32, s);
          tmem_ld_32x32_issue(trow + AB_TDP + half * 64 + c * 32, dp);
          tmem_ld_wait();
          const int col0 = half * 64 + c * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t pk[4], dk[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int e = q * 8 + 2 * t;
              float p0 = ex2_approx(fmaf(s[e], p.scale_log2e, nlse)), p1 = ex2_approx(fmaf(s[e + 1], p.scale_log2e, nlse));
              if (col0 + e >= nvalid) p0 = 0.f;
              if (col0 + e + 1 >= nvalid) p1 = 0.f;
              pk[t] = pack_bf16x2(p0, p1);
              dk[t] = pack_bf16x2(p0 * (dp[e] - dlt) * p.scale, p1 * (dp[e + 1] - dlt) * p.scale);
            }
            const uint32_t off = ((((uint32_t)c * 4u + (uint32_t)q) ^ sw) << 4);
            sts128(p_row + off, pk[0], pk[1], pk[2], pk[3]);
            sts128(ds_row + off, dk[0], dk[1], dk[2], dk[3]);
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(s_empty);
          mbar_arrive(p_full);
        }
      }
    }
  } else if (warp < 12) {
    // drain warps: dQ_i partials -> fp32 accumulator (vector reductions), dV_j / dK_j -> bf16 rows at the end of the item
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(qd * 32) << 16);
    uint32_t qc = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const int j = item % p.nblk, h = (item / p.nblk) % p.heads, b = item / (p.nblk * p.heads);
#pragma unroll 1
      for (int i = 0; i < NQ; ++i, ++qc) {
        mbar_wait(dq_full, qc & 1u);
        tc_fence_after();
        float v0[32], v1[32];
        tmem_ld_32x32_issue(trow + AB_TDQ, v0);
        tmem_ld_32x32_issue(trow + AB_TDQ + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty);
        const int qi = i * 128 + row;
        if (qi < p.T) {
          float* dst = p.dq + ((long long)b * p.T + qi) * p.H + h * 64;
#pragma unroll
          for (int t = 0; t < 32; t += 4) red_add_v4(dst + t, v0[t], v0[t + 1], v0[t + 2], v0[t + 3]);
#pragma unroll
          for (int t = 0; t < 32; t += 4) red_add_v4(dst + 32 + t, v1[t], v1[t + 1], v1[t + 2], v1[t + 3]);
        }
      }
      mbar_wait(dkv_full, (uint32_t)it & 1u);
      tc_fence_after();
      const int kj = j * 128 + row;
      __nv_bfloat16* orow = p.dqkv + ((long long)b * p.T + kj) * p.ld + h * 64;
#pragma unroll
      for (int part = 0; part < 2; ++part) {   // 0: dV, 1: dK
        float v0[32], v1[32];
        tmem_ld_32x32_issue(trow + (part ? AB_TDK : AB_TDV), v0);
        tmem_ld_32x32_issue(trow + (part ? AB_TDK : AB_TDV) + 32, v1);
        tmem_ld_wait();
        if (part == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dkv_empty);
        }
        if (kj < p.T) {
          __nv_bfloat16* dst = orow + (part ? p.k_off : p.v_off);
#pragma unroll
          for (int t = 0; t < 32; t += 8) {
            uint4 u;
            u.x = pack_bf16x2(v0[t], v0[t + 1]); u.y = pack_bf16x2(v0[t + 2], v0[t + 3]);
            u.z = pack_bf16x2(v0[t + 4], v0[t + 5]); u.w = pack_bf16x2(v0[t + 6], v0[t + 7]);
            *reinterpret_cast<uint4*>(dst + t) = u;
            u.x = pack_bf16x2(v1[t], v1[t + 1]); u.y = pack_bf16x2(v1[t + 2], v1[t + 3]);
            u.z = pack_bf16x2(v1[t + 4], v1[t + 5]); u.w = pack_bf16x2(v1[t + 6], v1[t + 7]);
            *reinterpret_cast<uint4*>(dst + 32 + t) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// D[b, h, i] = sum_d dO[b, i, h, d] O[b, i, h, d]: one thread per (row, head)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                                          int B, int T, int H, int heads, float* __restrict__ delta) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * T * heads) return;
  const int h = (int)(idx % heads);
  const long long r = idx / heads;   // b * T + i
  const uint4* a = reinterpret_cast<const uint4*>(dO + r * H + h * 64);
  const uint4* o = reinterpret_cast<const uint4*>(O + r * H + h * 64);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint4 x = a[k], y = o[k];
    acc += bf16_lo(x.x) * bf16_lo(y.x) + bf16_hi(x.x) * bf16_hi(y.x) + bf16_lo(x.y) * bf16_lo(y.y) + bf16_hi(x.y) * bf16_hi(y.y) +
           bf16_lo(x.z) * bf16_lo(y.z) + bf16_hi(x.z) * bf16_hi(y.z) + bf16_lo(x.w) * bf16_lo(y.w) + bf16_hi(x.w) * bf16_hi(y.w);
  }
  const long long b = r / T, i = r - b * T;
  delta[(b * heads + h) * T + i] = acc;
}

// dqkv[r, q_off + c] = bf16(dq[r, c])
__global__ void __launch_bounds__(256) attn_dq_cast_kernel(const float* __restrict__ dq, long long rows, int H, int ld, int q_off,
                                                            __nv_bfloat16* __restrict__ dqkv) {
  const int H4 = H / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * H4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H4;
    const int c = (int)(i - r * H4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(dq + r * H + c);
    *reinterpret_cast<uint2*>(dqkv + r * ld + q_off + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

bool attention_bwd_supported(const AttnParams& p) { return p.hd == 64 && p.pos_proj == nullptr && (p.H % 8 == 0); }

std::string attention_bwd_init() {
  cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM);
  if (e != cudaSuccess) return std::string("cudaFuncSetAttribute(attention_bwd_kernel): ") + cudaGetErrorString(e);
  return "";
}

std::string attention_bwd_prepare(const AttnParams& p, const __nv_bfloat16* dctx, const float* lse, float* delta, float* dq,
                                  __nv_bfloat16* dqkv, int num_sms, AttnBwdPlan** out) {
  if (!attention_bwd_supported(p)) return "attention backward (tcgen05): unsupported shape";
  AttnBwdPlan* pl = new AttnBwdPlan();
  AttnBwdDev& d = pl->dev;
  d.lse = lse; d.delta = delta; d.dq = dq; d.dqkv = dqkv;
  d.B = p.B; d.T = p.T; d.H = p.H; d.heads = p.heads; d.ld = p.ld; d.k_off = p.k_off; d.v_off = p.v_off;
  d.nblk = (p.T + 127) / 128;
  d.num_items = d.nblk * p.heads * p.B;
  d.scale = p.scale;
  d.scale_log2e = p.scale * 1.4426950408889634f;
  pl->grid = d.num_items < num_sms ? d.num_items : num_sms;
  const uint64_t ld = (uint64_t)p.ld;
  uint64_t dims[4] = {64, (uint64_t)p.T, (uint64_t)p.heads, (uint64_t)p.B};
  uint64_t str[3] = {ld * 2, 128, (uint64_t)p.T * ld * 2};
  uint64_t strdo[3] = {(uint64_t)p.H * 2, 128, (uint64_t)p.T * p.H * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  std::string err = make_tensor_map_bf16(&pl->mapQ, p.qkv + p.q_off, 4, dims, str, box);
  if (err.empty()) err = make_tensor_map_bf16(&pl->mapK, p.qkv + p.k_off, 4, dims, str, box);
  if (err.empty()) err = make_tensor_map_bf16(&pl->mapV, p.qkv + p.v_off, 4, dims, str, box);
  if (err.empty()) err = make_tensor_map_bf16(&pl->mapDO, dctx, 4, dims, strdo, box);
  if (!err.empty()) {
    delete pl;
    return err;
  }
  *out = pl;
  return "";
}

// delta, zeroed dQ accumulator, the fused kernel, dQ -> bf16 columns of d(q | k | v)
std::string attention_bwd_launch(const AttnBwdPlan* pl, const __nv_bfloat16* dctx, const __nv_bfloat16* ctx, int q_off,
                                 cudaStream_t s) {
  const AttnBwdDev& d = pl->dev;
  const long long rows = (long long)d.B * d.T;
  if (rows == 0) return "";
  W2S_CUDA_OK(cudaMemsetAsync(d.dq, 0, sizeof(float) * rows * d.H, s));
  attn_delta_kernel<<<(unsigned)((rows * d.heads + 255) / 256), 256, 0, s>>>(dctx, ctx, d.B, d.T, d.H, d.heads,
                                                                             const_cast<float*>(d.delta));
  attention_bwd_kernel<<<pl->grid, AB_THREADS, AB_SMEM, s>>>(pl->mapQ, pl->mapK, pl->mapV, pl->mapDO, d);
  const long long n4 = rows * (d.H / 4);
  attn_dq_cast_kernel<<<(unsigned)((n4 + 255) / 256 > 148 * 16 ? 148 * 16 : (n4 + 255) / 256), 256, 0, s>>>(d.dq, rows, d.H, d.ld,
                                                                                                         q_off, d.dqkv);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

void attention_bwd_free(AttnBwdPlan* pl) { delete pl; }

}  // namespace w2s
