// Self-attention of the encoder (HF wav2vec2/modeling_wav2vec2.py:438-463, :500-549; conformer relative
// positions wav2vec2_conformer/modeling_wav2vec2_conformer.py:509-565).
//
//  * attention_tc_kernel  : tcgen05 path.  One CTA per (128-query tile, head, coalition).  Q, K and V^T tiles
//                           arrive by TMA (128B swizzle); S = Q K^T accumulates in TMEM (up to 512 fp32
//                           columns = the whole key range, T' <= 512); the four warps run the softmax out of
//                           TMEM (thread = query row), write P as bf16 into swizzled shared memory and a
//                           second tcgen05.mma chain forms O = P V in TMEM.
//  * attention_simt_kernel: CUDA-core validation kernel (one warp per query row), also carries the
//                           conformer relative-position term.
#include "kernels.cuh"
#include "gemm.cuh"

namespace w2s {

// =================================================================================================
// validation kernel
// =================================================================================================
constexpr int SIMT_MAXC = 16;  // key chunks of 32 -> T <= 512

__global__ void __launch_bounds__(128) attention_simt_kernel(const AttnParams p) {
  __shared__ float qs[4][2][128];  // per warp: q + u, q + v  (hd <= 128)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 4 + warp;
  const long long total = (long long)p.B * p.heads * p.T;
  if (gw >= total) return;
  const int i = (int)(gw % p.T);
  const int h = (int)((gw / p.T) % p.heads);
  const int b = (int)(gw / ((long long)p.T * p.heads));
  const int H3 = 3 * p.H;
  const __nv_bfloat16* qrow = p.qkv + ((long long)b * p.T + i) * H3 + h * p.hd;
  for (int d = lane; d < p.hd; d += 32) {
    const float q = __bfloat162float(qrow[d]);
    qs[warp][0][d] = q + (p.bias_u ? p.bias_u[h * p.hd + d] : 0.f);
    qs[warp][1][d] = q + (p.bias_v ? p.bias_v[h * p.hd + d] : 0.f);
  }
  __syncwarp();
  float s[SIMT_MAXC];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < SIMT_MAXC; ++c) {
    const int j = c * 32 + lane;
    float acc = -INFINITY;
    if (j < p.T) {
      const __nv_bfloat16* krow = p.qkv + ((long long)b * p.T + j) * H3 + p.H + h * p.hd;
      acc = 0.f;
      for (int d = 0; d < p.hd; d += 2) {
        const uint32_t kk = *reinterpret_cast<const uint32_t*>(krow + d);
        acc = fmaf(qs[warp][0][d], bf16_lo(kk), acc);
        acc = fmaf(qs[warp][0][d + 1], bf16_hi(kk), acc);
      }
      if (p.pos_proj) {
        // shift trick as index arithmetic: bd[i, j] = (q_i + v) . pos_proj[T - 1 - i + j]
        const __nv_bfloat16* prow = p.pos_proj + (long long)(p.T - 1 - i + j) * p.H + h * p.hd;
        for (int d = 0; d < p.hd; d += 2) {
          const uint32_t pp = *reinterpret_cast<const uint32_t*>(prow + d);
          acc = fmaf(qs[warp][1][d], bf16_lo(pp), acc);
          acc = fmaf(qs[warp][1][d + 1], bf16_hi(pp), acc);
        }
      }
      acc *= p.scale;
    }
    s[c] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < SIMT_MAXC; ++c) {
    const float e = (c * 32 + lane < p.T) ? __expf(s[c] - mx) : 0.f;
    s[c] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  float o[4] = {0.f, 0.f, 0.f, 0.f};  // hd <= 128
#pragma unroll
  for (int c = 0; c < SIMT_MAXC; ++c) {
    if (c * 32 >= p.T) break;
    for (int l = 0; l < 32; ++l) {
      const int j = c * 32 + l;
      const float pj = __shfl_sync(0xffffffffu, s[c], l);
      if (j < p.T) {
        const __nv_bfloat16* vrow = p.qkv + ((long long)b * p.T + j) * H3 + 2 * p.H + h * p.hd;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int d = lane + 32 * r;
          if (d < p.hd) o[r] = fmaf(pj, __bfloat162float(vrow[d]), o[r]);
        }
      }
    }
  }
  __nv_bfloat16* orow = p.ctx + ((long long)b * p.T + i) * p.H + h * p.hd;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = lane + 32 * r;
    if (d < p.hd) orow[d] = __float2bfloat16_rn(o[r] * inv);
  }
}

std::string launch_attention_simt(const AttnParams& p, cudaStream_t s) {
  if (p.T > 32 * SIMT_MAXC) return "attention (validation kernel): T' > 512 not supported";
  if (p.hd > 128 || (p.hd & 1)) return "attention (validation kernel): head_dim must be even and <= 128";
  const long long total = (long long)p.B * p.heads * p.T;
  if (total == 0) return "";
  attention_simt_kernel<<<(unsigned)((total + 3) / 4), 128, 0, s>>>(p);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

// =================================================================================================
// tcgen05 kernel
// =================================================================================================
struct AttnTcDev {
  __nv_bfloat16* ctx;
  int B, T, Tp, H, heads;
  int nblk;       // 64-key blocks = Tp / 64
  int n0, n1;     // keys in the first / second S half (multiples of 64, <= 256 each)
  int sv_off;     // smem offset of the V tiles
  float scale_log2e;
};
struct AttnTcPlan {
  CUtensorMap mapQ, mapK, mapV;
  AttnTcDev dev;
  dim3 grid;
  size_t smem;
  int tmem_cols;
};

// shared memory map (bytes from the 1024-aligned base):
//   [0, 16K)          Q   [128 x 64]            \  dead once S = Q K^T has completed;
//   [16K, 16K + Ksz)  K   1 or 2 x [256 x 64]   /  P (4 x [128 x 64 keys] = 64 KB) is written over them
//   [sv_off, ...)     V   nblk x [64 keys x 64] straight from the qkv buffer (MN-major B operand)
constexpr int ATT_SQ = 0;
constexpr int ATT_SK = 16384;
constexpr int ATT_SP = 0;

__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  // MN-major operand (rows = K index, 128 contiguous bytes = 64 MN elements, 128B swizzle):
  // 8-row (K) groups are 1024 B apart (SBO); LBO (next 64-wide MN block) unused for N = 64.
  return umma_desc_sw128(smem_addr);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int M, int N) {
  return umma_idesc_bf16(M, N) | (1u << 16);  // b_major = MN
}

template <int TMEM_COLS>
__global__ void __launch_bounds__(256, (TMEM_COLS <= 256) ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapV, const AttnTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_red[2][128];
  __shared__ uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_load = smem_u32(&s_bar[0]), bar_mma = smem_u32(&s_bar[1]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qd = warp & 3, hf = warp >> 2;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row = qd * 32 + lane;  // query row inside the tile == TMEM lane

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) {
    tmem_alloc<TMEM_COLS>(smem_u32(&s_tmem));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&s_tmem);
  const uint32_t sV = base + p.sv_off;

  // ---- loads ---------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    const uint32_t bytes = 16384u + (p.n1 > 0 ? 65536u : 32768u) + 8192u * p.nblk;
    mbar_expect_tx(bar_load, bytes);
    tma_load_4d(base + ATT_SQ, &mapQ, bar_load, 0, qt * 128, h, b);
    tma_load_4d(base + ATT_SK, &mapK, bar_load, 0, 0, h, b);
    if (p.n1 > 0) tma_load_4d(base + ATT_SK + 32768, &mapK, bar_load, 0, 256, h, b);
    for (int kb = 0; kb < p.nblk; ++kb) tma_load_4d(sV + kb * 8192, &mapV, bar_load, 0, kb * 64, h, b);
    mbar_wait(bar_load, 0);
    // ---- S = Q K^T ---------------------------------------------------------------------------------
    tc_fence_after();
    const uint64_t dq = umma_desc_sw128(base + ATT_SQ);
    {
      const uint64_t dk = umma_desc_sw128(base + ATT_SK);
      const uint32_t idesc = umma_idesc_bf16(128, p.n0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, dq + 2u * k, dk + 2u * k, idesc, k != 0);
    }
    if (p.n1 > 0) {
      const uint64_t dk = umma_desc_sw128(base + ATT_SK + 32768);
      const uint32_t idesc = umma_idesc_bf16(128, p.n1);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + 256, dq + 2u * k, dk + 2u * k, idesc, k != 0);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  const uint32_t trow = tmem + (static_cast<uint32_t>(qd * 32) << 16);

  // ---- pass 1: row maximum over the valid keys (each half-warp-group scans its key blocks) -----------------
  float mx = -INFINITY;
  for (int r0 = 0; r0 < p.nblk; r0 += 4) {
    const int cnt = min(4, p.nblk - r0);
    const int per = (cnt + 1) >> 1;
    const int b0 = r0 + hf * per, b1 = min(r0 + cnt, b0 + per);
    for (int c = b0 * 64; c < b1 * 64; c += 32) {
      float v[32];
      tmem_ld_32x32(trow + c, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c + j < p.T) mx = fmaxf(mx, v[j]);
    }
  }
  s_red[hf][row] = mx;
  __syncthreads();
  mx = fmaxf(s_red[0][row], s_red[1][row]);
  const float mscaled = mx * p.scale_log2e;

  // ---- pass 2: P = exp2(S*scale - max) as bf16 into swizzled smem (over Q/K), then O (+)= P V -----------------
  float sum = 0.f;
  uint32_t mma_parity = 1;
  for (int r0 = 0; r0 < p.nblk; r0 += 4) {
    const int cnt = min(4, p.nblk - r0);
    const int per = (cnt + 1) >> 1;
    const int b0 = r0 + hf * per, b1 = min(r0 + cnt, b0 + per);
    if (r0 > 0) {
      mbar_wait(bar_mma, mma_parity);  // previous P V chain has finished reading sP
      mma_parity ^= 1u;
      tc_fence_after();
    }
    for (int kb = b0; kb < b1; ++kb) {
      const uint32_t sp_row = base + ATT_SP + (kb - r0) * 16384 + row * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        const int c = kb * 64 + half * 32;
        tmem_ld_32x32(trow + c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = (c + j < p.T) ? ex2_approx(fmaf(v[j], p.scale_log2e, -mscaled)) : 0.f;
          v[j] = e;
          sum += e;
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = half * 4 + q4;  // 16-byte chunk index inside the 128-byte row
          const uint32_t addr = sp_row + (((uint32_t)chunk ^ ((uint32_t)row & 7u)) << 4);
          const uint32_t u0 = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]);
          const uint32_t u1 = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
          const uint32_t u2 = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]);
          const uint32_t u3 = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u0), "r"(u1), "r"(u2), "r"(u3)
                       : "memory");
        }
      }
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_bf16_bmn(128, 64);
      for (int kb = r0; kb < r0 + cnt; ++kb) {
        const uint64_t dp = umma_desc_sw128(base + ATT_SP + (kb - r0) * 16384);
        const uint64_t dv = umma_desc_sw128_mn(sV + kb * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 keys per MMA: +32 B along P's rows, +16 rows (2048 B) down V
          umma_bf16(tmem, dp + 2u * k, dv + 128u * k, idesc, (kb | k) != 0);
      }
      umma_commit(bar_mma);
    }
  }
  s_red[hf][row] = sum;   // safe: every thread passed the barrier above after reading the maxima
  mbar_wait(bar_mma, mma_parity);
  tc_fence_after();
  __syncthreads();
  sum = s_red[0][row] + s_red[1][row];

  // ---- epilogue: O / rowsum -> ctx (each half takes 32 of the 64 output columns) ----------------------------------
  const int i = qt * 128 + row;
  const float inv = 1.0f / sum;
  __nv_bfloat16* orow = p.ctx + ((long long)b * p.T + i) * p.H + h * 64 + hf * 32;
  {
    float v[32];
    tmem_ld_32x32(trow + hf * 32, v);
    if (i < p.T) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_bf16x2(v[j] * inv, v[j + 1] * inv);
        u.y = pack_bf16x2(v[j + 2] * inv, v[j + 3] * inv);
        u.z = pack_bf16x2(v[j + 4] * inv, v[j + 5] * inv);
        u.w = pack_bf16x2(v[j + 6] * inv, v[j + 7] * inv);
        *reinterpret_cast<uint4*>(orow + j) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

bool attention_tc_supported(const AttnParams& p) {
  return p.hd == 64 && p.T <= 512 && p.pos_proj == nullptr && (p.H % 8 == 0);
}

static size_t att_smem(int Tp) { return (Tp <= 256 ? 65536 : 81920) + (size_t)(Tp / 64) * 8192 + 1024; }

std::string attention_tc_init() {
  const int mx = (int)att_smem(512);
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e != cudaSuccess) return std::string("cudaFuncSetAttribute(attention_tc_kernel): ") + cudaGetErrorString(e);
  return "";
}

std::string attention_tc_prepare(const AttnParams& p, AttnTcPlan** out) {
  if (!attention_tc_supported(p)) return "attention (tcgen05): unsupported shape";
  AttnTcPlan* pl = new AttnTcPlan();
  const int Tp = p.Tp;
  if (Tp % 64 || Tp < p.T || Tp > 512) {
    delete pl;
    return "attention (tcgen05): Tp must be a multiple of 64 covering T";
  }
  pl->dev.ctx = p.ctx;
  pl->dev.B = p.B; pl->dev.T = p.T; pl->dev.Tp = Tp; pl->dev.H = p.H; pl->dev.heads = p.heads;
  pl->dev.nblk = Tp / 64;
  pl->dev.n0 = Tp < 256 ? Tp : 256;
  pl->dev.n1 = Tp > 256 ? Tp - 256 : 0;
  pl->dev.sv_off = Tp <= 256 ? 65536 : 81920;
  pl->dev.scale_log2e = p.scale * 1.4426950408889634f;
  pl->grid = dim3((p.T + 127) / 128, p.heads, p.B);
  pl->smem = att_smem(Tp);
  pl->tmem_cols = Tp <= 64 ? 64 : (Tp <= 128 ? 128 : (Tp <= 256 ? 256 : 512));
  const uint64_t H3 = 3ull * p.H;
  std::string err;
  {
    uint64_t dims[4] = {64, (uint64_t)p.T, (uint64_t)p.heads, (uint64_t)p.B};
    uint64_t str[3] = {H3 * 2, 128, (uint64_t)p.T * H3 * 2};
    uint32_t boxq[4] = {64, 128, 1, 1};
    uint32_t boxk[4] = {64, 256, 1, 1};
    uint32_t boxv[4] = {64, 64, 1, 1};
    err = make_tensor_map_bf16(&pl->mapQ, p.qkv, 4, dims, str, boxq);
    if (err.empty()) err = make_tensor_map_bf16(&pl->mapK, p.qkv + p.H, 4, dims, str, boxk);
    if (err.empty()) err = make_tensor_map_bf16(&pl->mapV, p.qkv + 2 * p.H, 4, dims, str, boxv);
  }
  if (!err.empty()) {
    delete pl;
    return err;
  }
  *out = pl;
  return "";
}

std::string attention_tc_launch(const AttnTcPlan* pl, cudaStream_t s) {
  switch (pl->tmem_cols) {
    case 64: attention_tc_kernel<64><<<pl->grid, 256, pl->smem, s>>>(pl->mapQ, pl->mapK, pl->mapV, pl->dev); break;
    case 128: attention_tc_kernel<128><<<pl->grid, 256, pl->smem, s>>>(pl->mapQ, pl->mapK, pl->mapV, pl->dev); break;
    case 256: attention_tc_kernel<256><<<pl->grid, 256, pl->smem, s>>>(pl->mapQ, pl->mapK, pl->mapV, pl->dev); break;
    default: attention_tc_kernel<512><<<pl->grid, 256, pl->smem, s>>>(pl->mapQ, pl->mapK, pl->mapV, pl->dev); break;
  }
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

void attention_tc_free(AttnTcPlan* plan) { delete plan; }

}  // namespace w2s
