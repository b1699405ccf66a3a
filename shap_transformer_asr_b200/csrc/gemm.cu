// tcgen05 / TMEM / TMA contraction kernel (sm_100a) and its CUDA-core validation twin.
//
// Roles inside one 256-thread CTA (persistent over output tiles, static round-robin schedule):
// (the warp scheduler favours high warp ids, so the two latency-critical single-thread roles sit above the
//  ALU-heavy epilogue warps)
//   warp 8 lane 0 : TMA producer   -- fills a STAGES-deep ring of {A 128x64, W BNx64} bf16 tiles (128B swizzle)
//   warp 9 lane 0 : MMA issuer     -- tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16 per instruction,
//                                     accumulating in one of two TMEM accumulator slots
//   warp 10       : TMEM allocator
//   warps 0..7    : epilogue       -- tcgen05.ld the finished accumulator (thread = row), bias / activation /
//                                     GLU / residual, vectorised global stores, while the MMA warp already
//                                     works on the next tile in the other TMEM slot
// Pipelines: full/empty mbarriers per smem stage (TMA <-> MMA), tmem_full/tmem_empty per accumulator slot
// (MMA <-> epilogue).
#include "gemm.cuh"
#include "gemm_epi.cuh"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>

namespace w2s {

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
template <int BN>
struct TcCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int CH = (BN % 32 == 0) ? 32 : 16;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};

struct TileCoord {
  int g, b, m0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmDev& p, int unit, int BN) {
  TileCoord c;
  int nt = unit % p.tiles_n;
  int r = unit / p.tiles_n;
  int mu = r % p.units_m;
  r /= p.units_m;
  c.b = r % p.Bz;
  c.g = r / p.Bz;
  c.m0 = mu * 128;
  c.n0 = nt * BN;
  return c;
}

template <int BN>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
               const GemmDev p) {
  using C = TcCfg<BN>;
  const int unit0 = (int)blockIdx.x;
  const int unit_step = (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  uint8_t* tiles_ptr = smem_raw + (tiles - raw);
  const uint32_t bars = tiles + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(tiles_ptr + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 10) {
    tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 8) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < p.num_units; unit += unit_step) {
        const TileCoord tc = decode_tile(p, unit, BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          const int krow = kb / p.a_kb_per_row;
          const int kcol = kb - krow * p.a_kb_per_row;
          const uint32_t sa = tiles + stage * C::STAGE_BYTES;
          tma_load_3d(sa, &mapA, full_bar(stage), tc.g * p.a_g_col + kcol * 64, tc.g * p.a_g_row + tc.m0 + krow, tc.b);
          tma_load_4d(sa + C::A_BYTES, &mapW, full_bar(stage), kb * 64, tc.n0, tc.g, p.w_batched ? tc.b : 0);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = unit0; unit < p.num_units; unit += unit_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = tiles + stage * C::STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 bytes per K=16 step inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit0; unit < p.num_units; unit += unit_step) {
      const TileCoord tc = decode_tile(p, unit, BN);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int m = tc.m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::ACC_COLS;
      // warps 0..3 take the even column chunks of their lane quadrant, warps 4..7 the odd ones
#pragma unroll 1
      for (int c = (warp >> 2) * C::CH; c < BN; c += 2 * C::CH) {
        float v[C::CH];
        if constexpr (C::CH == 32) tmem_ld_32x32(t0 + c, v);
        else tmem_ld_32x16(t0 + c, v);
        if (row_ok) epi_store<C::CH>(p.epi, p.N, tc.g, tc.b, m, tc.n0 + c, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// CUDA-core validation kernel: same operands, same addressing rules (incl. zero fill outside the view),
// fp32 accumulation.  64x64 output tile, 16x16 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDev p) {
  __shared__ float As[16][64 + 1];
  __shared__ float Ws[16][64 + 1];
  int tile = blockIdx.x;
  const int tiles_n = (p.N + 63) / 64, tiles_m = (p.M + 63) / 64;
  const int nt = tile % tiles_n;
  tile /= tiles_n;
  const int mt = tile % tiles_m;
  tile /= tiles_m;
  const int b = tile % p.Bz, g = tile / p.Bz;
  const int m0 = mt * 64, n0 = nt * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const int rowlen = p.a_kb_per_row * 64;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      const int k = k0 + kk;
      float av = 0.f, wv = 0.f;
      if (k < p.K) {
        const long long row = (long long)g * p.a_g_row + m0 + r + k / rowlen;
        const long long col = (long long)g * p.a_g_col + k % rowlen;
        if (row < p.a_rows && col < p.a_cols)
          av = __bfloat162float(p.a[(long long)b * p.a_batch_stride + row * p.a_row_stride + col]);
        if (n0 + r < p.N) wv = __bfloat162float(p.w[((long long)g * p.N + n0 + r) * p.K + k]);
      }
      As[kk][r] = av;
      Ws[kk][r] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[kk][ty * 4 + i];
        w[i] = Ws[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m < p.M) epi_store<4>(p.epi, p.N, g, b, m, n0 + tx * 4, acc[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_init_once;
static std::string g_init_err;

template <int BN>
static cudaError_t set_attr() {
  return cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<BN>::SMEM);
}

std::string gemm_init() {
  std::call_once(g_init_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      g_init_err = "cuTensorMapEncodeTiled entry point not available (driver too old?)";
      return;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    cudaError_t a = set_attr<256>();
    if (a == cudaSuccess) a = set_attr<128>();
    if (a == cudaSuccess) a = set_attr<64>();
    if (a == cudaSuccess) a = set_attr<48>();
    if (a == cudaSuccess) a = set_attr<32>();
    if (a != cudaSuccess) g_init_err = std::string("cudaFuncSetAttribute(gemm_tc_kernel): ") + cudaGetErrorString(a);
    if (g_init_err.empty()) g_init_err = gemm2_init();
  });
  return g_init_err;
}

std::string make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tensor_map(map, base, 0, rank, dims, strides_bytes, box);
}

std::string make_tensor_map(CUtensorMap* map, const void* base, int fp32, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box) {
  if (!g_encode) return "tensor map encoder not initialised";
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string s = "cuTensorMapEncodeTiled failed (CUresult " + std::to_string((int)r) + ") rank=" +
                    std::to_string(rank) + " dims=";
    for (int i = 0; i < rank; ++i) s += std::to_string(dims[i]) + ",";
    s += " strides=";
    for (int i = 0; i + 1 < rank; ++i) s += std::to_string(strides_bytes[i]) + ",";
    return s;
  }
  return "";
}

static int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  if (N % 48 == 0) return 48;
  if (N % 32 == 0) return 32;
  return 0;
}

std::string gemm_prepare(const GemmProblem& p, int num_sms, GemmLaunch* out) {
  if (p.K % 64 != 0) return "gemm: K must be a multiple of 64 (got " + std::to_string(p.K) + ")";
  const int bn = pick_bn(p.N);
  if (bn == 0) return "gemm: N must be a multiple of 32 or 48 (got " + std::to_string(p.N) + ")";
  if (p.epi.glu && (bn % 32 != 0)) return "gemm: GLU epilogue needs N % 32 == 0";
  GemmDev& d = out->dev;
  d.M = p.M; d.N = p.N; d.K = p.K; d.Bz = p.Bz; d.G = p.G;
  d.tiles_m = (p.M + 127) / 128;
  d.tiles_n = p.N / bn;
  d.num_tiles = d.tiles_m * d.tiles_n * p.Bz * p.G;
  d.num_kb = p.K / 64;
  // CTA pairs (tcgen05 cta_group::2, 256-row tiles) when there are enough tiles to give every pair one: below that the
  // single-CTA kernel keeps more SMs busy
  out->mc = (bn >= 128 && d.tiles_m >= 2 && d.num_tiles >= num_sms) ? 2 : 0;
  d.units_m = out->mc ? (d.tiles_m + 1) / 2 : d.tiles_m;
  d.num_units = d.units_m * d.tiles_n * p.Bz * p.G;
  d.a_kb_per_row = p.a_kb_per_row;
  d.a_g_col = p.a_g_col;
  d.a_g_row = p.a_g_row;
  d.w_batched = p.w_batch_stride != 0;
  d.a = p.a; d.w = p.w;
  d.a_cols = p.a_cols; d.a_rows = p.a_rows; d.a_row_stride = p.a_row_stride; d.a_batch_stride = p.a_batch_stride;
  d.epi = p.epi;
  out->bn = bn;
  if (out->mc) {
    const int clusters = d.num_units < num_sms / 2 ? d.num_units : num_sms / 2;
    out->grid = 2 * clusters;
  } else {
    out->grid = d.num_units < num_sms ? d.num_units : num_sms;
  }
  out->smem = bn == 256 ? TcCfg<256>::SMEM : bn == 128 ? TcCfg<128>::SMEM : bn == 64 ? TcCfg<64>::SMEM
              : bn == 48 ? TcCfg<48>::SMEM : TcCfg<32>::SMEM;
  if ((p.a_row_stride * 2) % 16 || (p.a_batch_stride * 2) % 16 || (reinterpret_cast<uintptr_t>(p.a) % 16))
    return "gemm: A view must be 16-byte aligned in base and strides";
  {
    uint64_t dims[3] = {(uint64_t)p.a_cols, (uint64_t)p.a_rows, (uint64_t)p.a_batches};
    uint64_t str[2] = {(uint64_t)p.a_row_stride * 2, (uint64_t)(p.a_batches > 1 ? p.a_batch_stride : p.a_row_stride * p.a_rows) * 2};
    uint32_t box[3] = {64, 128, 1};
    W2S_TRY(make_tensor_map_bf16(&out->mapA, p.a, 3, dims, str, box));
  }
  {
    const uint64_t rs = (uint64_t)(p.w_row_stride ? p.w_row_stride : p.K);
    const uint64_t wr = (uint64_t)(p.w_rows ? p.w_rows : p.N);
    const uint64_t gs = (uint64_t)(p.w_g_stride ? p.w_g_stride : (long long)p.K * p.N);
    const uint64_t bs = p.w_batch_stride ? (uint64_t)p.w_batch_stride : gs * (uint64_t)p.G;
    uint64_t dims[4] = {(uint64_t)p.K, wr, (uint64_t)p.G, (uint64_t)(p.w_batch_stride ? p.Bz : 1)};
    uint64_t str[3] = {rs * 2, gs * 2, bs * 2};
    uint32_t box[4] = {64, (uint32_t)(out->mc ? bn / 2 : bn), 1, 1};
    W2S_TRY(make_tensor_map_bf16(&out->mapW, p.w, 4, dims, str, box));
  }
  // pair kernel: TMA-store epilogue (not for the GLU epilogue, whose output width differs from the tile width)
  out->tma_out = 0;
  if (out->mc == 2 && !p.epi.glu) {
    const uint64_t es = p.epi.out_fp32 ? 4 : 2;
    const uint64_t ldb = p.Bz > 1 ? (uint64_t)p.epi.ldb : (uint64_t)p.epi.ldm * p.M;
    const uint64_t ldg = p.G > 1 ? (uint64_t)p.epi.ldg : ldb * p.Bz;
    uint64_t dims[4] = {(uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.Bz, (uint64_t)p.G};
    uint64_t str[3] = {(uint64_t)p.epi.ldm * es, ldb * es, ldg * es};
    uint32_t box[4] = {p.epi.out_fp32 ? 32u : 64u, 32, 1, 1};
    if (str[0] % 16 == 0 && str[1] % 16 == 0 && str[2] % 16 == 0 && reinterpret_cast<uintptr_t>(p.epi.out) % 16 == 0) {
      W2S_TRY(make_tensor_map(&out->mapOut, p.epi.out, p.epi.out_fp32, 4, dims, str, box));
      out->tma_out = 1;
    }
  }
  return "";
}

std::string gemm_launch_tc(const GemmLaunch& l, cudaStream_t s) {
  if (l.mc == 2) return gemm2_launch(l, s);
  switch (l.bn) {
    case 256: W2S_CUDA_OK(launch_pdl(gemm_tc_kernel<256>, dim3(l.grid), dim3(384), l.smem, s, 1, l.mapA, l.mapW, l.dev)); break;
    case 128: W2S_CUDA_OK(launch_pdl(gemm_tc_kernel<128>, dim3(l.grid), dim3(384), l.smem, s, 1, l.mapA, l.mapW, l.dev)); break;
    case 64: W2S_CUDA_OK(launch_pdl(gemm_tc_kernel<64>, dim3(l.grid), dim3(384), l.smem, s, 1, l.mapA, l.mapW, l.dev)); break;
    case 48: W2S_CUDA_OK(launch_pdl(gemm_tc_kernel<48>, dim3(l.grid), dim3(384), l.smem, s, 1, l.mapA, l.mapW, l.dev)); break;
    case 32: W2S_CUDA_OK(launch_pdl(gemm_tc_kernel<32>, dim3(l.grid), dim3(384), l.smem, s, 1, l.mapA, l.mapW, l.dev)); break;
    default: return "gemm: bad BN";
  }
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

std::string gemm_launch_simt(const GemmLaunch& l, cudaStream_t s) {
  const GemmDev& d = l.dev;
  const long long tiles = (long long)((d.M + 63) / 64) * ((d.N + 63) / 64) * d.Bz * d.G;
  gemm_simt_kernel<<<(unsigned)tiles, 256, 0, s>>>(d);
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

}  // namespace w2s
