// K4: grouped positional convolution (HF wav2vec2/modeling_wav2vec2.py:326-379: Conv1d(H, H, k=128, pad=64,
// groups=16), last frame dropped, GELU) + residual, as a tcgen05 kernel specialised for its Toeplitz structure.
//
// For one (coalition, group) the 128 taps all read the SAME input window shifted by one frame per tap.  The generic
// contraction kernel reloads a 128x64 A tile for every tap; here the window (up to 256 + 127 frames x 64 channels,
// 48 KB, 128B swizzle) stays in shared memory and a tap's A operand is a descriptor whose start address is advanced
// by one 128-byte row per tap (a start inside a swizzle atom costs nothing measurable).  Only the 4-8 KB weight slice
// of a tap streams through the TMA ring, and the K steps that would multiply the zero padding of a 32/48-channel
// group are skipped.
//
// Measured on B200: a 128x48x16 MMA costs ~46 cycles here (it is bound by the ~5.5 KB of operands it pulls from
// shared memory, not by the tensor pipe), and with one coalition per unit a tap cost another ~310 cycles of
// weight-ring latency (12 x 6 KB in flight is less than latency x consumption rate).  So one unit is now
// (TWO coalitions, group, 256-frame chunk): every weight slice feeds 4 accumulators (2 coalitions x 2 row halves),
// the ring is as deep as shared memory allows (120 KB in flight) and the weight traffic out of L2 halves.
// A ring stage holds up to 8 taps (one full/empty barrier round trip per stage: ~300 cycles that the MMA queue does
// not hide when it is paid per tap; measured 10.3 / 8.7 / 7.9 / 7.5 ms per C2 step for 1 / 2 / 4 / 8 taps per stage).
#include "gemm.cuh"
#include "gemm_epi.cuh"
#include "kernels.cuh"

namespace w2s {

struct PosConvDev {
  int T, cpg, G, B, kpos, chunks, num_units;
  int tps;   // taps per ring stage (divides kpos): one barrier round trip per tps weight slices
  EpiParams epi;
};

constexpr int PC_WIN_BYTES = 3 * 16384;   // 384 rows x 128 B

template <int NG>
struct PcCfg {
  static constexpr int W_BYTES = NG * 128;
  static constexpr int WSTAGES = NG == 64 ? 15 : 20;
  static constexpr size_t SMEM = 2 * PC_WIN_BYTES + (size_t)WSTAGES * W_BYTES + 1024 + 512;
};

template <int NG>
__global__ void __launch_bounds__(384, 1)
posconv_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const PosConvDev p) {
  using C = PcCfg<NG>;
  constexpr int WS = C::WSTAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sWin = base, sW = base + 2 * PC_WIN_BYTES;
  const uint32_t bars = sW + WS * C::W_BYTES;
  auto wfull = [&](int s) { return bars + 8u * s; };
  auto wempty = [&](int s) { return bars + 8u * (WS + s); };
  const uint32_t winfull = bars + 8u * (2 * WS), winempty = bars + 8u * (2 * WS + 1);
  auto tfull = [&](int a) { return bars + 8u * (2 * WS + 2 + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * WS + 4 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * WS + 6);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + 2 * PC_WIN_BYTES + WS * C::W_BYTES + 8 * (2 * WS + 6));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < WS; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 1);
    }
    mbar_init(winfull, 1);
    mbar_init(winempty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull(i), 1);
      mbar_init(tempty(i), 8);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 10) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns: accumulator set a (alternating units) at 256 a; coalition u of the pair at + 128 u; row half at + 64 half

  if (warp == 8) {
    if (elect_one()) {   // weight slices: one stream across units, never blocked by the windows
      const int nws = WS / p.tps * p.tps;
      int ws = 0;
      uint32_t wphase = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const int g = (unit / p.chunks) % p.G;
        for (int j = 0; j < p.kpos; j += p.tps) {
          mbar_wait(wempty(ws), wphase ^ 1u);
          mbar_expect_tx(wfull(ws), C::W_BYTES * p.tps);
          for (int tt = 0; tt < p.tps; ++tt)
            tma_load_3d(sW + (ws + tt) * C::W_BYTES, &mapW, wfull(ws), (j + tt) * 64, 0, g);
          if ((ws += p.tps) >= nws) {
            ws = 0;
            wphase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 11) {
    if (elect_one()) {   // the two windows of a unit
      int it = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
        const int ch = unit % p.chunks;
        const int g = (unit / p.chunks) % p.G;
        const int b0 = 2 * (unit / (p.chunks * p.G));
        mbar_wait(winempty, ((uint32_t)it & 1u) ^ 1u);
        mbar_expect_tx(winfull, 2 * PC_WIN_BYTES);
        for (int u = 0; u < 2; ++u)   // b0 + 1 == B (odd batch): the box is out of bounds and arrives as zeros
          for (int i = 0; i < 3; ++i)
            tma_load_3d(sWin + u * PC_WIN_BYTES + i * 16384, &mapX, winfull, g * 64, ch * 256 + i * 128, b0 + u);
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, NG);
      constexpr int KSTEPS = NG / 16;   // channels beyond NG are zero padding in both operands
      const int nws = WS / p.tps * p.tps;
      int ws = 0;
      uint32_t wphase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
        const int abuf = it & 1;
        mbar_wait(tempty(abuf), (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_wait(winfull, (uint32_t)it & 1u);
        tc_fence_after();
        for (int j0 = 0; j0 < p.kpos; j0 += p.tps) {
          mbar_wait(wfull(ws), wphase);
          tc_fence_after();
          for (int tt = 0; tt < p.tps; ++tt) {
            const int j = j0 + tt;
            const uint64_t dw = umma_desc_sw128(sW + (ws + tt) * C::W_BYTES);
#pragma unroll
            for (int uh = 0; uh < 4; ++uh) {
              // coalition uh / 2, output rows [128 (uh % 2), + 128): tap j reads window rows 128 (uh % 2) + j ...
              const uint64_t da =
                  umma_desc_sw128(sWin + (uint32_t)(uh >> 1) * PC_WIN_BYTES + (uint32_t)((uh & 1) * 128 + j) * 128u);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16(tmem_base + abuf * 256 + uh * 64, da + 2u * k, dw + 2u * k, idesc, (j | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(wempty(ws));
          if ((ws += p.tps) >= nws) {
            ws = 0;
            wphase ^= 1u;
          }
        }
        umma_commit(winempty);
        umma_commit(tfull(abuf));
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, half = warp >> 2;
    int it = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
      const int ch = unit % p.chunks;
      const int g = (unit / p.chunks) % p.G;
      const int b0 = 2 * (unit / (p.chunks * p.G));
      const int abuf = it & 1;
      mbar_wait(tfull(abuf), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      const int t = ch * 256 + half * 128 + q * 32 + lane;
      const bool row_ok = t < p.T;
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        if (b0 + u >= p.B) break;   // warp-uniform
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + abuf * 256 + u * 128 + half * 64;
#pragma unroll 1
        for (int c = 0; c < NG; c += 16) {
          float v[16];
          tmem_ld_32x16(t0 + c, v);
          if (row_ok) epi_store<16>(p.epi, p.cpg, g, b0 + u, t, c, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(abuf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

struct PosConvPlan {
  CUtensorMap mapX, mapW;
  PosConvDev dev;
  int ng, grid;
};

std::string posconv_init() {
  cudaError_t e = cudaFuncSetAttribute(posconv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PcCfg<32>::SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(posconv_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PcCfg<48>::SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(posconv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PcCfg<64>::SMEM);
  if (e != cudaSuccess) return std::string("cudaFuncSetAttribute(posconv_kernel): ") + cudaGetErrorString(e);
  return "";
}

bool posconv_supported(int H, int G, int kpos) {
  const int cpg = H / G;
  return (cpg == 32 || cpg == 48 || cpg == 64) && kpos >= 1 && kpos <= 128;
}

// x: padded input [B, T + kpos, G*64] bf16; w: [G][cpg][kpos*64] bf16; epilogue as for the generic kernel
std::string posconv_prepare(const __nv_bfloat16* x, const __nv_bfloat16* w, int B, int T, int H, int G, int kpos,
                            const EpiParams& epi, int num_sms, PosConvPlan** out) {
  if (!posconv_supported(H, G, kpos)) return "positional conv kernel: unsupported channels per group";
  PosConvPlan* pl = new PosConvPlan();
  const int cpg = H / G;
  pl->ng = cpg;
  pl->dev.T = T; pl->dev.cpg = cpg; pl->dev.G = G; pl->dev.B = B; pl->dev.kpos = kpos;
  {
    int tps = 8;   // taps per weight-ring stage (one barrier round trip per stage)
    const int slots = cpg == 64 ? PcCfg<64>::WSTAGES : PcCfg<48>::WSTAGES;
    while (tps > 1 && (kpos % tps || slots / tps < 2)) --tps;
    pl->dev.tps = tps;
  }
  pl->dev.chunks = (T + 255) / 256;
  pl->dev.num_units = pl->dev.chunks * G * ((B + 1) / 2);   // a unit covers two coalitions
  pl->dev.epi = epi;
  pl->grid = pl->dev.num_units < num_sms ? pl->dev.num_units : num_sms;
  std::string err;
  {
    uint64_t dims[3] = {(uint64_t)G * 64, (uint64_t)(T + kpos), (uint64_t)B};
    uint64_t str[2] = {(uint64_t)G * 64 * 2, (uint64_t)(T + kpos) * G * 64 * 2};
    uint32_t box[3] = {64, 128, 1};
    err = make_tensor_map_bf16(&pl->mapX, x, 3, dims, str, box);
  }
  if (err.empty()) {
    uint64_t dims[3] = {(uint64_t)kpos * 64, (uint64_t)cpg, (uint64_t)G};
    uint64_t str[2] = {(uint64_t)kpos * 64 * 2, (uint64_t)kpos * 64 * cpg * 2};
    uint32_t box[3] = {64, (uint32_t)cpg, 1};
    err = make_tensor_map_bf16(&pl->mapW, w, 3, dims, str, box);
  }
  if (!err.empty()) {
    delete pl;
    return err;
  }
  *out = pl;
  return "";
}

std::string posconv_launch(const PosConvPlan* pl, cudaStream_t s) {
  switch (pl->ng) {
    case 32: W2S_CUDA_OK(launch_pdl(posconv_kernel<32>, dim3(pl->grid), dim3(384), PcCfg<32>::SMEM, s, 1, pl->mapX, pl->mapW, pl->dev)); break;
    case 48: W2S_CUDA_OK(launch_pdl(posconv_kernel<48>, dim3(pl->grid), dim3(384), PcCfg<48>::SMEM, s, 1, pl->mapX, pl->mapW, pl->dev)); break;
    case 64: W2S_CUDA_OK(launch_pdl(posconv_kernel<64>, dim3(pl->grid), dim3(384), PcCfg<64>::SMEM, s, 1, pl->mapX, pl->mapW, pl->dev)); break;
    default: return "positional conv kernel: bad group width";
  }
  W2S_CUDA_OK(cudaGetLastError());
  return "";
}

void posconv_free(PosConvPlan* pl) { delete pl; }

}  // namespace w2s
