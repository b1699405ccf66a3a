// Pipe-rate / latency microbenchmarks for the open questions of the attention and epilogue kernels:
// which pipe does cvt.rn.bf16x2 share, what does a satisfied mbarrier.try_wait cost the issuing thread, what are the
// tcgen05.ld latency and the per-SM MUFU / packed-FP32 rates on THIS part.  Not part of the product; build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
// Every test runs `warps` warps on ONE SM-sized CTA per SM and reports cycles per warp-instruction per scheduler (SMSP),
// i.e. 1.0 = one instruction per cycle per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t pack(float a, float b) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// mode 0: 8 independent ex2 chains; 1: 8 cvt.bf16x2; 2: 4 ex2 + 4 cvt interleaved; 3: 8 FFMA2; 4: 8 FADD2;
// 5: 4 ex2 + 4 FFMA2; 6: 8 rcp; 7: 8 FMNMX; 8: 4 ex2 + 4 FMNMX
template <int mode>
__global__ void pipe_kernel(float seed, long long* cycles, float* sink) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
  float2 w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = make_float2(v[i], v[i] * 0.5f);
  uint32_t u[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ex2(v[i]);
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { u[i] ^= pack(v[i], v[(i + 1) & 7]); v[i] += 1.0f; }   // (the add is a second pipe: see mode 9)
    } else if (mode == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] = ex2(v[i]); u[i] ^= pack(v[i + 4], v[i]); }
    } else if (mode == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __ffma2_rn(w[i], make_float2(0.999f, 0.999f), make_float2(0.001f, 0.001f));
    } else if (mode == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __fadd2_rn(w[i], make_float2(0.001f, 0.002f));
    } else if (mode == 5) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] = ex2(v[i]); w[i] = __ffma2_rn(w[i], make_float2(0.999f, 0.999f), make_float2(0.001f, 0.001f)); }
    } else if (mode == 6) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = rcp(v[i]);
    } else if (mode == 7) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], v[(i + 3) & 7] * 0.5f);
    } else if (mode == 8) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] = ex2(v[i]); v[i + 4] = fmaxf(v[i + 4], v[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += 1.0f;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + w[i].x + w[i].y + __uint_as_float(u[i] & 0x3f800000u);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// a satisfied mbarrier.try_wait.parity, back to back from one thread
__global__ void mbar_kernel(long long* cycles) {
  __shared__ uint64_t bar;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(addr));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");   // phase 0 complete
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ok_all = 0;
    const long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(addr), "r"(0u) : "memory");
      ok_all += ok;
      if (ok == 0) break;   // make every wait depend on the previous result, like a real wait loop
    }
    const long long t1 = clock64();
    cycles[0] = t1 - t0;
    cycles[1] = ok_all;
  }
}

// tcgen05.ld 32x32b.x32: latency of one load + wait, and back-to-back throughput, from `warps` warps
__global__ void tmem_kernel(long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
  const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < 64; ++it) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr + (uint32_t)((it & 3) * 32))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(r[it & 31] & 0x3f800000u);
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
  CK(cudaMalloc(&sink, sizeof(float) * 1024 * 1024));
  const char* names[] = {"8 x MUFU.EX2", "8 x cvt.rn.bf16x2 (+FADD)", "4 x EX2 + 4 x cvt.bf16x2", "8 x FFMA2", "8 x FADD2",
                         "4 x EX2 + 4 x FFMA2", "8 x MUFU.RCP", "8 x FMNMX (+FMUL)", "4 x EX2 + 4 x FMNMX", "8 x FADD"};
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 10; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {   // first launch warms the instruction cache
        switch (mode) {
          case 0: pipe_kernel<0><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 1: pipe_kernel<1><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 2: pipe_kernel<2><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 3: pipe_kernel<3><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 4: pipe_kernel<4><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 5: pipe_kernel<5><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 6: pipe_kernel<6><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 7: pipe_kernel<7><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          case 8: pipe_kernel<8><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
          default: pipe_kernel<9><<<sms, warps * 32>>>(0.3f, cyc, sink); break;
        }
        CK(cudaDeviceSynchronize());
      }
      long long h[8];
      CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      const double per_smsp_instr = (double)ITERS * 8 * warps / 4.0;   // 8 instructions of the named mix per iteration
      printf("warps/SM %2d  %-28s %7.2f cycles per named instruction per scheduler\n", warps, names[mode], (double)h[0] / per_smsp_instr);
    }
  }
  mbar_kernel<<<1, 32>>>(cyc);
  CK(cudaDeviceSynchronize());
  {
    long long h[2];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("satisfied mbarrier.try_wait, dependent chain: %.1f cycles each (%lld succeeded)\n", (double)h[0] / 256.0, h[1]);
  }
  for (int warps : {1, 4, 8}) {
    tmem_kernel<<<sms, warps * 32>>>(cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h[1];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("tcgen05.ld 32x32b.x32 + wait, %d warps/SM: %.1f cycles per load per warp\n", warps, (double)h[0] / 64.0);
  }
  return 0;
}
