// Micro-benchmark: issue rate of tcgen05.mma kind::f16 as a function of N, operand source
// (SS / TS) and accumulator dependency (1, 2 or 4 independent accumulators round-robin).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../enhancing_transformers_b200/csrc -o ubench_mma.bin ubench_mma.cu
#include <cstdio>
#include "common.cuh"
using namespace b200;
namespace b200 { int set_error(int c, const char*, ...) { return c; } void count_launch() {} int num_sms() { return 148; } }


__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N, int NACC, int TS, int BMN>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<1>(&slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, 0, BMN);
    const uint32_t sa = smem_u32(smem), sb = sa + 32768;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int acc = i % NACC;
      const uint32_t d = tb + acc * N;
      const uint64_t bd = BMN ? make_smem_desc(sb + (i & 3) * 2048, 8192, 1024, kLayoutSw128)
                              : make_smem_desc(sb + (i & 3) * 32, 16, 1024, kLayoutSw128);
      if (TS) umma_f16_ts(d, tb + 384 + (i & 7) * 8, bd, idesc, 1);
      else    umma_f16_ss(d, make_smem_desc(sa + (i & 3) * 32, 16, 1024, kLayoutSw128), bd, idesc, 1);
    }
    const long long t1 = clock64();
    umma_commit<1>(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<1>(tb, 512); }
}

template <int N, int NACC, int TS, int BMN>
void run(const char* name) {
  long long* out; cudaMalloc(&out, 16);
  auto kern = k<N, NACC, TS, BMN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  kern<<<148, 128, 100 * 1024>>>(out, iters);
  kern<<<148, 128, 100 * 1024>>>(out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d acc=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %.0f)  %s\n", name, N, NACC, (double)h[0] / iters,
         (double)h[1] / iters, N / 2.0 /* f16 K=16: same cycles as tf32 K=8 if the rate is 2x */, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  run<256, 1, 0, 0>("SS K-major");  run<256, 2, 0, 0>("SS K-major");
  run<128, 1, 0, 0>("SS K-major");  run<128, 2, 0, 0>("SS K-major");  run<128, 3, 0, 0>("SS K-major");
  run<64, 1, 0, 0>("SS K-major");   run<64, 2, 0, 0>("SS K-major");   run<64, 4, 0, 0>("SS K-major");
  run<64, 1, 0, 1>("SS B MN-major"); run<64, 2, 0, 1>("SS B MN-major"); run<64, 4, 0, 1>("SS B MN-major");
  run<256, 1, 0, 1>("SS B MN-major");
  run<64, 1, 1, 1>("TS B MN-major"); run<64, 2, 1, 1>("TS B MN-major"); run<64, 4, 1, 1>("TS B MN-major");
  run<128, 1, 1, 0>("TS B K-major"); run<128, 2, 1, 0>("TS B K-major");
  run<32, 1, 0, 0>("SS K-major");   run<32, 4, 0, 0>("SS K-major");
  return 0;
}
