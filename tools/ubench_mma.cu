// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 as a function of N, operand source
// (SS / TS) and accumulator dependency (1, 2 or 4 independent accumulators round-robin).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../enhancing-transformers_b200/csrc -o ubench_mma.bin ubench_mma.cu
#include <cstdio>
#include "common.cuh"
using namespace b200;
namespace b200 { int set_error(int c, const char*, ...) { return c; } void count_launch() {} int num_sms() { return 148; } }

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
    constexpr uint32_t idesc = make_idesc_tf32(128, N, 0, BMN);
    const uint32_t sa = smem_u32(smem), sb = sa + 32768;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int acc = i % NACC;
      const uint32_t d = tb + acc * N;
      const uint64_t bd = BMN ? make_smem_desc(sb + (i & 3) * 1024, 4096, 512, kLayoutSw128Base32)
                              : make_smem_desc(sb + (i & 3) * 32, 16, 1024, kLayoutSw128);
      if (TS) umma_tf32_ts(d, tb + 384 + (i & 7) * 8, bd, idesc, 1);
      else    umma_tf32<1>(d, make_smem_desc(sa + (i & 3) * 32, 16, 1024, kLayoutSw128), bd, idesc, 1);
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
         (double)h[1] / iters, N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
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
