// Micro-benchmark: does the ~48-cycle floor of an M=128, N<=64 tcgen05.mma (tools/ubench_mma_f16.cu) belong to the
// instruction or to each SM's datapath?  A 2-CTA cluster issues cta_group::2 MMAs (M = 256: 128 rows per SM, the N columns
// of B split between the two CTAs' shared memories) and the issue / completion cycles per instruction are compared with
// the cta_group::1 figures.  If the floor is per instruction, a pair does twice the work per 48 cycles at N = 64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../enhancing_transformers_b200/csrc -o ubench_mma_cg2.bin ubench_mma_cg2.cu
#include <cstdio>
#include "common.cuh"
using namespace b200;
namespace b200 { int set_error(int c, const char*, ...) { return c; } void count_launch() {} int num_sms() { return 148; } }

__device__ __forceinline__ void umma_f16_ts2(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int NACC, int TS, int BMN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<2>(&slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tb = slot;
  if (leader && warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_f16(256, N, 0, BMN);
    const uint32_t sa = smem_u32(smem), sb = sa + 32768;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int acc = i % NACC;
      const uint32_t d = tb + acc * N;
      const uint64_t bd = BMN ? make_smem_desc(sb + (i & 3) * 2048, 8192, 1024, kLayoutSw128)
                              : make_smem_desc(sb + (i & 3) * 32, 16, 1024, kLayoutSw128);
      if (TS) umma_f16_ts2(d, tb + 384 + (i & 7) * 8, bd, idesc, 1);
      else    umma_f16<2>(d, make_smem_desc(sa + (i & 3) * 32, 16, 1024, kLayoutSw128), bd, idesc, 1);
    }
    const long long t1 = clock64();
    umma_commit<2>(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (!leader && warp == 1 && lane == 0) {
    mbar_wait(&bar, 0);          // the multicast commit arrives on both CTAs' barriers
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<2>(tb, 512); }
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
  printf("cg2 M=256 %-22s N=%3d acc=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %.0f per SM)  %s\n", name, N, NACC,
         (double)h[0] / iters, (double)h[1] / iters, N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  run<256, 1, 0, 0>("SS K-major");  run<128, 1, 0, 0>("SS K-major"); run<128, 2, 0, 0>("SS K-major");
  run<64, 1, 0, 0>("SS K-major");   run<64, 2, 0, 0>("SS K-major");   run<64, 4, 0, 0>("SS K-major");
  run<64, 1, 0, 1>("SS B MN-major"); run<64, 2, 0, 1>("SS B MN-major");
  run<64, 1, 1, 1>("TS B MN-major"); run<64, 2, 1, 1>("TS B MN-major"); run<64, 4, 1, 1>("TS B MN-major");
  run<128, 1, 1, 0>("TS B K-major"); run<128, 2, 1, 0>("TS B K-major");
  run<32, 1, 0, 0>("SS K-major");   run<32, 4, 0, 0>("SS K-major");
  return 0;
}
