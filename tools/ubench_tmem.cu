// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of issuing warps
// (1, 2 or 4 per TMEM lane quarter).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../enhancing_transformers_b200/csrc -o ubench_tmem.bin ubench_tmem.cu
#include <cstdio>
#include "common.cuh"
using namespace b200;
namespace b200 { int set_error(int c, const char*, ...) { return c; } void count_launch() {} int num_sms() { return 148; } int current_device() { return 0; } int sm_limit() { return 0; } }

template <int MODE>   // 0: ld x32, 1: st x32, 2: ld x16
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<1>(&slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { tmem_ld_32x32(tb + (i & 1) * 32, v); tmem_ld_wait(); acc += v[i & 31]; }
    else if (MODE == 1) { v[0] = acc + i; tmem_st_32x32(tb + (i & 1) * 32, v); tmem_st_wait(); }
    else { uint32_t w[16]; tmem_ld_32x16(tb + (i & 3) * 16, w); tmem_ld_wait(); acc += w[i & 15]; }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0xdeadbeef) out[1] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<1>(slot, 512); }
}

template <int MODE>
void run(const char* name, int warps, int bytes_per_instr) {
  long long* out; cudaMalloc(&out, 16);
  const int iters = 2048;
  k<MODE><<<148, warps * 32>>>(out, iters);
  k<MODE><<<148, warps * 32>>>(out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)h / iters;
  printf("%-10s %2d warps: %.1f cyc per instr per warp, %.1f B/clk/SM %s\n", name, warps, cyc, (double)bytes_per_instr * warps / cyc,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) run<0>("ld.x32", w, 4096);
  for (int w : {4, 8, 16}) run<2>("ld.x16", w, 2048);
  for (int w : {4, 8, 16}) run<1>("st.x32", w, 4096);
  return 0;
}
