// Micro-benchmark 2: tcgen05.mma N=64 rate under concurrent TMEM traffic / smem traffic / dq-like issue pattern.
#include <cstdio>
#include "common.cuh"
using namespace b200;
namespace b200 { int set_error(int c, const char*, ...) { return c; } void count_launch() {} int num_sms() { return 148; } }

// MODE bit0: warps 2,3 hammer TMEM with tcgen05.ld/st; bit1: warps 4..7 hammer smem with st.shared (other region);
// bit2: dq-like pattern (8 SS acc0, 8 SS acc1, commit, 8 TS acc2, commit); bit3: warps 2,3 hammer ld only
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 256) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc<1>(&slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_tf32(128, 64, 0, 0);
      constexpr uint32_t idesc_g = make_idesc_tf32(128, 64, 0, 1);
      const uint64_t ad = make_smem_desc(smem_u32(smem), 16, 1024, kLayoutSw128);
      const uint64_t bd = make_smem_desc(smem_u32(smem) + 32768, 16, 1024, kLayoutSw128);
      const uint64_t md = make_smem_desc(smem_u32(smem) + 49152, 8192, 512, kLayoutSw128Base32);
      const long long t0 = clock64();
      int n = 0;
      uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) {
        if (MODE & 4) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_tf32<1>(tb, desc_advance(ad, (kk >> 2) * 16384 + (kk & 3) * 32), desc_advance(bd, (kk >> 2) * 8192 + (kk & 3) * 32), idesc_s, kk != 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_tf32<1>(tb + 128, desc_advance(ad, (kk >> 2) * 16384 + (kk & 3) * 32), desc_advance(bd, (kk >> 2) * 8192 + (kk & 3) * 32), idesc_s, kk != 0);
          umma_commit<1>(&bar[0]);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_tf32_ts(tb + 256, tb + 192 + kk * 8, desc_advance(md, kk * 1024), idesc_g, 1);
          umma_commit<1>(&bar[1]);
          mbar_wait(&bar[0], ph); mbar_wait(&bar[1], ph); ph ^= 1;   // keep the barriers cycling
          n += 24;
        } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_tf32<1>(tb + (i & 1) * 64, desc_advance(ad, (kk & 3) * 32), desc_advance(bd, (kk & 3) * 32), idesc_s, 1);
          n += 8;
        }
      }
      if (!(MODE & 4)) { umma_commit<1>(&bar[0]); mbar_wait(&bar[0], 0); }
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t2 - t0; out[1] = n; }
      stop = 1;
    }
  } else if (warp == 2 || warp == 3) {
    if (MODE & (1 | 8)) {
      uint32_t v[32];
      const uint32_t a = tb + ((uint32_t)((warp & 3) * 32) << 16) + 384;
      while (!stop) {
        tmem_ld_32x32(a, v);
        tmem_ld_wait();
        if (MODE & 1) { tmem_st_32x32(a + 32, v); tmem_st_wait(); }
      }
    }
  } else if (warp >= 4) {
    if (MODE & 2) {
      float4* dst = reinterpret_cast<float4*>(smem + 65536) + (warp - 4) * 512;
      const float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
      while (!stop) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j * 32 + lane] = val;   // 16 x 512 B per warp iteration
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<1>(tb, 512); }
}

template <int MODE>
void run(const char* name) {
  long long* out; cudaMalloc(&out, 16);
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  kern<<<148, 256, 120 * 1024>>>(out, 256);
  kern<<<148, 256, 120 * 1024>>>(out, 256);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("%-60s : %.1f cyc/mma  %s\n", name, (double)h[0] / (double)h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  run<0>("SS N=64 alone");
  run<8>("SS N=64 + 2 warps tcgen05.ld loop");
  run<1>("SS N=64 + 2 warps tcgen05.ld/st loop");
  run<2>("SS N=64 + 4 warps st.shared loop");
  run<3>("SS N=64 + tmem ld/st + st.shared");
  run<4>("dq pattern (8 SS, 8 SS, commit, 8 TS, commit, wait) alone");
  run<5>("dq pattern + tmem ld/st");
  run<6>("dq pattern + st.shared");
  run<7>("dq pattern + both");
  return 0;
}
