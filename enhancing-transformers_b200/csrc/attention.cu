// Attention core entry points (reference enhancing/modules/stage1/layers.py:124-130:
// q k^T * scale -> softmax -> . v) and the row-wise helper of its backward.
//
// The reference materialises and saves a [B, heads, N, N] fp32 probability tensor per layer
// (6.4 GB at B=128, base).  Here the scores live in tensor memory only (attention_tc.cu:
// tcgen05 / TMEM flash-style forward and backward); backward recomputes them from the saved
// log-sum-exp and needs delta = rowsum(dO * O), computed below.
#include "common.cuh"

namespace b200 {

int attention_forward_tc(const float*, float*, float*, int, int, int, int, float, int, cudaStream_t);
int attention_backward_tc(const float*, const float*, const float*, const float*, float*, int, int, int, int, float, int,
                          cudaStream_t);

// delta[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]; one warp per (b, n, h), fully coalesced
__global__ void attn_delta_kernel(const float* __restrict__ o, const float* __restrict__ dout, float* __restrict__ delta,
                                  int B, int N, int heads, int dh) {
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * N * heads;
  if (gw >= total) return;
  const int h = (int)(gw % heads);
  const long long bn = gw / heads;
  const float* po = o + bn * heads * dh + h * dh;
  const float* pd = dout + bn * heads * dh + h * dh;
  float s = 0.f;
  for (int d = lane; d < dh; d += 32) s += po[d] * pd[d];
  s = warp_sum(s);
  if (lane == 0) {
    const long long b = bn / N, n = bn % N;
    delta[(b * heads + h) * N + n] = s;
  }
}

int attention_forward(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale, int round_out,
                      cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  return attention_forward_tc(qkv, out, lse, B, N, heads, dh, scale, round_out, stream);
}

int attention_backward(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv, float* delta,
                       int B, int N, int heads, int dh, float scale, int round_out, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  const long long warps = (long long)B * N * heads;
  attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(out, dout, delta, B, N, heads, dh);
  B200_LAUNCH_OK("attn_delta_kernel");
  return attention_backward_tc(qkv, dout, lse, delta, dqkv, B, N, heads, dh, scale, round_out, stream);
}

}  // namespace b200
