// Attention core entry points (reference enhancing/modules/stage1/layers.py:124-130:
// q k^T * scale -> softmax -> . v) and the row-wise helper of its backward.
//
// The reference materialises and saves a [B, heads, N, N] fp32 probability tensor per layer
// (6.4 GB at B=128, base).  Here the scores live in tensor memory only (attention_tc.cu:
// tcgen05 / TMEM flash-style forward and backward); backward recomputes them from the saved
// log-sum-exp and needs delta = rowsum(dO * O), computed below.
#include "common.cuh"

namespace b200 {

int attention_forward_tc(const float*, void*, int, float*, int, int, int, int, float, int, int, cudaStream_t);
int attention_backward_tc(const float*, const float*, const float*, const float*, void*, int, const float*, int, int, int, int,
                          float, int, int, cudaStream_t);
int attention_exact_forward(const float*, float*, float*, int, int, int, int, float, int, cudaStream_t);
int attention_exact_backward(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, float, int,
                             cudaStream_t);

// delta[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d].  HBM-bound: one float4 of O and dO per thread, dh/4 adjacent
// lanes per (row, head) reduced with shuffles, so a warp streams 512 contiguous bytes of each tensor.
// O is fp32, or fp16 (OHALF) when the forward stored it for an fp16 to_out GEMM.
template <int LANES, int OHALF, int DHALF>   // lanes per (row, head) = dh / 4: 16 (dh 64) or 8 (dh 32)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const void* __restrict__ o, const void* __restrict__ dout, float* __restrict__ delta, long long groups,
                  int N, int heads) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long total = groups * LANES;          // one thread per float4
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ((total + 31) / 32) * 32; i += stride) {
    float s = 0.f;
    if (i < total) {
      auto load4 = [i](const void* ptr, bool half) {
        if (half) {
          const uint2 h = reinterpret_cast<const uint2*>(ptr)[i];
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
          return make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        return reinterpret_cast<const float4*>(ptr)[i];
      };
      const float4 a = load4(o, OHALF != 0);
      const float4 b = load4(dout, DHALF != 0);
      s = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (i < total && (i & (LANES - 1)) == 0) {
      const long long g = i / LANES;                // (b*N + n) * heads + h
      const int h = (int)(g % heads);
      const long long bn = g / heads;
      const long long b = bn / N, n = bn % N;
      delta[(b * heads + h) * N + n] = s;
    }
  }
}

int attention_forward(const float* qkv, void* out, int out_half, float* lse, int B, int N, int heads, int dh, float scale,
                      int round_out, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  return attention_forward_tc(qkv, out, out_half, lse, B, N, heads, dh, scale, round_out, -1, stream);
}

int attention_delta(const void* out, int out_half, const void* dout, int dout_half, float* delta, int B, int N, int heads, int dh,
                    cudaStream_t stream) {
  const long long groups = (long long)B * N * heads;
  const long long threads = groups * (dh / 4);
  long long blocks = (threads + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  const unsigned g = (unsigned)blocks;
#define B200_DELTA(L, OH, DHF) attn_delta_kernel<L, OH, DHF><<<g, 256, 0, stream>>>(out, dout, delta, groups, N, heads)
  if (dh == 64) {
    if (out_half && dout_half) B200_DELTA(16, 1, 1);
    else if (out_half)         B200_DELTA(16, 1, 0);
    else if (dout_half)        B200_DELTA(16, 0, 1);
    else                       B200_DELTA(16, 0, 0);
  } else {
    if (out_half && dout_half) B200_DELTA(8, 1, 1);
    else if (out_half)         B200_DELTA(8, 1, 0);
    else if (dout_half)        B200_DELTA(8, 0, 1);
    else                       B200_DELTA(8, 0, 0);
  }
#undef B200_DELTA
  B200_LAUNCH_OK("attn_delta_kernel");
  return 0;
}

int attention_backward(const float* qkv, const void* out, int out_half, const float* lse, const float* dout, void* dqkv,
                       int dqkv_half, const float* dqkv_scale, float* delta, int B, int N, int heads, int dh, float scale,
                       int round_out, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  int rc = attention_delta(out, out_half, dout, 0, delta, B, N, heads, dh, stream);
  if (rc) return rc;
  return attention_backward_tc(qkv, dout, lse, delta, dqkv, dqkv_half, dqkv_scale, B, N, heads, dh, scale, round_out, -1, stream);
}

// Stage-2 attention (reference enhancing/modules/stage2/layers.py:76-89): causal mask whose first cond_len tokens (the
// condition prefix) see each other fully -- query q attends to key k iff k <= max(q, cond_len - 1).  Same packed qkv layout
// as the stage-1 core (the reference's (T, B*nh, hs) views are that layout transposed).  exact != 0: 3xTF32 mma.sync kernels
// (precision="parity"); otherwise the tcgen05 kind::tf32 kernels with the mask folded into their padding logic.
int attention_causal_forward(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale, int cond_len,
                             int exact, int round_out, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: head size must be 32 or 64 (got %d)", dh);
  B200_CHECK_ARG(cond_len >= 0 && cond_len <= N, "attention: cond_len %d outside [0, %d]", cond_len, N);
  if (exact) return attention_exact_forward(qkv, out, lse, B, N, heads, dh, scale, cond_len, stream);
  return attention_forward_tc(qkv, out, 0, lse, B, N, heads, dh, scale, round_out, cond_len, stream);
}

int attention_causal_backward(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv, float* delta,
                              int B, int N, int heads, int dh, float scale, int cond_len, int exact, int round_out,
                              cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: head size must be 32 or 64 (got %d)", dh);
  B200_CHECK_ARG(cond_len >= 0 && cond_len <= N, "attention: cond_len %d outside [0, %d]", cond_len, N);
  if (exact) return attention_exact_backward(qkv, out, lse, dout, dqkv, delta, B, N, heads, dh, scale, cond_len, stream);
  int rc = attention_delta(out, 0, dout, 0, delta, B, N, heads, dh, stream);
  if (rc) return rc;
  return attention_backward_tc(qkv, dout, lse, delta, dqkv, 0, nullptr, B, N, heads, dh, scale, round_out, cond_len, stream);
}

}  // namespace b200
