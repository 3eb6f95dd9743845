// "Exact" attention core for the parity mode: the same flash-style algorithm as attention_tc.cu
// (reference enhancing/modules/stage1/layers.py:124-130 and its autograd), but every tensor-core product is the
// error-compensated 3xTF32 form  a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi  (hi = the 10-bit-mantissa truncation the
// tensor core would apply anyway, lo = x - hi, exact in fp32), issued as three mma.sync m16n8k8 instructions with the
// operands split in registers.  fp32-grade results (~2^-21 relative) at a fraction of the tcgen05 kernels' speed:
// it exists so that `precision="parity"` can be checked against the reference with margin, not to be fast.
// Backward is split in two kernels so that no atomics are needed: one CTA per key tile (dK, dV) and one per
// query tile (dQ); scores are recomputed from the saved log-sum-exp.
#include "common.cuh"

namespace b200 {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr uint32_t kTf32Mask = 0xFFFFE000u;
__device__ __forceinline__ uint32_t tf32_hi(uint32_t x) { return x & kTf32Mask; }
__device__ __forceinline__ uint32_t tf32_lo(uint32_t x) { return __float_as_uint(__uint_as_float(x) - __uint_as_float(x & kTf32Mask)) & kTf32Mask; }
// c += a . b with both operands given as raw fp32 bit patterns: small terms first
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  const uint32_t ah[4] = {tf32_hi(a[0]), tf32_hi(a[1]), tf32_hi(a[2]), tf32_hi(a[3])};
  const uint32_t al[4] = {tf32_lo(a[0]), tf32_lo(a[1]), tf32_lo(a[2]), tf32_lo(a[3])};
  mma_tf32(c, al, tf32_hi(b0), tf32_hi(b1));
  mma_tf32(c, ah, tf32_lo(b0), tf32_lo(b1));
  mma_tf32(c, ah, tf32_hi(b0), tf32_hi(b1));
}
__device__ __forceinline__ void cpa16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage a [ROWS x DH] tile (row stride ld floats in global) into smem with row pitch DH+4;
// rows >= nvalid are zero-filled
template <int DH, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(float* s, const float* g, long long ld, int row0, int nvalid, int tid) {
  constexpr int V = DH / 4;
  constexpr int LDS = DH + 4;
#pragma unroll
  for (int i = tid; i < ROWS * V; i += THREADS) {
    const int r = i / V, c = (i % V) * 4;
    const bool ok = row0 + r < nvalid;
    cpa16(s + r * LDS + c, g + (long long)(ok ? row0 + r : 0) * ld + c, ok);
  }
}

// A-operand fragments of a 16 x DH row block read straight from global memory (row stride ld)
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[DH / 8][4], const float* base, long long ld, int row_lo,
                                             int nvalid, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = row_lo + g, r1 = row_lo + g + 8;
  const float* p0 = base + (long long)(r0 < nvalid ? r0 : 0) * ld;
  const float* p1 = base + (long long)(r1 < nvalid ? r1 : 0) * ld;
  const bool ok0 = r0 < nvalid, ok1 = r1 < nvalid;
#pragma unroll
  for (int k = 0; k < DH / 8; ++k) {
    f[k][0] = ok0 ? __float_as_uint(p0[k * 8 + t]) : 0u;
    f[k][1] = ok1 ? __float_as_uint(p1[k * 8 + t]) : 0u;
    f[k][2] = ok0 ? __float_as_uint(p0[k * 8 + t + 4]) : 0u;
    f[k][3] = ok1 ? __float_as_uint(p1[k * 8 + t + 4]) : 0u;
  }
}

// acc[nt] (16 x 8 each, nt over 64/8 column tiles) += A(16 x DH, regs) . T^T where T is a smem
// tile [64][DH+4] whose rows are the output columns ("K-like" read: b = T[n0+g][k0+t])
template <int DH>
__device__ __forceinline__ void mma_rows_x_tileT(float (&acc)[8][4], const uint32_t (&a)[DH / 8][4], const float* T, int lane) {
  constexpr int LDS = DH + 4;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int k = 0; k < DH / 8; ++k)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float* p = T + (nt * 8 + g) * LDS + k * 8 + t;
      mma_3xtf32(acc[nt], a[k], __float_as_uint(p[0]), __float_as_uint(p[4]));
    }
}
// acc[dt] (16 x 8 each, dt over DH/8 column tiles) += P(16 x 64, given as accumulator-layout
// fragments, full fp32) . T where T is a smem tile [64][DH+4] whose rows are the
// contraction index ("V-like" read).  Accumulator columns (2t, 2t+1) of each 8-wide tile serve
// as k-slots (t, t+4), so the tile rows are read in the matching order 2t, 2t+1.
template <int DH>
__device__ __forceinline__ void mma_p_x_tile(float (&acc)[DH / 8][4], const uint32_t (&p)[8][4], const float* T, int lane) {
  constexpr int LDS = DH + 4;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kt = 0; kt < 8; ++kt)
#pragma unroll
    for (int dt = 0; dt < DH / 8; ++dt) {
      const float* q = T + (kt * 8 + 2 * t) * LDS + dt * 8 + g;
      mma_3xtf32(acc[dt], p[kt], __float_as_uint(q[0]), __float_as_uint(q[LDS]));
    }
}
// accumulator (c0,c1,c2,c3) -> A fragment (a0,a1,a2,a3) = (c0,c2,c1,c3), kept in full fp32 (split at the MMA)
__device__ __forceinline__ void acc_to_a(uint32_t (&a)[4], float c0, float c1, float c2, float c3) {
  a[0] = __float_as_uint(c0); a[1] = __float_as_uint(c2); a[2] = __float_as_uint(c1); a[3] = __float_as_uint(c3);
}

// ---------------------------------------------------------------------------------------------
// forward: grid (ceil(N/128), heads, B), 256 threads = 8 warps x 16 query rows
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256)
attn_exact_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse, int N, int heads,
                float scale, int cond) {
  constexpr int LDS = DH + 4;
  constexpr int TILE = 64 * LDS;
  extern __shared__ __align__(16) float sm[];   // K[2][TILE], V[2][TILE]
  float* Ks = sm;
  float* Vs = sm + 2 * TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = heads * DH;
  const long long ld = 3ll * inner;
  const float* qbase = qkv + (long long)b * N * ld + h * DH;
  const float* kbase = qbase + inner;
  const float* vbase = qbase + 2 * inner;
  const int q0 = blockIdx.x * 128 + warp * 16;
  const float c = scale * kLog2e;

  uint32_t qf[DH / 8][4];
  load_a_frags<DH>(qf, qbase, ld, q0, N, lane);

  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  // cond >= 0: stage-2 mask (reference stage2/layers.py:43-48,83-85): query q sees key k iff k <= max(q, cond - 1)
  // (causal, with the first `cond` condition tokens fully visible to each other).  Key 0 is visible to every row, so the
  // running maximum is finite after the first tile and fully masked tiles contribute exp2(-inf) = 0.
  const bool masked = cond >= 0;
  const int lim0 = masked ? max(q0 + g, cond - 1) : N, lim1 = masked ? max(q0 + g + 8, cond - 1) : N;
  int ntiles = (N + 63) / 64;
  if (masked) ntiles = min(ntiles, max(min(N - 1, (int)blockIdx.x * 128 + 127), cond - 1) / 64 + 1);   // tiles past the CTA's last visible key
  load_tile<DH, 64, 256>(Ks, kbase, ld, 0, N, tid);
  load_tile<DH, 64, 256>(Vs, vbase, ld, 0, N, tid);
  cpa_commit();
  for (int j = 0; j < ntiles; ++j) {
    const int buf = j & 1;
    cpa_wait<0>();
    __syncthreads();
    if (j + 1 < ntiles) {
      load_tile<DH, 64, 256>(Ks + (buf ^ 1) * TILE, kbase, ld, (j + 1) * 64, N, tid);
      load_tile<DH, 64, 256>(Vs + (buf ^ 1) * TILE, vbase, ld, (j + 1) * 64, N, tid);
      cpa_commit();
    }
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    mma_rows_x_tileT<DH>(s, qf, Ks + buf * TILE, lane);
    if ((j + 1) * 64 > N) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = j * 64 + nt * 8 + 2 * t;
        if (key >= N) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (key + 1 >= N) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
    }
    if (masked && j * 64 + 63 > min(lim0, lim1)) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = j * 64 + nt * 8 + 2 * t;
        if (key > lim0) s[nt][0] = -INFINITY;
        if (key + 1 > lim0) s[nt][1] = -INFINITY;
        if (key > lim1) s[nt][2] = -INFINITY;
        if (key + 1 > lim1) s[nt][3] = -INFINITY;
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float al0 = exp2f((m0 - mn0) * c), al1 = exp2f((m1 - mn1) * c);   // exp2(-inf) = 0 on the first tile
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
    uint32_t pf[8][4];
    const float mc0 = m0 * c, mc1 = m1 * c;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] * c - mc0), p1 = exp2f(s[nt][1] * c - mc0);
      const float p2 = exp2f(s[nt][2] * c - mc1), p3 = exp2f(s[nt][3] * c - mc1);
      l0 += p0 + p1; l1 += p2 + p3;
      acc_to_a(pf[nt], p0, p1, p2, p3);
    }
    mma_p_x_tile<DH>(o, pf, Vs + buf * TILE, lane);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  float* ob = out + (long long)b * N * inner + h * DH;
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) {
    float2 a = make_float2(o[dt][0] * inv0, o[dt][1] * inv0);
    float2 bb = make_float2(o[dt][2] * inv1, o[dt][3] * inv1);
    if (r0 < N) *reinterpret_cast<float2*>(ob + (long long)r0 * inner + dt * 8 + 2 * t) = a;
    if (r1 < N) *reinterpret_cast<float2*>(ob + (long long)r1 * inner + dt * 8 + 2 * t) = bb;
  }
  if (t == 0) {
    float* lb = lse + ((long long)b * heads + h) * N;
    if (r0 < N) lb[r0] = m0 * scale + logf(l0);
    if (r1 < N) lb[r1] = m1 * scale + logf(l1);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, key side: grid (ceil(N/64), heads, B), 128 threads = 4 warps x 16 keys.
// Works on transposed score tiles S^T[key, query] so that P^T / dS^T come out of the tensor
// cores already in A-operand position for dV += P^T dO and dK += dS^T Q.
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128)
attn_exact_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta, float* __restrict__ dqkv, int N, int heads, float scale, int cond) {
  constexpr int LDS = DH + 4;
  constexpr int TILE = 64 * LDS;
  extern __shared__ __align__(16) float sm[];   // Q[2][TILE], dO[2][TILE], lse[2][64], delta[2][64]
  float* Qs = sm;
  float* Ds = sm + 2 * TILE;
  float* Ls = sm + 4 * TILE;
  float* Es = Ls + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = heads * DH;
  const long long ld = 3ll * inner;
  const float* qbase = qkv + (long long)b * N * ld + h * DH;
  const float* dobase = dout + (long long)b * N * inner + h * DH;
  const float* lb = lse + ((long long)b * heads + h) * N;
  const float* eb = delta + ((long long)b * heads + h) * N;
  const int k0 = blockIdx.x * 64 + warp * 16;
  const float c = scale * kLog2e;

  uint32_t kf[DH / 8][4], vf[DH / 8][4];
  load_a_frags<DH>(kf, qbase + inner, ld, k0, N, lane);
  load_a_frags<DH>(vf, qbase + 2 * inner, ld, k0, N, lane);
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

  auto stage = [&](int j, int buf) {
    load_tile<DH, 64, 128>(Qs + buf * TILE, qbase, ld, j * 64, N, tid);
    load_tile<DH, 64, 128>(Ds + buf * TILE, dobase, inner, j * 64, N, tid);
    if (tid < 64) {
      const int q = j * 64 + tid;
      Ls[buf * 64 + tid] = q < N ? lb[q] * kLog2e : INFINITY;   // +inf -> P = 0 for padded queries
      Es[buf * 64 + tid] = q < N ? eb[q] : 0.f;
    }
    cpa_commit();
  };
  const int ntiles = (N + 63) / 64;
  // stage-2 mask: key k receives from query q iff k <= max(q, cond - 1).  A key tile that starts at or past `cond` is
  // invisible to every query tile before its own.
  const bool masked = cond >= 0;
  const int jbeg = (masked && (int)blockIdx.x * 64 >= cond) ? (int)blockIdx.x : 0;
  const int kr0 = k0 + g, kr1 = k0 + g + 8;
  stage(jbeg, 0);
  for (int j = jbeg; j < ntiles; ++j) {
    const int buf = (j - jbeg) & 1;
    cpa_wait<0>();
    __syncthreads();
    if (j + 1 < ntiles) stage(j + 1, buf ^ 1);
    const float* Q = Qs + buf * TILE;
    const float* dO = Ds + buf * TILE;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    mma_rows_x_tileT<DH>(s, kf, Q, lane);            // S^T = K Q^T
    uint32_t pf[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float la = Ls[buf * 64 + nt * 8 + 2 * t], lbb = Ls[buf * 64 + nt * 8 + 2 * t + 1];
      s[nt][0] = exp2f(s[nt][0] * c - la);  s[nt][1] = exp2f(s[nt][1] * c - lbb);
      s[nt][2] = exp2f(s[nt][2] * c - la);  s[nt][3] = exp2f(s[nt][3] * c - lbb);
      if (masked) {
        const int va = max(j * 64 + nt * 8 + 2 * t, cond - 1), vb = max(j * 64 + nt * 8 + 2 * t + 1, cond - 1);   // last key each query sees
        if (kr0 > va) s[nt][0] = 0.f;
        if (kr0 > vb) s[nt][1] = 0.f;
        if (kr1 > va) s[nt][2] = 0.f;
        if (kr1 > vb) s[nt][3] = 0.f;
      }
      acc_to_a(pf[nt], s[nt][0], s[nt][1], s[nt][2], s[nt][3]);
    }
    mma_p_x_tile<DH>(dv, pf, dO, lane);              // dV += P^T dO
    float dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    mma_rows_x_tileT<DH>(dp, vf, dO, lane);          // dP^T = V dO^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float ea = Es[buf * 64 + nt * 8 + 2 * t], ebb = Es[buf * 64 + nt * 8 + 2 * t + 1];
      acc_to_a(pf[nt], s[nt][0] * (dp[nt][0] - ea), s[nt][1] * (dp[nt][1] - ebb), s[nt][2] * (dp[nt][2] - ea),
               s[nt][3] * (dp[nt][3] - ebb));
    }
    mma_p_x_tile<DH>(dk, pf, Q, lane);               // dK += dS^T Q
  }
  const int r0 = k0 + g, r1 = k0 + g + 8;
  float* dkb = dqkv + (long long)b * N * ld + inner + h * DH;
  float* dvb = dkb + inner;
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) {
    float2 a = make_float2(dk[dt][0] * scale, dk[dt][1] * scale), a2 = make_float2(dk[dt][2] * scale, dk[dt][3] * scale);
    float2 v = make_float2(dv[dt][0], dv[dt][1]), v2 = make_float2(dv[dt][2], dv[dt][3]);
    if (r0 < N) {
      *reinterpret_cast<float2*>(dkb + (long long)r0 * ld + dt * 8 + 2 * t) = a;
      *reinterpret_cast<float2*>(dvb + (long long)r0 * ld + dt * 8 + 2 * t) = v;
    }
    if (r1 < N) {
      *reinterpret_cast<float2*>(dkb + (long long)r1 * ld + dt * 8 + 2 * t) = a2;
      *reinterpret_cast<float2*>(dvb + (long long)r1 * ld + dt * 8 + 2 * t) = v2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, query side: grid (ceil(N/64), heads, B), 128 threads = 4 warps x 16 queries
// ---------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128)
attn_exact_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                   const float* __restrict__ delta, float* __restrict__ dqkv, int N, int heads, float scale, int cond) {
  constexpr int LDS = DH + 4;
  constexpr int TILE = 64 * LDS;
  extern __shared__ __align__(16) float sm[];   // K[2][TILE], V[2][TILE]
  float* Ks = sm;
  float* Vs = sm + 2 * TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = heads * DH;
  const long long ld = 3ll * inner;
  const float* qbase = qkv + (long long)b * N * ld + h * DH;
  const float* kbase = qbase + inner;
  const float* vbase = qbase + 2 * inner;
  const float* dobase = dout + (long long)b * N * inner + h * DH;
  const int q0 = blockIdx.x * 64 + warp * 16;
  const float c = scale * kLog2e;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  const float* lb = lse + ((long long)b * heads + h) * N;
  const float* eb = delta + ((long long)b * heads + h) * N;
  const float lse0 = r0 < N ? lb[r0] * kLog2e : INFINITY, lse1 = r1 < N ? lb[r1] * kLog2e : INFINITY;
  const float e0 = r0 < N ? eb[r0] : 0.f, e1 = r1 < N ? eb[r1] : 0.f;

  uint32_t qf[DH / 8][4], dof[DH / 8][4];
  load_a_frags<DH>(qf, qbase, ld, q0, N, lane);
  load_a_frags<DH>(dof, dobase, inner, q0, N, lane);
  float dq[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

  const bool masked = cond >= 0;
  const int lim0 = masked ? max(r0, cond - 1) : N, lim1 = masked ? max(r1, cond - 1) : N;
  int ntiles = (N + 63) / 64;
  if (masked) ntiles = min(ntiles, max(min(N - 1, (int)blockIdx.x * 64 + 63), cond - 1) / 64 + 1);
  load_tile<DH, 64, 128>(Ks, kbase, ld, 0, N, tid);
  load_tile<DH, 64, 128>(Vs, vbase, ld, 0, N, tid);
  cpa_commit();
  for (int j = 0; j < ntiles; ++j) {
    const int buf = j & 1;
    cpa_wait<0>();
    __syncthreads();
    if (j + 1 < ntiles) {
      load_tile<DH, 64, 128>(Ks + (buf ^ 1) * TILE, kbase, ld, (j + 1) * 64, N, tid);
      load_tile<DH, 64, 128>(Vs + (buf ^ 1) * TILE, vbase, ld, (j + 1) * 64, N, tid);
      cpa_commit();
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
    mma_rows_x_tileT<DH>(s, qf, Ks + buf * TILE, lane);     // S = Q K^T
    mma_rows_x_tileT<DH>(dp, dof, Vs + buf * TILE, lane);   // dP = dO V^T
    uint32_t pf[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = j * 64 + nt * 8 + 2 * t;
      const bool ka = key < N, kb = key + 1 < N;
      const float p0 = (ka && key <= lim0) ? exp2f(s[nt][0] * c - lse0) : 0.f, p1 = (kb && key + 1 <= lim0) ? exp2f(s[nt][1] * c - lse0) : 0.f;
      const float p2 = (ka && key <= lim1) ? exp2f(s[nt][2] * c - lse1) : 0.f, p3 = (kb && key + 1 <= lim1) ? exp2f(s[nt][3] * c - lse1) : 0.f;
      acc_to_a(pf[nt], p0 * (dp[nt][0] - e0), p1 * (dp[nt][1] - e0), p2 * (dp[nt][2] - e1), p3 * (dp[nt][3] - e1));
    }
    mma_p_x_tile<DH>(dq, pf, Ks + buf * TILE, lane);        // dQ += dS K
  }
  float* dqb = dqkv + (long long)b * N * ld + h * DH;
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) {
    float2 a = make_float2(dq[dt][0] * scale, dq[dt][1] * scale), a2 = make_float2(dq[dt][2] * scale, dq[dt][3] * scale);
    if (r0 < N) *reinterpret_cast<float2*>(dqb + (long long)r0 * ld + dt * 8 + 2 * t) = a;
    if (r1 < N) *reinterpret_cast<float2*>(dqb + (long long)r1 * ld + dt * 8 + 2 * t) = a2;
  }
}


// ---------------------------------------------------------------------------------------------
template <int DH>
static int exact_fwd_launch(const float* qkv, float* out, float* lse, int B, int N, int heads, float scale, int cond, cudaStream_t s) {
  constexpr size_t smem = 4 * 64 * (DH + 4) * sizeof(float);
  auto kern = attn_exact_fwd_kernel<DH>;
  B200_CONFIGURE_SMEM_ONCE(kern, smem);
  kern<<<dim3((N + 127) / 128, heads, B), 256, smem, s>>>(qkv, out, lse, N, heads, scale, cond);
  B200_LAUNCH_OK("attn_exact_fwd_kernel");
  return 0;
}

template <int DH>
static int exact_bwd_launch(const float* qkv, const float* dout, const float* lse, const float* delta, float* dqkv, int B,
                            int N, int heads, float scale, int cond, cudaStream_t s) {
  constexpr size_t smem_kv = (4 * 64 * (DH + 4) + 256) * sizeof(float);
  constexpr size_t smem_q = 4 * 64 * (DH + 4) * sizeof(float);
  auto k1 = attn_exact_bwd_dkv_kernel<DH>;
  auto k2 = attn_exact_bwd_dq_kernel<DH>;
  B200_CONFIGURE_SMEM_ONCE(k1, smem_kv);
  B200_CONFIGURE_SMEM_ONCE(k2, smem_q);
  const dim3 grid((N + 63) / 64, heads, B);
  k1<<<grid, 128, smem_kv, s>>>(qkv, dout, lse, delta, dqkv, N, heads, scale, cond);
  B200_LAUNCH_OK("attn_exact_bwd_dkv_kernel");
  k2<<<grid, 128, smem_q, s>>>(qkv, dout, lse, delta, dqkv, N, heads, scale, cond);
  B200_LAUNCH_OK("attn_exact_bwd_dq_kernel");
  return 0;
}

// delta = rowsum(dO * O) through the shared helper of attention.cu
int attention_delta(const void* out, int out_half, const void* dout, int dout_half, float* delta, int B, int N, int heads, int dh,
                    cudaStream_t stream);

// cond_len < 0: no mask (stage 1).  cond_len >= 0: the stage-2 mask, causal with a fully visible prefix of cond_len tokens.
int attention_exact_forward(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale,
                            int cond_len, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  B200_CHECK_ARG(B <= 65535 && heads <= 65535, "attention: grid too large");
  B200_CHECK_ARG(cond_len <= N, "attention: cond_len %d exceeds the sequence length %d", cond_len, N);
  if (dh == 64) return exact_fwd_launch<64>(qkv, out, lse, B, N, heads, scale, cond_len, stream);
  return exact_fwd_launch<32>(qkv, out, lse, B, N, heads, scale, cond_len, stream);
}

int attention_exact_backward(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv, float* delta,
                             int B, int N, int heads, int dh, float scale, int cond_len, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(cond_len <= N, "attention: cond_len %d exceeds the sequence length %d", cond_len, N);
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  B200_CHECK_ARG(B <= 65535 && heads <= 65535, "attention: grid too large");
  int rc = attention_delta(out, 0, dout, 0, delta, B, N, heads, dh, stream);
  if (rc) return rc;
  if (dh == 64) return exact_bwd_launch<64>(qkv, dout, lse, delta, dqkv, B, N, heads, scale, cond_len, stream);
  return exact_bwd_launch<32>(qkv, dout, lse, delta, dqkv, B, N, heads, scale, cond_len, stream);
}

}  // namespace b200
