// Fused vector-quantiser kernels (reference enhancing/modules/stage1/quantizers.py:38-92).
//
//   vq_prep : normalise the codebook once per call (quantizers.py:76), emit it transposed
//             [D][K] so the lookup streams it with coalesced 16-byte loads, plus |e_n|^2.
//   vq_fwd  : per 128-token tile: normalise z (:75), distance d = (|z|^2 + |e|^2) - 2 z.e in
//             the reference's association (:78-80) in exact fp32 FMA arithmetic, argmin with
//             lowest-index tie-break (:82), gather + normalise the winner (:85-86), loss partial
//             sums (:89-90), straight-through value z + (q - z) (:60-61); the residual mode's
//             depth loop (:42-57) runs inside the kernel.  The [tokens, n_embed] distance
//             matrix the reference materialises never exists.
//   vq_loss : deterministic reduction of the per-CTA partial sums into the scalar loss.
//   vq_bwd  : closed-form gradients (see oracle/vitvq_oracle.py:vq_backward_np).
#include "common.cuh"

namespace b200 {

constexpr int kVqD = 32;          // embed_dim of every shipped config (configs/*.yaml quantizer.embed_dim)
constexpr int kVqTileM = 128;     // tokens per CTA
constexpr int kVqTileN = 128;     // codes per smem chunk
constexpr int kVqThreads = 256;
constexpr int kVqMaxDepth = 8;
constexpr float kNormEps = 1e-12f;

// ---------------------------------------------------------------------------------------------
__global__ void vq_prep_kernel(const float* __restrict__ E, float* __restrict__ EnT, float* __restrict__ ee, int K, int use_norm) {
  // 8 lanes per code row, one float4 each
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int code = gid >> 3, part = gid & 7;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (code < K) v = *reinterpret_cast<const float4*>(E + (size_t)code * kVqD + part * 4);
  float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float nrm = use_norm ? fmaxf(sqrtf(s), kNormEps) : 1.f;   // use_norm=False: norm is the identity (quantizers.py:24)
  if (use_norm) { v.x /= nrm; v.y /= nrm; v.z /= nrm; v.w /= nrm; }   // true division, as F.normalize does
  float s2 = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
  if (code < K) {
    EnT[(size_t)(part * 4 + 0) * K + code] = v.x;
    EnT[(size_t)(part * 4 + 1) * K + code] = v.y;
    EnT[(size_t)(part * 4 + 2) * K + code] = v.z;
    EnT[(size_t)(part * 4 + 3) * K + code] = v.w;
    if (part == 0) ee[code] = s2;
  }
}

__device__ __forceinline__ float group8_sum(float s) {
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  return s;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
// F.normalize: x / max(|x|, eps).  True division, as ATen does.  use_norm == 0: identity (nrm = 1).
__device__ __forceinline__ float4 normalize8(const float4& v, float& nrm, int use_norm = 1) {
  if (!use_norm) { nrm = 1.f; return v; }
  nrm = fmaxf(sqrtf(group8_sum(dot4(v, v))), kNormEps);
  return make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm);
}
// two independent fp32 FMAs in one instruction (FFMA2, sm_100): d.x += a * b.x, d.y += a * b.y -- each lane rounds
// exactly like a scalar fmaf, so the distances (and therefore the argmin) are bit-identical to the scalar loop,
// at half the FMA issue slots.
__device__ __forceinline__ void ffma2(float2& d, float a, const float2 b) {
  const float2 aa = make_float2(a, a);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(aa)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = smem_u32(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------
// z [M,32] fp32; EnT [32][K]; ee [K]; E [K,32] (un-normalised codebook, gathered for the winner)
// out [M,32]; idx [M,depth] int64; loss_part [depth][gridDim.x]
__global__ void __launch_bounds__(kVqThreads, 2)
vq_fwd_kernel(const float* __restrict__ z, const float* __restrict__ EnT, const float* __restrict__ ee,
              const float* __restrict__ E, float* __restrict__ out, long long* __restrict__ idx_out,
              float* __restrict__ loss_part, int M, int K, int depth, int use_norm) {
  extern __shared__ __align__(16) uint8_t vq_smem[];
  float (*zT)[kVqTileM] = reinterpret_cast<float (*)[kVqTileM]>(vq_smem);                       // normalised residual, transposed
  float (*eT)[kVqD][kVqTileN] = reinterpret_cast<float (*)[kVqD][kVqTileN]>(vq_smem + sizeof(float) * kVqD * kVqTileM);  // chunk ring
  float (*ee_s)[kVqTileN] = reinterpret_cast<float (*)[kVqTileN]>(vq_smem + sizeof(float) * (kVqD * kVqTileM + 2 * kVqD * kVqTileN));
  float* zz_s = reinterpret_cast<float*>(ee_s) + 2 * kVqTileN;
  int* best_s = reinterpret_cast<int*>(zz_s + kVqTileM);
  float* red_s = reinterpret_cast<float*>(best_s + kVqTileM);
  // residual / accumulated-code state of the depth loop, parked in shared memory while the distance loop runs: it is only
  // touched in phases A and C, and keeping its 32 registers live across phase B spilled the 8x8 accumulator tile
  float4* park = reinterpret_cast<float4*>(red_s + 8);          // [8][256]: rreg[ps] at ps*256 + tid, acc[ps] at (4+ps)*256 + tid

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * kVqTileM;
  // phase A/C mapping: 8 lanes per token, 4 passes of 32 tokens
  const int part = tid & 7;
  // phase B mapping: 16 x 16 threads, 8 tokens x 8 codes each
  const int tx = tid & 15, ty = tid >> 4;

  // residual and accumulated code per (token, 4-float part); z itself is re-read at the end
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int tok = m0 + ps * 32 + (tid >> 3);
    park[ps * kVqThreads + tid] = tok < M ? *reinterpret_cast<const float4*>(z + (size_t)tok * kVqD + part * 4) : make_float4(1.f, 0.f, 0.f, 0.f);
    park[(4 + ps) * kVqThreads + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int nchunks = (K + kVqTileN - 1) / kVqTileN;

  for (int t = 0; t < depth; ++t) {
    // ---- phase A: normalise the residual, stage it transposed -------------------------------
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      float nrm;
      const float4 rn = normalize8(park[ps * kVqThreads + tid], nrm, use_norm);
      const int lt = ps * 32 + (tid >> 3);
      zT[part * 4 + 0][lt] = rn.x;
      zT[part * 4 + 1][lt] = rn.y;
      zT[part * 4 + 2][lt] = rn.z;
      zT[part * 4 + 3][lt] = rn.w;
      const float s = group8_sum(dot4(rn, rn));
      if (part == 0) zz_s[lt] = s;
    }
    // ---- phase B: distances + running argmin over the codebook ------------------------------
    auto load_chunk = [&](int c, int buf) {
      const int c0 = c * kVqTileN;
      // 32 rows x 128 floats = 1024 float4; 256 threads x 4
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int f = tid + i * kVqThreads;
        const int row = f >> 5, col4 = (f & 31) * 4;
        const bool ok = c0 + col4 < K;   // K % 4 == 0 enforced by the host
        cp_async16(&eT[buf][row][col4], EnT + (size_t)row * K + (ok ? c0 + col4 : 0), ok);
      }
      if (tid < kVqTileN / 4) {
        const bool ok = c0 + tid * 4 < K;
        cp_async16(&ee_s[buf][tid * 4], ee + (ok ? c0 + tid * 4 : 0), ok);
      }
      cp_async_commit();
    };
    float bestd[8];
    int besti[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { bestd[i] = INFINITY; besti[i] = 0x7fffffff; }
    load_chunk(0, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      if (c + 1 < nchunks) { load_chunk(c + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncthreads();   // chunk c (and, for c == 0, zT/zz_s) visible
      float2 dacc[8][4];     // 8 tokens x 8 codes, code pairs packed for FFMA2
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dacc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 8
      for (int k = 0; k < kVqD; ++k) {
        const float4 za = *reinterpret_cast<const float4*>(&zT[k][ty * 4]);
        const float4 zb = *reinterpret_cast<const float4*>(&zT[k][64 + ty * 4]);
        const float4 ea = *reinterpret_cast<const float4*>(&eT[buf][k][tx * 4]);
        const float4 eb = *reinterpret_cast<const float4*>(&eT[buf][k][64 + tx * 4]);
        const float zv[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
        const float2 ev[4] = {make_float2(ea.x, ea.y), make_float2(ea.z, ea.w), make_float2(eb.x, eb.y), make_float2(eb.z, eb.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) ffma2(dacc[i][j], zv[i], ev[j]);   // same k-order and rounding as sequential fmaf
      }
      const int c0 = c * kVqTileN;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float zz = zz_s[(i < 4 ? 0 : 64) + ty * 4 + (i & 3)];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int lc = (j < 4 ? 0 : 64) + tx * 4 + (j & 3);
          const int code = c0 + lc;
          const float dot = (j & 1) ? dacc[i][j >> 1].y : dacc[i][j >> 1].x;
          const float d = fmaf(-2.f, dot, zz + ee_s[buf][lc]);   // (|z|^2+|e|^2) - 2 z.e
          if (code < K && d < bestd[i]) { bestd[i] = d; besti[i] = code; }   // codes ascend: first min kept
        }
      }
      __syncthreads();   // everyone done with buf before it is refilled
    }
    // reduce (d, idx) lexicographically over the 16 tx lanes that share a token
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bestd[i], o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti[i], o);
        if (od < bestd[i] || (od == bestd[i] && oi < besti[i])) { bestd[i] = od; besti[i] = oi; }
      }
      if (tx == 0) best_s[(i < 4 ? 0 : 64) + ty * 4 + (i & 3)] = besti[i];
    }
    __syncthreads();
    // ---- phase C: gather, normalise, loss, residual update ----------------------------------
    float lsum = 0.f;
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const int lt = ps * 32 + (tid >> 3);
      const int tok = m0 + lt;
      int code = best_s[lt];
      if (code >= K) code = 0;   // only for rows of NaNs: torch.argmin would also return some index
      const float4 e = *reinterpret_cast<const float4*>(E + (size_t)code * kVqD + part * 4);
      float nrm, nrm_r;
      const float4 q = normalize8(e, nrm, use_norm);
      float4 rr = park[ps * kVqThreads + tid], ac = park[(4 + ps) * kVqThreads + tid];
      const float4 rn = normalize8(rr, nrm_r, use_norm);   // same arithmetic as phase A: identical value
      const float dx = q.x - rn.x, dy = q.y - rn.y, dz = q.z - rn.z, dw = q.w - rn.w;
      if (tok < M) {
        lsum += dx * dx + dy * dy + dz * dz + dw * dw;
        if (part == 0) idx_out[(size_t)tok * depth + t] = code;
      }
      rr.x -= q.x; rr.y -= q.y; rr.z -= q.z; rr.w -= q.w;
      ac.x += q.x; ac.y += q.y; ac.z += q.z; ac.w += q.w;
      park[ps * kVqThreads + tid] = rr;
      park[(4 + ps) * kVqThreads + tid] = ac;
    }
    lsum = warp_sum(lsum);
    if ((tid & 31) == 0) red_s[tid >> 5] = lsum;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < kVqThreads / 32; ++w) s += red_s[w];
      loss_part[(size_t)t * gridDim.x + blockIdx.x] = s;
    }
    // (the next depth's phase A rewrites zT only after the __syncthreads above; best_s/red_s are
    //  rewritten only after later barriers)
  }
  // straight-through value: z + (z_q - z)
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int tok = m0 + ps * 32 + (tid >> 3);
    if (tok < M) {
      const float4 zr = *reinterpret_cast<const float4*>(z + (size_t)tok * kVqD + part * 4);
      const float4 ac = park[(4 + ps) * kVqThreads + tid];
      float4 o;
      o.x = zr.x + (ac.x - zr.x);
      o.y = zr.y + (ac.y - zr.y);
      o.z = zr.z + (ac.z - zr.z);
      o.w = zr.w + (ac.w - zr.w);
      *reinterpret_cast<float4*>(out + (size_t)tok * kVqD + part * 4) = o;
    }
  }
}

// loss = mean_t( beta * m_t + m_t ),  m_t = sum_t / (M * D)      (quantizers.py:56-57,89-90)
__global__ void vq_loss_kernel(const float* __restrict__ loss_part, int nparts, int depth, float inv_count, float beta,
                               float* __restrict__ loss_out) {
  __shared__ float red[32];
  float total = 0.f;
  for (int t = 0; t < depth; ++t) {
    float s = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += loss_part[(size_t)t * nparts + i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w];
      const float m = a * inv_count;
      total += beta * m + m;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = depth > 1 ? total / (float)depth : total;
}

// J_n(x)^T v = (v - xh (xh . v)) / max(|x|, eps), 8 lanes per vector (identity when use_norm == 0)
__device__ __forceinline__ float4 norm_jt8(const float4& xh, float nrm, const float4& v, int use_norm) {
  if (!use_norm) return v;
  const float p = group8_sum(dot4(xh, v));
  return make_float4((v.x - xh.x * p) / nrm, (v.y - xh.y * p) / nrm, (v.z - xh.z * p) / nrm, (v.w - xh.w * p) / nrm);
}

// Codebook-gradient scatter.  The reference's nn.Embedding backward is an atomic scatter-add into [K, D]; at
// initialisation an encoder hits only a few dozen codes per 4096 tokens, i.e. thousands of atomics land on the same
// few rows.  Each CTA therefore accumulates into a shared-memory table keyed by code (open addressing, 256 slots x
// 32 floats, shared-memory atomics) and flushes each live slot to global memory once: the global atomic count drops
// from tokens x 32 to CTAs x live codes x 32, whatever the code distribution.  A full table falls back to
// direct global atomics (uniform random codes: correct, just not better than before).
constexpr int kVqBwdSlots = 256;
constexpr int kVqBwdThreads = 256;
struct VqBwdTable {
  int keys[kVqBwdSlots];
  float acc[kVqBwdSlots][kVqD];
};
__device__ __forceinline__ void vq_scatter(VqBwdTable& tb, float* __restrict__ gE, long long code, int part, const float4& v) {
  // lane `part == 0` of the 8-lane group finds / claims the slot, the group shares it
  int slot = -1;
  if (part == 0) {
    unsigned h = ((unsigned)code * 2654435761u) >> 24;
#pragma unroll 1
    for (int probe = 0; probe < 8; ++probe) {
      const int old = atomicCAS(&tb.keys[h], -1, (int)code);
      if (old == -1 || old == (int)code) { slot = (int)h; break; }
      h = (h + 1) & (kVqBwdSlots - 1);
    }
  }
  slot = __shfl_sync(0xffffffffu, slot, (threadIdx.x & 31) & ~7);
  float* dst = slot >= 0 ? &tb.acc[slot][part * 4] : gE + (size_t)code * kVqD + part * 4;
  atomicAdd(dst + 0, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
}

// gz [M,32] (written), gE [K,32] (atomically accumulated; caller zero-fills)
__global__ void __launch_bounds__(kVqBwdThreads)
vq_bwd_kernel(const float* __restrict__ z, const float* __restrict__ E, const long long* __restrict__ idx,
              const float* __restrict__ g_out, const float* __restrict__ g_loss_ptr, float* __restrict__ gz,
              float* __restrict__ gE, int M, int K, int depth, int residual, float beta, int use_norm) {
  __shared__ VqBwdTable tb;
  for (int i = threadIdx.x; i < kVqBwdSlots; i += kVqBwdThreads) tb.keys[i] = -1;
  for (int i = threadIdx.x; i < kVqBwdSlots * kVqD; i += kVqBwdThreads) (&tb.acc[0][0])[i] = 0.f;
  __syncthreads();
  const int part = threadIdx.x & 7;
  const float g_loss = g_loss_ptr ? *g_loss_ptr : 0.f;
  const float c = 2.f / ((float)M * (float)kVqD);
  // 32 tokens per CTA pass; M % 4 == 0 (host) keeps every 8-lane group and every warp uniform
  for (int tok0 = blockIdx.x * (kVqBwdThreads / 8); tok0 < M; tok0 += gridDim.x * (kVqBwdThreads / 8)) {
    const int tok = tok0 + (threadIdx.x >> 3);
    const bool live = tok < M;
    const int tk = live ? tok : 0;
    const float4 zv = *reinterpret_cast<const float4*>(z + (size_t)tk * kVqD + part * 4);
    float4 go = g_out ? *reinterpret_cast<const float4*>(g_out + (size_t)tk * kVqD + part * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!residual) {
      const long long code = idx[tk];
      const float4 e = *reinterpret_cast<const float4*>(E + (size_t)code * kVqD + part * 4);
      float nz, ne;
      const float4 zn = normalize8(zv, nz, use_norm);
      const float4 qn = normalize8(e, ne, use_norm);
      const float4 dzq = make_float4(zn.x - qn.x, zn.y - qn.y, zn.z - qn.z, zn.w - qn.w);
      const float4 jz = norm_jt8(zn, nz, dzq, use_norm);
      const float sz = g_loss * beta * c;
      go.x += sz * jz.x; go.y += sz * jz.y; go.z += sz * jz.z; go.w += sz * jz.w;
      const float4 dqz = make_float4(-dzq.x, -dzq.y, -dzq.z, -dzq.w);
      const float4 je = norm_jt8(qn, ne, dqz, use_norm);
      const float se = live ? g_loss * c : 0.f;
      vq_scatter(tb, gE, code, part, make_float4(se * je.x, se * je.y, se * je.z, se * je.w));
    } else {
      const float gl = g_loss / (float)depth;
      float4 r = zv;
      float4 gq[kVqMaxDepth], gr[kVqMaxDepth], qh[kVqMaxDepth];
      float qn_norm[kVqMaxDepth];
#pragma unroll
      for (int t = 0; t < kVqMaxDepth; ++t) {
        if (t < depth) {
          const long long code = idx[(size_t)tk * depth + t];
          const float4 e = *reinterpret_cast<const float4*>(E + (size_t)code * kVqD + part * 4);
          float nr, ne;
          const float4 rn = normalize8(r, nr, use_norm);
          const float4 qn = normalize8(e, ne, use_norm);
          qh[t] = qn; qn_norm[t] = ne;
          const float s1 = gl * c;
          gq[t] = make_float4(s1 * (qn.x - rn.x), s1 * (qn.y - rn.y), s1 * (qn.z - rn.z), s1 * (qn.w - rn.w));
          const float4 d = make_float4(rn.x - qn.x, rn.y - qn.y, rn.z - qn.z, rn.w - qn.w);
          const float4 j = norm_jt8(rn, nr, d, use_norm);
          const float s2 = gl * beta * c;
          gr[t] = make_float4(s2 * j.x, s2 * j.y, s2 * j.z, s2 * j.w);
          r.x -= qn.x; r.y -= qn.y; r.z -= qn.z; r.w -= qn.w;
        }
      }
      float4 suffix = make_float4(0.f, 0.f, 0.f, 0.f);   // sum_{t > s} gr[t]
#pragma unroll
      for (int s = kVqMaxDepth - 1; s >= 0; --s) {
        if (s < depth) {
          const float4 tot = make_float4(gq[s].x - suffix.x, gq[s].y - suffix.y, gq[s].z - suffix.z, gq[s].w - suffix.w);
          float4 je = norm_jt8(qh[s], qn_norm[s], tot, use_norm);
          if (!live) je = make_float4(0.f, 0.f, 0.f, 0.f);
          const long long code = idx[(size_t)tk * depth + s];
          vq_scatter(tb, gE, code, part, je);
          suffix.x += gr[s].x; suffix.y += gr[s].y; suffix.z += gr[s].z; suffix.w += gr[s].w;
        }
      }
    }
    if (live) *reinterpret_cast<float4*>(gz + (size_t)tok * kVqD + part * 4) = go;
  }
  __syncthreads();
  // flush: one global atomic per (live slot, column)
  for (int i = threadIdx.x; i < kVqBwdSlots * kVqD; i += kVqBwdThreads) {
    const int slot = i / kVqD;
    const int key = tb.keys[slot];
    if (key >= 0) {
      const float v = tb.acc[slot][i % kVqD];
      if (v != 0.f) atomicAdd(gE + (size_t)key * kVqD + (i % kVqD), v);
    }
  }
}

// decode_codes support (vitvqgan.py:81-86): out[m] = sum_t normalize(E[code[m,t]])
__global__ void vq_embed_kernel(const float* __restrict__ E, const long long* __restrict__ codes, float* __restrict__ out,
                                int M, int K, int depth, int use_norm) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int tok = gid >> 3, part = gid & 7;
  if (tok >= M) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < depth; ++t) {
    long long code = codes[(size_t)tok * depth + t];
    if (code < 0 || code >= K) code = 0;
    const float4 e = *reinterpret_cast<const float4*>(E + (size_t)code * kVqD + part * 4);
    float n;
    const float4 q = normalize8(e, n, use_norm);
    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
  }
  *reinterpret_cast<float4*>(out + (size_t)tok * kVqD + part * 4) = a;
}

// ---------------------------------------------------------------------------------------------
constexpr size_t kVqSmemBytes =
    sizeof(float) * (kVqD * kVqTileM + 2 * kVqD * kVqTileN + 2 * kVqTileN + kVqTileM + kVqTileM + kVqThreads / 32) +
    sizeof(float4) * 8 * kVqThreads;

size_t vq_workspace_bytes(int M, int K, int depth) {
  const size_t nblk = (M + kVqTileM - 1) / kVqTileM;
  return ((size_t)kVqD * K + K + (size_t)depth * nblk) * sizeof(float);
}

int vq_forward(const float* z, const float* E, float* out, long long* idx, float* loss, int M, int K, int D, int depth,
               float beta, int use_norm, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  B200_CHECK_ARG(D == kVqD, "vq: embed_dim must be %d (got %d)", kVqD, D);
  B200_CHECK_ARG(M > 0 && K > 0 && K % 4 == 0 && M % 4 == 0, "vq: need M %% 4 == 0 and n_embed %% 4 == 0 (M=%d K=%d)", M, K);
  B200_CHECK_ARG(depth >= 1 && depth <= kVqMaxDepth, "vq: num_quantizers must be in [1,%d]", kVqMaxDepth);
  B200_CHECK_ARG(ws_bytes >= vq_workspace_bytes(M, K, depth), "vq: workspace too small");
  float* EnT = static_cast<float*>(workspace);
  float* ee = EnT + (size_t)kVqD * K;
  float* part = ee + K;
  const int nblk = (M + kVqTileM - 1) / kVqTileM;
  vq_prep_kernel<<<(K * 8 + 255) / 256, 256, 0, stream>>>(E, EnT, ee, K, use_norm);
  B200_LAUNCH_OK("vq_prep_kernel");
  B200_CONFIGURE_SMEM_ONCE(vq_fwd_kernel, kVqSmemBytes);
  vq_fwd_kernel<<<nblk, kVqThreads, kVqSmemBytes, stream>>>(z, EnT, ee, E, out, idx, part, M, K, depth, use_norm);
  B200_LAUNCH_OK("vq_fwd_kernel");
  vq_loss_kernel<<<1, 256, 0, stream>>>(part, nblk, depth, 1.f / ((float)M * (float)kVqD), beta, loss);
  B200_LAUNCH_OK("vq_loss_kernel");
  return 0;
}

int vq_backward(const float* z, const float* E, const long long* idx, const float* g_out, const float* g_loss, float* gz,
                float* gE, int M, int K, int D, int depth, int residual, float beta, int use_norm, cudaStream_t stream) {
  B200_CHECK_ARG(D == kVqD, "vq: embed_dim must be %d (got %d)", kVqD, D);
  B200_CHECK_ARG(M > 0 && M % 4 == 0, "vq: need M %% 4 == 0");
  B200_CHECK_ARG(depth >= 1 && depth <= kVqMaxDepth, "vq: bad depth");
  B200_CUDA_OK(cudaMemsetAsync(gE, 0, (size_t)K * kVqD * sizeof(float), stream));
  int blocks = (M + 31) / 32;
  const int cap = num_sms() * 2;          // few, long-lived CTAs: each aggregates ~M / cap tokens in its table
  if (blocks > cap) blocks = cap;
  vq_bwd_kernel<<<blocks, kVqBwdThreads, 0, stream>>>(z, E, idx, g_out, g_loss, gz, gE, M, K, depth, residual, beta, use_norm);
  B200_LAUNCH_OK("vq_bwd_kernel");
  return 0;
}

int vq_embed(const float* E, const long long* codes, float* out, int M, int K, int D, int depth, int use_norm,
             cudaStream_t stream) {
  B200_CHECK_ARG(D == kVqD, "vq: embed_dim must be %d", kVqD);
  B200_CHECK_ARG(M > 0 && M % 4 == 0 && depth >= 1, "vq_embed: bad sizes");
  vq_embed_kernel<<<(M * 8 + 255) / 256, 256, 0, stream>>>(E, codes, out, M, K, depth, use_norm);
  B200_LAUNCH_OK("vq_embed_kernel");
  return 0;
}

}  // namespace b200
