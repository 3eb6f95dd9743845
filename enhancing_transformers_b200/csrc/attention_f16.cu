// fp16-operand flash-attention forward and backward on tcgen05 / TMEM (reference layers.py:124-130), dim_head = 64.
//
// The same warp-specialised structure as attention_tc.cu (TMA producer, two MMA-issuing warps, eight softmax
// warps with a query row shared by two threads) with every tensor-core operand in fp16 (kind::f16, fp32
// accumulation in TMEM): q/k/v arrive as the fp16 output of the to_qkv GEMM, dO as the fp16 (gradient-scaled) output of
// the to_out dgrad GEMM, and P / dS are written back to TMEM as packed fp16 pairs.  fp16 carries the same 11-bit
// significand as the tf32 operands of attention_tc.cu, so the numerics are unchanged; what changes is
//   * every MMA does twice the work per issue slot (UMMA_K = 16) -- the backward kernels sat on their MMA-issue floor;
//   * a 64-wide head row is exactly one 128-byte swizzle row, so ONE shared-memory image of a tile serves both as a
//     K-major operand (rows = M/N index) and as an MN-major operand (rows = contraction index): the second TMA fetch of
//     Q / dO / K with a different swizzle that the tf32 kernels need disappears, and tiles are half the bytes;
//   * P / dS take half the TMEM columns and half the tcgen05.st traffic.
// Gradients: dO carries the power-of-two gradient scale S of the backward segment (functional.py), and so do delta,
// dS and the dq / dk / dv this kernel stores (fp16): everything is linear in dO, nothing is rescaled here.
#include <cstdlib>
#include "common.cuh"

namespace b200 {

namespace {

// With fp16 operands the MMAs take half the issue slots and the softmax warps become the critical path (measured: the
// 8-softmax-warp port of the tf32 kernels ran at exactly the tf32 kernels' speed).  Hence SIXTEEN softmax warps: four per
// TMEM lane quarter, a row shared by four threads, each owning a quarter of the tile's columns -- half the serial work per
// warp and twice the warps per scheduler to hide tcgen05.ld/st, barrier and ex2 latencies behind each other.
constexpr int kThreads = 608;      // warp 0 TMA, warp 1 MMA issuer A, warps 2..17 softmax (four per TMEM lane quarter), warp 18 MMA issuer B
constexpr int kIssuerB = 18;
constexpr int kSoftmaxWarps = 16;
constexpr int DH = 64;             // head dim: one 128-byte row of fp16
constexpr int kTile128 = 128 * 128;   // bytes of a 128-row tile
constexpr int kTile64 = 64 * 128;     // bytes of a 64-row tile
constexpr float kLog2eF = 1.4426950408889634f;
constexpr int kBoxBytes = kSoftmaxWarps * 2048;   // one 32-row x 64-byte store box per softmax warp
// Shared-memory rings.  The loop-carried tiles are small in fp16 (16 KB per 128 rows), so the rings are deep enough to
// cover a full TMA round trip (~1500 cycles from "stage free" to "bytes landed"): with the 2-deep rings inherited from
// the tf32 kernels every sub-tile waited for its own load (ncu: the softmax warps' top stall is the wait for S).
constexpr int kFwdStages = 4;      // K and V rings of the forward kernel (128-row tiles)
// Backward rings ([Q64 | dO64] of the dKV kernel, [K64 | V64] of the dQ kernel).  A stage is released by the LAST MMA that
// reads it (the gradient MMA of its sub-tile, which runs two to three sub-tiles after the score MMA), so of N stages only
// N - 3 are ahead of the score issuer: with 4 the issuer waited ~370 cycles for every stage (attn16_trace).
#ifndef ATTN_DKV_STAGES
#define ATTN_DKV_STAGES 6
#endif
#ifndef ATTN_DQ_STAGES
#define ATTN_DQ_STAGES 6
#endif
constexpr int kDkvStages = ATTN_DKV_STAGES;
constexpr int kDqStages = ATTN_DQ_STAGES;

// named barrier of the four softmax warps that share TMEM lane quarter q (ids 2..5, 128 threads)
__device__ __forceinline__ void quad_bar(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 2) : "memory"); }

// D[tmem] (+)= A[tmem] * B[smem], fp16 inputs: A holds M = 128 rows (lanes) x K = 16 as 8 columns of packed pairs
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {      // values known to be in fp16 range
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}


// Development trace (only with -DB200_ATTN16_TRACE, built into a separate .so by `make trace16`): block 0 of the dKV kernel
// logs (event id, clock64) pairs of the two issuer warps and one softmax warp into shared memory (tools/gpu_probe.py
// attn16_trace prints the merged timeline).
#ifdef B200_ATTN16_TRACE
#ifndef B200_TRACE_SKIP
#define B200_TRACE_SKIP 320
#endif
constexpr int kTraceCap = 64;
__device__ long long g_trace16[3 * 2 * kTraceCap];
#define TRACE_DECL                                                                                \
  __shared__ long long tr_buf[3][2 * kTraceCap]; int tr_n = 0;                                    \
  for (int i_ = threadIdx.x; i_ < 3 * 2 * kTraceCap; i_ += blockDim.x) (&tr_buf[0][0])[i_] = 0
#define TRACE(role, ev)                                                                           \
  do {                                                                                            \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                                             \
      const int k_ = tr_n++ - B200_TRACE_SKIP;                                                    \
      if (k_ >= 0 && k_ < kTraceCap) { tr_buf[role][2 * k_] = (ev); tr_buf[role][2 * k_ + 1] = clock64(); } \
    }                                                                                             \
  } while (0)
#define TRACE_DUMP()                                                                              \
  do {                                                                                            \
    __syncthreads();                                                                              \
    if (blockIdx.x == 0) for (int i_ = threadIdx.x; i_ < 3 * 2 * kTraceCap; i_ += blockDim.x) g_trace16[i_] = (&tr_buf[0][0])[i_]; \
  } while (0)
#else
#define TRACE_DECL
#define TRACE(role, ev) do {} while (0)
#define TRACE_DUMP() do {} while (0)
#endif

// ---- 2^x on the FMA pipe (tools/exp2_poly.py) --------------------------------------------------------------
// The softmax of all three kernels is bounded by MUFU.EX2 (16 lanes / clk / SM: ncu shows the XU pipe saturated while
// the exp phase runs and every warp stalled on its queue), while the FMA pipe idles.  A fixed share of each thread's
// exponentials is therefore evaluated as a degree-4 polynomial (max rel. error 2.7e-6, below the 4.9e-4 of the fp16
// rounding P gets anyway) with packed f32x2 arithmetic: 11 issue slots per PAIR of values instead of 16 XU cycles.
//     t = max(x, -126);  r = t + 1.5*2^23 (round to integer n in the low mantissa bits);  f = t - (r - 1.5*2^23)
//     p = c0 + f (c1 + f (c2 + f (c3 + f c4)));   2^x = as_float(as_int(p) + (as_int(r) << 23))
#ifndef ATTN_FWD_POLY_PAIRS
#define ATTN_FWD_POLY_PAIRS 4       // of the 16 pairs a forward thread exponentiates per key tile
#endif
#ifndef ATTN_BWD_POLY_PAIRS
#define ATTN_BWD_POLY_PAIRS 1       // of the 8 pairs a backward thread exponentiates per sub-tile
#endif
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void ex2_poly_pair(float x0, float x1, float& y0, float& y1) {
  constexpr float kMagic = 12582912.f;
  const unsigned long long t = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const unsigned long long r = f2_add(t, f2_pack(kMagic, kMagic));
  const unsigned long long nf = f2_add(r, f2_pack(-kMagic, -kMagic));
  const unsigned long long f = f2_fma(nf, f2_pack(-1.f, -1.f), t);
  unsigned long long pl = f2_fma(f2_pack(0.00957007f, 0.00957007f), f, f2_pack(0.05591786f, 0.05591786f));
  pl = f2_fma(pl, f, f2_pack(0.24024746f, 0.24024746f));
  pl = f2_fma(pl, f, f2_pack(0.6931218f, 0.6931218f));
  pl = f2_fma(pl, f, f2_pack(0.9999993f, 0.9999993f));
  float p0, p1, r0, r1;
  f2_unpack(pl, p0, p1);
  f2_unpack(r, r0, r1);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}
// pair index i of n: the polynomial pairs are spread evenly between the MUFU pairs so both pipes stay fed
template <int POLY, int N>
__device__ __forceinline__ constexpr bool poly_slot(int i) { return POLY > 0 && ((i + 1) * POLY) / N != (i * POLY) / N; }
// use_poly is a compile-time constant at every call site once the surrounding loop is unrolled
__device__ __forceinline__ void ex2_pair(bool use_poly, float x0, float x1, float& y0, float& y1) {
  if (use_poly) ex2_poly_pair(x0, x1, y0, y1);
  else { y0 = ex2_approx(x0); y1 = ex2_approx(x1); }
}

// 32 rows x 32 values -> fp16 (x mul, saturating) -> 32-row x 64-byte un-swizzled box -> one TMA store
__device__ __forceinline__ void store_box_h(uint8_t* box, const CUtensorMap* tm, const float (&r)[32], float mul, int c0, int c1, int c2,
                                            int lane) {
  if (lane == 0) bulk_wait_group_read<0>();
  __syncwarp();
  uint8_t* rowp = box + lane * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 pk;
    pk.x = pack_half2_sat(r[8 * j] * mul, r[8 * j + 1] * mul);
    pk.y = pack_half2_sat(r[8 * j + 2] * mul, r[8 * j + 3] * mul);
    pk.z = pack_half2_sat(r[8 * j + 4] * mul, r[8 * j + 5] * mul);
    pk.w = pack_half2_sat(r[8 * j + 6] * mul, r[8 * j + 7] * mul);
    *reinterpret_cast<uint4*>(rowp + j * 16) = pk;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(tm, box, c0, c1, c2);
    bulk_commit_group();
  }
}

// 32 rows x 16 values -> 32-row x 32-byte box
__device__ __forceinline__ void store_box_h16(uint8_t* box, const CUtensorMap* tm, const float (&r)[16], float mul, int c0, int c1, int c2,
                                              int lane) {
  if (lane == 0) bulk_wait_group_read<0>();
  __syncwarp();
  uint8_t* rowp = box + lane * 32;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 pk;
    pk.x = pack_half2_sat(r[8 * j] * mul, r[8 * j + 1] * mul);
    pk.y = pack_half2_sat(r[8 * j + 2] * mul, r[8 * j + 3] * mul);
    pk.z = pack_half2_sat(r[8 * j + 4] * mul, r[8 * j + 5] * mul);
    pk.w = pack_half2_sat(r[8 * j + 6] * mul, r[8 * j + 7] * mul);
    *reinterpret_cast<uint4*>(rowp + j * 16) = pk;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(tm, box, c0, c1, c2);
    bulk_commit_group();
  }
}

// K-major operand descriptor of a tile whose rows are 128 bytes (64 fp16): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t kmajor_desc(const void* tile) { return make_smem_desc(smem_u32(tile), 16, 1024, kLayoutSw128); }
// the same image read as an MN-major B operand (rows = contraction index, 64 columns = N): one 64-wide atom
__device__ __forceinline__ uint64_t mnmajor_desc(const void* tile) { return make_smem_desc(smem_u32(tile), 8192, 1024, kLayoutSw128); }
constexpr uint32_t kKStepK = 32;        // K-major: 16 fp16 along the row
constexpr uint32_t kKStepMN = 2048;     // MN-major: 16 rows of 128 bytes

struct FwdParams {
  float* lse;        // [B*heads*N]
  int N, heads, q_tiles, kv_tiles, total_items;
  float scale;
};

// =============================================================================================
// forward: work item = (batch, head, 128-query tile)
// TMEM columns: S buffers [0,128) [128,256) (P packed over the first 64 columns of its S), PV [256,320) [320,384)
// =============================================================================================
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_f16_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  constexpr int NST = kFwdStages;
  uint8_t* Qs = smem;                              // [2]: the next item's Q tile loads while this item computes
  uint8_t* Ks = smem + 2 * kTile128;               // [NST]
  uint8_t* Vs = smem + (2 + NST) * kTile128;       // [NST]
  uint8_t* obox = smem + (2 + 2 * NST) * kTile128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(obox + kBoxBytes);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;    // [2]
  uint64_t* p_full = bars + 6;    // [2]
  uint64_t* o_full = bars + 8;    // [2]
  uint64_t* o_empty = bars + 10;  // [2]
  uint64_t* sfree = bars + 12;    // [2]
  uint64_t* k_full = bars + 14;             // [NST]
  uint64_t* k_empty = k_full + NST;         // [NST]
  uint64_t* v_full = k_full + 2 * NST;      // [NST]
  uint64_t* v_empty = k_full + 3 * NST;     // [NST]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k_full + 4 * NST);
  float* xch = reinterpret_cast<float*>(k_full + 4 * NST + 2);   // [2][4][128] row-max (double buffered) + [4][128] row-sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TRACE_DECL;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < NST; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], kSoftmaxWarps);
      mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], kSoftmaxWarps);
      mbar_init(&sfree[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int inner = p.heads * DH;
  const int T = p.kv_tiles;

  if (warp == 0) {
    uint32_t kv_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int qt = w % p.q_tiles;
      const int bh = w / p.q_tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      const int qs = item_it & 1;
      mbar_wait(&q_empty[qs], ((item_it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[qs], kTile128);
        tma_load_3d(Qs + qs * kTile128, &tmQKV, &q_full[qs], h * DH, qt * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j, ++kv_it) {
        const int s = kv_it % NST;
        const uint32_t ph = (kv_it / NST) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[s], kTile128);
          tma_load_3d(Ks + s * kTile128, &tmQKV, &k_full[s], inner + h * DH, j * 128, b);
        }
        __syncwarp();
        mbar_wait(&v_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[s], kTile128);
          tma_load_3d(Vs + s * kTile128, &tmQKV, &v_full[s], 2 * inner + h * DH, j * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- issuer A: S_j = Q K_j^T
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
    const uint64_t qd0 = kmajor_desc(Qs);
    const uint64_t kd0 = kmajor_desc(Ks);
    uint32_t s_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int qs = item_it & 1;
      mbar_wait(&q_full[qs], (item_it >> 1) & 1);
      const uint64_t qd = desc_advance(qd0, qs * kTile128);
      for (int j = 0; j < T; ++j, ++s_it) {
        const int s = s_it & 1;                       // TMEM S buffer
        const uint32_t ph = (s_it >> 1) & 1;
        const int r = s_it % NST;                     // K ring stage
        TRACE(0, 200);
        mbar_wait2(&k_full[r], (s_it / NST) & 1, &sfree[s], ph ^ 1);
        TRACE(0, 201);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t kd = desc_advance(kd0, r * kTile128);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_f16<1>(tmem_base + s * 128, desc_advance(qd, k * kKStepK), desc_advance(kd, k * kKStepK), idesc_s, k != 0);
          umma_commit<1>(&s_full[s]);
          umma_commit<1>(&k_empty[r]);
          if (j == T - 1) umma_commit<1>(&q_empty[qs]);
        }
        __syncwarp();
        TRACE(0, 202);
      }
    }
  } else if (warp == kIssuerB) {
    // ---- issuer B: O_j = P_j V_j (A = packed fp16 P from TMEM, B = V read MN-major)
    constexpr uint32_t idesc_o = make_idesc_f16(128, DH, 0, 1);
    const uint64_t vd0 = mnmajor_desc(Vs);
    uint32_t pv_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      for (int j = 0; j < T; ++j, ++pv_it) {
        const int s = pv_it & 1;
        const uint32_t ph = (pv_it >> 1) & 1;
        const int r = pv_it % NST;                    // V ring stage
        TRACE(1, 210);
        mbar_wait3(&v_full[r], (pv_it / NST) & 1, &o_empty[s], ph ^ 1, &p_full[s], ph);
        TRACE(1, 211);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t vd = desc_advance(vd0, r * kTile128);
#pragma unroll
          for (int k = 0; k < 8; ++k)      // 128 keys = 8 x UMMA_K
            umma_f16_ts(tmem_base + 256 + s * 64, tmem_base + s * 128 + k * 8, desc_advance(vd, k * kKStepMN), idesc_o, k != 0);
          umma_commit<1>(&o_full[s]);
          umma_commit<1>(&v_empty[r]);
          umma_commit<1>(&sfree[s]);
        }
        __syncwarp();
        TRACE(1, 212);
      }
    }
  } else {
    // ---- softmax / output warps: a query row (TMEM lane) is shared by four threads (warps of the same lane quarter),
    // each owning 32 of the tile's 128 scores and 16 of the 64 output columns; the row max (per tile) and the row sum
    // (once per item) are exchanged through shared memory under a 128-thread named barrier.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    constexpr int OC = DH / 4;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const float c = p.scale * kLog2eF;
    uint32_t t_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      const int qt = w % p.q_tiles;
      const int bh = w / p.q_tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      float o[OC];
#pragma unroll
      for (int i = 0; i < OC; ++i) o[i] = 0.f;
      float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
      auto fold_pv = [&](uint32_t it, float alpha) {
        const int sp = it & 1;
        uint32_t v[OC];
        tmem_ld_32x16(tmem_base + lane_off + 256 + sp * 64 + sub * OC, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < OC; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(v[i]));
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[sp]);
      };
      uint32_t v0[32];
      mbar_wait(&s_full[t_it & 1], (t_it >> 1) & 1);
      tcgen05_fence_after();
      tmem_ld_32x32(tmem_base + lane_off + (t_it & 1) * 128 + sub * 32, v0);
      tmem_ld_wait();
      for (int j = 0; j < T; ++j, ++t_it) {
        const int s = t_it & 1;
        const int kv_left = p.N - j * 128 - sub * 32;       // this thread's columns >= kv_left are padding
        if (kv_left < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= kv_left) v0[i] = 0xff800000u;          // -inf
        }
        if (warp == 2) TRACE(2, 220);
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v0[i]));
        float* xs = xch + (t_it & 1) * 512;
        xs[sub * 128 + row_in_tile] = mx;
        quad_bar(q);        // also orders the partners' S loads before this thread's packed P stores (they overlap their columns)
        mx = fmaxf(fmaxf(xs[row_in_tile], xs[128 + row_in_tile]), fmaxf(xs[256 + row_in_tile], xs[384 + row_in_tile]));
        if (warp == 2) TRACE(2, 221);
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2_approx((m - m_new) * c);
        const float mc = m_new * c;
        float sum = 0.f, sum1 = 0.f;
        uint32_t pk[16];       // this thread's 32 probabilities as 16 packed fp16 pairs (keys sub*32 + 2i, +1)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0, a1;
          ex2_pair(poly_slot<ATTN_FWD_POLY_PAIRS, 16>(i), fmaf(__uint_as_float(v0[2 * i]), c, -mc), fmaf(__uint_as_float(v0[2 * i + 1]), c, -mc), a0, a1);
          sum += a0; sum1 += a1;
          pk[i] = pack_h2(a0, a1);
        }
        sum += sum1;
        if (warp == 2) TRACE(2, 222);
        tmem_st_32x16(tmem_base + lane_off + s * 128 + sub * 16, pk);     // P columns [0,64) of the S buffer
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (warp == 2) TRACE(2, 223);
        l = fmaf(l, alpha, sum);
        const bool more = j + 1 < T;
        const uint32_t ph_s = ((t_it + 1) >> 1) & 1, ph_o = ((t_it - 1) >> 1) & 1;
        if (more && j >= 1) mbar_wait2(&s_full[s ^ 1], ph_s, &o_full[s ^ 1], ph_o);
        else if (more)      mbar_wait(&s_full[s ^ 1], ph_s);
        else if (j >= 1)    mbar_wait(&o_full[s ^ 1], ph_o);
        tcgen05_fence_after();
        if (warp == 2) TRACE(2, 224);
        if (more) tmem_ld_32x32(tmem_base + lane_off + (s ^ 1) * 128 + sub * 32, v0);
        if (j >= 1) fold_pv(t_it - 1, alpha_prev);
        else        tmem_ld_wait();
        if (warp == 2) TRACE(2, 225);
        alpha_prev = alpha;
        m = m_new;
      }
      mbar_wait(&o_full[(t_it - 1) & 1], ((t_it - 1) >> 1) & 1);
      tcgen05_fence_after();
      fold_pv(t_it - 1, alpha_prev);
      float* ls = xch + 1024;
      ls[sub * 128 + row_in_tile] = l;
      quad_bar(q);
      l = (ls[row_in_tile] + ls[128 + row_in_tile]) + (ls[256 + row_in_tile] + ls[384 + row_in_tile]);
      const int row = qt * 128 + row_in_tile;
      store_box_h16(obox + (warp - 2) * 2048, &tmO, o, 1.f / l, h * DH + sub * OC, qt * 128 + q * 32, b, lane);
      if (row < p.N && sub == 0) p.lse[((long long)b * p.heads + h) * p.N + row] = m * p.scale + logf(l);
      quad_bar(q);     // ls is rewritten by the next item only after every partner has read it
    }
    if (lane == 0) bulk_wait_group_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
#ifdef B200_ATTN16_TRACE_FWD
  TRACE_DUMP();
#endif
}


// =============================================================================================
// forward, second design: TWO 128-query tiles per CTA, one softmax group of eight warps per tile, O accumulated in TMEM
//
// In the kernel above all 16 softmax warps work on the same score tile and move in lock-step (row max -> barrier -> exp ->
// store -> fold): the XU pipe is saturated during the exp phase and idle otherwise (attn16_trace: 1100 of a tile's 2500
// cycles), and both MMA issuers wait ~1800 cycles per tile.  Here a CTA owns a PAIR of query tiles that share every K / V
// tile (half the K / V traffic per query), each tile has its own softmax group (two threads per query row, 64 scores each)
// and the groups run out of phase: one exponentiates while the other waits for its P V / next Q K^T.  What makes the
// per-thread state fit is that O never enters registers: P V accumulates in tensor memory across key tiles
// (accumulate = 1) and is rescaled in place only when a row's running maximum has grown by more than 2^8 since the scale in
// use was chosen ("lazy rescaling": P is then at most 256, well inside fp16; l is kept relative to the same stale maximum,
// so O / l is exact) -- a rare event after the first tiles of a row.
//   TMEM columns: S_g at g*128 (P packed fp16 over its first 64 columns), O_g at 256 + g*64
//   warps: 0 TMA, 1 MMA issuer, 2..9 group 0, 10..17 group 1 (lane quarter = warp & 3, column half = ((warp - 2) & 7) >> 2)
// =============================================================================================
constexpr int kThreads2 = 576;
constexpr int kV2Stages = 3;         // K and V rings (128-row tiles)
constexpr float kRescaleLog2 = 8.f;  // rescale O when the row maximum has grown by more than this many powers of two
#ifndef ATTN_FWD2_POLY_PAIRS
#define ATTN_FWD2_POLY_PAIRS 8       // of the 32 pairs a thread exponentiates per key tile
#endif

__device__ __forceinline__ void pair_bar2(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

struct Fwd2Params {
  float* lse;        // [B*heads*N]
  int N, heads, q_pairs, kv_tiles, total_items;
  float scale;
};

__global__ void __launch_bounds__(kThreads2, 1)
attn_fwd2_f16_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const Fwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  constexpr int NST = kV2Stages;
  uint8_t* Qs = smem;                                  // [2 item buffers][2 query tiles]
  uint8_t* Ks = smem + 4 * kTile128;                   // [NST]
  uint8_t* Vs = smem + (4 + NST) * kTile128;           // [NST]
  uint8_t* obox = smem + (4 + 2 * NST) * kTile128;     // 16 x 2 KB store boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(obox + kBoxBytes);
  uint64_t* q_full = bars + 0;     // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* s_full = bars + 4;     // [2 groups]
  uint64_t* p_full = bars + 6;     // [2]
  uint64_t* pv_done = bars + 8;    // [2]
  uint64_t* o_free = bars + 10;    // [2]
  uint64_t* k_full = bars + 12;           // [NST]
  uint64_t* k_empty = k_full + NST;
  uint64_t* v_full = k_full + 2 * NST;
  uint64_t* v_empty = k_full + 3 * NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k_full + 4 * NST);
  float* xch = reinterpret_cast<float*>(bars + 64);    // [2 tile parities][2 groups][2 halves][128] row max, then [2][2][128] row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 8); mbar_init(&pv_done[s], 1); mbar_init(&o_free[s], 8);
    }
    for (int s = 0; s < NST; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int inner = p.heads * DH;
  const int T = p.kv_tiles;

  if (warp == 0) {
    // ---- TMA producer
    uint32_t kv_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int qp = w % p.q_pairs;
      const int bh = w / p.q_pairs;
      const int h = bh % p.heads, b = bh / p.heads;
      const int qs = item_it & 1;
      mbar_wait(&q_empty[qs], ((item_it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[qs], 2 * kTile128);
        tma_load_3d(Qs + qs * 2 * kTile128, &tmQKV, &q_full[qs], h * DH, qp * 256, b);
        tma_load_3d(Qs + qs * 2 * kTile128 + kTile128, &tmQKV, &q_full[qs], h * DH, qp * 256 + 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j, ++kv_it) {
        const int s = kv_it % NST;
        const uint32_t ph = (kv_it / NST) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[s], kTile128);
          tma_load_3d(Ks + s * kTile128, &tmQKV, &k_full[s], inner + h * DH, j * 128, b);
        }
        __syncwarp();
        mbar_wait(&v_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[s], kTile128);
          tma_load_3d(Vs + s * kTile128, &tmQKV, &v_full[s], 2 * inner + h * DH, j * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer.  Order per item: S_0(0) S_1(0), then for every key tile j and group g: P_g(j) V_j, S_g(j+1) --
    // a group's next scores are issued right behind the P V that consumed its probabilities (P lives in the S buffer), so
    // the two groups drift half a period apart and fill each other's tensor-pipe and XU gaps.
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(128, DH, 0, 1);
    const uint64_t qd0 = kmajor_desc(Qs);
    const uint64_t kd0 = kmajor_desc(Ks);
    const uint64_t vd0 = mnmajor_desc(Vs);
    uint32_t kv_it = 0, item_it = 0, tile_it = 0;      // tile_it: key tiles issued so far (per group)
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int qs = item_it & 1;
      mbar_wait(&q_full[qs], (item_it >> 1) & 1);
      tcgen05_fence_after();
      auto issue_s = [&](int g, uint32_t kv) {          // S_g = Q_g K^T of ring stage kv % NST
        const int r = kv % NST;
        if (elect_one()) {
          const uint64_t qd = desc_advance(qd0, (qs * 2 + g) * kTile128);
          const uint64_t kd = desc_advance(kd0, r * kTile128);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_f16<1>(tmem_base + g * 128, desc_advance(qd, k * kKStepK), desc_advance(kd, k * kKStepK), idesc_s, k != 0);
          umma_commit<1>(&s_full[g]);
          if (g == 1) umma_commit<1>(&k_empty[r]);
        }
        __syncwarp();
      };
      mbar_wait(&k_full[kv_it % NST], (kv_it / NST) & 1);
      tcgen05_fence_after();
      issue_s(0, kv_it);
      issue_s(1, kv_it);
      for (int j = 0; j < T; ++j, ++kv_it, ++tile_it) {
        const int r = kv_it % NST;
        for (int g = 0; g < 2; ++g) {
          if (g == 0) mbar_wait2(&v_full[r], (kv_it / NST) & 1, &p_full[0], tile_it & 1);
          else        mbar_wait(&p_full[1], tile_it & 1);
          if (j == 0) mbar_wait(&o_free[g], (item_it & 1) ^ 1);     // the previous item's output has left TMEM
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t vd = desc_advance(vd0, r * kTile128);
#pragma unroll
            for (int k = 0; k < 8; ++k)      // 128 keys = 8 x UMMA_K
              umma_f16_ts(tmem_base + 256 + g * 64, tmem_base + g * 128 + k * 8, desc_advance(vd, k * kKStepMN), idesc_o, (j | k) != 0);
            umma_commit<1>(&pv_done[g]);
            if (g == 1) umma_commit<1>(&v_empty[r]);
          }
          __syncwarp();
          if (j + 1 < T) {
            if (g == 0) { mbar_wait(&k_full[(kv_it + 1) % NST], ((kv_it + 1) / NST) & 1); tcgen05_fence_after(); }
            issue_s(g, kv_it + 1);
          } else if (g == 1) {
            if (elect_one()) umma_commit<1>(&q_empty[qs]);          // every Q K^T of the item has been issued
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ---- softmax groups
    const int g = (warp - 2) >> 3;
    const int q = warp & 3;
    const int hlf = ((warp - 2) & 7) >> 2;
    const int bar_id = 1 + g * 4 + q;                       // named barrier of the two warps that share these 32 query rows
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + g * 128;
    const uint32_t o_addr = tmem_base + lane_off + 256 + g * 64 + hlf * 32;
    const int row_in_tile = q * 32 + lane;
    const float c = p.scale * kLog2eF;
    uint32_t t_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int qp = w % p.q_pairs;
      const int bh = w / p.q_pairs;
      const int h = bh % p.heads, b = bh / p.heads;
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < T; ++j, ++t_it) {
        uint32_t v[64];
        mbar_wait(&s_full[g], t_it & 1);
        tcgen05_fence_after();
        tmem_ld_32x32(s_addr + hlf * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_32x32(s_addr + hlf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        const int kv_left = p.N - j * 128 - hlf * 64;       // this thread's columns >= kv_left are padding
        if (kv_left < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= kv_left) v[i] = 0xff800000u;           // -inf
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        float* xs = xch + (t_it & 1) * 512 + g * 256;
        xs[hlf * 128 + row_in_tile] = mx;
        pair_bar2(bar_id);      // also orders the partner's S loads before this thread's packed P stores (they overlap its columns)
        mx = fmaxf(mx, xs[(hlf ^ 1) * 128 + row_in_tile]);
        if (j == 0) {
          m_used = mx;
        } else {
          const bool grow = (mx - m_used) * c > kRescaleLog2;
          if (__any_sync(0xffffffffu, grow)) {              // rare: bring O and l to the new maximum
            const float m_new = grow ? mx : m_used;
            const float alpha = ex2_approx((m_used - m_new) * c);
            mbar_wait(&pv_done[g], (t_it - 1) & 1);          // P V of the previous key tile has landed in O
            tcgen05_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 32; cc += 8) {
              uint32_t ov[8];
              tmem_ld_32x8(o_addr + cc, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st_32x8(o_addr + cc, ov);
            }
            tmem_st_wait();
            l *= alpha;
            m_used = m_new;
          }
        }
        const float mc = m_used * c;
        float sum = 0.f, sum1 = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int e = hh * 32 + 2 * i;
            float a0, a1;
            ex2_pair(poly_slot<ATTN_FWD2_POLY_PAIRS, 32>(hh * 16 + i), fmaf(__uint_as_float(v[e]), c, -mc), fmaf(__uint_as_float(v[e + 1]), c, -mc), a0, a1);
            sum += a0; sum1 += a1;
            pk[i] = pack_half2_sat(a0, a1);                  // <= 2^8 unless a row jumps by more than the threshold inside one tile
          }
          tmem_st_32x16(s_addr + hlf * 32 + hh * 16, pk);    // packed P: keys hlf*64 + hh*32 .. +32 -> columns hlf*32 + hh*16 .. +16
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        l += sum + sum1;
      }
      // ---- item epilogue: O / l -> fp16 -> TMA store; log-sum-exp
      float* ls = xch + 1024 + g * 256;
      ls[hlf * 128 + row_in_tile] = l;
      mbar_wait(&pv_done[g], (t_it - 1) & 1);
      tcgen05_fence_after();
      uint32_t ov[32];
      tmem_ld_32x32(o_addr, ov);
      tmem_ld_wait();
      tcgen05_fence_before();
      pair_bar2(bar_id);
      l += ls[(hlf ^ 1) * 128 + row_in_tile];
      if (lane == 0) mbar_arrive(&o_free[g]);
      float r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(ov[i]);
      const int row0 = qp * 256 + g * 128 + q * 32;
      store_box_h(obox + (warp - 2) * 2048, &tmO, r, 1.f / l, h * DH + hlf * 32, row0, b, lane);
      const int row = row0 + lane;
      if (row < p.N && hlf == 0) p.lse[((long long)b * p.heads + h) * p.N + row] = m_used * p.scale + logf(l);
      pair_bar2(bar_id);     // ls is rewritten by the next item only after the partner has read it
    }
    if (lane == 0) bulk_wait_group_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// =============================================================================================
// backward (two kernels, no atomics; scores recomputed from the saved log-sum-exp)
//   dKV: item = (batch, head, 128-key tile), loop over 64-query sub-tiles:
//        S^T = K Q^T, dP^T = V dO^T (SS, 128 x 64); P^T, dS^T packed fp16 in TMEM; dV += P^T dO, dK += dS^T Q (TS)
//   dQ : item = (batch, head, 128-query tile), loop over 64-key sub-tiles:
//        S = Q K^T, dP = dO V^T; dS packed; dQ += dS K
// A 64-row fp16 tile is used both K-major (score MMAs) and MN-major (gradient MMAs) from one shared-memory image.
// Packed operands: thread `sub` (0..3) of a row owns 16 of the 64 columns of a sub-tile and writes its 8 packed columns
// at the start of its own 16-column range, so no thread overwrites scores another one has not loaded yet:
// A-operand k-step k (16 contraction indices = thread k's columns) lives at column k * 16 of the buffer.
// =============================================================================================
struct BwdParams {
  const float* lse;
  const float* delta;    // rowsum(dO * O): written by the dQ kernel (delta_w), read by the dKV kernel that runs after it
  float* delta_w;
  int N, heads, tiles128, sub64, total_items;
  float scale;
};

__device__ __forceinline__ uint32_t packed_a_col(int k) { return (uint32_t)(k * 16); }
// TMEM plan of both backward kernels: THREE score buffers (the loop is a chain S-MMA -> softmax -> gradient MMA -> S-MMA of
// the same buffer, dominated by commit / mbarrier / tcgen05.ld-st latencies: its rate is buffers / round-trip time).
//   S or S^T buffers at columns [0,64) [64,128) [128,192); dP or dP^T buffers at [192,256) [256,320) [320,384);
//   accumulators at [384,448) (dV or dQ) and [448,512) (dK)
constexpr int kTB = 3;
constexpr int kPArrivals = kSoftmaxWarps / 2;     // two groups of eight softmax warps take alternate sub-tiles
constexpr uint32_t kColDP = 192, kColAcc = 384;

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkv_f16_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                        const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmOut, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  constexpr int NST = kDkvStages;
  uint8_t* KVs = smem;                        // [2] item buffers [K128 | V128]: the next item's tiles load while this one computes
  uint8_t* St = smem + 4 * kTile128;          // ring stage r at St + r * 2 * kTile64: [Q64 | dO64]
  uint8_t* obox = St + NST * 2 * kTile64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(obox + kBoxBytes);
  uint64_t* kv_full = bars + 0;    // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* acc_full = bars + 4;
  uint64_t* acc_empty = bars + 5;
  uint64_t* s_full = bars + 6;     // [kTB]
  uint64_t* p_full = bars + 9;     // [kTB]
  uint64_t* sfree = bars + 12;     // [kTB]
  uint64_t* qd_full = bars + 15;          // [NST]
  uint64_t* qd_empty = qd_full + NST;     // [NST] two arrivals: the score MMAs and the gradient MMAs have both read the stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qd_full + 2 * NST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TRACE_DECL;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV128); tma_prefetch_desc(&tmQKV64); tma_prefetch_desc(&tmDO64);
    for (int s = 0; s < NST; ++s) { mbar_init(&qd_full[s], 1); mbar_init(&qd_empty[s], 2); }
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < kTB; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], kPArrivals); mbar_init(&sfree[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, kSoftmaxWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int inner = p.heads * DH;
  const int NS = p.sub64;

  if (warp == 0) {
    uint32_t sub_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int kt = w % p.tiles128;
      const int bh = w / p.tiles128;
      const int h = bh % p.heads, b = bh / p.heads;
      const int ib = item_it & 1;
      mbar_wait(&kv_empty[ib], ((item_it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[ib], 2 * kTile128);
        tma_load_3d(KVs + ib * 2 * kTile128, &tmQKV128, &kv_full[ib], inner + h * DH, kt * 128, b);
        tma_load_3d(KVs + ib * 2 * kTile128 + kTile128, &tmQKV128, &kv_full[ib], 2 * inner + h * DH, kt * 128, b);
      }
      __syncwarp();
      for (int i = 0; i < NS; ++i, ++sub_it) {
        const int s = sub_it % NST;
        const uint32_t ph = (sub_it / NST) & 1;
        uint8_t* st = St + s * 2 * kTile64;
        mbar_wait(&qd_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&qd_full[s], 2 * kTile64);
          tma_load_3d(st, &tmQKV64, &qd_full[s], h * DH, i * 64, b);
          tma_load_3d(st + kTile64, &tmDO64, &qd_full[s], h * DH, i * 64, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- issuer A: S^T = K Q^T, dP^T = V dO^T
    constexpr uint32_t idesc_s = make_idesc_f16(128, 64, 0, 0);
    const uint64_t kd0 = kmajor_desc(KVs);
    const uint64_t qd0 = kmajor_desc(St);
    uint32_t sd_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int ib = item_it & 1;
      mbar_wait(&kv_full[ib], (item_it >> 1) & 1);
      const uint64_t kd = desc_advance(kd0, ib * 2 * kTile128);
      const uint64_t vd = desc_advance(kd, kTile128);
      for (int i = 0; i < NS; ++i, ++sd_it) {
        const int s = sd_it % kTB;                    // TMEM S^T / dP^T buffer
        const uint32_t ph = (sd_it / kTB) & 1;
        const int r = sd_it % NST;                    // ring stage
        TRACE(0, 100);
        mbar_wait2(&qd_full[r], (sd_it / NST) & 1, &sfree[s], ph ^ 1);   // probes overlap: each costs ~200 cycles under MMA operand traffic
        TRACE(0, 102);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t qd = desc_advance(qd0, r * 2 * kTile64);
          const uint64_t dd = desc_advance(qd, kTile64);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_f16<1>(tmem_base + s * 64, desc_advance(kd, k * kKStepK), desc_advance(qd, k * kKStepK), idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_f16<1>(tmem_base + kColDP + s * 64, desc_advance(vd, k * kKStepK), desc_advance(dd, k * kKStepK), idesc_s, k != 0);
          umma_commit<1>(&s_full[s]);
          umma_commit<1>(&qd_empty[r]);
          if (i == NS - 1) umma_commit<1>(&kv_empty[ib]);
        }
        __syncwarp();
        TRACE(0, 103);
      }
    }
  } else if (warp == kIssuerB) {
    // ---- issuer B: dV += P^T dO, dK += dS^T Q (A packed fp16 from TMEM, B = the stage's tiles read MN-major)
    constexpr uint32_t idesc_g = make_idesc_f16(128, DH, 0, 1);
    const uint64_t qm0 = mnmajor_desc(St);
    uint32_t dv_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      for (int i = 0; i < NS; ++i, ++dv_it) {
        const int s = dv_it % kTB;
        const uint32_t ph = (dv_it / kTB) & 1;
        const int r = dv_it % NST;
        if (i == 0) mbar_wait(acc_empty, (item_it & 1) ^ 1);
        TRACE(1, 110);
        mbar_wait2(&qd_full[r], (dv_it / NST) & 1, &p_full[s], ph);
        TRACE(1, 111);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t qmd = desc_advance(qm0, r * 2 * kTile64);
          const uint64_t dmd = desc_advance(qmd, kTile64);
          const uint32_t acc_on = i > 0;
#pragma unroll
          for (int k = 0; k < 4; ++k)   // dV += P^T dO      (64 queries = 4 x UMMA_K)
            umma_f16_ts(tmem_base + kColAcc, tmem_base + s * 64 + packed_a_col(k), desc_advance(dmd, k * kKStepMN), idesc_g, acc_on | (k != 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // dK += dS^T Q
            umma_f16_ts(tmem_base + kColAcc + 64, tmem_base + kColDP + s * 64 + packed_a_col(k), desc_advance(qmd, k * kKStepMN), idesc_g, acc_on | (k != 0));
          umma_commit<1>(&qd_empty[r]);
          umma_commit<1>(&sfree[s]);
          if (i == NS - 1) umma_commit<1>(acc_full);
        }
        __syncwarp();
        TRACE(1, 112);
      }
    }
  } else {
    // 16 softmax warps: the four warps of a TMEM lane quarter (key rows) split the 64 query columns of a sub-tile;
    // at the end of an item warps sub 0,1 store dV (32 columns each), warps sub 2,3 store dK.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c = p.scale * kLog2eF;
    float* bc = reinterpret_cast<float*>(bars + 64) + (warp - 2) * 64;     // this warp's lse / delta broadcast slot (2 x 32 floats)
    uint32_t t_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int kt = w % p.tiles128;
      const int bh = w / p.tiles128;
      const int h = bh % p.heads, b = bh / p.heads;
      const float* lb = p.lse + ((long long)b * p.heads + h) * p.N;
      const float* eb = p.delta + ((long long)b * p.heads + h) * p.N;
      // per-column lse / delta: the 16 query columns of this warp's slice are the same for all 32 key rows, so the warp
      // stages them (lanes 0..15: lse * log2e, lanes 16..31: delta; fetched one sub-tile ahead) in a 32-float slot of
      // shared memory and every thread reads them back as eight broadcast 128-bit loads -- the warp-shuffle broadcast
      // this replaces cost two SHFL per score (ncu: 32 of the 260 issue slots a warp spent per sub-tile).
      // Two groups of eight warps take alternate sub-tiles (grp = parity), a warp owning 32 of the 64 query columns as two
      // 16-column chunks: the second chunk's tcgen05.ld travels while the first is computed, and while one group is in its
      // exp phase the other is loading / storing -- with all 16 warps on every sub-tile the sub-tile rate was one over a
      // warp's own ld -> exp -> st -> arrive latency (~1450 cycles, attn16_trace), not the MMA floor (768).
      const int grp = sub & 1, half = sub >> 1;
      const int qcol = half * 32 + lane;
      const uint32_t t_base = item_it * (uint32_t)NS;
      float rawL = 0.f, rawE = 0.f;
      bool nvalid = grp < NS && grp * 64 + qcol < p.N;
      if (nvalid) { rawL = lb[grp * 64 + qcol]; rawE = eb[grp * 64 + qcol]; }
      auto chunk = [&](const uint32_t (&v)[16], const uint32_t (&g)[16], int col, int boff) {
        uint32_t pp[8], ds[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {        // four query columns per step: one broadcast 128-bit load of lse and of delta
          const float4 L = reinterpret_cast<const float4*>(bc + boff)[j];
          const float4 E = reinterpret_cast<const float4*>(bc + 32 + boff)[j];
          float p0, p1, p2, p3;
          ex2_pair(poly_slot<ATTN_BWD_POLY_PAIRS, 8>(2 * j), fmaf(__uint_as_float(v[4 * j]), c, -L.x), fmaf(__uint_as_float(v[4 * j + 1]), c, -L.y), p0, p1);
          ex2_pair(poly_slot<ATTN_BWD_POLY_PAIRS, 8>(2 * j + 1), fmaf(__uint_as_float(v[4 * j + 2]), c, -L.z), fmaf(__uint_as_float(v[4 * j + 3]), c, -L.w), p2, p3);
          pp[2 * j] = pack_h2(p0, p1);
          pp[2 * j + 1] = pack_h2(p2, p3);
          ds[2 * j] = pack_half2_sat(p0 * (__uint_as_float(g[4 * j]) - E.x), p1 * (__uint_as_float(g[4 * j + 1]) - E.y));
          ds[2 * j + 1] = pack_half2_sat(p2 * (__uint_as_float(g[4 * j + 2]) - E.z), p3 * (__uint_as_float(g[4 * j + 3]) - E.w));
        }
        tmem_st_32x8(tmem_base + lane_off + col, pp);
        tmem_st_32x8(tmem_base + lane_off + kColDP + col, ds);
      };
      for (int i = grp; i < NS; i += 2) {
        const uint32_t t = t_base + i;
        const int s = t % kTB;
        bc[lane] = nvalid ? rawL * kLog2eF : INFINITY;   // +inf -> P = 0 for padded queries
        bc[32 + lane] = nvalid ? rawE : 0.f;
        {
          const int qi = (i + 2) * 64 + qcol;
          nvalid = i + 2 < NS && qi < p.N;
          if (nvalid) { rawL = lb[qi]; rawE = eb[qi]; }
        }
        if (warp == 2) TRACE(2, 120);
        mbar_wait(&s_full[s], (t / kTB) & 1);
        if (warp == 2) TRACE(2, 121);
        tcgen05_fence_after();
        __syncwarp();
        const int col = s * 64 + half * 32;
        uint32_t vA[16], gA[16], vB[16], gB[16];
        tmem_ld_32x16(tmem_base + lane_off + col, vA);
        tmem_ld_32x16(tmem_base + lane_off + kColDP + col, gA);
        tmem_ld_wait();
        if (warp == 2) TRACE(2, 122);
        tmem_ld_32x16(tmem_base + lane_off + col + 16, vB);
        tmem_ld_32x16(tmem_base + lane_off + kColDP + col + 16, gB);
        chunk(vA, gA, col, 0);
        tmem_ld_wait();
        chunk(vB, gB, col + 16, 16);
        if (warp == 2) TRACE(2, 123);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (warp == 2) TRACE(2, 124);
      }
      t_it += NS;
      // item epilogue: dV lives in TMEM columns [384,448), dK in [448,512): warp `sub` takes columns 384 + sub*32
      mbar_wait(acc_full, item_it & 1);
      tcgen05_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_off + kColAcc + sub * 32, v);
        tmem_ld_wait();
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
        const int col0 = (sub < 2 ? 2 * inner : inner) + h * DH + (sub & 1) * 32;
        store_box_h(obox + (warp - 2) * 2048, &tmOut, r, sub < 2 ? 1.f : p.scale, col0, kt * 128 + q * 32, b, lane);   // key rows >= N are clipped
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (lane == 0) bulk_wait_group_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
#ifndef B200_ATTN16_TRACE_FWD
  TRACE_DUMP();
#endif
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_f16_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmDO128,
                       const __grid_constant__ CUtensorMap tmO128, const __grid_constant__ CUtensorMap tmQKV64,
                       const __grid_constant__ CUtensorMap tmOut16, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  constexpr int NST = kDqStages;
  constexpr int kItem = 3 * kTile128;
  uint8_t* QDs = smem;                        // [2] item buffers [Q128 | dO128 | O128] (O only feeds delta = rowsum(dO * O))
  uint8_t* St = smem + 2 * kItem;             // ring stage r at St + r * 2 * kTile64: [K64 | V64]
  uint8_t* obox = St + NST * 2 * kTile64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(obox + kBoxBytes / 2);   // 32-row x 32-byte boxes: 1 KB per softmax warp
  uint64_t* q_full = bars + 0;     // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* acc_full = bars + 4;
  uint64_t* acc_empty = bars + 5;
  uint64_t* s_full = bars + 6;     // [kTB]
  uint64_t* p_full = bars + 9;     // [kTB]
  uint64_t* sfree = bars + 12;     // [kTB]
  uint64_t* kv_full = bars + 15;          // [NST]
  uint64_t* kv_empty = kv_full + NST;     // [NST] two arrivals (score MMAs, dQ MMAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_full + 2 * NST);
  float* xch = reinterpret_cast<float*>(bars + 64);   // [2][4][128] partial delta exchange (double buffered by item)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV128); tma_prefetch_desc(&tmDO128); tma_prefetch_desc(&tmO128); tma_prefetch_desc(&tmQKV64);
    for (int s = 0; s < NST; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < kTB; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], kPArrivals); mbar_init(&sfree[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, kSoftmaxWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int inner = p.heads * DH;
  const int NS = p.sub64;

  if (warp == 0) {
    uint32_t sub_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int qt = w % p.tiles128;
      const int bh = w / p.tiles128;
      const int h = bh % p.heads, b = bh / p.heads;
      const int ib = item_it & 1;
      mbar_wait(&q_empty[ib], ((item_it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[ib], kItem);
        tma_load_3d(QDs + ib * kItem, &tmQKV128, &q_full[ib], h * DH, qt * 128, b);
        tma_load_3d(QDs + ib * kItem + kTile128, &tmDO128, &q_full[ib], h * DH, qt * 128, b);
        tma_load_3d(QDs + ib * kItem + 2 * kTile128, &tmO128, &q_full[ib], h * DH, qt * 128, b);
      }
      __syncwarp();
      for (int i = 0; i < NS; ++i, ++sub_it) {
        const int s = sub_it % NST;
        const uint32_t ph = (sub_it / NST) & 1;
        uint8_t* st = St + s * 2 * kTile64;
        mbar_wait(&kv_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[s], 2 * kTile64);
          tma_load_3d(st, &tmQKV64, &kv_full[s], inner + h * DH, i * 64, b);
          tma_load_3d(st + kTile64, &tmQKV64, &kv_full[s], 2 * inner + h * DH, i * 64, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- issuer A: S = Q K^T, dP = dO V^T
    constexpr uint32_t idesc_s = make_idesc_f16(128, 64, 0, 0);
    const uint64_t qd0 = kmajor_desc(QDs);
    const uint64_t kk0 = kmajor_desc(St);
    uint32_t sd_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int ib = item_it & 1;
      mbar_wait(&q_full[ib], (item_it >> 1) & 1);
      const uint64_t qd = desc_advance(qd0, ib * kItem);
      const uint64_t dd = desc_advance(qd, kTile128);
      for (int i = 0; i < NS; ++i, ++sd_it) {
        const int s = sd_it % kTB;
        const uint32_t ph = (sd_it / kTB) & 1;
        const int r = sd_it % NST;
        mbar_wait2(&kv_full[r], (sd_it / NST) & 1, &sfree[s], ph ^ 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t kkd = desc_advance(kk0, r * 2 * kTile64);
          const uint64_t vkd = desc_advance(kkd, kTile64);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_f16<1>(tmem_base + s * 64, desc_advance(qd, k * kKStepK), desc_advance(kkd, k * kKStepK), idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_f16<1>(tmem_base + kColDP + s * 64, desc_advance(dd, k * kKStepK), desc_advance(vkd, k * kKStepK), idesc_s, k != 0);
          umma_commit<1>(&s_full[s]);
          umma_commit<1>(&kv_empty[r]);
          if (i == NS - 1) umma_commit<1>(&q_empty[ib]);
        }
        __syncwarp();
      }
    }
  } else if (warp == kIssuerB) {
    // ---- issuer B: dQ += dS K (A packed fp16 from TMEM, B = the K sub-tile read MN-major)
    constexpr uint32_t idesc_g = make_idesc_f16(128, DH, 0, 1);
    const uint64_t km0 = mnmajor_desc(St);
    uint32_t dq_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      for (int i = 0; i < NS; ++i, ++dq_it) {
        const int s = dq_it % kTB;
        const uint32_t ph = (dq_it / kTB) & 1;
        const int r = dq_it % NST;
        if (i == 0) mbar_wait(acc_empty, (item_it & 1) ^ 1);
        mbar_wait2(&kv_full[r], (dq_it / NST) & 1, &p_full[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t kmd = desc_advance(km0, r * 2 * kTile64);
          const uint32_t acc_on = i > 0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(tmem_base + kColAcc, tmem_base + kColDP + s * 64 + packed_a_col(k), desc_advance(kmd, k * kKStepMN), idesc_g, acc_on | (k != 0));
          umma_commit<1>(&kv_empty[r]);
          umma_commit<1>(&sfree[s]);
          if (i == NS - 1) umma_commit<1>(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // 16 softmax warps: thread = query row, the four warps of a lane quarter split the 64 key columns of a sub-tile
    // and, at the end of an item, the 64 columns of dQ
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    constexpr int OC = DH / 4;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c = p.scale * kLog2eF;
    uint32_t t_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int qt = w % p.tiles128;
      const int bh = w / p.tiles128;
      const int h = bh % p.heads, b = bh / p.heads;
      const int row = qt * 128 + q * 32 + lane;
      const long long sidx = ((long long)b * p.heads + h) * p.N + row;
      const float lse2 = row < p.N ? p.lse[sidx] * kLog2eF : INFINITY;
      // delta = rowsum(dO * O) of this thread's query row, from the item's dO and O tiles in shared memory (128-byte rows,
      // 16-byte chunks XOR-swizzled by row & 7): each of the four threads of a row takes 16 of the 64 columns and the
      // partial sums meet in shared memory under the lane quarter's named barrier.  Rows >= N are zero-filled by TMA.
      float dl;
      {
        const int ib = item_it & 1;
        mbar_wait(&q_full[ib], (item_it >> 1) & 1);
        const int r = q * 32 + lane;
        const uint8_t* dOt = QDs + ib * kItem + kTile128 + r * 128;
        const uint8_t* Ot = dOt + kTile128;
        float acc = 0.f;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int off = ((sub * 2 + ch) ^ (r & 7)) << 4;
          const uint4 a = *reinterpret_cast<const uint4*>(dOt + off);
          const uint4 o = *reinterpret_cast<const uint4*>(Ot + off);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[k]));
            const float2 fo = __half22float2(*reinterpret_cast<const __half2*>(&ow[k]));
            acc = fmaf(fa.x, fo.x, acc);
            acc = fmaf(fa.y, fo.y, acc);
          }
        }
        float* xs = xch + ib * 512;
        xs[sub * 128 + r] = acc;
        quad_bar(q);
        dl = (xs[r] + xs[128 + r]) + (xs[256 + r] + xs[384 + r]);
        if (sub == 0 && row < p.N) p.delta_w[sidx] = dl;
      }
      // two groups of eight warps on alternate key sub-tiles, 32 key columns per warp in two 16-column chunks (see the dKV kernel)
      const int grp = sub & 1, half = sub >> 1;
      const uint32_t t_base = item_it * (uint32_t)NS;
      auto chunk = [&](const uint32_t (&v)[16], const uint32_t (&g)[16], int col, int kv_left) {
        uint32_t ds[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float p0, p1;
          ex2_pair(poly_slot<ATTN_BWD_POLY_PAIRS, 8>(j), fmaf(__uint_as_float(v[2 * j]), c, -lse2), fmaf(__uint_as_float(v[2 * j + 1]), c, -lse2), p0, p1);
          ds[j] = pack_half2_sat(p0 * (__uint_as_float(g[2 * j]) - dl), p1 * (__uint_as_float(g[2 * j + 1]) - dl));
        }
        if (kv_left < 16) {                            // ragged last tile (warp-uniform): padded key columns contribute nothing
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (2 * j >= kv_left) ds[j] = 0u;
            else if (2 * j + 1 >= kv_left) ds[j] &= 0xffffu;
          }
        }
        tmem_st_32x8(tmem_base + lane_off + kColDP + col, ds);
      };
      for (int i = grp; i < NS; i += 2) {
        const uint32_t t = t_base + i;
        const int s = t % kTB;
        mbar_wait(&s_full[s], (t / kTB) & 1);
        tcgen05_fence_after();
        const int col = s * 64 + half * 32;
        const int kv_left = p.N - i * 64 - half * 32;
        uint32_t vA[16], gA[16], vB[16], gB[16];
        tmem_ld_32x16(tmem_base + lane_off + col, vA);
        tmem_ld_32x16(tmem_base + lane_off + kColDP + col, gA);
        tmem_ld_wait();
        tmem_ld_32x16(tmem_base + lane_off + col + 16, vB);
        tmem_ld_32x16(tmem_base + lane_off + kColDP + col + 16, gB);
        chunk(vA, gA, col, kv_left);
        tmem_ld_wait();
        chunk(vB, gB, col + 16, kv_left - 16);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
      }
      t_it += NS;
      mbar_wait(acc_full, item_it & 1);
      tcgen05_fence_after();
      {
        uint32_t v[OC];
        tmem_ld_32x16(tmem_base + lane_off + kColAcc + sub * OC, v);
        tmem_ld_wait();
        float r[OC];
#pragma unroll
        for (int j = 0; j < OC; ++j) r[j] = __uint_as_float(v[j]);
        store_box_h16(obox + (warp - 2) * 1024, &tmOut16, r, p.scale, h * DH + sub * OC, qt * 128 + q * 32, b, lane);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (lane == 0) bulk_wait_group_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}


// fp16 [B*N, ld] matrix viewed as {ld, N, B}; box {64 (one 128-byte row), rows, 1}, SWIZZLE_128B
int make_tile_map(CUtensorMap* out, const void* ptr, long long ld, int N, int B, int box_rows) {
  const unsigned long long dims[3] = {(unsigned long long)ld, (unsigned long long)N, (unsigned long long)B};
  const unsigned long long strides[2] = {(unsigned long long)ld * 2, (unsigned long long)N * ld * 2};
  const unsigned box[3] = {64, (unsigned)box_rows, 1};
  return make_tensor_map(out, ptr, 2, 3, dims, strides, box, 0);
}
// store view: 32-row x `cols`-column un-swizzled fp16 boxes
int make_store_map(CUtensorMap* out, const void* ptr, long long ld, int N, int B, int cols) {
  const unsigned long long dims[3] = {(unsigned long long)ld, (unsigned long long)N, (unsigned long long)B};
  const unsigned long long strides[2] = {(unsigned long long)ld * 2, (unsigned long long)N * ld * 2};
  const unsigned box[3] = {(unsigned)cols, 32, 1};
  return make_tensor_map(out, ptr, 2, 3, dims, strides, box, 3);
}

int persistent_grid(int items) {
  int grid = num_sms();
  if (sm_limit() > 0 && grid > sm_limit()) grid = sm_limit();
  return grid > items ? items : grid;
}

}  // namespace

// qkv fp16 [B*N, 3*heads*64]; out fp16 [B*N, heads*64]; lse fp32 [B*heads*N]
int attention_f16_forward(const void* qkv, void* out, float* lse, int B, int N, int heads, int dh, float scale, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == DH, "attention_f16: dim_head must be 64 (got %d); other head sizes use the tf32 core", dh);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "attention: qkv/out must be 16-byte aligned");
  const int inner = heads * DH;
  CUtensorMap tmQKV, tmO;
  int rc;
  if ((rc = make_tile_map(&tmQKV, qkv, 3ll * inner, N, B, 128))) return rc;
  if ((rc = make_store_map(&tmO, out, inner, N, B, 16))) return rc;
#ifndef ATTN_FWD_V2
#define ATTN_FWD_V2 1
#endif
  static const int use_v2 = [] { const char* e = getenv("B200VQ_ATTN_FWD_V2"); return e ? atoi(e) : ATTN_FWD_V2; }();
  if (use_v2) {
    if ((rc = make_store_map(&tmO, out, inner, N, B, 32))) return rc;
    Fwd2Params p2;
    p2.lse = lse; p2.N = N; p2.heads = heads;
    p2.q_pairs = (N + 255) / 256; p2.kv_tiles = (N + 127) / 128;
    p2.total_items = p2.q_pairs * heads * B;
    p2.scale = scale;
    constexpr int smem2 = (4 + 2 * kV2Stages) * kTile128 + kBoxBytes + 512 + 1536 * 4 + 1024;
    static_assert(smem2 <= 227 * 1024, "attention forward v2: shared-memory plan exceeds 227 KB");
    B200_CONFIGURE_SMEM_ONCE(attn_fwd2_f16_kernel, smem2);
    attn_fwd2_f16_kernel<<<persistent_grid(p2.total_items), kThreads2, smem2, stream>>>(tmQKV, tmO, p2);
    B200_LAUNCH_OK("attn_fwd2_f16_kernel");
    return 0;
  }
  FwdParams p;
  p.lse = lse; p.N = N; p.heads = heads;
  p.q_tiles = (N + 127) / 128; p.kv_tiles = (N + 127) / 128;
  p.total_items = p.q_tiles * heads * B;
  p.scale = scale;
  constexpr int smem = (2 + 2 * kFwdStages) * kTile128 + kBoxBytes + 512 + 1536 * 4 + 1024;
  B200_CONFIGURE_SMEM_ONCE(attn_fwd_f16_kernel, smem);
  attn_fwd_f16_kernel<<<persistent_grid(p.total_items), kThreads, smem, stream>>>(tmQKV, tmO, p);
  B200_LAUNCH_OK("attn_fwd_f16_kernel");
  return 0;
}

int attention_f16_backward(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, float* delta, int B, int N,
                           int heads, int dh, float scale, cudaStream_t stream) {
  B200_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attention: empty problem");
  B200_CHECK_ARG(dh == DH, "attention_f16: dim_head must be 64 (got %d)", dh);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "attention: qkv/out/dout/dqkv must be 16-byte aligned");
  int rc;
  const int inner = heads * DH;
  const long long ld = 3ll * inner;
  CUtensorMap tmQKV128, tmQKV64, tmDO128, tmDO64, tmO128, tmOut, tmOut16;
  if ((rc = make_tile_map(&tmQKV128, qkv, ld, N, B, 128))) return rc;
  if ((rc = make_tile_map(&tmQKV64, qkv, ld, N, B, 64))) return rc;
  if ((rc = make_tile_map(&tmDO128, dout, inner, N, B, 128))) return rc;
  if ((rc = make_tile_map(&tmDO64, dout, inner, N, B, 64))) return rc;
  if ((rc = make_tile_map(&tmO128, out, inner, N, B, 128))) return rc;
  if ((rc = make_store_map(&tmOut, dqkv, ld, N, B, 32))) return rc;
  if ((rc = make_store_map(&tmOut16, dqkv, ld, N, B, 16))) return rc;
  BwdParams p;
  p.lse = lse; p.delta = delta; p.delta_w = delta; p.N = N; p.heads = heads;
  p.tiles128 = (N + 127) / 128; p.sub64 = (N + 63) / 64;
  p.total_items = p.tiles128 * heads * B;
  p.scale = scale;
  constexpr int smem_kv = 4 * kTile128 + kDkvStages * 2 * kTile64 + kBoxBytes + 512 + kSoftmaxWarps * 256 + 1024;
  constexpr int smem_q = 6 * kTile128 + kDqStages * 2 * kTile64 + kBoxBytes / 2 + 512 + 1024 * 4 + 1024;
  static_assert(smem_kv <= 227 * 1024 && smem_q <= 227 * 1024, "attention backward: shared-memory plan exceeds 227 KB");
  B200_CONFIGURE_SMEM_ONCE(attn_bwd_dkv_f16_kernel, smem_kv);
  B200_CONFIGURE_SMEM_ONCE(attn_bwd_dq_f16_kernel, smem_q);
  const int grid = persistent_grid(p.total_items);
  // dQ first: it also produces delta = rowsum(dO * O) from the O tile it loads next to dO (no separate pass over dO and O).
  // (A two-query-tiles-per-CTA variant in the style of the forward v2 kernel -- one S / dP buffer per group, the groups
  // alternating -- was built and measured: 0.45 ms against this kernel's 0.37 ms at B=64.  With 64-key sub-tiles a group's
  // commit -> wait -> ld -> st -> arrive -> issue round trip is as long as its MMAs, and TMEM has no room for a second buffer
  // per group or for 128-key sub-tiles: 2 x (128 + 128 + 64) columns.)
  attn_bwd_dq_f16_kernel<<<grid, kThreads, smem_q, stream>>>(tmQKV128, tmDO128, tmO128, tmQKV64, tmOut16, p);
  B200_LAUNCH_OK("attn_bwd_dq_f16_kernel");
  attn_bwd_dkv_f16_kernel<<<grid, kThreads, smem_kv, stream>>>(tmQKV128, tmQKV64, tmDO64, tmOut, p);
  B200_LAUNCH_OK("attn_bwd_dkv_f16_kernel");
  return 0;
}

}  // namespace b200
