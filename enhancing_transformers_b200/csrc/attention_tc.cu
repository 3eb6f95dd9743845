// tcgen05 / TMEM flash-attention forward and backward for the ViT blocks (reference layers.py:124-130).
//
// Forward.  Persistent, warp-specialised (352 threads), one work item = (batch, head, 128-query tile):
//   warp 0      TMA producer: Q tile once per item, K and V tiles (128 keys) through 2-stage rings.
//               q/k are K-major operands (SWIZZLE_128B, one 128-byte block per 32 head dims);
//               V is the MN-major B operand of P.V (SWIZZLE_128B_BASE32B, 4-D tensor map).
//   warp 1      MMA issuer A: S_j = Q K_j^T (kind::tf32, 128x128 accumulator in TMEM, two S buffers).
//   warp 10     MMA issuer B: O_j = P_j V_j with the A operand read from TMEM (P_j overwrites S_j in place).
//               Two issuing warps because tcgen05.mma issue blocks while the tensor pipe executes (measured:
//               the queue is one or two instructions deep), so a single issuer would serialise its own mbarrier
//               polling with MMA execution; cross-issuer hazards are covered by mbarriers (sfree).
//   warps 2..9  softmax: a query row (TMEM lane) is shared by two threads (warps w and w+4 address the same lane
//               quarter), each owning 64 of the tile's 128 scores and half of the output columns.  Scores stay in
//               registers between the max and the exp2 pass; the row max is exchanged through shared memory under a
//               64-thread named barrier; P is written back tf32-rounded with tcgen05.st; the running output lives
//               in registers: o = o * alpha + (P_j V_j read back from TMEM), folded in while the next tile's scores
//               are already being loaded.  Outputs leave through swizzled smem boxes + TMA stores.
// TMEM columns: S/P buffers [0,128) [128,256), PV buffers [256,256+DH) [320,320+DH).
#include "common.cuh"
#include <cstdlib>

namespace b200 {

// Development trace (compiled only with -DB200_ATTN_TRACE into a separate .so, `make trace`): block 0 of
// the forward and dKV kernels logs (event id, clock64) pairs of three warps into a small shared-memory
// ring (global-memory logging costs ~500 cycles per event and hides what is being measured).
#ifdef B200_ATTN_TRACE
#ifndef B200_TRACE_SKIP
#define B200_TRACE_SKIP 300          // events to let pass before recording (steady state, past the first items)
#endif
constexpr int kTraceCap = 36;        // events kept per role: the dKV kernel has < 2 KB of shared memory to spare
__device__ long long g_trace[3 * 2 * kTraceCap];
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
    if (blockIdx.x == 0) for (int i_ = threadIdx.x; i_ < 3 * 2 * kTraceCap; i_ += blockDim.x) g_trace[i_] = (&tr_buf[0][0])[i_]; \
  } while (0)
#else
#define TRACE_DECL
#define TRACE(role, ev) do {} while (0)
#define TRACE_DUMP() do {} while (0)
#endif

constexpr int kAtcThreads = 352;      // warp 0 TMA, warp 1 MMA issuer A, warps 2..9 softmax (two per TMEM lane quarter), warp 10 MMA issuer B
constexpr int kIssuerB = 10;

template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[NC]) {
  if constexpr (NC == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
}
// named barrier of the two softmax warps that share TMEM lane quarter q (ids 2..5, 64 threads)
__device__ __forceinline__ void pair_bar(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 2) : "memory"); }
constexpr float kLog2eF = 1.4426950408889634f;

constexpr int kOutBoxBytes = 8 * 4096;   // one 32-row x 128-byte store box per softmax warp

// One warp stores 32 rows x 32 fp32 columns (lane = row, values r[0..31]) with a single TMA store:
// registers -> 128B-swizzled 4 KB smem box -> cp.async.bulk.tensor, so HBM sees whole 128-byte lines and
// rows / columns past the tensor edge are clipped.  (A thread writing its own row straight to global
// memory makes every store instruction touch 32 different lines; at the end of a work item that
// serialises in the L1 store path for thousands of cycles and back-pressures the mbarrier traffic of
// the TMA / MMA warps.)
__device__ __forceinline__ void warp_store_box(uint8_t* box, const CUtensorMap* tm, const float (&r)[32], int c0, int c1, int c2,
                                               int lane) {
  if (lane == 0) bulk_wait_group_read<0>();   // the box's previous store has been read out
  __syncwarp();
  uint8_t* rowp = box + lane * 128;
  const uint32_t swz = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(rowp + ((j ^ swz) << 4)) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(tm, box, c0, c1, c2);
    bulk_commit_group();
  }
}

// fp16 flavour: the same 32 x 32 values, multiplied by `mul`, leave as a 32-row x 64-byte box (saturating
// fp16; un-swizzled -- this runs once per work item, the bank conflicts of the 16-byte row stores are noise).
__device__ __forceinline__ void warp_store_box_h(uint8_t* box, const CUtensorMap* tm, const float (&r)[32], float mul, int c0, int c1,
                                                 int c2, int lane) {
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

struct AttnTcParams {
  float* out;        // [B*N, heads*DH] fp32, or fp16 when out_half
  int out_half;      // 1: the output feeds an fp16 GEMM (to_out): store fp16
  float* lse;        // [B*heads*N]
  int B, N, heads;
  int q_tiles, kv_tiles, total_items;
  float scale;
  int round_out;
  int cond;          // < 0: no mask; >= 0: stage-2 mask, query q sees key k iff k <= max(q, cond - 1) (stage2/layers.py:43-48)
};

template <int DH>
__global__ void __launch_bounds__(kAtcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmO, const AttnTcParams p) {
  constexpr int KB = DH / 32;                 // 128-byte k-blocks per row
  constexpr int TILE_BYTES = 128 * DH * 4;    // one Q / K / V tile
  constexpr int KBLK_BYTES = 128 * 128;       // one k-block (or one MN atom of V): 128 rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + TILE_BYTES;            // [2]
  uint8_t* Vs = smem + 3 * TILE_BYTES;        // [2]
  uint8_t* obox = smem + 5 * TILE_BYTES;      // [8 softmax warps][4 KB] output store boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(obox + kOutBoxBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;    // [2]
  uint64_t* k_empty = bars + 4;   // [2]
  uint64_t* v_full = bars + 6;    // [2]
  uint64_t* v_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;   // [2]
  uint64_t* p_full = bars + 12;   // [2]
  uint64_t* o_full = bars + 14;   // [2]
  uint64_t* o_empty = bars + 16;  // [2]
  uint64_t* sfree = bars + 18;    // [2] S/P buffer consumed by the P.V MMAs (issuer B -> issuer A)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* xch = reinterpret_cast<float*>(bars + 22);   // [3][2][128] row-max (double buffered) and row-sum exchange
  TRACE_DECL;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 8);
      mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], 8);
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
  // key tiles an item visits: all of them, or -- under the stage-2 mask -- those up to the last key any of its rows sees
  auto tiles_of = [&](int w) -> int {
    if (p.cond < 0) return p.kv_tiles;
    const int last = max(min(p.N - 1, (w % p.q_tiles) * 128 + 127), p.cond - 1);
    return min(p.kv_tiles, last / 128 + 1);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform, one lane issues)
    uint32_t kv_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int T = tiles_of(w);
      const int qt = w % p.q_tiles;
      const int bh = w / p.q_tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      mbar_wait(q_empty, (item_it & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, TILE_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(Qs + kb * KBLK_BYTES, &tmQK, q_full, h * DH + kb * 32, qt * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j, ++kv_it) {
        const int s = kv_it & 1;
        const uint32_t ph = (kv_it >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(Ks + s * TILE_BYTES + kb * KBLK_BYTES, &tmQK, &k_full[s], inner + h * DH + kb * 32, j * 128, b);
        }
        __syncwarp();
        mbar_wait(&v_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
          tma_load_4d(Vs + s * TILE_BYTES, &tmV, &v_full[s], 0, j * 128, (2 * inner + h * DH) / 32, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer A: S_j = Q K_j^T
    constexpr uint32_t idesc_s = make_idesc_tf32(128, 128, 0, 0);   // both operands K-major
    const uint64_t qd = make_smem_desc(smem_u32(Qs), 16, 1024, kLayoutSw128);
    const uint64_t kd0 = make_smem_desc(smem_u32(Ks), 16, 1024, kLayoutSw128);
    uint32_t s_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int T = tiles_of(w);
      mbar_wait(q_full, item_it & 1);
      for (int j = 0; j < T; ++j, ++s_it) {
        const int s = s_it & 1;
        const uint32_t ph = (s_it >> 1) & 1;
        mbar_wait(&k_full[s], ph);
        mbar_wait(&sfree[s], ph ^ 1);       // P_{j-2} (same buffer) has been consumed by issuer B
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t kd = desc_advance(kd0, s * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < DH / 8; ++k) {
            const uint32_t off = (k >> 2) * KBLK_BYTES + (k & 3) * 32;
            umma_tf32<1>(tmem_base + s * 128, desc_advance(qd, off), desc_advance(kd, off), idesc_s, k != 0);
          }
          umma_commit<1>(&s_full[s]);
          umma_commit<1>(&k_empty[s]);
          if (j == T - 1) umma_commit<1>(q_empty);
        }
        __syncwarp();
      }
    }
  } else if (warp == kIssuerB) {
    // ------------------------------------------------------------------ MMA issuer B: O_j = P_j V_j (A operand from TMEM)
    constexpr uint32_t idesc_o = make_idesc_tf32(128, DH, 0, 1);    // B MN-major
    const uint64_t vd0 = make_smem_desc(smem_u32(Vs), KBLK_BYTES, 512, kLayoutSw128Base32);
    uint32_t pv_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      const int T = tiles_of(w);
      for (int j = 0; j < T; ++j, ++pv_it) {
        const int s = pv_it & 1;
        const uint32_t ph = (pv_it >> 1) & 1;
        TRACE(2, 103);
        mbar_wait(&v_full[s], ph);
        mbar_wait(&o_empty[s], ph ^ 1);
        TRACE(2, 105);
        mbar_wait(&p_full[s], ph);
        TRACE(2, 104);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t vd = desc_advance(vd0, s * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_tf32_ts(tmem_base + 256 + s * 64, tmem_base + s * 128 + k * 8, desc_advance(vd, k * 1024), idesc_o, k != 0);
          umma_commit<1>(&o_full[s]);
          umma_commit<1>(&v_empty[s]);
          umma_commit<1>(&sfree[s]);
        }
        __syncwarp();
        TRACE(2, 106);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / output warps
    // A query row is shared by two threads (warps w and w+4 address the same TMEM lanes): each
    // takes 64 of the 128 scores of a tile and half of the head dim of the output; the row max
    // (per tile) and the row sum (once per item) are exchanged through shared memory.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int OC = DH / 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const float c = p.scale * kLog2eF;
    uint32_t t_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      const int T = tiles_of(w);
      const int qt = w % p.q_tiles;
      const int bh = w / p.q_tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      float o[OC];
#pragma unroll
      for (int i = 0; i < OC; ++i) o[i] = 0.f;
      float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
      // fold the P.V partial product of tile `it` into the running output (its barrier has been waited on)
      auto fold_pv = [&](uint32_t it, float alpha) {
        const int sp = it & 1;
        uint32_t v[OC];
        tmem_ld_cols<OC>(tmem_base + lane_off + 256 + sp * 64 + half * OC, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < OC; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(v[i]));
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[sp]);
      };
      // the thread's 64 scores of a tile stay in registers between the max and the exp pass; those of tile
      // j+1 are requested as soon as P_j has been handed over, so the TMEM load overlaps the P.V fold below
      uint32_t v0[32], v1[32];
      if (warp == 2) TRACE(0, 200); else if (warp == 6) TRACE(1, 200);
      mbar_wait(&s_full[t_it & 1], (t_it >> 1) & 1);
      tcgen05_fence_after();
      {
        const uint32_t s0 = tmem_base + lane_off + (t_it & 1) * 128 + half * 64;
        tmem_ld_32x32(s0, v0);
        tmem_ld_32x32(s0 + 32, v1);
      }
      tmem_ld_wait();
      for (int j = 0; j < T; ++j, ++t_it) {
        const int s = t_it & 1;
        const uint32_t sa = tmem_base + lane_off + s * 128 + half * 64;
        int kv_left = p.N - j * 128 - half * 64;         // this thread's columns >= kv_left are padding
        // stage-2 mask: the row's last visible key is max(row, cond - 1).  Key 0 is visible to every row, so the running
        // maximum is finite after the first tile and a fully masked tile contributes exp2(-inf) = 0 with alpha = 1.
        if (p.cond >= 0) kv_left = min(kv_left, max(qt * 128 + row_in_tile, p.cond - 1) + 1 - j * 128 - half * 64);
        if (warp == 2) TRACE(0, 230); else if (warp == 6) TRACE(1, 230);
        if (kv_left < 64) {                                 // only the last tile of a ragged sequence has padded keys
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= kv_left) v0[i] = 0xff800000u;          // -inf
            if (32 + i >= kv_left) v1[i] = 0xff800000u;
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
        if (warp == 2) TRACE(0, 240); else if (warp == 6) TRACE(1, 240);
        float* xs = xch + (t_it & 1) * 256;
        xs[half * 128 + row_in_tile] = mx;
        pair_bar(q);                                        // only the two warps that share these rows
        mx = fmaxf(mx, xs[(half ^ 1) * 128 + row_in_tile]);
        const float m_new = fmaxf(m, mx);
        if (warp == 2) TRACE(0, 250); else if (warp == 6) TRACE(1, 250);
        const float alpha = ex2_approx((m - m_new) * c);
        const float mc = m_new * c;
        float sum = 0.f, sum1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(v0[i]), c, -mc));
          const float e1 = ex2_approx(fmaf(__uint_as_float(v1[i]), c, -mc));
          sum += e0; sum1 += e1;
          v0[i] = tf32_bits_for_mma(e0);
          v1[i] = tf32_bits_for_mma(e1);
        }
        sum += sum1;
        if (warp == 2) TRACE(0, 260); else if (warp == 6) TRACE(1, 260);
        tmem_st_32x32(sa, v0);
        tmem_st_32x32(sa + 32, v1);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (warp == 2) TRACE(0, 280); else if (warp == 6) TRACE(1, 280);
        l = fmaf(l, alpha, sum);
        const bool more = j + 1 < T;
        const uint32_t ph_s = ((t_it + 1) >> 1) & 1, ph_o = ((t_it - 1) >> 1) & 1;   // both live in buffer s ^ 1
        if (more && j >= 1) mbar_wait2(&s_full[s ^ 1], ph_s, &o_full[s ^ 1], ph_o);
        else if (more)      mbar_wait(&s_full[s ^ 1], ph_s);
        else if (j >= 1)    mbar_wait(&o_full[s ^ 1], ph_o);
        tcgen05_fence_after();
        if (warp == 2) TRACE(0, 220); else if (warp == 6) TRACE(1, 220);
        if (more) {
          const uint32_t sn = tmem_base + lane_off + (s ^ 1) * 128 + half * 64;
          tmem_ld_32x32(sn, v0);
          tmem_ld_32x32(sn + 32, v1);
        }
        if (j >= 1) fold_pv(t_it - 1, alpha_prev);
        else        tmem_ld_wait();
        alpha_prev = alpha;
        m = m_new;
        if (warp == 2) TRACE(0, 290); else if (warp == 6) TRACE(1, 290);
      }
      mbar_wait(&o_full[(t_it - 1) & 1], ((t_it - 1) >> 1) & 1);
      tcgen05_fence_after();
      fold_pv(t_it - 1, alpha_prev);
      float* ls = xch + 512;
      ls[half * 128 + row_in_tile] = l;
      pair_bar(q);
      l += ls[(half ^ 1) * 128 + row_in_tile];
      const int row = qt * 128 + row_in_tile;
      const float inv = 1.f / l;
      if constexpr (OC == 32) {
        if (p.out_half) {
          warp_store_box_h(obox + (warp - 2) * 4096, &tmO, o, inv, h * DH + half * OC, qt * 128 + q * 32, b, lane);
        } else {
#pragma unroll
          for (int i = 0; i < OC; ++i) o[i] = p.round_out ? round_tf32(o[i] * inv) : o[i] * inv;
          warp_store_box(obox + (warp - 2) * 4096, &tmO, o, h * DH + half * OC, qt * 128 + q * 32, b, lane);
        }
      } else if (row < p.N && p.out_half) {
        __half* op = reinterpret_cast<__half*>(p.out) + ((long long)b * p.N + row) * inner + h * DH + half * OC;
#pragma unroll
        for (int i = 0; i < OC; i += 8) {
          uint4 pk;
          pk.x = pack_half2_sat(o[i] * inv, o[i + 1] * inv); pk.y = pack_half2_sat(o[i + 2] * inv, o[i + 3] * inv);
          pk.z = pack_half2_sat(o[i + 4] * inv, o[i + 5] * inv); pk.w = pack_half2_sat(o[i + 6] * inv, o[i + 7] * inv);
          *reinterpret_cast<uint4*>(op + i) = pk;
        }
      } else if (row < p.N) {
        float* op = p.out + ((long long)b * p.N + row) * inner + h * DH + half * OC;
#pragma unroll
        for (int i = 0; i < OC; i += 4) {
          float4 r = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
          if (p.round_out) { r.x = round_tf32(r.x); r.y = round_tf32(r.y); r.z = round_tf32(r.z); r.w = round_tf32(r.w); }
          *reinterpret_cast<float4*>(op + i) = r;
        }
      }
      if (row < p.N && half == 0) p.lse[((long long)b * p.heads + h) * p.N + row] = m * p.scale + logf(l);
      pair_bar(q);     // ls is rewritten by the next item only after the partner has read it
    }
    if (lane == 0) bulk_wait_group_read<0>();   // the store boxes must outlive the last TMA reads
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
  TRACE_DUMP();
}


template <int DH>
static int attn_fwd_tc_launch(const float* qkv, void* out, int out_half, float* lse, int B, int N, int heads, float scale,
                              int round_out, int cond, cudaStream_t stream) {
  const int inner = heads * DH;
  const long long ld = 3ll * inner;
  CUtensorMap tmQK, tmV;
  {
    const unsigned long long dims[3] = {(unsigned long long)ld, (unsigned long long)N, (unsigned long long)B};
    const unsigned long long strides[2] = {(unsigned long long)ld * 4, (unsigned long long)N * ld * 4};
    const unsigned box[3] = {32, 128, 1};
    int rc = make_tensor_map_f32(&tmQK, qkv, 3, dims, strides, box, 0);
    if (rc) return rc;
  }
  {
    const unsigned long long dims[4] = {32, (unsigned long long)N, (unsigned long long)(ld / 32), (unsigned long long)B};
    const unsigned long long strides[3] = {(unsigned long long)ld * 4, 128, (unsigned long long)N * ld * 4};
    const unsigned box[4] = {32, 128, (unsigned)(DH / 32), 1};
    int rc = make_tensor_map_f32(&tmV, qkv, 4, dims, strides, box, 1);
    if (rc) return rc;
  }
  AttnTcParams p;
  p.out = static_cast<float*>(out); p.out_half = out_half; p.lse = lse; p.B = B; p.N = N; p.heads = heads;
  p.q_tiles = (N + 127) / 128; p.kv_tiles = (N + 127) / 128;
  p.total_items = p.q_tiles * heads * B;
  p.scale = scale; p.round_out = round_out; p.cond = cond;
  CUtensorMap tmO;
  {
    const unsigned long long dims[3] = {(unsigned long long)inner, (unsigned long long)N, (unsigned long long)B};
    const unsigned long long esz = out_half ? 2 : 4;
    const unsigned long long strides[2] = {(unsigned long long)inner * esz, (unsigned long long)N * inner * esz};
    const unsigned box[3] = {32, 32, 1};
    int rc = make_tensor_map(&tmO, out, (int)esz, 3, dims, strides, box, out_half ? 3 : 0);
    if (rc) return rc;
  }
  int grid = num_sms();
  if (sm_limit() > 0 && grid > sm_limit()) grid = sm_limit();
  if (grid > p.total_items) grid = p.total_items;
  constexpr int smem = 5 * 128 * DH * 4 + kOutBoxBytes + 512 + 3 * 256 * 4 + 1024;
  auto kern = attn_fwd_tc_kernel<DH>;
  B200_CONFIGURE_SMEM_ONCE(kern, smem);
  kern<<<grid, kAtcThreads, smem, stream>>>(tmQK, tmV, tmO, p);
  B200_LAUNCH_OK("attn_fwd_tc_kernel");
  return 0;
}

// =============================================================================================
// backward on tcgen05 (two kernels, no atomics; scores recomputed from the saved log-sum-exp)
//
//   dKV kernel: work item = (batch, head, 128-key tile); loops over 64-query sub-tiles:
//       S^T = K Q^T, dP^T = V dO^T            (SS MMAs, 128 x 64 accumulators, two buffers each)
//       P^T = exp2(S^T c - lse[q]),  dS^T = P^T (dP^T - delta[q])     (softmax warps, thread = key row,
//                                                    written back over S^T / dP^T as tf32)
//       dV += P^T dO,  dK += dS^T Q           (TS MMAs: A from TMEM, B = dO / Q as MN-major tiles)
//   dQ kernel: work item = (batch, head, 128-query tile); loops over 64-key sub-tiles:
//       S = Q K^T, dP = dO V^T ; dS = exp2(S c - lse) (dP - delta) ; dQ += dS K
// Q / dO (resp. K) are needed both as K-major operands (SWIZZLE_128B) and as MN-major operands
// (SWIZZLE_128B_BASE32B): TMA fetches the same global tile twice with two tensor maps.
// =============================================================================================

struct AttnBwdParams {
  const float* lse;     // [B*heads*N]
  const float* delta;   // [B*heads*N]
  float* dqkv;          // [B*N, 3*heads*DH] fp32, or fp16 when out_half
  int out_half;         // 1: dqkv feeds fp16 GEMMs: store fp16, multiplied by *out_scale (the gradient scale)
  const float* out_scale;
  int B, N, heads;
  int tiles128, sub64, total_items;
  float scale;
  int round_out;
  int cond;             // stage-2 mask, as in AttnTcParams
};

template <int DH>
__global__ void __launch_bounds__(kAtcThreads, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmQ64,
                       const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmQM,
                       const __grid_constant__ CUtensorMap tmDOM, const __grid_constant__ CUtensorMap tmOut,
                       const AttnBwdParams p) {
  constexpr int KB = DH / 32;
  constexpr int T128 = 128 * DH * 4;   // K or V tile
  constexpr int T64 = 64 * DH * 4;     // one 64-row operand copy
  constexpr int KBLK128 = 128 * 128, KBLK64 = 64 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* Ks = smem;
  uint8_t* Vs = smem + T128;
  uint8_t* St = smem + 2 * T128;        // stage s at St + s*4*T64: [QK | DK | QM | DM]
  uint8_t* obox = smem + 2 * T128 + 8 * T64;   // [8 softmax warps][4 KB] output store boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(obox + kOutBoxBytes);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  TRACE_DECL;
  uint64_t* qk_full = bars + 2;    // [2]
  uint64_t* qk_empty = bars + 4;   // [2]
  uint64_t* qm_full = bars + 6;    // [2]
  uint64_t* qm_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* p_full = bars + 12;    // [2]
  uint64_t* acc_full = bars + 14;
  uint64_t* acc_empty = bars + 15;
  uint64_t* sfree = bars + 16;     // [2] P^T / dS^T buffers consumed by issuer B
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(&sfree[0], 1); mbar_init(&sfree[1], 1);
    tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmQ64); tma_prefetch_desc(&tmDO64);
    tma_prefetch_desc(&tmQM); tma_prefetch_desc(&tmDOM);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qk_full[s], 1); mbar_init(&qk_empty[s], 1);
      mbar_init(&qm_full[s], 1); mbar_init(&qm_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 8);
    }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
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
      mbar_wait(kv_empty, (item_it & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * T128);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(Ks + kb * KBLK128, &tmKV, kv_full, inner + h * DH + kb * 32, kt * 128, b);
          tma_load_3d(Vs + kb * KBLK128, &tmKV, kv_full, 2 * inner + h * DH + kb * 32, kt * 128, b);
        }
      }
      __syncwarp();
      for (int i = 0; i < NS; ++i, ++sub_it) {
        const int s = sub_it & 1;
        const uint32_t ph = (sub_it >> 1) & 1;
        uint8_t* st = St + s * 4 * T64;
        mbar_wait(&qk_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&qk_full[s], 2 * T64);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            tma_load_3d(st + kb * KBLK64, &tmQ64, &qk_full[s], h * DH + kb * 32, i * 64, b);
            tma_load_3d(st + T64 + kb * KBLK64, &tmDO64, &qk_full[s], h * DH + kb * 32, i * 64, b);
          }
        }
        __syncwarp();
        mbar_wait(&qm_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&qm_full[s], 2 * T64);
          tma_load_4d(st + 2 * T64, &tmQM, &qm_full[s], 0, i * 64, (h * DH) / 32, b);
          tma_load_4d(st + 3 * T64, &tmDOM, &qm_full[s], 0, i * 64, (h * DH) / 32, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- issuer A: S^T = K Q^T, dP^T = V dO^T
    constexpr uint32_t idesc_s = make_idesc_tf32(128, 64, 0, 0);
    const uint64_t kd = make_smem_desc(smem_u32(Ks), 16, 1024, kLayoutSw128);
    const uint64_t vd = make_smem_desc(smem_u32(Vs), 16, 1024, kLayoutSw128);
    const uint64_t qd0 = make_smem_desc(smem_u32(St), 16, 1024, kLayoutSw128);
    uint32_t sd_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      mbar_wait(kv_full, item_it & 1);
      for (int i = 0; i < NS; ++i, ++sd_it) {
        const int s = sd_it & 1;
        const uint32_t ph = (sd_it >> 1) & 1;
        TRACE(0, 100);
        mbar_wait(&qk_full[s], ph);
        TRACE(0, 107);
        mbar_wait(&sfree[s], ph ^ 1);
        TRACE(0, 101);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t qd = desc_advance(qd0, s * 4 * T64);
          const uint64_t dd = desc_advance(qd, T64);
#pragma unroll
          for (int k = 0; k < DH / 8; ++k) {
            const uint32_t offa = (k >> 2) * KBLK128 + (k & 3) * 32, offb = (k >> 2) * KBLK64 + (k & 3) * 32;
            umma_tf32<1>(tmem_base + s * 64, desc_advance(kd, offa), desc_advance(qd, offb), idesc_s, k != 0);
          }
#pragma unroll
          for (int k = 0; k < DH / 8; ++k) {
            const uint32_t offa = (k >> 2) * KBLK128 + (k & 3) * 32, offb = (k >> 2) * KBLK64 + (k & 3) * 32;
            umma_tf32<1>(tmem_base + 128 + s * 64, desc_advance(vd, offa), desc_advance(dd, offb), idesc_s, k != 0);
          }
          umma_commit<1>(&s_full[s]);
          umma_commit<1>(&qk_empty[s]);
          if (i == NS - 1) umma_commit<1>(kv_empty);
        }
        __syncwarp();
        TRACE(0, 102);
      }
    }
  } else if (warp == kIssuerB) {
    // ---- issuer B: dV += P^T dO, dK += dS^T Q (A operands from TMEM)
    constexpr uint32_t idesc_g = make_idesc_tf32(128, DH, 0, 1);
    const uint64_t qmd0 = make_smem_desc(smem_u32(St + 2 * T64), KBLK64, 512, kLayoutSw128Base32);
    uint32_t dv_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      for (int i = 0; i < NS; ++i, ++dv_it) {
        const int s = dv_it & 1;
        const uint32_t ph = (dv_it >> 1) & 1;
        TRACE(1, 103);
        mbar_wait(&qm_full[s], ph);
        if (i == 0) mbar_wait(acc_empty, (item_it & 1) ^ 1);
        TRACE(1, 105);
        mbar_wait(&p_full[s], ph);
        TRACE(1, 104);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t qmd = desc_advance(qmd0, s * 4 * T64);
          const uint64_t dmd = desc_advance(qmd, T64);
          const uint32_t acc_on = i > 0;
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dV += P^T dO
            umma_tf32_ts(tmem_base + 256, tmem_base + s * 64 + k * 8, desc_advance(dmd, k * 1024), idesc_g, acc_on | (k != 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dK += dS^T Q
            umma_tf32_ts(tmem_base + 320, tmem_base + 128 + s * 64 + k * 8, desc_advance(qmd, k * 1024), idesc_g, acc_on | (k != 0));
          umma_commit<1>(&qm_empty[s]);
          umma_commit<1>(&sfree[s]);
          if (i == NS - 1) umma_commit<1>(acc_full);
        }
        __syncwarp();
        TRACE(1, 106);
      }
    }
  } else {
    // 8 softmax warps: warps w and w+4 share a TMEM lane quarter (key rows) and split the 64
    // query columns of a sub-tile; at the end of an item one half stores dV, the other dK.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c = p.scale * kLog2eF;
    uint32_t t_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int kt = w % p.tiles128;
      const int bh = w / p.tiles128;
      const int h = bh % p.heads, b = bh / p.heads;
      const float* lb = p.lse + ((long long)b * p.heads + h) * p.N;
      const float* eb = p.delta + ((long long)b * p.heads + h) * p.N;
      // Per-column lse / delta: lane l of a warp keeps the values of query column (half*32 + l) of the
      // current sub-tile in registers (fetched from global one sub-tile ahead; the raw values are only
      // *used* at the top of the next iteration so the load latency is off the critical path) and the
      // 32 columns are broadcast with warp shuffles.  Shared-memory / L1 reads are not an option here:
      // while the N=64 tf32 MMAs run they take the SM's whole SRAM bandwidth and such reads cost
      // ~1000 extra cycles per sub-tile (measured).
      const int qcol = half * 32 + lane;
      float raw_l = 0.f, raw_e = 0.f;
      bool nvalid = qcol < p.N;
      if (nvalid) { raw_l = lb[qcol]; raw_e = eb[qcol]; }
      for (int i = 0; i < NS; ++i, ++t_it) {
        const int s = t_it & 1;
        const float myL = nvalid ? raw_l * kLog2eF : INFINITY;   // +inf -> P = 0 for padded queries
        const float myE = nvalid ? raw_e : 0.f;
        {
          const int qi = (i + 1) * 64 + qcol;
          nvalid = i + 1 < NS && qi < p.N;
          if (nvalid) { raw_l = lb[qi]; raw_e = eb[qi]; }
        }
        if (warp == 2) TRACE(2, 200);
        mbar_wait(&s_full[s], (t_it >> 1) & 1);
        if (warp == 2) TRACE(2, 220);
        tcgen05_fence_after();
        const int col = s * 64 + half * 32;
        uint32_t v[32], g[32];
        tmem_ld_32x32(tmem_base + lane_off + col, v);
        tmem_ld_32x32(tmem_base + lane_off + 128 + col, g);
        tmem_ld_wait();
        if (warp == 2) TRACE(2, 240);
        if (p.cond < 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float Lj = __shfl_sync(0xffffffffu, myL, j);
            const float Ej = __shfl_sync(0xffffffffu, myE, j);
            const float pr = ex2_approx(fmaf(__uint_as_float(v[j]), c, -Lj));
            const float ds = pr * (__uint_as_float(g[j]) - Ej);
            v[j] = tf32_bits_for_mma(pr);
            g[j] = tf32_bits_for_mma(ds);
          }
        } else {      // stage-2 mask: this thread's key row receives from query column qj iff key <= max(qj, cond - 1)
          const int key = kt * 128 + q * 32 + lane;
          const int q0c = i * 64 + half * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float Lj = __shfl_sync(0xffffffffu, myL, j);
            const float Ej = __shfl_sync(0xffffffffu, myE, j);
            const float pr = key <= max(q0c + j, p.cond - 1) ? ex2_approx(fmaf(__uint_as_float(v[j]), c, -Lj)) : 0.f;
            const float ds = pr * (__uint_as_float(g[j]) - Ej);
            v[j] = tf32_bits_for_mma(pr);
            g[j] = tf32_bits_for_mma(ds);
          }
        }
        if (warp == 2) TRACE(2, 260);
        tmem_st_32x32(tmem_base + lane_off + col, v);
        tmem_st_32x32(tmem_base + lane_off + 128 + col, g);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (warp == 2) TRACE(2, 280);
      }
      // item epilogue: this thread's key row of dV (half 0) or dK (half 1)
      mbar_wait(acc_full, item_it & 1);
      tcgen05_fence_after();
      const int col0 = (half == 0 ? 2 * inner : inner) + h * DH;
      const float mul = (half == 0 ? 1.f : p.scale) * (p.out_half && p.out_scale ? __ldg(p.out_scale) : 1.f);
#pragma unroll 1
      for (int cc = 0; cc < DH / 32; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_off + (half == 0 ? 256 : 320) + cc * 32, v);
        tmem_ld_wait();
        float r[32];
        if (p.out_half) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
          warp_store_box_h(obox + (warp - 2) * 4096, &tmOut, r, mul, col0 + cc * 32, kt * 128 + q * 32, b, lane);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = p.round_out ? round_tf32(__uint_as_float(v[j]) * mul) : __uint_as_float(v[j]) * mul;
          warp_store_box(obox + (warp - 2) * 4096, &tmOut, r, col0 + cc * 32, kt * 128 + q * 32, b, lane);   // key rows >= N are clipped
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (lane == 0) bulk_wait_group_read<0>();   // the store boxes must outlive the last TMA reads
  }
  TRACE_DUMP();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int DH>
__global__ void __launch_bounds__(kAtcThreads, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ128, const __grid_constant__ CUtensorMap tmDO128,
                      const __grid_constant__ CUtensorMap tmKV64, const __grid_constant__ CUtensorMap tmKM,
                      const __grid_constant__ CUtensorMap tmOut, const AttnBwdParams p) {
  constexpr int KB = DH / 32;
  constexpr int T128 = 128 * DH * 4;
  constexpr int T64 = 64 * DH * 4;
  constexpr int KBLK128 = 128 * 128, KBLK64 = 64 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* Qs = smem;
  uint8_t* Ds = smem + T128;
  uint8_t* St = smem + 2 * T128;        // stage s at St + s*3*T64: [KK | VK | KM]
  uint8_t* obox = smem + 2 * T128 + 6 * T64;   // [8 softmax warps][4 KB] output store boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(obox + kOutBoxBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kk_full = bars + 2;    // [2]
  uint64_t* kk_empty = bars + 4;   // [2]
  uint64_t* km_full = bars + 6;    // [2]
  uint64_t* km_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* p_full = bars + 12;    // [2]
  uint64_t* acc_full = bars + 14;
  uint64_t* acc_empty = bars + 15;
  uint64_t* sfree = bars + 16;     // [2] dS buffer consumed by issuer B
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(&sfree[0], 1); mbar_init(&sfree[1], 1);
    tma_prefetch_desc(&tmQ128); tma_prefetch_desc(&tmDO128); tma_prefetch_desc(&tmKV64); tma_prefetch_desc(&tmKM);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kk_full[s], 1); mbar_init(&kk_empty[s], 1);
      mbar_init(&km_full[s], 1); mbar_init(&km_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 8);
    }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int inner = p.heads * DH;
  // 64-key sub-tiles an item visits (stage-2 mask: up to the last key any of its 128 rows sees)
  auto subs_of = [&](int w) -> int {
    if (p.cond < 0) return p.sub64;
    const int last = max(min(p.N - 1, (w % p.tiles128) * 128 + 127), p.cond - 1);
    return min(p.sub64, last / 64 + 1);
  };

  if (warp == 0) {
    uint32_t sub_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int NS = subs_of(w);
      const int qt = w % p.tiles128;
      const int bh = w / p.tiles128;
      const int h = bh % p.heads, b = bh / p.heads;
      mbar_wait(q_empty, (item_it & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * T128);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(Qs + kb * KBLK128, &tmQ128, q_full, h * DH + kb * 32, qt * 128, b);
          tma_load_3d(Ds + kb * KBLK128, &tmDO128, q_full, h * DH + kb * 32, qt * 128, b);
        }
      }
      __syncwarp();
      for (int i = 0; i < NS; ++i, ++sub_it) {
        const int s = sub_it & 1;
        const uint32_t ph = (sub_it >> 1) & 1;
        uint8_t* st = St + s * 3 * T64;
        mbar_wait(&kk_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kk_full[s], 2 * T64);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            tma_load_3d(st + kb * KBLK64, &tmKV64, &kk_full[s], inner + h * DH + kb * 32, i * 64, b);
            tma_load_3d(st + T64 + kb * KBLK64, &tmKV64, &kk_full[s], 2 * inner + h * DH + kb * 32, i * 64, b);
          }
        }
        __syncwarp();
        mbar_wait(&km_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&km_full[s], T64);
          tma_load_4d(st + 2 * T64, &tmKM, &km_full[s], 0, i * 64, (inner + h * DH) / 32, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- issuer A: S = Q K^T, dP = dO V^T
    constexpr uint32_t idesc_s = make_idesc_tf32(128, 64, 0, 0);
    const uint64_t qd = make_smem_desc(smem_u32(Qs), 16, 1024, kLayoutSw128);
    const uint64_t dd = make_smem_desc(smem_u32(Ds), 16, 1024, kLayoutSw128);
    const uint64_t kkd0 = make_smem_desc(smem_u32(St), 16, 1024, kLayoutSw128);
    uint32_t sd_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int NS = subs_of(w);
      mbar_wait(q_full, item_it & 1);
      for (int i = 0; i < NS; ++i, ++sd_it) {
        const int s = sd_it & 1;
        const uint32_t ph = (sd_it >> 1) & 1;
        mbar_wait(&kk_full[s], ph);
        mbar_wait(&sfree[s], ph ^ 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t kkd = desc_advance(kkd0, s * 3 * T64);
          const uint64_t vkd = desc_advance(kkd, T64);
#pragma unroll
          for (int k = 0; k < DH / 8; ++k) {
            const uint32_t offa = (k >> 2) * KBLK128 + (k & 3) * 32, offb = (k >> 2) * KBLK64 + (k & 3) * 32;
            umma_tf32<1>(tmem_base + s * 64, desc_advance(qd, offa), desc_advance(kkd, offb), idesc_s, k != 0);
          }
#pragma unroll
          for (int k = 0; k < DH / 8; ++k) {
            const uint32_t offa = (k >> 2) * KBLK128 + (k & 3) * 32, offb = (k >> 2) * KBLK64 + (k & 3) * 32;
            umma_tf32<1>(tmem_base + 128 + s * 64, desc_advance(dd, offa), desc_advance(vkd, offb), idesc_s, k != 0);
          }
          umma_commit<1>(&s_full[s]);
          umma_commit<1>(&kk_empty[s]);
          if (i == NS - 1) umma_commit<1>(q_empty);
        }
        __syncwarp();
      }
    }
  } else if (warp == kIssuerB) {
    // ---- issuer B: dQ += dS K (A operand from TMEM)
    constexpr uint32_t idesc_g = make_idesc_tf32(128, DH, 0, 1);
    const uint64_t kmd0 = make_smem_desc(smem_u32(St + 2 * T64), KBLK64, 512, kLayoutSw128Base32);
    uint32_t dq_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int NS = subs_of(w);
      for (int i = 0; i < NS; ++i, ++dq_it) {
        const int s = dq_it & 1;
        const uint32_t ph = (dq_it >> 1) & 1;
        mbar_wait(&km_full[s], ph);
        if (i == 0) mbar_wait(acc_empty, (item_it & 1) ^ 1);
        mbar_wait(&p_full[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t kmd = desc_advance(kmd0, s * 3 * T64);
          const uint32_t acc_on = i > 0;
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dQ += dS K
            umma_tf32_ts(tmem_base + 256, tmem_base + 128 + s * 64 + k * 8, desc_advance(kmd, k * 1024), idesc_g, acc_on | (k != 0));
          umma_commit<1>(&km_empty[s]);
          umma_commit<1>(&sfree[s]);
          if (i == NS - 1) umma_commit<1>(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int OC = DH / 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float c = p.scale * kLog2eF;
    uint32_t t_it = 0, item_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
      const int NS = subs_of(w);
      const int qt = w % p.tiles128;
      const int bh = w / p.tiles128;
      const int h = bh % p.heads, b = bh / p.heads;
      const int row = qt * 128 + q * 32 + lane;
      const long long sidx = ((long long)b * p.heads + h) * p.N + row;
      const float lse2 = row < p.N ? p.lse[sidx] * kLog2eF : INFINITY;
      const float dl = row < p.N ? p.delta[sidx] : 0.f;
      for (int i = 0; i < NS; ++i, ++t_it) {
        const int s = t_it & 1;
        mbar_wait(&s_full[s], (t_it >> 1) & 1);
        tcgen05_fence_after();
        const int col = s * 64 + half * 32;
        int kv_left = p.N - i * 64 - half * 32;
        if (p.cond >= 0) kv_left = min(kv_left, max(row, p.cond - 1) + 1 - i * 64 - half * 32);   // stage-2 mask
        uint32_t v[32], g[32];
        tmem_ld_32x32(tmem_base + lane_off + col, v);
        tmem_ld_32x32(tmem_base + lane_off + 128 + col, g);
        tmem_ld_wait();
        if (kv_left >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            g[j] = tf32_bits_for_mma(ex2_approx(fmaf(__uint_as_float(v[j]), c, -lse2)) * (__uint_as_float(g[j]) - dl));
        } else {                                    // ragged last tile: padded key columns contribute nothing
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pr = (j < kv_left) ? ex2_approx(fmaf(__uint_as_float(v[j]), c, -lse2)) : 0.f;
            g[j] = tf32_bits_for_mma(pr * (__uint_as_float(g[j]) - dl));
          }
        }
        tmem_st_32x32(tmem_base + lane_off + 128 + col, g);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
      }
      mbar_wait(acc_full, item_it & 1);
      tcgen05_fence_after();
      float* dqp = p.dqkv + ((long long)b * p.N + row) * (3ll * inner) + h * DH + half * OC;
      {
        uint32_t v[OC];
        tmem_ld_cols<OC>(tmem_base + lane_off + 256 + half * OC, v);
        tmem_ld_wait();
        if constexpr (OC == 32) {
          float r[32];
          if (p.out_half) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
            warp_store_box_h(obox + (warp - 2) * 4096, &tmOut, r, p.scale * (p.out_scale ? __ldg(p.out_scale) : 1.f), h * DH + half * OC,
                             qt * 128 + q * 32, b, lane);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = p.round_out ? round_tf32(__uint_as_float(v[j]) * p.scale) : __uint_as_float(v[j]) * p.scale;
            warp_store_box(obox + (warp - 2) * 4096, &tmOut, r, h * DH + half * OC, qt * 128 + q * 32, b, lane);
          }
        } else if (row < p.N && p.out_half) {
          const float mul = p.scale * (p.out_scale ? __ldg(p.out_scale) : 1.f);
          __half* hp = reinterpret_cast<__half*>(p.dqkv) + ((long long)b * p.N + row) * (3ll * inner) + h * DH + half * OC;
#pragma unroll
          for (int j = 0; j < OC; j += 8) {
            uint4 pk;
            pk.x = pack_half2_sat(__uint_as_float(v[j]) * mul, __uint_as_float(v[j + 1]) * mul);
            pk.y = pack_half2_sat(__uint_as_float(v[j + 2]) * mul, __uint_as_float(v[j + 3]) * mul);
            pk.z = pack_half2_sat(__uint_as_float(v[j + 4]) * mul, __uint_as_float(v[j + 5]) * mul);
            pk.w = pack_half2_sat(__uint_as_float(v[j + 6]) * mul, __uint_as_float(v[j + 7]) * mul);
            *reinterpret_cast<uint4*>(hp + j) = pk;
          }
        } else if (row < p.N) {
#pragma unroll
          for (int j = 0; j < OC; j += 4) {
            float4 a = make_float4(__uint_as_float(v[j]) * p.scale, __uint_as_float(v[j + 1]) * p.scale,
                                   __uint_as_float(v[j + 2]) * p.scale, __uint_as_float(v[j + 3]) * p.scale);
            if (p.round_out) { a.x = round_tf32(a.x); a.y = round_tf32(a.y); a.z = round_tf32(a.z); a.w = round_tf32(a.w); }
            *reinterpret_cast<float4*>(dqp + j) = a;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (lane == 0) bulk_wait_group_read<0>();   // the store boxes must outlive the last TMA reads
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// K-major (SWIZZLE_128B) view of a [B*N, ld] matrix as {ld, N, B}, box {32, rows, 1}
static int make_kmajor_map(CUtensorMap* out, const float* ptr, long long ld, int N, int B, int box_rows) {
  const unsigned long long dims[3] = {(unsigned long long)ld, (unsigned long long)N, (unsigned long long)B};
  const unsigned long long strides[2] = {(unsigned long long)ld * 4, (unsigned long long)N * ld * 4};
  const unsigned box[3] = {32, (unsigned)box_rows, 1};
  return make_tensor_map_f32(out, ptr, 3, dims, strides, box, 0);
}
// MN-major (SWIZZLE_128B_BASE32B) view as {32, N, ld/32, B}, box {32, rows, atoms, 1}
static int make_mnmajor_map(CUtensorMap* out, const float* ptr, long long ld, int N, int B, int box_rows, int atoms) {
  const unsigned long long dims[4] = {32, (unsigned long long)N, (unsigned long long)(ld / 32), (unsigned long long)B};
  const unsigned long long strides[3] = {(unsigned long long)ld * 4, 128, (unsigned long long)N * ld * 4};
  const unsigned box[4] = {32, (unsigned)box_rows, (unsigned)atoms, 1};
  return make_tensor_map_f32(out, ptr, 4, dims, strides, box, 1);
}

template <int DH>
static int attn_bwd_tc_launch(const float* qkv, const float* dout, const float* lse, const float* delta, void* dqkv, int out_half,
                              const float* out_scale, int B, int N, int heads, float scale, int round_out, int cond,
                              cudaStream_t stream) {
  const int inner = heads * DH;
  const long long ld = 3ll * inner;
  CUtensorMap tmKV128, tmQ64, tmDO64, tmQM, tmDOM, tmDO128, tmOut;
  int rc;
  if (out_half) {   // {ld, N, B} fp16 view, 32 x 32 un-swizzled store boxes
    const unsigned long long dims[3] = {(unsigned long long)ld, (unsigned long long)N, (unsigned long long)B};
    const unsigned long long strides[2] = {(unsigned long long)ld * 2, (unsigned long long)N * ld * 2};
    const unsigned box[3] = {32, 32, 1};
    if ((rc = make_tensor_map(&tmOut, dqkv, 2, 3, dims, strides, box, 3))) return rc;
  } else if ((rc = make_kmajor_map(&tmOut, static_cast<const float*>(dqkv), ld, N, B, 32))) return rc;
  if ((rc = make_kmajor_map(&tmKV128, qkv, ld, N, B, 128))) return rc;
  if ((rc = make_kmajor_map(&tmQ64, qkv, ld, N, B, 64))) return rc;
  if ((rc = make_kmajor_map(&tmDO64, dout, inner, N, B, 64))) return rc;
  if ((rc = make_kmajor_map(&tmDO128, dout, inner, N, B, 128))) return rc;
  if ((rc = make_mnmajor_map(&tmQM, qkv, ld, N, B, 64, DH / 32))) return rc;
  if ((rc = make_mnmajor_map(&tmDOM, dout, inner, N, B, 64, DH / 32))) return rc;
  AttnBwdParams p;
  p.lse = lse; p.delta = delta; p.dqkv = static_cast<float*>(dqkv); p.out_half = out_half; p.out_scale = out_scale;
  p.B = B; p.N = N; p.heads = heads;
  p.tiles128 = (N + 127) / 128; p.sub64 = (N + 63) / 64;
  p.total_items = p.tiles128 * heads * B;
  p.scale = scale; p.round_out = round_out; p.cond = cond;
  constexpr int smem_kv = 2 * 128 * DH * 4 + 8 * 64 * DH * 4 + kOutBoxBytes + 256 + 1024;
  constexpr int smem_q = 2 * 128 * DH * 4 + 6 * 64 * DH * 4 + kOutBoxBytes + 256 + 1024;
  auto k1 = attn_bwd_dkv_tc_kernel<DH>;
  auto k2 = attn_bwd_dq_tc_kernel<DH>;
  B200_CONFIGURE_SMEM_ONCE(k1, smem_kv);
  B200_CONFIGURE_SMEM_ONCE(k2, smem_q);
  int grid = num_sms();
  if (sm_limit() > 0 && grid > sm_limit()) grid = sm_limit();
  if (grid > p.total_items) grid = p.total_items;
  k1<<<grid, kAtcThreads, smem_kv, stream>>>(tmKV128, tmQ64, tmDO64, tmQM, tmDOM, tmOut, p);
  B200_LAUNCH_OK("attn_bwd_dkv_tc_kernel");
  k2<<<grid, kAtcThreads, smem_q, stream>>>(tmKV128, tmDO128, tmQ64, tmQM, tmOut, p);
  B200_LAUNCH_OK("attn_bwd_dq_tc_kernel");
  return 0;
}

#ifdef B200_ATTN_TRACE
extern "C" int b200vq_trace_read(long long* out) {   // out[3][2 * 36]: (event, clock) pairs per role, zero padded
  cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace));
  return 0;
}
#endif

int attention_backward_tc(const float* qkv, const float* dout, const float* lse, const float* delta, void* dqkv, int out_half,
                          const float* out_scale, int B, int N, int heads, int dh, float scale, int round_out, int cond_len,
                          cudaStream_t stream) {
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0,
                 "attention: qkv/dout must be 16-byte aligned");
  B200_CHECK_ARG(cond_len <= N, "attention: cond_len %d exceeds the sequence length %d", cond_len, N);
  if (dh == 64) return attn_bwd_tc_launch<64>(qkv, dout, lse, delta, dqkv, out_half, out_scale, B, N, heads, scale, round_out, cond_len, stream);
  return attn_bwd_tc_launch<32>(qkv, dout, lse, delta, dqkv, out_half, out_scale, B, N, heads, scale, round_out, cond_len, stream);
}

int attention_forward_tc(const float* qkv, void* out, int out_half, float* lse, int B, int N, int heads, int dh, float scale,
                         int round_out, int cond_len, cudaStream_t stream) {
  B200_CHECK_ARG(cond_len <= N, "attention: cond_len %d exceeds the sequence length %d", cond_len, N);
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "attention: qkv must be 16-byte aligned");
  if (dh == 64) return attn_fwd_tc_launch<64>(qkv, out, out_half, lse, B, N, heads, scale, round_out, cond_len, stream);
  return attn_fwd_tc_launch<32>(qkv, out, out_half, lse, B, N, heads, scale, round_out, cond_len, stream);
}

}  // namespace b200
