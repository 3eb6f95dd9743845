// tcgen05 / TMEM flash-attention forward for the ViT blocks (reference layers.py:124-130).
//
// Persistent, warp-specialised, one work item = (batch, head, 128-query tile):
//   warp 0      TMA producer: Q tile once per item, K and V tiles (128 keys) through 2-stage rings.
//               q/k are K-major operands (SWIZZLE_128B, one 128-byte block per 32 head dims);
//               V is the MN-major B operand of P.V (SWIZZLE_128B_BASE32B, 4-D tensor map).
//   warp 1      MMA issuer: S_j = Q K_j^T (kind::tf32, 128x128 accumulator in TMEM, two S buffers),
//               then O_j = P_j V_j with the A operand read from TMEM (P_j overwrites S_j in place).
//               Issue order S_0 S_1 PV_0 S_2 PV_1 ... so the tensor pipe works on S_{j+1} while the
//               softmax warps process S_j.
//   warps 2..5  softmax: thread = query row (TMEM lane), two passes over the 128 scores of a tile
//               (max, then exp2 + sum + tf32-rounded P written back with tcgen05.st); the running
//               output lives in registers: o = o * alpha + (P_j V_j read back from TMEM).
// TMEM columns: S/P buffers [0,128) [128,256), PV buffers [256,256+DH) [320,320+DH).
#include "common.cuh"

namespace b200 {

constexpr int kAtcThreads = 192;
constexpr float kLog2eF = 1.4426950408889634f;

struct AttnTcParams {
  float* out;        // [B*N, heads*DH]
  float* lse;        // [B*heads*N]
  int B, N, heads;
  int q_tiles, kv_tiles, total_items;
  float scale;
  int round_out;
};

template <int DH>
__global__ void __launch_bounds__(kAtcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  constexpr int KB = DH / 32;                 // 128-byte k-blocks per row
  constexpr int TILE_BYTES = 128 * DH * 4;    // one Q / K / V tile
  constexpr int KBLK_BYTES = 128 * 128;       // one k-block (or one MN atom of V): 128 rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + TILE_BYTES;            // [2]
  uint8_t* Vs = smem + 3 * TILE_BYTES;        // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TILE_BYTES);
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], 4);
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t kv_it = 0, item_it = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
        const int qt = w % p.q_tiles;
        const int bh = w / p.q_tiles;
        const int h = bh % p.heads, b = bh / p.heads;
        mbar_wait(q_empty, (item_it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, TILE_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(Qs + kb * KBLK_BYTES, &tmQK, q_full, h * DH + kb * 32, qt * 128, b);
        for (int j = 0; j < T; ++j, ++kv_it) {
          const int s = kv_it & 1;
          const uint32_t ph = (kv_it >> 1) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(Ks + s * TILE_BYTES + kb * KBLK_BYTES, &tmQK, &k_full[s], inner + h * DH + kb * 32, j * 128, b);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
          tma_load_4d(Vs + s * TILE_BYTES, &tmV, &v_full[s], 0, j * 128, (2 * inner + h * DH) / 32, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_tf32(128, 128, 0, 0);   // S = Q K^T, both K-major
      constexpr uint32_t idesc_o = make_idesc_tf32(128, DH, 0, 1);    // O = P V, B MN-major
      uint32_t s_it = 0, pv_it = 0, item_it = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++item_it) {
        mbar_wait(q_full, item_it & 1);
        tcgen05_fence_after();
        const uint32_t qa = smem_u32(Qs);
        for (int j = 0; j <= T; ++j) {
          if (j < T) {
            const int s = s_it & 1;
            const uint32_t ph = (s_it >> 1) & 1;
            mbar_wait(&k_full[s], ph);
            tcgen05_fence_after();
            const uint32_t ka = smem_u32(Ks + s * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < DH / 8; ++k) {
              const uint32_t off = (k >> 2) * KBLK_BYTES + (k & 3) * 32;
              umma_tf32<1>(tmem_base + s * 128, make_smem_desc(qa + off, 16, 1024, kLayoutSw128),
                           make_smem_desc(ka + off, 16, 1024, kLayoutSw128), idesc_s, k != 0);
            }
            umma_commit<1>(&s_full[s]);
            umma_commit<1>(&k_empty[s]);
            if (j == T - 1) umma_commit<1>(q_empty);
            ++s_it;
          }
          if (j >= 1) {
            const int s = pv_it & 1;
            const uint32_t ph = (pv_it >> 1) & 1;
            mbar_wait(&p_full[s], ph);
            mbar_wait(&v_full[s], ph);
            mbar_wait(&o_empty[s], ph ^ 1);
            tcgen05_fence_after();
            const uint32_t va = smem_u32(Vs + s * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 16; ++k)
              umma_tf32_ts(tmem_base + 256 + s * 64, tmem_base + s * 128 + k * 8,
                           make_smem_desc(va + k * 1024, KBLK_BYTES, 512, kLayoutSw128Base32), idesc_o, k != 0);
            umma_commit<1>(&o_full[s]);
            umma_commit<1>(&v_empty[s]);
            ++pv_it;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / output warps
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int row_in_tile = q * 32 + lane;
    const float c = p.scale * kLog2eF;
    uint32_t t_it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
      const int qt = w % p.q_tiles;
      const int bh = w / p.q_tiles;
      const int h = bh % p.heads, b = bh / p.heads;
      float o[DH];
#pragma unroll
      for (int i = 0; i < DH; ++i) o[i] = 0.f;
      float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
      auto accumulate_pv = [&](uint32_t it, float alpha) {
        const int sp = it & 1;
        mbar_wait(&o_full[sp], (it >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int cc = 0; cc < DH / 32; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + lane_off + 256 + sp * 64 + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[cc * 32 + i] = fmaf(o[cc * 32 + i], alpha, __uint_as_float(v[i]));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[sp]);
      };
      for (int j = 0; j < T; ++j, ++t_it) {
        const int s = t_it & 1;
        mbar_wait(&s_full[s], (t_it >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t sa = tmem_base + lane_off + s * 128;
        const int kv_left = p.N - j * 128;       // keys >= kv_left in this tile are padding
        float mx = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32(sa + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = (cc * 32 + i < kv_left) ? __uint_as_float(v[i]) : -INFINITY;
            mx = fmaxf(mx, x);
          }
        }
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2_approx((m - m_new) * c);
        const float mc = m_new * c;
        float sum = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32(sa + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = (cc * 32 + i < kv_left) ? __uint_as_float(v[i]) : -INFINITY;
            const float e = ex2_approx(fmaf(x, c, -mc));
            sum += e;
            v[i] = __float_as_uint(round_tf32(e));
          }
          tmem_st_32x32(sa + cc * 32, v);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        l = fmaf(l, alpha, sum);
        if (j >= 1) accumulate_pv(t_it - 1, alpha_prev);
        alpha_prev = alpha;
        m = m_new;
      }
      accumulate_pv(t_it - 1, alpha_prev);
      const int row = qt * 128 + row_in_tile;
      if (row < p.N) {
        const float inv = 1.f / l;
        float* op = p.out + ((long long)b * p.N + row) * inner + h * DH;
#pragma unroll
        for (int i = 0; i < DH; i += 4) {
          float4 r = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
          if (p.round_out) { r.x = round_tf32(r.x); r.y = round_tf32(r.y); r.z = round_tf32(r.z); r.w = round_tf32(r.w); }
          *reinterpret_cast<float4*>(op + i) = r;
        }
        p.lse[((long long)b * p.heads + h) * p.N + row] = m * p.scale + logf(l);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int DH>
static int attn_fwd_tc_launch(const float* qkv, float* out, float* lse, int B, int N, int heads, float scale, int round_out,
                              cudaStream_t stream) {
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
  p.out = out; p.lse = lse; p.B = B; p.N = N; p.heads = heads;
  p.q_tiles = (N + 127) / 128; p.kv_tiles = (N + 127) / 128;
  p.total_items = p.q_tiles * heads * B;
  p.scale = scale; p.round_out = round_out;
  constexpr int smem = 5 * 128 * DH * 4 + 1024 + 256;
  auto kern = attn_fwd_tc_kernel<DH>;
  static bool configured = false;
  if (!configured) { B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); configured = true; }
  int grid = num_sms();
  if (grid > p.total_items) grid = p.total_items;
  kern<<<grid, kAtcThreads, smem, stream>>>(tmQK, tmV, p);
  B200_LAUNCH_OK("attn_fwd_tc_kernel");
  return 0;
}

int attention_forward_tc(const float* qkv, float* out, float* lse, int B, int N, int heads, int dh, float scale, int round_out,
                         cudaStream_t stream) {
  B200_CHECK_ARG(dh == 64 || dh == 32, "attention: dim_head must be 32 or 64 (got %d)", dh);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "attention: qkv must be 16-byte aligned");
  if (dh == 64) return attn_fwd_tc_launch<64>(qkv, out, lse, B, N, heads, scale, round_out, stream);
  return attn_fwd_tc_launch<32>(qkv, out, lse, B, N, heads, scale, round_out, stream);
}

}  // namespace b200
