// Persistent TF32 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (TMEM
// accumulators, double buffered) -> fused epilogue.  Replaces every cuBLAS/cuDNN call the
// reference issues for nn.Linear / Conv2d(k=s) / ConvTranspose2d(k=s) on the hot path
// (reference enhancing/modules/stage1/layers.py:99-101,118-132,168-171,202-205) and their
// dgrad / wgrad.
//
//   C[M,N] = epilogue( sum_k A[m,k] * B[n,k] )                      (per split z)
//
// Operand storage ("major"):
//   a_major = 0 : A is [M, K]        row-major, K contiguous   (K-major)
//   a_major = 1 : A is [K_total, M]  row-major, M contiguous   (MN-major; wgrad reads dY^T)
//   b_major = 0 : B is [N, K]        row-major                 (nn.Linear weight)
//   b_major = 1 : B is [K_total, N]  row-major                 (dgrad reads W, wgrad reads X)
// Split-K: split z contracts rows/cols [z*K, (z+1)*K) and writes C + z*c_split_stride.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter take alternate
// 32-column chunks so that one warp's TMEM-load / TMA latencies hide behind the other's math).  CG = 2 pairs two CTAs on one
// 256 x BN tile (tcgen05 cta_group::2): each CTA stages its 128 rows of A and half of B.
#include "common.cuh"

#include <mutex>
#include <unordered_map>

namespace b200 {

struct GemmParams {
  int M, N, K;                 // output rows / cols, contraction length per split
  int num_m_blocks;            // ceil(M / (128*CG)) cluster tiles along M
  int num_n_blocks;            // ceil(N / BN)
  int num_splits;
  int k_blocks;                // ceil(K / 32)
  float* C;
  long long ldc, c_split_stride;
  const float* bias;           // [N] or null
  const float* res;            // residual added in the epilogue through direct loads (row % res_row_mod), or null
  long long ldres;
  int res_row_mod;             // >0: residual row = row % res_row_mod (positional table)
  int in_mode;                 // 0 none, 1: + tile of tmIn (residual), 2: * (1 - tile^2) (tanh backward)
  int act;                     // 0 none, 1 tanh
  int round_out;               // 1: round the stored value to tf32 (it only feeds another GEMM)
  float* colsum_part;          // null, or [ceil(M/32)][N]: column sums of every 32-row group of the stored C
};

constexpr int kBM = 128;
constexpr int kBK = 32;        // 32 fp32 = one 128-byte swizzle row
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

template <int BN, int CG>
struct GemmCfg {
  static constexpr int BN_CTA = BN / CG;
  static constexpr int A_BYTES = kBM * kBK * 4;
  static constexpr int B_BYTES = BN_CTA * kBK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 8 * 8192;      // per epilogue warp: out[4 KB] + in[4 KB]
  static constexpr int MAX_SMEM = 232448;             // 227 KB opt-in limit
  static constexpr int RING_BUDGET = MAX_SMEM - STAGING_BYTES - 1024 - 512;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

template <int BN, int CG, int AMAJ, int BMAJ>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmIn, const GemmParams p) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
  uint64_t* in_full_all = bars + 2 * STAGES + 4;  // [8 epilogue warps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (p.in_mode) tma_prefetch_desc(&tmIn);
    for (int i = 0; i < 8; ++i) mbar_init(&in_full_all[i], 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // the leader's producer arrives once, expecting the bytes of the whole pair
      mbar_init(&empty_bar[s], 1);   // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8 * CG);  // one arrive per epilogue warp of every CTA in the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_split = p.num_m_blocks * p.num_n_blocks;
  const int total_tiles = tiles_per_split * p.num_splits;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (warp-uniform loop, one lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int z = t / tiles_per_split;
      const int r = t - z * tiles_per_split;
      const int mb = r / p.num_n_blocks, nb = r - mb * p.num_n_blocks;
      const int m0 = (mb * CG + (int)cta_rank) * kBM;
      const int n0 = nb * BN + (int)cta_rank * Cfg::BN_CTA;
      const int kbase = z * p.K;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = kbase + kb * kBK;
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if constexpr (AMAJ == 0) tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
            else                     tma_load_3d(sa, &tmA, &full_bar[stage], 0, k0, m0 / 32);
            if constexpr (BMAJ == 0) tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
            else                     tma_load_3d(sb, &tmB, &full_bar[stage], 0, k0, n0 / 32);
          } else {
            // Only the leader arrives (expecting both CTAs' bytes).  The follower's TMA credits
            // the leader's barrier directly; it may land before the leader's expect_tx of the
            // same phase (tx-count goes transiently negative), which is legal: the phase cannot
            // complete before the leader's single pending arrival.
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            if constexpr (AMAJ == 0) tma_load_2d_2sm(sa, &tmA, &full_bar[stage], k0, m0);
            else                     tma_load_3d_2sm(sa, &tmA, &full_bar[stage], 0, k0, m0 / 32);
            if constexpr (BMAJ == 0) tma_load_2d_2sm(sb, &tmB, &full_bar[stage], k0, n0);
            else                     tma_load_3d_2sm(sb, &tmB, &full_bar[stage], 0, k0, n0 / 32);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform loop, one lane issues)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM * CG, BN, AMAJ, BMAJ);
      // K-major (SWIZZLE_128B): 8-row groups 1024 B apart (SBO), LBO unused.
      // MN-major (SWIZZLE_128B_BASE32B, mandatory for tf32): smem holds [mn atom of 32][32 k-rows][128 B];
      // atoms 4096 B apart (LBO), 4-k-row swizzle groups 512 B apart (SBO).
      constexpr uint32_t A_LBO = AMAJ ? 4096 : 16, B_LBO = BMAJ ? 4096 : 16;
      constexpr uint32_t A_SBO = AMAJ ? 512 : 1024, B_SBO = BMAJ ? 512 : 1024;
      constexpr uint32_t A_LAY = AMAJ ? kLayoutSw128Base32 : kLayoutSw128, B_LAY = BMAJ ? kLayoutSw128Base32 : kLayoutSw128;
      constexpr uint32_t A_KSTEP = AMAJ ? 1024 : 32, B_KSTEP = BMAJ ? 1024 : 32;  // bytes per UMMA_K = 8
      const uint64_t a_desc0 = make_smem_desc(smem_u32(smem), A_LBO, A_SBO, A_LAY);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(smem) + Cfg::A_BYTES, B_LBO, B_SBO, B_LAY);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t ad = desc_advance(a_desc0, stage * Cfg::STAGE_BYTES);
            const uint64_t bd = desc_advance(b_desc0, stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k)
              umma_tf32<CG>(d_tmem, desc_advance(ad, k * A_KSTEP), desc_advance(bd, k * B_KSTEP), idesc, (kb | k) != 0);
            umma_commit<CG>(&empty_bar[stage]);  // frees the slot in both CTAs once the MMAs retire
            if (kb == p.k_blocks - 1) umma_commit<CG>(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    // Each warp owns 32 rows of the tile (TMEM lane quarter q) and every other 32-column chunk.
    // Per chunk: tcgen05.ld -> registers -> fused math -> 128B-swizzled smem box -> TMA store (fully
    // coalesced, clipped at the matrix edge).  Residual / tanh' inputs arrive the same way through
    // TMA loads issued one chunk ahead.
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int ew = warp - 2;
    const int ehalf = ew >> 2;
    uint8_t* out_buf = staging + ew * 8192;
    uint8_t* in_buf = out_buf + 4096;
    uint64_t* in_full = in_full_all + ew;
    const uint32_t swz = lane & 7;
    const bool tma_in = p.in_mode != 0;
    uint32_t in_cnt = 0;
    constexpr int NCHUNK = BN / 32;
    int it = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters, ++it) {
      const int z = t / tiles_per_split;
      const int r = t - z * tiles_per_split;
      const int mb = r / p.num_n_blocks, nb = r - mb * p.num_n_blocks;
      const int row0 = (mb * CG + (int)cta_rank) * kBM + q * 32;
      const int row = row0 + lane;
      const int n0 = nb * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (tma_in && ehalf < NCHUNK && lane == 0) {   // this warp's first input chunk of the tile (buffer free since the last tile)
        mbar_arrive_expect_tx(in_full, 4096);
        tma_load_2d(in_buf, &tmIn, in_full, n0 + ehalf * 32, row0);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const float* rrow = nullptr;
      if (p.res && row < p.M) rrow = p.res + (long long)(p.res_row_mod > 0 ? row % p.res_row_mod : row) * p.ldres;
#pragma unroll 1
      for (int c = ehalf; c < NCHUNK; c += 2) {
        const int col0 = n0 + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_STRIDE + c * 32, v);
        float4 a[8];
        if (tma_in) {
          mbar_wait(in_full, in_cnt & 1);
          ++in_cnt;
          const uint8_t* inb = in_buf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(inb + ((j ^ swz) << 4));
        }
        tmem_ld_wait();
        if (tma_in) {   // placed after the TMEM wait so that the stall on the loads below is already paid for
          // The refill below overwrites in_buf through the async proxy, so every lane's loads of this chunk must
          // have *completed*, not merely issued: while tcgen05.mma operand fetches saturate shared memory a
          // load can stay in flight longer than a TMA round trip (seen as one stale 16-byte group per few
          // thousand tiles).  The ballot consumes a loaded register of every 16-byte group of every lane, and
          // the refill is control-dependent on its result (which is never 0 in practice).
          uint32_t bits = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) bits |= __float_as_uint(a[j].x);
          const uint32_t landed = __ballot_sync(0xffffffffu, bits != 0x7fc0deadu);
          if (lane == 0 && c + 2 < NCHUNK && landed != 0u) {
            mbar_arrive_expect_tx(in_full, 4096);
            tma_load_2d(in_buf, &tmIn, in_full, col0 + 64, row0);
          }
        }
        // the staging buffer must have been read by the TMA store of this warp's previous chunk
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
        uint8_t* outb = out_buf + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
          const int col = col0 + 4 * j;
          if (p.bias && col < p.N) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (p.act == 1) { o.x = fast_tanh(o.x); o.y = fast_tanh(o.y); o.z = fast_tanh(o.z); o.w = fast_tanh(o.w); }
          if (tma_in) {
            if (p.in_mode == 2) {
              o.x *= 1.f - a[j].x * a[j].x; o.y *= 1.f - a[j].y * a[j].y; o.z *= 1.f - a[j].z * a[j].z; o.w *= 1.f - a[j].w * a[j].w;
            } else {
              o.x += a[j].x; o.y += a[j].y; o.z += a[j].z; o.w += a[j].w;
            }
          }
          if (rrow && col < p.N) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(rrow + col));
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
          *reinterpret_cast<float4*>(outb + ((j ^ swz) << 4)) = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (col0 < p.N && row0 < p.M) tma_store_3d(&tmC, out_buf, col0, row0, z);
          bulk_commit_group();
        }
        if (p.colsum_part) {
          // bias gradient for free: lane c sums column c of the 32 x 32 box just staged (rows past M hold
          // epilogue constants and are skipped); the row-group partials are reduced by a colsum launch
          const int rmax = p.M - row0;
          const uint8_t* colp = out_buf + ((lane & 3) << 2);
          const uint32_t cj = lane >> 2;
          float cs0 = 0.f, cs1 = 0.f;
#pragma unroll 1
          for (int rb = 0; rb < 32; rb += 8) {       // 8 rows = one swizzle period; kept rolled so the 32 addresses are not hoisted
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
              if (rb + k < rmax) cs0 += *reinterpret_cast<const float*>(colp + (rb + k) * 128 + ((cj ^ k) << 4));
              if (rb + k + 1 < rmax) cs1 += *reinterpret_cast<const float*>(colp + (rb + k + 1) * 128 + ((cj ^ (k + 1)) << 4));
            }
          }
          if (col0 + lane < p.N && row0 < p.M) p.colsum_part[(size_t)(row0 >> 5) * p.N + col0 + lane] = cs0 + cs1;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || leader) mbar_arrive(&tmem_empty[acc]);
        else                   mbar_arrive_remote(&tmem_empty[acc], 0);
      }
    }
    if (lane == 0) bulk_wait_group_read<0>();   // smem must outlive the last stores' reads
  }

  // ------------------------------------------------------------------ teardown
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so
// the library has no link-time dependency on libcuda) and launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr; int rank, swz; unsigned long long dims[5], strides[4]; unsigned box[5];
  bool operator==(const TmapKey& o) const {
    if (ptr != o.ptr || rank != o.rank || swz != o.swz) return false;
    for (int i = 0; i < rank; ++i) if (dims[i] != o.dims[i] || box[i] != o.box[i]) return false;
    for (int i = 0; i + 1 < rank; ++i) if (strides[i] != o.strides[i]) return false;
    return true;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.rank * 2 + k.swz);
    for (int i = 0; i < k.rank; ++i) { mix((size_t)k.dims[i]); mix((size_t)k.box[i]); }
    for (int i = 0; i + 1 < k.rank; ++i) mix((size_t)k.strides[i]);
    return h;
  }
};

int make_tensor_map_f32(CUtensorMap* out, const float* ptr, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box, int swizzle_base32) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{};
  key.ptr = ptr; key.rank = rank; key.swz = swizzle_base32;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(-4, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(ptr), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_base32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(-4, "cuTensorMapEncodeTiled failed (%d) ptr=%p rank=%d dims0=%llu box0=%u", (int)r, ptr, rank, dims[0], box[0]);
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// major 0: matrix [rows = MN extent, cols = K extent], box {32 k, box_mn rows}
// major 1: matrix [rows = K extent, cols = MN extent], viewed as {32, rows, cols/32}, box {32, 32, box_mn/32}
static int make_operand_tmap(CUtensorMap* out, const float* ptr, long long ld, int rows, int cols, int major, int box_mn) {
  if (major == 0) {
    const unsigned long long dims[2] = {(unsigned long long)cols, (unsigned long long)rows};
    const unsigned long long strides[1] = {(unsigned long long)ld * 4};
    const unsigned box[2] = {32, (unsigned)box_mn};
    return make_tensor_map_f32(out, ptr, 2, dims, strides, box, 0);
  }
  const unsigned long long dims[3] = {32, (unsigned long long)rows, (unsigned long long)(cols / 32)};
  const unsigned long long strides[2] = {(unsigned long long)ld * 4, 128};
  const unsigned box[3] = {32, 32, (unsigned)(box_mn / 32)};
  return make_tensor_map_f32(out, ptr, 3, dims, strides, box, 1);
}

template <int BN, int CG, int AMAJ, int BMAJ>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmIn,
                       const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG>;
  auto kern = gemm_tf32_kernel<BN, CG, AMAJ, BMAJ>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int total = p.num_m_blocks * p.num_n_blocks * p.num_splits;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((num_sms() / CG) * CG);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent kernel: launch exactly as many clusters as can be co-resident (a CTA pair needs
  // two free SMs of one TPC; fewer than num_sms/2 pairs may fit) so that no cluster waits for
  // a second wave behind CTAs that never exit early
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = num_sms() / CG;
    max_clusters = n;
  }
  int clusters = max_clusters < total ? max_clusters : total;
  cfg.gridDim = dim3(clusters * CG);
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmIn, p));
  count_launch();
  return 0;
}

template <int BN, int CG>
static int dispatch_major(int am, int bm, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                          const CUtensorMap& in, const GemmParams& p, cudaStream_t s) {
  if (am == 0 && bm == 0) return launch_gemm<BN, CG, 0, 0>(a, b, c, in, p, s);
  if (am == 0 && bm == 1) return launch_gemm<BN, CG, 0, 1>(a, b, c, in, p, s);
  if (am == 1 && bm == 1) return launch_gemm<BN, CG, 1, 1>(a, b, c, in, p, s);
  if (am == 1 && bm == 0) return launch_gemm<BN, CG, 1, 0>(a, b, c, in, p, s);
  return set_error(-1, "bad operand major (%d,%d)", am, bm);
}

int gemm_tf32(const float* A, long long lda, int a_major, const float* B, long long ldb, int b_major, float* C,
              long long ldc, int M, int N, int K, int splits, long long c_split_stride, const float* bias,
              const float* res, long long ldres, int res_row_mod, const float* aux, long long ldaux,
              float* colsum_part, int act, int round_out, int cta_group, int bn, cudaStream_t stream) {
  B200_CHECK_ARG(!colsum_part || splits == 1, "gemm: colsum_part cannot be combined with split-K");
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0 && splits > 0, "gemm: empty problem M=%d N=%d K=%d splits=%d", M, N, K, splits);
  B200_CHECK_ARG(N % 4 == 0 && ldc % 4 == 0, "gemm: N and ldc must be multiples of 4 (N=%d ldc=%lld)", N, ldc);
  B200_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "gemm: lda/ldb must be multiples of 4 floats (16-byte TMA strides)");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm: operands must be 16-byte aligned");
  B200_CHECK_ARG(!(a_major == 1 && M % 32), "gemm: MN-major A needs M %% 32 == 0 (M=%d)", M);
  B200_CHECK_ARG(!(b_major == 1 && N % 32), "gemm: MN-major B needs N %% 32 == 0 (N=%d)", N);
  B200_CHECK_ARG(splits == 1 || K % 32 == 0, "gemm: split-K needs K %% 32 == 0");
  B200_CHECK_ARG(!res || ldres % 4 == 0, "gemm: ldres %% 4");
  B200_CHECK_ARG(!aux || ldaux % 4 == 0, "gemm: ldaux %% 4");
  if (cta_group != 2) cta_group = 1;
  if (bn <= 0) bn = N >= 256 ? 256 : (N > 128 ? (N <= 192 ? 192 : 256) : (N > 64 ? 128 : 64));
  B200_CHECK_ARG(bn == 64 || bn == 128 || bn == 192 || bn == 256, "gemm: unsupported BN %d", bn);

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = (M + kBM * cta_group - 1) / (kBM * cta_group);
  p.num_n_blocks = (N + bn - 1) / bn;
  p.num_splits = splits;
  p.k_blocks = (K + kBK - 1) / kBK;
  p.C = C; p.ldc = ldc; p.c_split_stride = c_split_stride;
  B200_CHECK_ARG(!(aux && res && res_row_mod == 0), "gemm: residual and tanh' inputs cannot be combined");
  p.bias = bias; p.act = act; p.round_out = round_out; p.colsum_part = colsum_part;
  p.res = (res && res_row_mod > 0) ? res : nullptr; p.ldres = ldres; p.res_row_mod = res_row_mod;
  p.in_mode = aux ? 2 : ((res && res_row_mod == 0) ? 1 : 0);

  CUtensorMap tmA, tmB, tmC, tmIn;
  {
    const unsigned long long dims[3] = {(unsigned long long)N, (unsigned long long)M, (unsigned long long)splits};
    const unsigned long long strides[2] = {(unsigned long long)ldc * 4,
                                           (unsigned long long)(splits > 1 ? c_split_stride : (long long)M * ldc) * 4};
    const unsigned box[3] = {32, 32, 1};
    B200_CHECK_ARG(splits == 1 || c_split_stride % 4 == 0, "gemm: c_split_stride %% 4");
    int rc0 = make_tensor_map_f32(&tmC, C, 3, dims, strides, box, 0);
    if (rc0) return rc0;
    tmIn = tmC;
    if (p.in_mode) {
      const float* src = aux ? aux : res;
      const long long ldin = aux ? ldaux : ldres;
      B200_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0, "gemm: epilogue input must be 16-byte aligned");
      const unsigned long long d2[2] = {(unsigned long long)N, (unsigned long long)M};
      const unsigned long long s2[1] = {(unsigned long long)ldin * 4};
      const unsigned b2[2] = {32, 32};
      rc0 = make_tensor_map_f32(&tmIn, src, 2, d2, s2, b2, 0);
      if (rc0) return rc0;
    }
  }
  const int ktot = K * splits;
  int rc;
  if (a_major == 0) rc = make_operand_tmap(&tmA, A, lda, M, ktot, 0, kBM);
  else              rc = make_operand_tmap(&tmA, A, lda, ktot, M, 1, kBM);
  if (rc) return rc;
  const int bn_cta = bn / cta_group;
  if (b_major == 0) rc = make_operand_tmap(&tmB, B, ldb, N, ktot, 0, bn_cta);
  else              rc = make_operand_tmap(&tmB, B, ldb, ktot, N, 1, bn_cta);
  if (rc) return rc;

#define B200_GEMM_CASE(BN_)                                                                         \
  case BN_:                                                                                         \
    return cta_group == 2 ? dispatch_major<BN_, 2>(a_major, b_major, tmA, tmB, tmC, tmIn, p, stream) \
                          : dispatch_major<BN_, 1>(a_major, b_major, tmA, tmB, tmC, tmIn, p, stream);
  switch (bn) {
    B200_GEMM_CASE(64)
    B200_GEMM_CASE(128)
    B200_GEMM_CASE(192)
    B200_GEMM_CASE(256)
  }
#undef B200_GEMM_CASE
  return set_error(-1, "gemm: unreachable");
}

}  // namespace b200
