// Causal flash attention FORWARD on tcgen05 tensor cores (head_dim 128): S = Q.K^T and O += P.V are tcgen05.mma with the
// accumulators in TMEM, Q/K/V tiles arrive by TMA into 128B-swizzled shared memory, the softmax runs on 128 threads
// that each own one query row (= one TMEM lane).  Same semantics and rounding points as attention.cu (the mma.sync
// implementation that still serves head_dim 32/64 and the backward): key j visible to query i iff j <= i and
// mask[b,j]; masked-out query rows give zeros; P is rounded to bf16 before the PV product, statistics are fp32.
// With a shared-prefix layout (FaParams::group > 0) key j is additionally hidden from query i when j lies in the
// suffix region (j >= P_b) but in another group than i: the R diffusion repeats of a sample then share one prefix.
//
// Two CTAs per (batch, head) split that sequence's 128-row q tiles between them in a zig-zag (heaviest first, so both
// get the same number of causal kv tiles) and walk them persistently: the TMEM allocation, barrier set-up and pipeline
// fill are paid once per ~15 kv tiles.  112 KB of shared memory and 256 TMEM columns per CTA let TWO CTAs share an SM,
// so one CTA's q-tile epilogue / barrier round trips are covered by the other's main loop.
//   warp 0      TMA loader   Q per q tile, K_j / V_j (64 kv rows per tile, 2-slot rings; a K slot is released as soon
//                            as its QK^T retired, a V slot after its PV)
//   warp 1      MMA issuer   S_{j+1} = Q.K_{j+1}^T is issued before P_j.V_j, so the tensor core works on the next tile
//                            while the softmax warps are busy with the current one (two S buffers in TMEM)
//   warps 2-5   softmax      TMEM -> registers once (tcgen05.ld 32x32b.x32), row max, ex2.approx, P -> swizzled smem
//                            (K-major A operand of the PV MMA); O is rescaled in TMEM lazily, only when a running
//                            maximum grew by more than 2^8 (a stale maximum is exact after the final normalisation)
// TMEM: S0 @ col 0, S1 @ col 64, O @ col 128 (fp32, 128 lanes).   smem: 32 KB Q + 2x16 KB K + 2x16 KB V + 16 KB P.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int FA_BM = 128, FA_BN = 64, FA_D = 128;
constexpr int FA_Q_BYTES = FA_BM * FA_D * 2;      // 32 KB: two 64-column chunks of [128 rows x 128 B]
constexpr int FA_KV_BYTES = FA_BN * FA_D * 2;     // 16 KB: two 64-column chunks of [64 rows x 128 B]
constexpr int FA_P_BYTES = FA_BM * FA_BN * 2;     // 16 KB: [128 rows x 128 B]
constexpr int FA_THREADS = 192;
constexpr float FA_RESCALE_THRESHOLD = 8.f;       // log2 units
constexpr int FA_TILES_BYTES = FA_Q_BYTES + 4 * FA_KV_BYTES + FA_P_BYTES;   // 112 KB
constexpr int FA_SMEM = FA_TILES_BYTES + 256 + 64 + 16;                     // tiles | barriers | mask tile | flag
constexpr float FA_LOG2E = 1.4426950408889634f;

struct FaParams {
  int B, S, H;
  float scale;
  __nv_bfloat16* o;
  int64_t ld_o;
  float* lse;           // [B, H, S]
  const uint8_t* mask;  // [B, S] or null
  // shared-prefix layout: rows [0, P_b) of sequence b are a causal prefix, rows from P_b on are groups of
  // `group` rows that see the whole prefix and (causally) their own group only.  group = 0: plain causal attention.
  const int32_t* prefix_len;   // [B] P_b, or null
  int group;
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(FA_THREADS, 2)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, FaParams p) {
  // No static shared memory: the dynamic window then starts 1024-byte aligned (SWIZZLE_128B tiles need it; checked).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + FA_Q_BYTES;                    // 2 slots
  uint8_t* sV = sK + 2 * FA_KV_BYTES;                 // 2 slots
  uint8_t* sP = sV + 2 * FA_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_TILES_BYTES);
  uint64_t *q_full = bars, *k_full = bars + 1, *v_full = bars + 3, *k_empty = bars + 5, *v_empty = bars + 7,
           *s_full = bars + 9, *s_empty = bars + 11, *p_full = bars + 13, *pv_done = bars + 14, *q_empty = bars + 15,
           *o_full = bars + 16, *o_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  uint8_t* s_mask_tile = smem + FA_TILES_BYTES + 256;
  int& s_any_masked = *reinterpret_cast<int*>(smem + FA_TILES_BYTES + 256 + 64);

  const int warp = int(warp_idx_uniform()), lane = threadIdx.x & 31;     // warp-uniform role index (see ptx.cuh)
  const int S = p.S;
  const int n_qt = (S + FA_BM - 1) / FA_BM;
  const int half_id = blockIdx.x & 1;                   // which of the two CTAs of this (batch, head)
  const int bh = blockIdx.x >> 1;
  const int b = bh / p.H, hd = bh % p.H;
  const int HD = p.H * FA_D;
  const int row_base = b * S;
  const int n_kv_max = (S + FA_BN - 1) / FA_BN;
  // q tiles in descending order are dealt A B B A A B B A ...: tile index of this CTA's k-th tile, or -1 when done
  auto my_tile = [&](int k) -> int {
    // positions (in the descending order) owned by CTA 0: 0,3,4,7,8,...; by CTA 1: 1,2,5,6,9,...
    const int pos = half_id == 0 ? (k == 0 ? 0 : 4 * ((k + 1) >> 1) - ((k & 1) ? 1 : 0)) : (4 * (k >> 1) + 1 + (k & 1));
    return pos < n_qt ? n_qt - 1 - pos : -1;
  };
  auto kv_tiles = [&](int t) { return min(2 * t + 2, n_kv_max); };   // causal: kv tiles 0 .. 2t+1

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4); mbar_init(pv_done, 1); mbar_init(q_empty, 1); mbar_init(o_full, 1); mbar_init(o_empty, 4);
    s_any_masked = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 0) {
    // ================================ TMA loader ================================
    if (lane == 0) {
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int t = my_tile(ti);
        if (t < 0) break;
        mbar_wait(q_empty, (ti & 1) ^ 1);          // previous q tile's QK^T products have all retired
        mbar_arrive_expect_tx(q_full, FA_Q_BYTES);
        tma_load_2d(sQ, &map_q, q_full, hd * FA_D, row_base + t * FA_BM);
        tma_load_2d(sQ + FA_Q_BYTES / 2, &map_q, q_full, hd * FA_D + 64, row_base + t * FA_BM);
        const int n_kv = kv_tiles(t);
        for (int j = 0; j < n_kv; ++j, ++it) {
          const int s = it & 1;
          const uint32_t ph = ((it >> 1) & 1) ^ 1;
          uint8_t* dk = sK + s * FA_KV_BYTES;
          uint8_t* dv = sV + s * FA_KV_BYTES;
          mbar_wait(&k_empty[s], ph);
          mbar_arrive_expect_tx(&k_full[s], FA_KV_BYTES);
          tma_load_2d(dk, &map_kv, &k_full[s], HD + hd * FA_D, row_base + j * FA_BN);
          tma_load_2d(dk + FA_KV_BYTES / 2, &map_kv, &k_full[s], HD + hd * FA_D + 64, row_base + j * FA_BN);
          mbar_wait(&v_empty[s], ph);
          mbar_arrive_expect_tx(&v_full[s], FA_KV_BYTES);
          tma_load_2d(dv, &map_kv, &v_full[s], 2 * HD + hd * FA_D, row_base + j * FA_BN);
          tma_load_2d(dv + FA_KV_BYTES / 2, &map_kv, &v_full[s], 2 * HD + hd * FA_D + 64, row_base + j * FA_BN);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // All 32 lanes run this loop (converged, warp-uniform values; each tcgen05.mma / commit elects its issuing lane):
    // the per-instruction issue cost drops from ~130 cycles (divergent `lane == 0` region) to a few uniform-datapath
    // instructions, which matters because a 128x64x16 product only occupies the tensor core for 32 cycles.
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(FA_BM, FA_BN, 0, 0);    // 128 x 64  (K = d)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(FA_BM, FA_D, 0, 1);    // 128 x 128 (K = kv), V is MN-major
      const uint64_t dQ = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024), dP = umma_smem_desc_sw128(smem_u32(sP), 16, 1024);
      const uint64_t dK0 = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t dV0 = umma_smem_desc_sw128(smem_u32(sV), FA_KV_BYTES / 2, 1024);
      auto issue_s = [&](int cur) {
        const int s = cur & 1;
        const uint32_t ph = (cur >> 1) & 1;
        mbar_wait(&k_full[s], ph);
        mbar_wait(&s_empty[s], ph ^ 1);
        tc_fence_after();
        const uint64_t dK = umma_desc_advance(dK0, s * FA_KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < FA_D / 16; ++kk) {
          // Q: [128 rows x 128 B] chunks 16 KB apart; K: [64 rows x 128 B] chunks 8 KB apart; +32 B per 16-wide k step
          umma_f16_ss_elect(tmem_base + s * FA_BN, umma_desc_advance(dQ, (kk >> 2) * (FA_Q_BYTES / 2) + (kk & 3) * 32),
                            umma_desc_advance(dK, (kk >> 2) * (FA_KV_BYTES / 2) + (kk & 3) * 32), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit_elect(&s_full[s]);
        umma_commit_elect(&k_empty[s]);
      };
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int t = my_tile(ti);
        if (t < 0) break;
        const int n_kv = kv_tiles(t);
        mbar_wait(q_full, ti & 1);
        issue_s(it);
        for (int j = 0; j < n_kv; ++j) {
          const int cur = it + j;
          if (j + 1 < n_kv) issue_s(cur + 1);
          else umma_commit_elect(q_empty);               // every QK^T of this q tile has been issued
          const int s = cur & 1;
          mbar_wait(p_full, cur & 1);
          if (j == 0) mbar_wait(o_empty, (ti & 1) ^ 1);   // previous q tile's O has been read out
          mbar_wait(&v_full[s], (cur >> 1) & 1);
          tc_fence_after();
          const uint64_t dV = umma_desc_advance(dV0, s * FA_KV_BYTES);
#pragma unroll
          for (int kk = 0; kk < FA_BN / 16; ++kk) {
            // P: K-major [128 x 64 kv]; V: MN-major, d chunks 8 KB apart (LBO), 16 kv rows per step = 2048 B
            umma_f16_ss_elect(tmem_o, umma_desc_advance(dP, kk * 32), umma_desc_advance(dV, kk * 2048), idesc_pv,
                              (j | kk) != 0 ? 1u : 0u);
          }
          umma_commit_elect(pv_done);
          umma_commit_elect(&v_empty[s]);
          if (j == n_kv - 1) umma_commit_elect(o_full);
        }
        it += n_kv;
      }
    }
  } else {
    // ================================ softmax (128 threads, one query row each) ================================
    const int quarter = warp & 3;       // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;  // row inside the q tile == TMEM lane
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const float sl2 = p.scale * FA_LOG2E;
    const uint8_t* gmask = p.mask ? p.mask + int64_t(b) * S : nullptr;
    const int st = threadIdx.x - 64;    // 0..127
    if (gmask) {
      int bad = 0;
      for (int i = st; i < S; i += 128) bad |= (gmask[i] == 0);
      if (bad) atomicOr(&s_any_masked, 1);
      named_bar_sync(1, 128);
    }
    const bool use_mask = gmask && s_any_masked;
    const int P = p.group > 0 ? p.prefix_len[b] : 0x3fffffff;     // keys >= P are visible to their own group only
    int it = 0;
    for (int ti = 0;; ++ti) {
    const int t = my_tile(ti);
    if (t < 0) break;
    const int n_kv = kv_tiles(t);
    const int row = t * FA_BM + r;  // position in the sequence
    // first key of this row's own group (rows of the prefix: 0, nothing is hidden by the group rule)
    const int glo = (p.group > 0 && row >= P) ? P + ((row - P) / p.group) * p.group : 0;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int cur = it + j;
      const int s = cur & 1;
      const uint32_t ts = tmem_base + lane_off + s * FA_BN;
      const int kv0 = j * FA_BN;
      const bool edge = (j >= 2 * t) || (kv0 + FA_BN > S) || use_mask || (kv0 + FA_BN > P);  // element-wise masking
      if (use_mask) {
        named_bar_sync(1, 128);  // previous tile's readers are done with s_mask_tile
        if (st < FA_BN) s_mask_tile[st] = (kv0 + st < S) ? gmask[kv0 + st] : 0;
        named_bar_sync(1, 128);
      }
      mbar_wait(&s_full[s], (cur >> 1) & 1);
      tc_fence_after();
      // ---- this row's 64 scores, kept in registers
      uint32_t v0[32], v1[32];
      tmem_ld_32x32b_x32(ts, v0);
      tmem_ld_32x32b_x32(ts + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);   // S buffer may be overwritten by the QK^T of tile j+2
      float mx = -INFINITY;
      if (!edge) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
      } else if (!use_mask) {
        // causal / sequence-end tile without padding: column kv0 + i is visible iff i <= lim — one compare + select per
        // element, no shared-memory mask, no branches
        const int lim = min(row, S - 1) - kv0;
        if (kv0 + FA_BN <= P) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = (i <= lim) ? v0[i] : 0xff800000u;   // -inf
            v1[i] = (i + 32 <= lim) ? v1[i] : 0xff800000u;
            mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
          }
        } else {
          // tile reaches into the suffix region: columns in [P, glo) belong to other groups
          const int pi = P - kv0, gi = glo - kv0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = ((i <= lim) && (i < pi || i >= gi)) ? v0[i] : 0xff800000u;
            v1[i] = ((i + 32 <= lim) && (i + 32 < pi || i + 32 >= gi)) ? v1[i] : 0xff800000u;
            mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
          }
        }
      } else {
        const int lim = min(row, S - 1) - kv0;
        const int pi = P - kv0, gi = glo - kv0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v0[i] = ((i <= lim) && (i < pi || i >= gi) && s_mask_tile[i]) ? v0[i] : 0xff800000u;
          v1[i] = ((i + 32 <= lim) && (i + 32 < pi || i + 32 >= gi) && s_mask_tile[32 + i]) ? v1[i] : 0xff800000u;
          mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
        }
      }
      // ---- lazy maximum: keep the stale one unless it grew by more than 2^8
      const float m_cand = fmaxf(m_run, mx);
      const bool grow = (m_run == -INFINITY) ? (m_cand != -INFINITY) : ((m_cand - m_run) * sl2 > FA_RESCALE_THRESHOLD);
      const float m_new = grow ? m_cand : m_run;
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float msub = m_use * sl2;
      const float corr = grow ? ex2_approx((m_run - m_use) * sl2) : 1.f;   // m_run = -inf -> 0
      // ---- P = exp2(S*scale - m) in registers first (overlaps the previous tile's PV on the tensor core) ...
      float rowsum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float a0 = ex2_approx(__fmaf_rn(__uint_as_float(v0[i]), sl2, -msub));       // -inf -> 0
        const float a1 = ex2_approx(__fmaf_rn(__uint_as_float(v0[i + 1]), sl2, -msub));
        const float b0 = ex2_approx(__fmaf_rn(__uint_as_float(v1[i]), sl2, -msub));
        const float b1 = ex2_approx(__fmaf_rn(__uint_as_float(v1[i + 1]), sl2, -msub));
        rowsum += (a0 + a1) + (b0 + b1);
        pk[i >> 1] = pack_bf16x2(a0, a1);
        pk[16 + (i >> 1)] = pack_bf16x2(b0, b1);
      }
      l_run = l_run * corr + rowsum;
      m_run = m_new;
      // ... then, once the previous PV has retired (P buffer free, O quiescent), into the swizzled K-major A tile
      if (cur > 0) mbar_wait(pv_done, (cur - 1) & 1);
      uint8_t* prow = sP + r * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) * 16)) = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
      // ---- O rescale in TMEM, only if some row of the warp moved its maximum
      if (j > 0) {
        const bool need = __any_sync(0xffffffffu, grow);
        if (need) {
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_o + lane_off + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
            tmem_st_32x32b_x32(tmem_o + lane_off + c * 32, v);
          }
          tmem_st_wait();
        }
      }
      fence_proxy_async();  // P stores (generic proxy) -> visible to the UMMA (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue of this q tile: wait for its last PV, normalise, store
    mbar_wait(o_full, ti & 1);
    tc_fence_after();
    const bool qvalid = row < S && (!use_mask || gmask[row]);
    const bool live = qvalid && l_run > 0.f;
    const float inv = live ? 1.f / l_run : 0.f;
    __nv_bfloat16* orow = p.o + int64_t(row_base + row) * p.ld_o + hd * FA_D;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_o + lane_off + c * 32, v);
      tmem_ld_wait();
      if (row < S) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = make_uint4(
              pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv),
              pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv),
              pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv),
              pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv));
      }
    }
    if (row < S)
      p.lse[(int64_t(b) * p.H + hd) * S + row] = live ? (m_run * sl2 + log2f(l_run)) * 0.6931471805599453f : INFINITY;
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(o_empty);     // O may be overwritten by the next q tile's first PV
    it += n_kv;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace mla

using namespace mla;

// Forward for head_dim 128 on the tcgen05 path; q/k/v must be the three column blocks of one [B*S, 3*H*128] buffer.
extern "C" int mla_attn_fwd_sm100_grouped(const void* qkv, int64_t ld_qkv, void* o, int64_t ld_o, void* lse,
                                          const void* mask, const void* prefix_len, int32_t group, int32_t batch,
                                          int32_t seq, int32_t heads, float scale, void* stream);

extern "C" int mla_attn_fwd_sm100(const void* qkv, int64_t ld_qkv, void* o, int64_t ld_o, void* lse, const void* mask,
                                  int32_t batch, int32_t seq, int32_t heads, float scale, void* stream) {
  return mla_attn_fwd_sm100_grouped(qkv, ld_qkv, o, ld_o, lse, mask, nullptr, 0, batch, seq, heads, scale, stream);
}

// Shared-prefix variant: prefix_len int32 [batch] (device), group > 0 = rows per suffix group; group = 0 / NULL = causal.
extern "C" int mla_attn_fwd_sm100_grouped(const void* qkv, int64_t ld_qkv, void* o, int64_t ld_o, void* lse,
                                          const void* mask, const void* prefix_len, int32_t group, int32_t batch,
                                          int32_t seq, int32_t heads, float scale, void* stream) {
  if (int rc = device_check()) return rc;
  if (group < 0 || (group > 0 && prefix_len == nullptr))
    return set_error(MLA_ERR_ARG, "attn_fwd_sm100: a shared-prefix layout needs prefix_len and group > 0");
  if (batch <= 0 || seq <= 0 || heads <= 0) return set_error(MLA_ERR_ARG, "attn_fwd_sm100: empty problem");
  if ((ld_qkv & 7) || (ld_o & 7) || (reinterpret_cast<uintptr_t>(qkv) & 15))
    return set_error(MLA_ERR_ARG, "attn_fwd_sm100: pitches must be multiples of 8 elements, base 16-byte aligned");
  CUtensorMap map_q, map_kv;
  const uint64_t dims[2] = {uint64_t(3) * heads * FA_D, uint64_t(batch) * seq};
  const uint64_t strides[1] = {uint64_t(ld_qkv) * 2};
  const uint32_t box_q[2] = {64, FA_BM};
  const uint32_t box_kv[2] = {64, FA_BN};
  if (int rc = encode_tmap_2d_bf16(&map_q, qkv, dims, strides, box_q)) return rc;
  if (int rc = encode_tmap_2d_bf16(&map_kv, qkv, dims, strides, box_kv)) return rc;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "attn_fwd_sm100 smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  FaParams p;
  p.B = batch; p.S = seq; p.H = heads; p.scale = scale;
  p.o = (__nv_bfloat16*)o; p.ld_o = ld_o; p.lse = (float*)lse; p.mask = (const uint8_t*)mask;
  p.prefix_len = group > 0 ? (const int32_t*)prefix_len : nullptr; p.group = group;
  attn_fwd_sm100_kernel<<<batch * heads * 2, FA_THREADS, FA_SMEM, (cudaStream_t)stream>>>(map_q, map_kv, p);
  MLA_CHECK_LAUNCH("attn_fwd_sm100");
  return MLA_OK;
}
