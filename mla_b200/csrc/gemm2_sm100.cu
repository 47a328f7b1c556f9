// bf16 GEMM on CTA PAIRS (tcgen05.mma.cta_group::2): two SMs of one TPC cooperate on a 256x256 output tile.
//
// Same contract, operand layouts and epilogue as gemm_sm100.cu (C = epilogue(alpha * A_op . B_op)), different machine
// mapping: a cluster of 2 CTAs owns a 256(M) x 256(N) tile; CTA r holds rows [128r, 128r+128) of A and HALF of B
// (columns [128r, 128r+128) of the tile) in its shared memory, the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (M=256, N=256, K=16) which reads A from both CTAs and the two B halves from both CTAs, and
// each CTA's TMEM receives its own 128 rows x 256 columns of the accumulator.  Per CTA and k-block this stages
// 16 KB (A) + 16 KB (half B) instead of 16 + 32 KB, so the L2->smem traffic and the smem operand reads of the tensor
// core drop by a third — on a power-capped part that is clock headroom for the MMAs — and the 32 KB stage allows a
// 6-deep ring instead of 4.
//
//   warp 0    TMA producer   both CTAs: own A tile + own half of B, completion signalled on the LEADER's full barrier
//   warp 1    MMA issuer     leader only; tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs
//   warps 2-5 epilogue       both CTAs: own TMEM -> registers -> global; "accumulator drained" arrives at the leader
#include <cstdlib>

#include "gemm_epilogue.cuh"
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int G2_BM = 128;            // rows per CTA (256 per cluster)
constexpr int G2_BN = 256;            // columns per cluster tile
constexpr int G2_BNH = 128;           // B rows staged per CTA
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;     // 16 KB
constexpr int G2_B_BYTES = G2_BNH * G2_BK * 2;    // 16 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_THREADS = 192;
constexpr int G2_ACC = 2;
constexpr int G2_TMEM_COLS = 512;
constexpr int G2_SCHED = 4;            // depth of the tile-id ring (dynamic scheduling)
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// wait on a LOCAL barrier whose arrivals (and the data they publish) come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load whose completion bytes are counted on an mbarrier that may live in the peer CTA of the pair.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// The same with an L2 eviction-priority hint (createpolicy): within a raster group the A row-panels are re-used by every
// column tile of the sweep (evict_last), each B column-panel only by the few tiles that run next to each other
// (evict_first).
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                      int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issued MMAs retired) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
}

// warp-uniform issue variants (see ptx.cuh: every lane of the converged MMA warp executes them, one elected lane issues)
__device__ __forceinline__ void umma_f16_ss_pair_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
}

__device__ __forceinline__ void tile_coords2(int tile, int tiles_m, int tiles_n, int group_m, int& tm, int& tn) {
  int group_size = group_m * tiles_n;
  int g = tile / group_size;
  int first_m = g * group_m;
  int gm = min(group_m, tiles_m - first_m);
  int r = tile - g * group_size;
  tm = first_m + r % gm;
  tn = r / gm;
}

// SWIGLU = 1 (gate|up projection, K-major operands): tile tn = 128 gate columns + the matching 128 up columns — the
// peer CTA's half of the B panel is fetched from the `up` rows — and the epilogue applies SwiGLU (gemm_epilogue.cuh).
template <int A_MN, int B_MN, int SWIGLU = 0, int UNI = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N,
                  int K, int group_m, int* sched, GemmEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full_bar = empty_bar + G2_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + G2_ACC;
  uint64_t* sched_full = tmem_empty_bar + G2_ACC;       // tile-id ring (dynamic scheduling): filled by the leader's
  uint64_t* sched_empty = sched_full + G2_SCHED;        // producer in BOTH CTAs, drained at the leader
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(sched_empty + G2_SCHED);
  int* sched_ids = reinterpret_cast<int*>(tmem_base_slot + 1);
  const bool dynamic = sched != nullptr;

  const int warp = int(warp_idx_uniform());      // warp-uniform role index: the MMA warp's code stays convergent
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m = (M + 2 * G2_BM - 1) / (2 * G2_BM);
  const int tiles_n = SWIGLU ? ep.swiglu_f / G2_BNH : (N + G2_BN - 1) / G2_BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + G2_BK - 1) / G2_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // leader's producer arrives (+ the bytes of both CTAs' loads)
      mbar_init(&empty_bar[s], 1);   // one multicast commit
    }
    for (int s = 0; s < G2_ACC; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);   // 4 epilogue warps of each CTA arrive at the leader
    }
    for (int s = 0; s < G2_SCHED; ++s) {
      mbar_init(&sched_full[s], 1);
      mbar_init(&sched_empty[s], 10);     // leader: MMA thread + 4 epilogue warps; peer: producer + 4 epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_base_slot, G2_TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // barriers of both CTAs are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
      const uint64_t pol_a = ep.l2_hints == 2 ? pol_first : pol_last, pol_b = ep.l2_hints == 2 ? pol_last : pol_first;
      // Tile sequence: static (cluster_id, +num_clusters, ...) or claimed dynamically by the leader's producer from a
      // global counter (a pair that starts late because an NCCL kernel holds one of its SMs then simply takes fewer
      // tiles) and published to every other role of both CTAs through a small ring of tile ids.
      int tile = cluster_id;
      for (int seq = 0;; ++seq) {
        if (dynamic) {
          const int sl = seq & (G2_SCHED - 1);
          const uint32_t par = (seq / G2_SCHED) & 1;
          if (leader) {
            mbar_wait_cluster(&sched_empty[sl], par ^ 1);
            const int v = tile < num_tiles ? tile : -1;
            sched_ids[sl] = v;
            st_cluster_u32(mapa_u32(smem_u32(&sched_ids[sl]), 1), uint32_t(v));
            mbar_arrive(&sched_full[sl]);
            mbar_arrive_cluster(mapa_u32(smem_u32(&sched_full[sl]), 1));
          } else {
            mbar_wait_cluster(&sched_full[sl], par);
            tile = sched_ids[sl];
            mbar_arrive_cluster(mapa_u32(smem_u32(&sched_empty[sl]), 0));
            if (tile < 0) break;
          }
        }
        if (tile >= num_tiles) break;
        int tm, tn;
        tile_coords2(tile, tiles_m, tiles_n, group_m, tm, tn);
        const int m0 = tm * 2 * G2_BM + int(rank) * G2_BM;
        const int n0 = SWIGLU ? tn * G2_BNH + int(rank) * ep.swiglu_f : tn * G2_BN + int(rank) * G2_BNH;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G2_STAGE_BYTES;
          uint8_t* sb = sa + G2_A_BYTES;
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          const int k0 = kb * G2_BK;
          if (ep.l2_hints) {
            if (A_MN == 0) {
              tma_load_2d_pair_hint(sa, &map_a, full_leader, k0, m0, pol_a);
            } else {
#pragma unroll
              for (int j = 0; j < G2_BM / 64; ++j)
                tma_load_2d_pair_hint(sa + j * (G2_BK * 128), &map_a, full_leader, m0 + j * 64, k0, pol_a);
            }
            if (B_MN == 0) {
              tma_load_2d_pair_hint(sb, &map_b, full_leader, k0, n0, pol_b);
            } else {
#pragma unroll
              for (int j = 0; j < G2_BNH / 64; ++j)
                tma_load_2d_pair_hint(sb + j * (G2_BK * 128), &map_b, full_leader, n0 + j * 64, k0, pol_b);
            }
          } else {
          if (A_MN == 0) {
            tma_load_2d_pair(sa, &map_a, full_leader, k0, m0);              // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < G2_BM / 64; ++j)
              tma_load_2d_pair(sa + j * (G2_BK * 128), &map_a, full_leader, m0 + j * 64, k0);
          }
          if (B_MN == 0) {
            tma_load_2d_pair(sb, &map_b, full_leader, k0, n0);              // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < G2_BNH / 64; ++j)
              tma_load_2d_pair(sb + j * (G2_BK * 128), &map_b, full_leader, n0 + j * 64, k0);
          }
          }
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        if (!dynamic) tile += num_clusters;
        else if (leader) tile = num_clusters + atomicAdd(sched, 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if constexpr (UNI) {
    // All 32 lanes of the warp run the loop (warp-uniform values, barrier polls warp-wide); tcgen05.mma / commit elect
    // their issuing lane themselves, so the operands sit in uniform registers and the ~20-instruction ELECT / R2UR /
    // BRA.U.ANY sequence a divergent `lane == 0` region costs per MMA (about as long as the 128-cycle MMA itself) is gone.
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * G2_BM, G2_BN, A_MN, B_MN);
      const uint64_t dA0 = A_MN == 0 ? umma_smem_desc_sw128(smem_u32(smem), 16, 1024)
                                     : umma_smem_desc_sw128(smem_u32(smem), G2_BK * 128, 1024);
      const uint64_t dB0 = B_MN == 0 ? umma_smem_desc_sw128(smem_u32(smem) + G2_A_BYTES, 16, 1024)
                                     : umma_smem_desc_sw128(smem_u32(smem) + G2_A_BYTES, G2_BK * 128, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tile = cluster_id;
      for (int seq = 0;; ++seq) {
        if (dynamic) {
          const int sl = seq & (G2_SCHED - 1);
          mbar_wait(&sched_full[sl], (seq / G2_SCHED) & 1);
          tile = __shfl_sync(0xffffffffu, sched_ids[sl], 0);      // same value in every lane; tells the compiler so
          __syncwarp();
          if (lane == 0) mbar_arrive(&sched_empty[sl]);
          if (tile < 0) break;
        } else {
          if (seq > 0) tile += num_clusters;
          if (tile >= num_tiles) break;
        }
        mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * G2_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t dA = umma_desc_advance(dA0, stage * G2_STAGE_BYTES), dB = umma_desc_advance(dB0, stage * G2_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k) {
            umma_f16_ss_pair_elect(tmem_d, umma_desc_advance(dA, A_MN == 0 ? k * 32 : k * 2048),
                                   umma_desc_advance(dB, B_MN == 0 ? k * 32 : k * 2048), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair_elect(&empty_bar[stage]);      // frees this slot in both CTAs
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair_elect(&tmem_full_bar[acc]);      // accumulator complete -> both epilogues
        if (++acc == G2_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
    } else {
    // first-generation issue path (one lane in a divergent region), kept for A/B measurements (MLA_GEMM_UNIFORM_ISSUE=0)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * G2_BM, G2_BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tile = cluster_id;
      for (int seq = 0;; ++seq) {
        if (dynamic) {
          const int sl = seq & (G2_SCHED - 1);
          mbar_wait(&sched_full[sl], (seq / G2_SCHED) & 1);
          tile = sched_ids[sl];
          mbar_arrive(&sched_empty[sl]);
          if (tile < 0) break;
        } else {
          if (seq > 0) tile += num_clusters;
          if (tile >= num_tiles) break;
        }
        mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * G2_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint32_t sb = sa + G2_A_BYTES;
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k) {
            uint64_t da = A_MN == 0 ? umma_smem_desc_sw128(sa + k * 32, 16, 1024)
                                    : umma_smem_desc_sw128(sa + k * 2048, G2_BK * 128, 1024);
            uint64_t db = B_MN == 0 ? umma_smem_desc_sw128(sb + k * 32, 16, 1024)
                                    : umma_smem_desc_sw128(sb + k * 2048, G2_BK * 128, 1024);
            umma_f16_ss_pair(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);      // frees this slot in both CTAs
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full_bar[acc]);      // accumulator complete -> both epilogues
        if (++acc == G2_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
    }
  } else {
    // ===================== epilogue (4 warps, both CTAs) =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tile = cluster_id;
    for (int seq = 0;; ++seq) {
      if (dynamic) {
        const int sl = seq & (G2_SCHED - 1);
        mbar_wait_cluster(&sched_full[sl], (seq / G2_SCHED) & 1);
        tile = sched_ids[sl];
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sched_empty[sl]), 0));
        if (tile < 0) break;
      } else {
        if (seq > 0) tile += num_clusters;
        if (tile >= num_tiles) break;
      }
      int tm, tn;
      tile_coords2(tile, tiles_m, tiles_n, group_m, tm, tn);
      const int64_t row = int64_t(tm) * 2 * G2_BM + int64_t(rank) * G2_BM + quarter * 32 + lane;
      const int n0 = tn * G2_BN;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * G2_BN;
      if constexpr (SWIGLU) gemm_store_tile_swiglu(ep, taddr, row, tn, M);
      else gemm_store_tile(ep, taddr, row, n0, M, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
      if (++acc == G2_ACC) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, G2_TMEM_COLS);
  }
  if (dynamic && leader && threadIdx.x == 0) {
    // the last cluster to finish re-arms the counters for the next launch on this stream
    __threadfence();
    if (atomicAdd(sched + 1, 1) == num_clusters - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

static int g_group_m_override = [] { const char* e = getenv("MLA_GEMM2_GROUP_M"); return e ? atoi(e) : 0; }();
void gemm2_set_group_m(int g) { g_group_m_override = g; }

static int encode_operand_map2(CUtensorMap* map, const void* ptr, int mn_major, int64_t rows_mn, int64_t k, int64_t ld) {
  uint64_t dims[2];
  uint64_t strides[1] = {uint64_t(ld) * 2};
  uint32_t box[2];
  if (!mn_major) {
    dims[0] = uint64_t(k); dims[1] = uint64_t(rows_mn);
    box[0] = G2_BK; box[1] = 128;
  } else {
    dims[0] = uint64_t(rows_mn); dims[1] = uint64_t(k);
    box[0] = 64; box[1] = G2_BK;
  }
  return encode_tmap_2d_bf16(map, ptr, dims, strides, box);
}

template <int A_MN, int B_MN, int SWIGLU = 0, int UNI = 1>
static int launch_gemm2(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const GemmEpilogue& ep,
                        int* sched, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = gemm2_bf16_kernel<A_MN, B_MN, SWIGLU, UNI>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "cudaFuncSetAttribute(gemm2 smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ((M + 2 * G2_BM - 1) / (2 * G2_BM)) * (SWIGLU ? ep.swiglu_f / G2_BNH : (N + G2_BN - 1) / G2_BN);
  const int max_clusters = num_sms() / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  // M-tiles per rasterisation group: ~64 MB of A panels (256 rows x K) per group, between 6 and 24 — from the sweep of
  // tools/bench_gemm_raster.py (profiles/r02_gemm_raster.json): K = 4096 is flat from 12 to 24, the K >= 16 K shapes
  // (weight gradients, gate|up dgrad) peak at 6-8 and lose 10 % at 24
  int group_m = int((64ll << 20) / (int64_t(2 * G2_BM) * K * 2));
  group_m = group_m < 6 ? 6 : (group_m > 24 ? 24 : group_m);
  if (g_group_m_override > 0) group_m = g_group_m_override;      // tuning switch (tools/bench_gemm_raster.py)
  kern<<<2 * clusters, G2_THREADS, G2_SMEM_BYTES, stream>>>(ma, mb, M, N, K, group_m, tiles > clusters ? sched : nullptr,
                                                            ep);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "gemm2 launch: %s", cudaGetErrorString(e));
  count_launch();
  return MLA_OK;
}

// Called by mla_gemm_bf16 (gemm_sm100.cu) when the pair kernel is selected; arguments are already validated.
int gemm2_dispatch(const mla_gemm_args* g, const GemmEpilogue& ep, cudaStream_t stream) {
  CUtensorMap ma, mb;
  if (int rc = encode_operand_map2(&ma, g->a, g->a_mn_major, g->m, g->k, g->lda)) return rc;
  if (int rc = encode_operand_map2(&mb, g->b, g->b_mn_major, g->n, g->k, g->ldb)) return rc;
  const int M = int(g->m), N = int(g->n), K = int(g->k);
  int* sched = static_cast<int*>(g->sched_ws);
  // MLA_GEMM_UNIFORM_ISSUE=0: the first-generation MMA issue path (A/B switch; default: warp-uniform issue)
  static const bool uni = [] { const char* e = getenv("MLA_GEMM_UNIFORM_ISSUE"); return e ? atoi(e) != 0 : true; }();
  if (!uni) {
    if (ep.swiglu_f > 0) return launch_gemm2<0, 0, 1, 0>(ma, mb, M, N, K, ep, sched, stream);
    if (!g->a_mn_major && !g->b_mn_major) return launch_gemm2<0, 0, 0, 0>(ma, mb, M, N, K, ep, sched, stream);
    if (!g->a_mn_major && g->b_mn_major) return launch_gemm2<0, 1, 0, 0>(ma, mb, M, N, K, ep, sched, stream);
    if (g->a_mn_major && !g->b_mn_major) return launch_gemm2<1, 0, 0, 0>(ma, mb, M, N, K, ep, sched, stream);
    return launch_gemm2<1, 1, 0, 0>(ma, mb, M, N, K, ep, sched, stream);
  }
  if (ep.swiglu_f > 0) return launch_gemm2<0, 0, 1>(ma, mb, M, N, K, ep, sched, stream);   // validated by the caller
  if (!g->a_mn_major && !g->b_mn_major) return launch_gemm2<0, 0>(ma, mb, M, N, K, ep, sched, stream);
  if (!g->a_mn_major && g->b_mn_major) return launch_gemm2<0, 1>(ma, mb, M, N, K, ep, sched, stream);
  if (g->a_mn_major && !g->b_mn_major) return launch_gemm2<1, 0>(ma, mb, M, N, K, ep, sched, stream);
  return launch_gemm2<1, 1>(ma, mb, M, N, K, ep, sched, stream);
}

}  // namespace mla
