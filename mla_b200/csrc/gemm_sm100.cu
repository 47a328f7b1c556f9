// bf16 GEMM on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), operands fed by TMA.
//
//   C[M,N] = epilogue( alpha * A_op[M,K] . B_op[K,N] )
//
// A is either K-contiguous ([M,K] row-major, "K-major") or M-contiguous ([K,M] row-major, "MN-major");
// B is either K-contiguous ([N,K] row-major, i.e. an nn.Linear weight) or N-contiguous ([K,N] row-major).
// The four combinations cover forward (x.W^T), dgrad (dy.W) and wgrad (dy^T.x) of every linear layer on the
// hot path (reference: transformers/models/llama/modeling_llama.py:240,:435-437,:495 run as cuBLAS calls).
//
// Kernel shape: persistent, one CTA per SM, 192 threads:
//   warp 0    TMA producer      (cp.async.bulk.tensor -> 128B-swizzled smem ring, 4 stages x 48 KB)
//   warp 1    MMA issuer        (one elected lane issues tcgen05.mma 128x256x16; owns the TMEM allocation)
//   warps 2-5 epilogue          (tcgen05.ld 32x32b -> registers -> fused bias/activation/residual -> global)
// Two 128x256 fp32 accumulators (2 x 256 TMEM columns) are double-buffered so the epilogue of tile i overlaps
// the main loop of tile i+1.
#include <cstdlib>
#include "gemm_epilogue.cuh"
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int SCHED_STAGES = 4;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// Tile rasterisation: groups of GROUP_M row-tiles sweep all column-tiles, so the CTAs that run concurrently share
// a small set of A row-panels while B streams through L2.
// The group height is chosen on the host so that one group's A row-panel (group_m x 128 x K bf16) stays L2 resident
// (~48 MB of the 126 MB): B is then re-read from HBM only tiles_m / group_m times.
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int group_m, int& tm, int& tn) {
  int group_size = group_m * tiles_n;
  int g = tile / group_size;
  int first_m = g * group_m;
  int gm = min(group_m, tiles_m - first_m);
  int r = tile - g * group_size;
  tm = first_m + r % gm;
  tn = r / gm;
}

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 int M, int N, int K, int group_m, int* sched, GemmEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_STAGES;
  uint64_t* sched_full = tmem_empty_bar + ACC_STAGES;      // tile-id ring (dynamic scheduling only)
  uint64_t* sched_empty = sched_full + SCHED_STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(sched_empty + SCHED_STAGES);
  int* sched_ids = reinterpret_cast<int*>(tmem_base_slot + 1);
  const bool dynamic = sched != nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);
    }
    for (int s = 0; s < SCHED_STAGES; ++s) {
      mbar_init(&sched_full[s], 1);
      mbar_init(&sched_empty[s], 5);   // MMA thread + 4 epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // Tile sequence: static (blockIdx.x, +gridDim.x, ...) or, with a scheduler counter, claimed dynamically so that
      // CTAs which start late (SMs busy with a concurrent NCCL kernel) simply take fewer tiles.  The producer publishes
      // every claimed tile id to the other warps through a small smem ring.
      int tile = blockIdx.x;
      for (int seq = 0;; ++seq) {
        if (dynamic) {
          const int sl = seq & (SCHED_STAGES - 1);
          mbar_wait(&sched_empty[sl], ((seq / SCHED_STAGES) & 1) ^ 1);
          sched_ids[sl] = tile < num_tiles ? tile : -1;
          mbar_arrive(&sched_full[sl]);
        }
        if (tile >= num_tiles) break;
        int tm, tn;
        tile_coords(tile, tiles_m, tiles_n, group_m, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int k0 = kb * BK;
          if (A_MN == 0) {
            tma_load_2d(sa, &map_a, &full_bar[stage], k0, m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)  // box {64 m, 64 k-rows}, one 8 KB slab per 64 columns of M
              tma_load_2d(sa + j * (BK * 128), &map_a, &full_bar[stage], m0 + j * 64, k0);
          }
          if (B_MN == 0) {
            tma_load_2d(sb, &map_b, &full_bar[stage], k0, n0);  // box {64 k, 256 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (BK * 128), &map_b, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tile = dynamic ? int(gridDim.x) + atomicAdd(sched, 1) : tile + int(gridDim.x);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tile = blockIdx.x;
      for (int seq = 0;; ++seq) {
        if (dynamic) {
          const int sl = seq & (SCHED_STAGES - 1);
          mbar_wait(&sched_full[sl], (seq / SCHED_STAGES) & 1);
          tile = sched_ids[sl];
          mbar_arrive(&sched_empty[sl]);
          if (tile < 0) break;
        } else {
          if (seq > 0) tile += gridDim.x;
          if (tile >= num_tiles) break;
        }
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: rows are 128 B apart, 8-row groups 1024 B apart; a 16-element K step is +32 B.
            // MN-major: k-rows are 128 B apart, 8-k groups 1024 B apart (SBO), 64-element MN slabs BK*128 B apart
            //           (LBO); a 16-element K step is two 8-k groups = +2048 B.
            uint64_t da = A_MN == 0 ? umma_smem_desc_sw128(sa + k * 32, 16, 1024)
                                    : umma_smem_desc_sw128(sa + k * 2048, BK * 128, 1024);
            uint64_t db = B_MN == 0 ? umma_smem_desc_sw128(sb + k * 32, 16, 1024)
                                    : umma_smem_desc_sw128(sb + k * 2048, BK * 128, 1024);
            umma_f16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (4 warps) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    int tile = blockIdx.x;
    for (int seq = 0;; ++seq) {
      if (dynamic) {
        const int sl = seq & (SCHED_STAGES - 1);
        mbar_wait(&sched_full[sl], (seq / SCHED_STAGES) & 1);
        tile = sched_ids[sl];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sched_empty[sl]);
        if (tile < 0) break;
      } else {
        if (seq > 0) tile += gridDim.x;
        if (tile >= num_tiles) break;
      }
      int tm, tn;
      tile_coords(tile, tiles_m, tiles_n, group_m, tm, tn);
      const int64_t row = int64_t(tm) * BM + quarter * 32 + lane;
      const int n0 = tn * BN;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
      gemm_store_tile(ep, taddr, row, n0, M, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (dynamic && threadIdx.x == 0) {
    // the last CTA to finish re-arms the counters for the next launch on this stream
    __threadfence();
    if (atomicAdd(sched + 1, 1) == int(gridDim.x) - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int encode_operand_map(CUtensorMap* map, const void* ptr, int mn_major, int64_t rows_mn, int64_t k,
                              int64_t ld, int tile_mn) {
  // K-major:  global [rows_mn, k] (k contiguous):  dims {k, rows_mn}, box {64, tile_mn}
  // MN-major: global [k, rows_mn] (mn contiguous): dims {rows_mn, k}, box {64, 64}
  uint64_t dims[2];
  uint64_t strides[1] = {uint64_t(ld) * 2};
  uint32_t box[2];
  if (!mn_major) {
    dims[0] = uint64_t(k); dims[1] = uint64_t(rows_mn);
    box[0] = BK; box[1] = uint32_t(tile_mn);
  } else {
    dims[0] = uint64_t(rows_mn); dims[1] = uint64_t(k);
    box[0] = 64; box[1] = BK;
  }
  return encode_tmap_2d_bf16(map, ptr, dims, strides, box);
}

template <int A_MN, int B_MN>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const GemmEpilogue& ep,
                       int* sched, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<A_MN, B_MN>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "cudaFuncSetAttribute(gemm smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  static const int group_override = [] { const char* e = getenv("MLA_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();
  int group_m = int((48ll << 20) / (int64_t(BM) * K * 2));
  group_m = group_m < 4 ? 4 : (group_m > 64 ? 64 : group_m);
  if (group_override > 0) group_m = group_override;   // tuning switch (tools/ab_bench.sh)
  kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(ma, mb, M, N, K, group_m, tiles > grid ? sched : nullptr, ep);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return MLA_OK;
}

int gemm2_dispatch(const mla_gemm_args* g, const GemmEpilogue& ep, cudaStream_t stream);   // gemm2_sm100.cu

// 0 = one CTA per tile (gemm_bf16_kernel), 1 = CTA pairs (gemm2_bf16_kernel) for problems of at least 1024 rows,
// 2 = CTA pairs always (tests).  Default from MLA_GEMM_2CTA, overridable with mla_gemm_set_mode().
static int g_gemm_mode = [] { const char* e = getenv("MLA_GEMM_2CTA"); return e ? atoi(e) : 1; }();

}  // namespace mla

using namespace mla;

namespace mla {
void gemm2_set_group_m(int g);   // gemm2_sm100.cu
}

/* Rasterisation override of the CTA-pair kernel (M-tiles per group; 0 = the built-in heuristic): a tuning switch. */
extern "C" int mla_gemm_set_group_m(int32_t group_m) {
  if (group_m < 0 || group_m > 1024) return set_error(MLA_ERR_ARG, "gemm_set_group_m: out of range");
  gemm2_set_group_m(group_m);
  return MLA_OK;
}

extern "C" int mla_gemm_set_mode(int32_t mode) {
  if (mode < 0 || mode > 2) return set_error(MLA_ERR_ARG, "gemm_set_mode: mode must be 0, 1 or 2");
  g_gemm_mode = mode;
  return MLA_OK;
}

extern "C" int mla_gemm_bf16(const mla_gemm_args* g, void* stream_) {
  if (g == nullptr) return set_error(MLA_ERR_ARG, "gemm: null args");
  if (int rc = device_check()) return rc;
  if (g->m <= 0 || g->n <= 0 || g->k <= 0) return set_error(MLA_ERR_ARG, "gemm: empty problem %lld x %lld x %lld",
                                                           (long long)g->m, (long long)g->n, (long long)g->k);
  if (g->m > INT32_MAX || g->n > INT32_MAX || g->k > INT32_MAX) return set_error(MLA_ERR_ARG, "gemm: dims exceed int32");
  if ((g->lda & 7) || (g->ldb & 7)) return set_error(MLA_ERR_ARG, "gemm: lda/ldb must be multiples of 8 elements (TMA 16-byte stride)");
  if ((reinterpret_cast<uintptr_t>(g->a) & 15) || (reinterpret_cast<uintptr_t>(g->b) & 15) ||
      (reinterpret_cast<uintptr_t>(g->c) & 15))
    return set_error(MLA_ERR_ARG, "gemm: pointers must be 16-byte aligned");
  if (g->c_dtype != 0 && g->c_dtype != 1) return set_error(MLA_ERR_ARG, "gemm: c_dtype must be 0 (bf16) or 1 (fp32)");
  if (g->c_dtype == 1 && (g->bias || g->residual || g->pre_act || g->activation != MLA_ACT_NONE))
    return set_error(MLA_ERR_ARG, "gemm: fp32 output supports only alpha/accumulate");
  if (g->c_dtype == 0 && g->accumulate) return set_error(MLA_ERR_ARG, "gemm: accumulate requires fp32 output");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap ma, mb;
  if (int rc = encode_operand_map(&ma, g->a, g->a_mn_major, g->m, g->k, g->lda, BM)) return rc;
  if (int rc = encode_operand_map(&mb, g->b, g->b_mn_major, g->n, g->k, g->ldb, BN)) return rc;
  GemmEpilogue ep;
  ep.c = g->c; ep.ldc = g->ldc;
  ep.bias = static_cast<const __nv_bfloat16*>(g->bias);
  ep.residual = static_cast<const __nv_bfloat16*>(g->residual); ep.ldr = g->ldr;
  ep.pre_act = static_cast<__nv_bfloat16*>(g->pre_act); ep.ldp = g->ldp;
  ep.alpha = g->alpha; ep.c_dtype = g->c_dtype; ep.accumulate = g->accumulate; ep.activation = g->activation;
  ep.rope_cos = static_cast<const __nv_bfloat16*>(g->rope_cos);
  ep.rope_sin = static_cast<const __nv_bfloat16*>(g->rope_sin);
  ep.rope_seq = g->rope_seq; ep.rope_cols = g->rope_cols;
  ep.rope_pos = static_cast<const int32_t*>(g->rope_pos);
  ep.sumsq = static_cast<float*>(g->sumsq);
  if (g->sumsq != nullptr && g->c_dtype != 1) return set_error(MLA_ERR_ARG, "gemm: sumsq applies to fp32 outputs");
  ep.swiglu_out = static_cast<__nv_bfloat16*>(g->swiglu_out); ep.ld_swiglu = g->ld_swiglu;
  ep.swiglu_f = g->swiglu_out ? int(g->n / 2) : 0;
  ep.sb_gu = static_cast<const __nv_bfloat16*>(g->swiglu_bwd_gu); ep.ld_sb_gu = g->ld_swiglu_bwd_gu;
  ep.sb_dgu = static_cast<__nv_bfloat16*>(g->swiglu_bwd_dgu); ep.ld_sb_dgu = g->ld_swiglu_bwd_dgu;
  ep.sb_act = static_cast<__nv_bfloat16*>(g->swiglu_bwd_act); ep.ld_sb_act = g->ld_swiglu_bwd_act;
  {
    // evict-first output stores for outputs that cannot stay in L2 anyway (>= 64 MB), MLA_GEMM_CS_STORES=1: measured
    // with ncu on the 12 layer GEMMs — DRAM bytes 23.64 vs 23.69 GB, step time unchanged — so the operand re-reads are
    // not caused by the output stream; kept as a switch, off
    static int cs = -1;
    if (cs < 0) {
      const char* e = getenv("MLA_GEMM_CS_STORES");
      cs = (e && e[0] == '1') ? 1 : 0;
    }
    const int64_t out_bytes = g->m * g->n * (g->c_dtype == 1 ? 4 : 2);
    ep.stream_stores = cs && out_bytes >= (int64_t(64) << 20);
    static int hints = -1;
    if (hints < 0) {
      const char* e = getenv("MLA_GEMM_L2_HINTS");
      hints = e ? atoi(e) : 0;
    }
    // operands that fit L2 together need no hint
    ep.l2_hints = ((g->m + g->n) * g->k * 2 >= (int64_t(96) << 20)) ? hints : 0;
  }
  if (g->swiglu_bwd_gu != nullptr) {
    if (g->swiglu_out || g->c_dtype != 0 || g->bias || g->residual || g->pre_act || g->activation != MLA_ACT_NONE ||
        g->rope_cols != 0 || g->accumulate)
      return set_error(MLA_ERR_ARG, "gemm: fused SwiGLU backward applies to a plain bf16 input-gradient GEMM");
    if (!g->swiglu_bwd_dgu || (g->n & 31) || (g->ld_swiglu_bwd_gu & 7) || (g->ld_swiglu_bwd_dgu & 7) ||
        (g->swiglu_bwd_act && (g->ld_swiglu_bwd_act & 7)) || g->ld_swiglu_bwd_gu < 2 * g->n || g->ld_swiglu_bwd_dgu < 2 * g->n ||
        (reinterpret_cast<uintptr_t>(g->swiglu_bwd_gu) & 15) || (reinterpret_cast<uintptr_t>(g->swiglu_bwd_dgu) & 15) ||
        (reinterpret_cast<uintptr_t>(g->swiglu_bwd_act) & 15))
      return set_error(MLA_ERR_ARG, "gemm: fused SwiGLU backward needs N = f a multiple of 32, gate|up and d(gate|up) of "
                                    "[M, 2f] with 16-byte aligned rows");
  }
  if (g->swiglu_out != nullptr) {
    if (g->a_mn_major || g->b_mn_major || g->c_dtype != 0 || g->bias || g->residual || g->pre_act ||
        g->activation != MLA_ACT_NONE || g->rope_cols != 0)
      return set_error(MLA_ERR_ARG, "gemm: fused SwiGLU applies to a plain K-major bf16 gate|up projection");
    if ((g->n & 255) || (g->ld_swiglu & 7) || (g->c && (g->ldc & 7)) || (reinterpret_cast<uintptr_t>(g->swiglu_out) & 15))
      return set_error(MLA_ERR_ARG, "gemm: fused SwiGLU needs N = 2f with f a multiple of 128 and 16-byte aligned rows");
    if (g->m > INT32_MAX) return set_error(MLA_ERR_ARG, "gemm: dims exceed int32");
    // C (the gate|up matrix itself) is optional here: NULL skips its store
    return gemm2_dispatch(g, ep, reinterpret_cast<cudaStream_t>(stream_));
  }
  if (g->rope_cols != 0) {
    if (g->rope_cols < 0 || (g->rope_cols & 255) || g->rope_cols > g->n || g->rope_seq <= 0 || !g->rope_cos || !g->rope_sin)
      return set_error(MLA_ERR_ARG, "gemm: fused RoPE needs rope_cols a multiple of 256 within N, rope_seq > 0 and both tables");
    if (g->c_dtype != 0 || g->bias || g->residual || g->pre_act || g->activation != MLA_ACT_NONE || (g->ldc & 7))
      return set_error(MLA_ERR_ARG, "gemm: fused RoPE applies to a plain bf16 projection (no bias/activation/residual), ldc % 8 == 0");
  }
  const int M = int(g->m), N = int(g->n), K = int(g->k);
  int* sched = static_cast<int*>(g->sched_ws);
  // tall-skinny products (one column tile, short K: the tokenizers' 1.3 M x 96..384 linears) are HBM- and epilogue-bound;
  // the one-CTA kernel's 128-row tiles give twice the tiles in flight (measured 0.24 / 0.37 ms against 0.28 / 0.51 ms)
  const bool tall_skinny = g->n <= BN && g->k <= 512 && !g->swiglu_out && !g->swiglu_bwd_gu;
  if (g_gemm_mode == 2 || (g_gemm_mode == 1 && M >= 1024 && !tall_skinny)) return gemm2_dispatch(g, ep, stream);
  if (!g->a_mn_major && !g->b_mn_major) return launch_gemm<0, 0>(ma, mb, M, N, K, ep, sched, stream);
  if (!g->a_mn_major && g->b_mn_major) return launch_gemm<0, 1>(ma, mb, M, N, K, ep, sched, stream);
  if (g->a_mn_major && !g->b_mn_major) return launch_gemm<1, 0>(ma, mb, M, N, K, ep, sched, stream);
  return launch_gemm<1, 1>(ma, mb, M, N, K, ep, sched, stream);
}
