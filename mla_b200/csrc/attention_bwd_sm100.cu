// Causal flash attention BACKWARD on tcgen05 tensor cores (head_dim 128).  Two kernels share one template:
//   MODE 0  dK/dV : a CTA owns 128 key rows (K_j, V_j resident in smem) and streams 64-row Q_i / dO_i tiles:
//                   S^T = K_j.Q_i^T, dP^T = V_j.dO_i^T  (TMEM)  ->  P^T = exp2(S^T*c - lse_i), dS^T = P^T o (dP^T - delta_i)
//                   (registers -> bf16 -> swizzled smem)  ->  dV += P^T.dO_i,  dK += dS^T.Q_i  (TMEM accumulators)
//   MODE 1  dQ    : a CTA owns 128 query rows (Q_t, dO_t resident) and streams 64-row K_j / V_j tiles:
//                   S = Q_t.K_j^T, dP = dO_t.V_j^T  ->  dS = P o (dP - delta_t)  ->  dQ += dS.K_j
// The transposed formulation of MODE 0 makes P^T / dS^T come out with the key row on the TMEM lane, i.e. directly as
// the A operand of the two accumulation MMAs; the streamed tiles are consumed twice from the same smem bytes, once
// K-major (as B of the score products) and once MN-major (as B of the accumulations).  Recomputing S in both kernels
// costs 7 GEMM-units instead of 5 but needs no atomics and is deterministic.  Semantics and rounding points as
// attention.cu (P and dS rounded to bf16 before the second product; masked / padded query rows carry lse = +inf and
// therefore contribute nothing).
//
// Per CTA (1 per SM, 320 threads): warp 0 TMA loader, warp 1 MMA issuer, warps 2-9 element-wise math (two threads
// per TMEM lane, 32 score columns each; the backward needs no row reductions).  S/dP are double-buffered in TMEM so
// the score products of tile i+1 run while tile i is in the math warps; two CTAs per (batch, head) take that
// sequence's outer tiles in a zig-zag and walk them persistently.
// TMEM (512 cols): [S0|dP0|S1|dP1] 4 x 64, accumulators @256 (dV or dQ) and @384 (dK).
// smem: fixed 2x32 KB | streamed 2 slots x (16+16) KB | P^T, dS^T 2x16 KB | barriers | per-column lse/delta | mask.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int BW_D = 128, BW_OUT = 128, BW_IN = 64;
constexpr int BW_FIX_BYTES = BW_OUT * BW_D * 2;     // 32 KB
constexpr int BW_STR_BYTES = BW_IN * BW_D * 2;      // 16 KB
constexpr int BW_T_BYTES = BW_OUT * BW_IN * 2;      // 16 KB
constexpr int BW_THREADS = 320;
constexpr int BW_MATH_WARPS = 8;
constexpr int BW_TILES_BYTES = 2 * BW_FIX_BYTES + 4 * BW_STR_BYTES + 2 * BW_T_BYTES;   // 160 KB
constexpr int BW_SMEM = BW_TILES_BYTES + 256 + 1024 + 128 + 16;
constexpr float BW_LOG2E = 1.4426950408889634f;

struct BwParams {
  int B, S, H, S_pad;
  float scale;
  const float* lse2;     // [B,H,S_pad]  lse * log2(e); +inf for rows that take no part
  const float* delta;    // [B,H,S_pad]
  const uint8_t* mask;   // [B,S] or null
  __nv_bfloat16* dqkv;   // [B*S, 3*H*D]
  int64_t ld_dqkv;
};

__device__ __forceinline__ void bw_named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float bw_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// lse2 = lse*log2(e) and delta = sum_d dO*O, both padded to S_pad per (b,h).  Half a warp per (b,h,s) row (16 lanes x
// 16-byte loads = the row's 256 bytes of O and of dO), four rows per half-warp with all eight loads issued before the
// first reduction: 128 bytes in flight per thread instead of 16 (the first version ran at 1.8 TB/s).
constexpr int PREP_ROWS = 4;
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int64_t ld_o,
                     const float* __restrict__ lse, float* __restrict__ lse2, float* __restrict__ delta, int B, int S,
                     int H, int S_pad) {
  // 32-bit index arithmetic throughout (B*H*S_pad < 2^31 is checked by the launcher): 64-bit div/mod are emulated in
  // ~100 instructions each and made the first version of this kernel instruction-bound
  const uint32_t hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;     // half-warp index
  const int l16 = threadIdx.x & 15;
  const uint32_t total = uint32_t(B) * H * S_pad;
  uint4 a[PREP_ROWS], c[PREP_ROWS];
  bool live[PREP_ROWS];
  const uint32_t w0 = hw * PREP_ROWS;
  // PREP_ROWS divides S_pad (a multiple of 128): the rows of one half-warp share (b, h)
  const uint32_t bh = w0 / uint32_t(S_pad), s0 = w0 - bh * uint32_t(S_pad);
  const uint32_t bb = bh / uint32_t(H), hh = bh - bb * uint32_t(H);
  const int64_t base = (int64_t(bb) * S + s0) * ld_o + int64_t(hh) * BW_D + l16 * 8;
#pragma unroll
  for (int u = 0; u < PREP_ROWS; ++u) {
    a[u] = c[u] = make_uint4(0, 0, 0, 0);
    live[u] = w0 + u < total && int(s0) + u < S;
    if (live[u]) {
      a[u] = *reinterpret_cast<const uint4*>(o + base + u * ld_o);
      c[u] = *reinterpret_cast<const uint4*>(d_o + base + u * ld_o);
    }
  }
#pragma unroll
  for (int u = 0; u < PREP_ROWS; ++u) {
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a[u]);
    const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&c[u]);
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = __bfloat1622float2(pa[t]), y = __bfloat1622float2(pc[t]);
      acc += x.x * y.x + x.y * y.y;
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (l16 == 0 && w0 + u < total) {
      lse2[w0 + u] = live[u] ? lse[bh * uint32_t(S) + s0 + u] * BW_LOG2E : INFINITY;
      delta[w0 + u] = acc;
    }
  }
}

// Host launcher of the prep kernel, shared with the pipelined generation (attention_bwd2_sm100.cu).
int attn_bwd_prep_launch(const void* o, const void* d_o, int64_t ld_o, const void* lse, float* lse2, float* delta, int B,
                         int S, int H, int S_pad, cudaStream_t s) {
  const int64_t rows = int64_t(B) * H * S_pad;
  if (rows >= (int64_t(1) << 31)) return set_error(MLA_ERR_ARG, "attn_bwd_prep: batch * heads * seq exceeds 2^31 rows");
  const int64_t threads = (rows + PREP_ROWS - 1) / PREP_ROWS * 16;
  attn_bwd_prep_kernel<<<unsigned((threads + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o,
                                                                        ld_o, (const float*)lse, lse2, delta, B, S, H, S_pad);
  MLA_CHECK_LAUNCH("attn_bwd_prep");
  return MLA_OK;
}

template <int MODE>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_sm100_kernel(const __grid_constant__ CUtensorMap map_qkv_fix, const __grid_constant__ CUtensorMap map_qkv_str,
                      const __grid_constant__ CUtensorMap map_do_fix, const __grid_constant__ CUtensorMap map_do_str,
                      BwParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sF1 = smem;                               // MODE 0: K_j   MODE 1: Q_t
  uint8_t* sF2 = smem + BW_FIX_BYTES;                // MODE 0: V_j   MODE 1: dO_t
  uint8_t* sX = smem + 2 * BW_FIX_BYTES;             // 2 slots  MODE 0: Q_i   MODE 1: K_j
  uint8_t* sY = sX + 2 * BW_STR_BYTES;               // 2 slots  MODE 0: dO_i  MODE 1: V_j
  uint8_t* sT1 = sY + 2 * BW_STR_BYTES;              // P^T (MODE 0 only)
  uint8_t* sT2 = sT1 + BW_T_BYTES;                   // dS^T / dS
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_TILES_BYTES);
  uint64_t *fix_full = bars, *fix_empty = bars + 1, *in_full = bars + 2, *in_empty = bars + 4, *sd_full = bars + 6,
           *sd_empty = bars + 8, *ds_full = bars + 10, *acc_done = bars + 11, *acc_full = bars + 12,
           *acc_empty = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* s_stat = reinterpret_cast<float*>(smem + BW_TILES_BYTES + 256);   // [slot][lse2|delta][64]
  uint8_t* s_mask_tile = smem + BW_TILES_BYTES + 256 + 1024;
  int& s_any_masked = *reinterpret_cast<int*>(smem + BW_TILES_BYTES + 256 + 1024 + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const int n_out = (S + BW_OUT - 1) / BW_OUT;
  const int n_in = (S + BW_IN - 1) / BW_IN;
  const int half_id = blockIdx.x & 1;
  const int bh = blockIdx.x >> 1;
  const int b = bh / p.H, hd = bh % p.H;
  const int HD = p.H * BW_D;
  const int row_base = b * S;
  // outer tiles in order of decreasing work are dealt A B B A A B B A ...
  auto my_tile = [&](int k) -> int {
    const int pos = half_id == 0 ? (k == 0 ? 0 : 4 * ((k + 1) >> 1) - ((k & 1) ? 1 : 0)) : (4 * (k >> 1) + 1 + (k & 1));
    if (pos >= n_out) return -1;
    return MODE == 0 ? pos : n_out - 1 - pos;      // dK/dV: key tile 0 sees every query; dQ: the last query tile sees every key
  };
  auto in_begin = [&](int t) { return MODE == 0 ? 2 * t : 0; };
  auto in_end = [&](int t) { return MODE == 0 ? n_in : min(2 * t + 2, n_in); };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_qkv_fix); tma_prefetch_desc(&map_qkv_str);
    tma_prefetch_desc(&map_do_fix); tma_prefetch_desc(&map_do_str);
    mbar_init(fix_full, 1); mbar_init(fix_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1);
      mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], BW_MATH_WARPS);
    }
    mbar_init(ds_full, BW_MATH_WARPS); mbar_init(acc_done, 1); mbar_init(acc_full, 1); mbar_init(acc_empty, BW_MATH_WARPS);
    s_any_masked = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc0 = tmem_base + 256, tmem_acc1 = tmem_base + 384;

  // column blocks of the fused [q | k | v] buffer
  const int col_q = hd * BW_D, col_k = HD + hd * BW_D, col_v = 2 * HD + hd * BW_D, col_do = hd * BW_D;

  if (warp == 0) {
    // ================================ TMA loader ================================
    if (lane == 0) {
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int t = my_tile(ti);
        if (t < 0) break;
        mbar_wait(fix_empty, (ti & 1) ^ 1);
        mbar_arrive_expect_tx(fix_full, 2 * BW_FIX_BYTES);
        const int r0 = row_base + t * BW_OUT;
        if (MODE == 0) {
          tma_load_2d(sF1, &map_qkv_fix, fix_full, col_k, r0);
          tma_load_2d(sF1 + BW_FIX_BYTES / 2, &map_qkv_fix, fix_full, col_k + 64, r0);
          tma_load_2d(sF2, &map_qkv_fix, fix_full, col_v, r0);
          tma_load_2d(sF2 + BW_FIX_BYTES / 2, &map_qkv_fix, fix_full, col_v + 64, r0);
        } else {
          tma_load_2d(sF1, &map_qkv_fix, fix_full, col_q, r0);
          tma_load_2d(sF1 + BW_FIX_BYTES / 2, &map_qkv_fix, fix_full, col_q + 64, r0);
          tma_load_2d(sF2, &map_do_fix, fix_full, col_do, r0);
          tma_load_2d(sF2 + BW_FIX_BYTES / 2, &map_do_fix, fix_full, col_do + 64, r0);
        }
        for (int i = in_begin(t); i < in_end(t); ++i, ++it) {
          const int s = it & 1;
          mbar_wait(&in_empty[s], ((it >> 1) & 1) ^ 1);
          uint8_t* dx = sX + s * BW_STR_BYTES;
          uint8_t* dy = sY + s * BW_STR_BYTES;
          const int c0 = row_base + i * BW_IN;
          if (MODE == 0) {
            mbar_arrive_expect_tx(&in_full[s], 2 * BW_STR_BYTES + 512);
            tma_load_2d(dx, &map_qkv_str, &in_full[s], col_q, c0);
            tma_load_2d(dx + BW_STR_BYTES / 2, &map_qkv_str, &in_full[s], col_q + 64, c0);
            tma_load_2d(dy, &map_do_str, &in_full[s], col_do, c0);
            tma_load_2d(dy + BW_STR_BYTES / 2, &map_do_str, &in_full[s], col_do + 64, c0);
            const int64_t so = int64_t(bh) * p.S_pad + i * BW_IN;
            bulk_copy_g2s(s_stat + s * 128, p.lse2 + so, 256, &in_full[s]);
            bulk_copy_g2s(s_stat + s * 128 + 64, p.delta + so, 256, &in_full[s]);
          } else {
            mbar_arrive_expect_tx(&in_full[s], 2 * BW_STR_BYTES);
            tma_load_2d(dx, &map_qkv_str, &in_full[s], col_k, c0);
            tma_load_2d(dx + BW_STR_BYTES / 2, &map_qkv_str, &in_full[s], col_k + 64, c0);
            tma_load_2d(dy, &map_qkv_str, &in_full[s], col_v, c0);
            tma_load_2d(dy + BW_STR_BYTES / 2, &map_qkv_str, &in_full[s], col_v + 64, c0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc_sd = umma_idesc_bf16(BW_OUT, BW_IN, 0, 0);    // 128 x 64,  K = d
      constexpr uint32_t idesc_acc = umma_idesc_bf16(BW_OUT, BW_D, 0, 1);    // 128 x 128, K = 64 streamed rows, B MN-major
      const uint32_t aF1 = smem_u32(sF1), aF2 = smem_u32(sF2), aT1 = smem_u32(sT1), aT2 = smem_u32(sT2);
      auto issue_sd = [&](int cur) {
        const int s = cur & 1;
        const uint32_t ph = (cur >> 1) & 1;
        mbar_wait(&in_full[s], ph);
        mbar_wait(&sd_empty[s], ph ^ 1);
        tc_fence_after();
        const uint32_t aX = smem_u32(sX + s * BW_STR_BYTES), aY = smem_u32(sY + s * BW_STR_BYTES);
        const uint32_t tS = tmem_base + s * 128, tP = tS + 64;
#pragma unroll
        for (int kk = 0; kk < BW_D / 16; ++kk) {
          const uint32_t of = (kk >> 2) * (BW_FIX_BYTES / 2) + (kk & 3) * 32;
          const uint32_t os = (kk >> 2) * (BW_STR_BYTES / 2) + (kk & 3) * 32;
          umma_f16_ss(tS, umma_smem_desc_sw128(aF1 + of, 16, 1024), umma_smem_desc_sw128(aX + os, 16, 1024), idesc_sd,
                      kk != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < BW_D / 16; ++kk) {
          const uint32_t of = (kk >> 2) * (BW_FIX_BYTES / 2) + (kk & 3) * 32;
          const uint32_t os = (kk >> 2) * (BW_STR_BYTES / 2) + (kk & 3) * 32;
          umma_f16_ss(tP, umma_smem_desc_sw128(aF2 + of, 16, 1024), umma_smem_desc_sw128(aY + os, 16, 1024), idesc_sd,
                      kk != 0 ? 1u : 0u);
        }
        umma_commit(&sd_full[s]);
      };
      int it = 0;
      for (int ti = 0;; ++ti) {
        const int t = my_tile(ti);
        if (t < 0) break;
        const int n = in_end(t) - in_begin(t);
        mbar_wait(fix_full, ti & 1);
        issue_sd(it);
        for (int j = 0; j < n; ++j) {
          const int cur = it + j;
          if (j + 1 < n) issue_sd(cur + 1);
          else umma_commit(fix_empty);                       // every score product of this outer tile has been issued
          const int s = cur & 1;
          mbar_wait(ds_full, cur & 1);
          if (j == 0) mbar_wait(acc_empty, (ti & 1) ^ 1);    // previous outer tile's accumulators were read out
          tc_fence_after();
          const uint32_t aX = smem_u32(sX + s * BW_STR_BYTES), aY = smem_u32(sY + s * BW_STR_BYTES);
#pragma unroll
          for (int kk = 0; kk < BW_IN / 16; ++kk) {
            const uint32_t acc = (j | kk) != 0 ? 1u : 0u;
            if (MODE == 0) {
              umma_f16_ss(tmem_acc0, umma_smem_desc_sw128(aT1 + kk * 32, 16, 1024),
                          umma_smem_desc_sw128(aY + kk * 2048, BW_STR_BYTES / 2, 1024), idesc_acc, acc);   // dV += P^T.dO
              umma_f16_ss(tmem_acc1, umma_smem_desc_sw128(aT2 + kk * 32, 16, 1024),
                          umma_smem_desc_sw128(aX + kk * 2048, BW_STR_BYTES / 2, 1024), idesc_acc, acc);   // dK += dS^T.Q
            } else {
              umma_f16_ss(tmem_acc0, umma_smem_desc_sw128(aT2 + kk * 32, 16, 1024),
                          umma_smem_desc_sw128(aX + kk * 2048, BW_STR_BYTES / 2, 1024), idesc_acc, acc);   // dQ += dS.K
            }
          }
          umma_commit(acc_done);
          umma_commit(&in_empty[s]);
          if (j == n - 1) umma_commit(acc_full);
        }
        it += n;
      }
    }
  } else {
    // ================================ element-wise math (256 threads, two per TMEM lane) ================================
    const int mw = warp - 2;
    const int quarter = warp & 3;
    const int half = mw >> 2;                 // which 32 of the 64 streamed columns
    const int r = quarter * 32 + lane;        // row of the outer tile == TMEM lane
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const float sl2 = p.scale * BW_LOG2E;
    const uint8_t* gmask = p.mask ? p.mask + int64_t(b) * S : nullptr;
    const int st = threadIdx.x - 64;          // 0..255
    if (gmask) {
      int bad = 0;
      for (int i = st; i < S; i += 256) bad |= (gmask[i] == 0);
      if (bad) atomicOr(&s_any_masked, 1);
      bw_named_bar_sync(1, 256);
    }
    const bool use_mask = gmask && s_any_masked;
    int it = 0;
    for (int ti = 0;; ++ti) {
      const int t = my_tile(ti);
      if (t < 0) break;
      const int i0 = in_begin(t), n = in_end(t) - i0;
      const int orow = t * BW_OUT + r;        // key row (MODE 0) / query row (MODE 1) in the sequence
      float row_lse2 = 0.f, row_delta = 0.f;
      bool row_ok = true;
      if (MODE == 1) {
        row_lse2 = p.lse2[int64_t(bh) * p.S_pad + orow];      // S_pad covers every row of every outer tile
        row_delta = p.delta[int64_t(bh) * p.S_pad + orow];
      } else {
        row_ok = orow < S && (!use_mask || gmask[orow]);      // masked / out-of-range keys get P = 0
      }
      for (int j = 0; j < n; ++j) {
        const int cur = it + j;
        const int s = cur & 1;
        const int in0 = (i0 + j) * BW_IN + half * 32;   // first streamed column (sequence position) of this thread
        const uint32_t tS = tmem_base + lane_off + s * 128 + half * 32;
        bool edge;
        if (MODE == 0) edge = ((i0 + j) * BW_IN < t * BW_OUT + BW_OUT);                       // diagonal: query < key possible
        else edge = ((i0 + j) * BW_IN + BW_IN - 1 > t * BW_OUT) || ((i0 + j + 1) * BW_IN > S) || use_mask;
        if (MODE == 1 && use_mask) {
          bw_named_bar_sync(1, 256);
          if (st < BW_IN) s_mask_tile[st] = ((i0 + j) * BW_IN + st < S) ? gmask[(i0 + j) * BW_IN + st] : 0;
          bw_named_bar_sync(1, 256);
        }
        mbar_wait(&sd_full[s], (cur >> 1) & 1);
        tc_fence_after();
        uint32_t vs[32], vp[32];
        tmem_ld_32x32b_x32(tS, vs);
        tmem_ld_32x32b_x32(tS + 64, vp);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sd_empty[s]);
        const float* stat = s_stat + s * 128 + half * 32;     // MODE 0: per-column lse2 | delta of this q tile
        uint32_t pk_p[16], pk_d[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pr[2], ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = i + e;
            const float l2 = MODE == 0 ? stat[c] : row_lse2;
            const float dl = MODE == 0 ? stat[64 + c] : row_delta;
            bool vis = true;
            if (MODE == 0) {
              vis = row_ok && (!edge || orow <= in0 + c);
            } else if (edge) {
              const int col = in0 + c;
              vis = (col <= orow) && (col < S) && (!use_mask || s_mask_tile[half * 32 + c]);
            }
            const float pv = vis ? bw_ex2(__fmaf_rn(__uint_as_float(vs[c]), sl2, -l2)) : 0.f;   // lse2 = +inf -> 0
            pr[e] = pv;
            ds[e] = pv * (__uint_as_float(vp[c]) - dl);
          }
          pk_p[i >> 1] = pack_bf16x2(pr[0], pr[1]);
          pk_d[i >> 1] = pack_bf16x2(ds[0], ds[1]);
        }
        // the accumulation MMAs of the previous tile have retired: P^T / dS^T buffers are free again
        if (cur > 0) mbar_wait(acc_done, (cur - 1) & 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int off = r * 128 + (((half * 4 + g) ^ (r & 7)) * 16);
          if (MODE == 0)
            *reinterpret_cast<uint4*>(sT1 + off) = make_uint4(pk_p[g * 4], pk_p[g * 4 + 1], pk_p[g * 4 + 2], pk_p[g * 4 + 3]);
          *reinterpret_cast<uint4*>(sT2 + off) = make_uint4(pk_d[g * 4], pk_d[g * 4 + 1], pk_d[g * 4 + 2], pk_d[g * 4 + 3]);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
      }
      // ---- epilogue of this outer tile: accumulators -> bf16 -> global (this thread: columns half*64 .. +63)
      mbar_wait(acc_full, ti & 1);
      tc_fence_after();
      const bool store = orow < S;
      __nv_bfloat16* base = p.dqkv + int64_t(row_base + orow) * p.ld_dqkv + half * 64;
#pragma unroll 1
      for (int a = 0; a < (MODE == 0 ? 2 : 1); ++a) {
        const uint32_t tacc = (a == 0 ? tmem_acc0 : tmem_acc1) + lane_off + half * 64;
        // MODE 0: acc0 = dV (v block), acc1 = dK (k block, scaled); MODE 1: acc0 = dQ (q block, scaled)
        const int col = MODE == 0 ? (a == 0 ? col_v : col_k) : col_q;
        const float sc = (MODE == 0 && a == 0) ? 1.f : p.scale;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tacc + c * 32, v);
          tmem_ld_wait();
          if (store) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(base + col + c * 32 + g * 8) = make_uint4(
                  pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * sc, __uint_as_float(v[g * 8 + 1]) * sc),
                  pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * sc, __uint_as_float(v[g * 8 + 3]) * sc),
                  pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * sc, __uint_as_float(v[g * 8 + 5]) * sc),
                  pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * sc, __uint_as_float(v[g * 8 + 7]) * sc));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      it += n;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mla

using namespace mla;

extern "C" size_t mla_attn_bwd_sm100_workspace(int32_t batch, int32_t seq, int32_t heads) {
  const size_t s_pad = size_t((seq + 127) / 128) * 128;
  // lse2 | delta ([batch, heads, s_pad] f32 each) | per-batch "has padding" flags (int32 [batch], pipelined generation)
  return 2 * sizeof(float) * size_t(batch) * heads * s_pad + (size_t(batch) * 4 + 15) / 16 * 16;
}

// Backward for head_dim 128 on the tcgen05 path.  qkv / dqkv: fused [B*S, 3*H*128] buffers (q | k | v column blocks).
extern "C" int mla_attn_bwd_sm100(const void* qkv, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o,
                                  const void* lse, const void* mask, void* dqkv, int64_t ld_dqkv, void* workspace,
                                  int32_t batch, int32_t seq, int32_t heads, float scale, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || seq <= 0 || heads <= 0) return set_error(MLA_ERR_ARG, "attn_bwd_sm100: empty problem");
  if ((ld_qkv & 7) || (ld_o & 7) || (ld_dqkv & 7) || (reinterpret_cast<uintptr_t>(qkv) & 15) ||
      (reinterpret_cast<uintptr_t>(d_o) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return set_error(MLA_ERR_ARG, "attn_bwd_sm100: pitches must be multiples of 8 elements, bases 16-byte aligned");
  auto s = (cudaStream_t)stream;
  const int s_pad = (seq + 127) / 128 * 128;
  float* lse2 = (float*)workspace;
  float* delta = lse2 + size_t(batch) * heads * s_pad;
  if (int rc = attn_bwd_prep_launch(o, d_o, ld_o, lse, lse2, delta, batch, seq, heads, s_pad, s)) return rc;
  CUtensorMap m_qkv_fix, m_qkv_str, m_do_fix, m_do_str;
  const uint64_t dims_qkv[2] = {uint64_t(3) * heads * BW_D, uint64_t(batch) * seq};
  const uint64_t dims_do[2] = {uint64_t(heads) * BW_D, uint64_t(batch) * seq};
  const uint64_t st_qkv[1] = {uint64_t(ld_qkv) * 2}, st_do[1] = {uint64_t(ld_o) * 2};
  const uint32_t box_fix[2] = {64, BW_OUT}, box_str[2] = {64, BW_IN};
  if (int rc = encode_tmap_2d_bf16(&m_qkv_fix, qkv, dims_qkv, st_qkv, box_fix)) return rc;
  if (int rc = encode_tmap_2d_bf16(&m_qkv_str, qkv, dims_qkv, st_qkv, box_str)) return rc;
  if (int rc = encode_tmap_2d_bf16(&m_do_fix, d_o, dims_do, st_do, box_fix)) return rc;
  if (int rc = encode_tmap_2d_bf16(&m_do_str, d_o, dims_do, st_do, box_str)) return rc;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_sm100_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_sm100_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "attn_bwd_sm100 smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  BwParams p;
  p.B = batch; p.S = seq; p.H = heads; p.S_pad = s_pad; p.scale = scale;
  p.lse2 = lse2; p.delta = delta; p.mask = (const uint8_t*)mask;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv;
  attn_bwd_sm100_kernel<0><<<batch * heads * 2, BW_THREADS, BW_SMEM, s>>>(m_qkv_fix, m_qkv_str, m_do_fix, m_do_str, p);
  MLA_CHECK_LAUNCH("attn_bwd_sm100_dkv");
  attn_bwd_sm100_kernel<1><<<batch * heads * 2, BW_THREADS, BW_SMEM, s>>>(m_qkv_fix, m_qkv_str, m_do_fix, m_do_str, p);
  MLA_CHECK_LAUNCH("attn_bwd_sm100_dq");
  return MLA_OK;
}
