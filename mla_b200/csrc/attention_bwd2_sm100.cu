// Causal flash attention BACKWARD, second generation (head_dim 128): same two-pass formulation, operand layouts and
// rounding points as attention_bwd_sm100.cu —
//   MODE 0  dK/dV : a CTA owns 128 key rows (K_j, V_j resident) and streams 64-row Q_i / dO_i tiles:
//                   S^T = K_j.Q_i^T, dP^T = V_j.dO_i^T (TMEM) -> P^T = exp2(S^T*c - lse_i), dS^T = P^T o (dP^T - delta_i)
//                   -> dV += P^T.dO_i, dK += dS^T.Q_i (TMEM accumulators)
//   MODE 1  dQ    : a CTA owns 128 query rows (Q_t, dO_t resident) and streams 64-row K_j / V_j tiles:
//                   S = Q_t.K_j^T, dP = dO_t.V_j^T -> dS = P o (dP - delta_t) -> dQ += dS.K_j
// — but pipelined so that the tensor core, not the element-wise math, sets the pace.  The first generation kept ONE
// P^T/dS^T staging buffer and ONE group of math warps, so every streamed tile ran  score MMAs -> math -> accumulation
// MMAs  as a serial chain (ncu: tensor pipe 18-20 % active, 179 TFLOP/s at [32,548,32,128]).  Here:
//   * TWO groups of 8 math warps alternate over the streamed tiles (group g takes tiles with parity g), each with its
//     own S/dP accumulator pair in TMEM and its own P^T/dS^T staging buffers in shared memory: while group 0 does the
//     exponentials of tile c, group 1 is already on tile c+1 and the MMA warp issues the accumulation products of
//     tile c-1 and the score products of tile c+2;
//   * the streamed Q/dO (K/V) tiles sit in a 3-slot TMA ring, so a slot's reload (it must wait for the accumulation
//     MMAs that read it MN-major) is a whole tile ahead of its use;
//   * the per-column softmax statistics of MODE 0 are read as float4 broadcasts, P/dS go to shared memory as soon as
//     8 of them are packed (no 32-register staging), which keeps 576 threads within 112 registers;
//   * the transpose of the rotary embedding (the backward of RoPE on dq and dk) is applied in the epilogue while the
//     accumulators leave TMEM, bit-identical to the separate in-place pass it replaces (rope_kernel, sign = -1).
// Per CTA (1 per SM, 576 threads): warp 0 TMA loader, warp 1 MMA issuer, warps 2-9 math group 0, warps 10-17 group 1
// (two threads per TMEM lane and group, 32 score columns each).  Work item = one half of a (batch, head): the sequence's
// outer tiles are dealt to the two halves in a zig-zag.  The CTAs are PERSISTENT (one per SM, items round-robin): the
// barriers, the TMEM allocation and the TMA / MMA / math pipelines stay alive across items, the loader runs ahead into
// the next item, and (TMEM hand-over variant) the fixed operands are double-buffered, so the next outer tile's K_j | V_j
// (Q_t | dO_t) arrive while the current one computes — at 548-token sequences an outer tile only has ~5 streamed tiles,
// and the launch / fill / drain of one-CTA-per-item used to cost as much as the steady state.
// TMEM (512 cols): [S0|dP0|S1|dP1] 4 x 64, accumulators @256 (dV or dQ) and @384 (dK).
// smem: fixed 2x32 KB | streamed 3 slots x (16+16) KB | 2 x (P^T 16 KB + dS^T 16 KB) | barriers | per-column lse/delta.
#include <cstdlib>
#include <type_traits>

#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int B2_D = 128, B2_OUT = 128, B2_IN = 64;
constexpr int B2_FIX_BYTES = B2_OUT * B2_D * 2;     // 32 KB
constexpr int B2_STR_BYTES = B2_IN * B2_D * 2;      // 16 KB
constexpr int B2_T_BYTES = B2_OUT * B2_IN * 2;      // 16 KB
constexpr int B2_SLOTS = 3;
constexpr int B2_GROUP_WARPS = 8;
constexpr int B2_THREADS = 64 + 2 * B2_GROUP_WARPS * 32;    // 576
constexpr int B2_TILES_BYTES = 2 * B2_FIX_BYTES + 2 * B2_SLOTS * B2_STR_BYTES + 4 * B2_T_BYTES;   // 224 KB
constexpr int B2_STAT_BYTES = B2_SLOTS * 128 * 4;
constexpr int B2_SMEM = B2_TILES_BYTES + 256 + B2_STAT_BYTES + 16;
constexpr float B2_LOG2E = 1.4426950408889634f;
static_assert(B2_SMEM <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

struct Bw2Params {
  int B, S, H, S_pad;
  float scale;
  const float* lse2;     // [B,H,S_pad]  lse * log2(e); +inf for rows that take no part
  const float* delta;    // [B,H,S_pad]
  const uint8_t* mask;   // [B,S] or null
  __nv_bfloat16* dqkv;   // [B*S, 3*H*D]
  int64_t ld_dqkv;
  const __nv_bfloat16* rope_cos;   // bf16 [S, 64] or null: apply the transposed rotation to dq / dk in the epilogue
  const __nv_bfloat16* rope_sin;
  // shared-prefix layout (see attention_fwd_sm100.cu): keys >= prefix_len[b] are visible to the `group` rows of their own
  // group only; rope_pos (int32 [B*S] or null) maps a row to its rotary position (suffix groups repeat positions)
  const int32_t* prefix_len;
  int group;
  const int32_t* rope_pos;
  const int32_t* batch_masked;   // int32 [B]: 1 when mask[b, :] has a zero (computed by the launcher's flag kernel), or null
};

__device__ __forceinline__ void b2_named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float b2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void b2_bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 18 warps are allocated as 20 (warp allocation granularity 4): 640 x 96 registers is what the register file allows
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void b2_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (bf16, two K-elements per 32-bit column, row = TMEM lane) is read
// from tensor memory, so it costs no shared-memory bandwidth.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TS = 1: P^T / dS^T (P / dS) are handed to the accumulation MMAs through TENSOR memory instead of shared memory: each
// math thread packs its 32 values to bf16 pairs and stores them (tcgen05.st) over the first half of the S / dP columns it
// has just read, and the accumulation products take their A operand from there.  The kernel is shared-memory-bandwidth
// bound (the tensor core re-reads its A operand from shared memory for every K = 16 step: 160 KB per streamed tile at
// 128 B/clk); this removes 32 KB of operand reads and 32 KB of staging writes per tile.  The S/dP buffer of group g is
// recycled by the score products of tile c + 2, which the in-order tensor pipe executes after the accumulation products
// of tile c that read it — no extra barrier.
template <int MODE, int TS>
__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd2_sm100_kernel(const __grid_constant__ CUtensorMap map_qkv_fix, const __grid_constant__ CUtensorMap map_qkv_str,
                       const __grid_constant__ CUtensorMap map_do_fix, const __grid_constant__ CUtensorMap map_do_str,
                       Bw2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // fixed operands (MODE 0: K_j | V_j, MODE 1: Q_t | dO_t): buffer 0 here, buffer 1 (TS only: the next outer tile is
  // prefetched while the current one computes) in the P / dS staging area the TS variant does not need
  uint8_t* sFix0 = smem;
  uint8_t* sX = smem + 2 * B2_FIX_BYTES;                 // 3 slots  MODE 0: Q_i   MODE 1: K_j
  uint8_t* sY = sX + B2_SLOTS * B2_STR_BYTES;            // 3 slots  MODE 0: dO_i  MODE 1: V_j
  uint8_t* sT1 = sY + B2_SLOTS * B2_STR_BYTES;           // 2 buffers: P^T (MODE 0 only)
  uint8_t* sT2 = sT1 + 2 * B2_T_BYTES;                   // 2 buffers: dS^T / dS
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B2_TILES_BYTES);
  uint64_t *fix_full = bars /*2*/, *fix_empty = bars + 2 /*2*/, *in_full = bars + 4 /*3*/, *in_empty = bars + 7 /*3*/,
           *sd_full = bars + 10 /*2*/, *sd_empty = bars + 12 /*2*/, *ds_full = bars + 14 /*2*/, *acc_done = bars + 16 /*2*/,
           *acc_full = bars + 18, *acc_empty = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  constexpr int NFIX = TS ? 2 : 1;
  uint8_t* sFix1 = TS ? sT1 : sFix0;
  auto fix_buf = [&](int ti) -> uint8_t* { return (NFIX == 2 && (ti & 1)) ? sFix1 : sFix0; };
  auto fix_idx = [&](int ti) { return NFIX == 2 ? (ti & 1) : 0; };
  auto fix_par = [&](int ti) -> uint32_t { return NFIX == 2 ? ((ti >> 1) & 1) : (ti & 1); };
  float* s_stat = reinterpret_cast<float*>(smem + B2_TILES_BYTES + 256);   // [slot][lse2|delta][64]

  const int warp = int(warp_idx_uniform()), lane = threadIdx.x & 31;     // warp-uniform role index (see ptx.cuh)
  const int S = p.S;
  const int n_out = (S + B2_OUT - 1) / B2_OUT;
  const int n_in = (S + B2_IN - 1) / B2_IN;
  const int HD = p.H * B2_D;
  // Persistent CTAs: work item w = (batch, head, half); a CTA walks items blockIdx.x, + gridDim.x, ... with its barriers,
  // TMEM allocation and pipelines alive across items (the loader runs ahead into the next item).  The halves of
  // consecutive rounds are swapped so that every CTA alternates between the heavier and the lighter half.
  const int n_items = p.B * p.H * 2;
  struct Item { int half_id, bh, b, hd, row_base, col_q, col_k, col_v, col_do; };
  auto item_of = [&](int w) {
    Item I;
    I.half_id = (w & 1) ^ ((w / int(gridDim.x)) & 1);
    I.bh = w >> 1;
    I.b = I.bh / p.H; I.hd = I.bh % p.H;
    I.row_base = I.b * S;
    I.col_q = I.hd * B2_D; I.col_k = HD + I.hd * B2_D; I.col_v = 2 * HD + I.hd * B2_D; I.col_do = I.hd * B2_D;
    return I;
  };
  // outer tiles in order of decreasing work are dealt A B B A A B B A ...
  auto my_tile = [&](int half_id, int k) -> int {
    const int pos = half_id == 0 ? (k == 0 ? 0 : 4 * ((k + 1) >> 1) - ((k & 1) ? 1 : 0)) : (4 * (k >> 1) + 1 + (k & 1));
    if (pos >= n_out) return -1;
    return MODE == 0 ? pos : n_out - 1 - pos;      // dK/dV: key tile 0 sees every query; dQ: the last query tile sees every key
  };
  auto in_begin = [&](int t) { return MODE == 0 ? 2 * t : 0; };
  auto in_end = [&](int t) { return MODE == 0 ? n_in : min(2 * t + 2, n_in); };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_qkv_fix); tma_prefetch_desc(&map_qkv_str);
    tma_prefetch_desc(&map_do_fix); tma_prefetch_desc(&map_do_str);
    for (int i = 0; i < 2; ++i) { mbar_init(&fix_full[i], 1); mbar_init(&fix_empty[i], 1); }
    for (int i = 0; i < B2_SLOTS; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], B2_GROUP_WARPS);
      mbar_init(&ds_full[i], B2_GROUP_WARPS); mbar_init(&acc_done[i], 1);
    }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 2 * B2_GROUP_WARPS);
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

  if (warp == 0) {
    // ================================ TMA loader ================================
    if (lane == 0) {
      int c = 0, tg = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item I = item_of(w);
      const int bh = I.bh, row_base = I.row_base, col_q = I.col_q, col_k = I.col_k, col_v = I.col_v, col_do = I.col_do;
      for (int k = 0;; ++k) {
        const int t = my_tile(I.half_id, k);
        if (t < 0) break;
        const int ti = tg++;
        uint64_t* ff = &fix_full[fix_idx(ti)];
        uint8_t* sF1 = fix_buf(ti);
        uint8_t* sF2 = sF1 + B2_FIX_BYTES;
        mbar_wait(&fix_empty[fix_idx(ti)], fix_par(ti) ^ 1);
        mbar_arrive_expect_tx(ff, 2 * B2_FIX_BYTES);
        const int r0 = row_base + t * B2_OUT;
        if (MODE == 0) {
          tma_load_2d(sF1, &map_qkv_fix, ff, col_k, r0);
          tma_load_2d(sF1 + B2_FIX_BYTES / 2, &map_qkv_fix, ff, col_k + 64, r0);
          tma_load_2d(sF2, &map_qkv_fix, ff, col_v, r0);
          tma_load_2d(sF2 + B2_FIX_BYTES / 2, &map_qkv_fix, ff, col_v + 64, r0);
        } else {
          tma_load_2d(sF1, &map_qkv_fix, ff, col_q, r0);
          tma_load_2d(sF1 + B2_FIX_BYTES / 2, &map_qkv_fix, ff, col_q + 64, r0);
          tma_load_2d(sF2, &map_do_fix, ff, col_do, r0);
          tma_load_2d(sF2 + B2_FIX_BYTES / 2, &map_do_fix, ff, col_do + 64, r0);
        }
        for (int i = in_begin(t); i < in_end(t); ++i, ++c) {
          const int s = c % B2_SLOTS;
          mbar_wait(&in_empty[s], ((c / B2_SLOTS) & 1) ^ 1);
          uint8_t* dx = sX + s * B2_STR_BYTES;
          uint8_t* dy = sY + s * B2_STR_BYTES;
          const int c0 = row_base + i * B2_IN;
          if (MODE == 0) {
            mbar_arrive_expect_tx(&in_full[s], 2 * B2_STR_BYTES + 512);
            tma_load_2d(dx, &map_qkv_str, &in_full[s], col_q, c0);
            tma_load_2d(dx + B2_STR_BYTES / 2, &map_qkv_str, &in_full[s], col_q + 64, c0);
            tma_load_2d(dy, &map_do_str, &in_full[s], col_do, c0);
            tma_load_2d(dy + B2_STR_BYTES / 2, &map_do_str, &in_full[s], col_do + 64, c0);
            const int64_t so = int64_t(bh) * p.S_pad + i * B2_IN;
            b2_bulk_copy_g2s(s_stat + s * 128, p.lse2 + so, 256, &in_full[s]);
            b2_bulk_copy_g2s(s_stat + s * 128 + 64, p.delta + so, 256, &in_full[s]);
          } else {
            mbar_arrive_expect_tx(&in_full[s], 2 * B2_STR_BYTES);
            tma_load_2d(dx, &map_qkv_str, &in_full[s], col_k, c0);
            tma_load_2d(dx + B2_STR_BYTES / 2, &map_qkv_str, &in_full[s], col_k + 64, c0);
            tma_load_2d(dy, &map_qkv_str, &in_full[s], col_v, c0);
            tma_load_2d(dy + B2_STR_BYTES / 2, &map_qkv_str, &in_full[s], col_v + 64, c0);
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // Executed by all 32 lanes (converged, warp-uniform values): barrier polls are warp-wide, every tcgen05.mma / commit
    // elects its issuing lane itself.  Descriptors are built once and advanced by constants.
    {
      constexpr uint32_t idesc_sd = umma_idesc_bf16(B2_OUT, B2_IN, 0, 0);    // 128 x 64,  K = d
      constexpr uint32_t idesc_acc = umma_idesc_bf16(B2_OUT, B2_D, 0, 1);    // 128 x 128, K = 64 streamed rows, B MN-major
      const uint64_t dFix0 = umma_smem_desc_sw128(smem_u32(sFix0), 16, 1024);
      const uint64_t dFix1 = umma_smem_desc_sw128(smem_u32(sFix1), 16, 1024);
      uint64_t dF1 = dFix0, dF2 = umma_desc_advance(dFix0, B2_FIX_BYTES);     // set per outer tile
      const uint64_t dX0 = umma_smem_desc_sw128(smem_u32(sX), 16, 1024), dY0 = umma_smem_desc_sw128(smem_u32(sY), 16, 1024);
      // the same streamed tiles read MN-major by the accumulation products (LBO = the other 64-column chunk)
      const uint64_t dXm0 = umma_smem_desc_sw128(smem_u32(sX), B2_STR_BYTES / 2, 1024);
      const uint64_t dYm0 = umma_smem_desc_sw128(smem_u32(sY), B2_STR_BYTES / 2, 1024);
      const uint64_t dT1_0 = umma_smem_desc_sw128(smem_u32(sT1), 16, 1024), dT2_0 = umma_smem_desc_sw128(smem_u32(sT2), 16, 1024);
      auto issue_sd = [&](int c) {
        const int s = c % B2_SLOTS, g = c & 1;
        mbar_wait(&in_full[s], (c / B2_SLOTS) & 1);
        if (!TS) mbar_wait(&sd_empty[g], ((c >> 1) & 1) ^ 1);   // TS: ordered behind ACC(c - 2) by the in-order tensor pipe
        tc_fence_after();
        const uint64_t dX = umma_desc_advance(dX0, s * B2_STR_BYTES), dY = umma_desc_advance(dY0, s * B2_STR_BYTES);
        const uint32_t tS = tmem_base + g * 128, tP = tS + 64;
#pragma unroll
        for (int kk = 0; kk < B2_D / 16; ++kk) {
          const uint32_t of = (kk >> 2) * (B2_FIX_BYTES / 2) + (kk & 3) * 32;
          const uint32_t os = (kk >> 2) * (B2_STR_BYTES / 2) + (kk & 3) * 32;
          umma_f16_ss_elect(tS, umma_desc_advance(dF1, of), umma_desc_advance(dX, os), idesc_sd, kk != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < B2_D / 16; ++kk) {
          const uint32_t of = (kk >> 2) * (B2_FIX_BYTES / 2) + (kk & 3) * 32;
          const uint32_t os = (kk >> 2) * (B2_STR_BYTES / 2) + (kk & 3) * 32;
          umma_f16_ss_elect(tP, umma_desc_advance(dF2, of), umma_desc_advance(dY, os), idesc_sd, kk != 0 ? 1u : 0u);
        }
        umma_commit_elect(&sd_full[g]);
      };
      int it = 0, tg = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item I = item_of(w);
      for (int k = 0;; ++k) {
        const int t = my_tile(I.half_id, k);
        if (t < 0) break;
        const int ti = tg++;
        const int n = in_end(t) - in_begin(t);
        uint64_t* fe = &fix_empty[fix_idx(ti)];
        dF1 = (NFIX == 2 && (ti & 1)) ? dFix1 : dFix0;
        dF2 = umma_desc_advance(dF1, B2_FIX_BYTES);
        mbar_wait(&fix_full[fix_idx(ti)], fix_par(ti));
        issue_sd(it);
        if (n > 1) issue_sd(it + 1);
        else umma_commit_elect(fe);
        for (int j = 0; j < n; ++j) {
          const int c = it + j;
          const int s = c % B2_SLOTS, g = c & 1;
          mbar_wait(&ds_full[g], (c >> 1) & 1);
          if (j == 0) mbar_wait(acc_empty, (ti & 1) ^ 1);    // previous outer tile's accumulators were read out
          tc_fence_after();
          const uint64_t dXm = umma_desc_advance(dXm0, s * B2_STR_BYTES), dYm = umma_desc_advance(dYm0, s * B2_STR_BYTES);
          const uint64_t dP = umma_desc_advance(dT1_0, g * B2_T_BYTES), dD = umma_desc_advance(dT2_0, g * B2_T_BYTES);
#pragma unroll
          for (int kk = 0; kk < B2_IN / 16; ++kk) {
            const uint32_t acc = (j | kk) != 0 ? 1u : 0u;
            if (TS) {
              // A from TMEM: 16 streamed columns = 8 packed TMEM columns; half h of the tile sits at column 32 h of its
              // S (P^T) / dP (dS^T) block
              const uint32_t tP = tmem_base + g * 128 + (kk >> 1) * 32 + (kk & 1) * 8, tD = tP + 64;
              if (MODE == 0) {
                umma_f16_ts_elect(tmem_acc0, tP, umma_desc_advance(dYm, kk * 2048), idesc_acc, acc);   // dV += P^T.dO
                umma_f16_ts_elect(tmem_acc1, tD, umma_desc_advance(dXm, kk * 2048), idesc_acc, acc);   // dK += dS^T.Q
              } else {
                umma_f16_ts_elect(tmem_acc0, tD, umma_desc_advance(dXm, kk * 2048), idesc_acc, acc);   // dQ += dS.K
              }
            } else if (MODE == 0) {
              umma_f16_ss_elect(tmem_acc0, umma_desc_advance(dP, kk * 32), umma_desc_advance(dYm, kk * 2048), idesc_acc, acc);
              umma_f16_ss_elect(tmem_acc1, umma_desc_advance(dD, kk * 32), umma_desc_advance(dXm, kk * 2048), idesc_acc, acc);
            } else {
              umma_f16_ss_elect(tmem_acc0, umma_desc_advance(dD, kk * 32), umma_desc_advance(dXm, kk * 2048), idesc_acc, acc);
            }
          }
          umma_commit_elect(&acc_done[g]);
          umma_commit_elect(&in_empty[s]);
          if (j == n - 1) umma_commit_elect(acc_full);
          // keep the score products two tiles ahead of the accumulation products
          if (j + 2 < n) issue_sd(c + 2);
          else if (j + 2 == n) umma_commit_elect(fe);              // every score product of this outer tile has been issued
        }
        it += n;
      }
      }
    }
  } else {
    // ================================ element-wise math: 2 groups x 256 threads, two threads per TMEM lane ================================
    const int mw = warp - 2;
    const int grp = mw >> 3;                  // which parity of streamed tiles this warp's group handles
    const int quarter = warp & 3;             // TMEM lane quarter this warp may touch
    const int half = (mw >> 2) & 1;           // which 32 of the 64 streamed columns
    const int r = quarter * 32 + lane;        // row of the outer tile == TMEM lane
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const float sl2 = p.scale * B2_LOG2E;
    uint8_t* myT1 = sT1 + grp * B2_T_BYTES;
    uint8_t* myT2 = sT2 + grp * B2_T_BYTES;
    int it = 0, tg = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
    const Item I = item_of(w);
    const int bh = I.bh, b = I.b, row_base = I.row_base, col_q = I.col_q, col_k = I.col_k, col_v = I.col_v;
    const uint8_t* gmask = p.mask ? p.mask + int64_t(b) * S : nullptr;
    const bool use_mask = gmask && (p.batch_masked == nullptr || p.batch_masked[b] != 0);
    const int P = p.group > 0 ? p.prefix_len[b] : 0x3fffffff;     // keys >= P are visible to their own group only
    for (int k = 0;; ++k) {
      const int t = my_tile(I.half_id, k);
      if (t < 0) break;
      const int ti = tg++;
      const int i0 = in_begin(t), n = in_end(t) - i0;
      const int orow = t * B2_OUT + r;        // key row (MODE 0) / query row (MODE 1) in the sequence
      float row_lse2 = 0.f, row_delta = 0.f;
      bool row_ok = true;
      // MODE 1: first key of this query row's own group (prefix rows: 0 = nothing hidden by the group rule)
      const int glo = (MODE == 1 && p.group > 0 && orow >= P) ? P + ((orow - P) / p.group) * p.group : 0;
      if (MODE == 1) {
        row_lse2 = p.lse2[int64_t(bh) * p.S_pad + orow];      // S_pad covers every row of every outer tile
        row_delta = p.delta[int64_t(bh) * p.S_pad + orow];
      } else {
        row_ok = orow < S && (!use_mask || gmask[orow]);      // masked / out-of-range keys get P = 0
      }
      for (int j = ((it & 1) == grp ? 0 : 1); j < n; j += 2) {
        const int c = it + j;                  // (c & 1) == grp
        const int slot = c % B2_SLOTS;
        const int in0 = (i0 + j) * B2_IN + half * 32;   // first streamed column (sequence position) of this thread
        const uint32_t tS = tmem_base + lane_off + grp * 128 + half * 32;
        bool edge;
        if (MODE == 0) edge = ((i0 + j) * B2_IN < t * B2_OUT + B2_OUT) || (t * B2_OUT + B2_OUT > P);   // diagonal / suffix keys
        else edge = ((i0 + j) * B2_IN + B2_IN - 1 > t * B2_OUT) || ((i0 + j + 1) * B2_IN > S) || use_mask ||
                    ((i0 + j + 1) * B2_IN > P);
        mbar_wait(&sd_full[grp], (c >> 1) & 1);
        tc_fence_after();
        uint32_t vs[32], vp[32];
        tmem_ld_32x32b_x32(tS, vs);
        tmem_ld_32x32b_x32(tS + 64, vp);
        tmem_ld_wait();
        if (!TS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sd_empty[grp]);
          // this group's staging buffers are free once the accumulation MMAs of its previous tile (c - 2) retired
          if (c >= 2) mbar_wait(&acc_done[grp], ((c >> 1) - 1) & 1);
        }
        const float4* stat4 = reinterpret_cast<const float4*>(s_stat + slot * 128 + half * 32);   // MODE 0: lse2 | delta
        // MASKED = false: every element of the tile is visible (no diagonal, no padding): no predicates, no branches
        auto tile_math = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        uint32_t pkp[8], pkd[8];
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float l2v[8], dlv[8];
          if (MODE == 0) {
            const float4 a0 = stat4[g8 * 2], a1 = stat4[g8 * 2 + 1], d0 = stat4[16 + g8 * 2], d1 = stat4[16 + g8 * 2 + 1];
            l2v[0] = a0.x; l2v[1] = a0.y; l2v[2] = a0.z; l2v[3] = a0.w; l2v[4] = a1.x; l2v[5] = a1.y; l2v[6] = a1.z; l2v[7] = a1.w;
            dlv[0] = d0.x; dlv[1] = d0.y; dlv[2] = d0.z; dlv[3] = d0.w; dlv[4] = d1.x; dlv[5] = d1.y; dlv[6] = d1.z; dlv[7] = d1.w;
          }
          float pr[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int cc = g8 * 8 + e;
            const float l2 = MODE == 0 ? l2v[e] : row_lse2;
            const float dl = MODE == 0 ? dlv[e] : row_delta;
            float x = __fmaf_rn(__uint_as_float(vs[cc]), sl2, -l2);     // lse2 = +inf (rows that take no part) -> P = 0
            if (MASKED) {
              bool vis;
              if (MODE == 0) {
                const int col = in0 + cc;                      // query position; orow = key position
                vis = row_ok && (!edge || orow <= col);
                if (p.group > 0 && orow >= P) vis = vis && orow >= P + ((col - P) / p.group) * p.group;   // same group
              } else {
                const int col = in0 + cc;                      // key position; orow = query position
                vis = (col <= orow) && (col < S) && (col < P || col >= glo) && (!use_mask || gmask[col]);
              }
              x = vis ? x : -INFINITY;                                   // select, not a branch: ex2(-inf) = 0
            }
            const float pv = b2_ex2(x);
            pr[e] = pv;
            ds[e] = pv * (__uint_as_float(vp[cc]) - dl);
          }
          if (TS) {
            if (MODE == 0) {
              pkp[(g8 & 1) * 4 + 0] = pack_bf16x2(pr[0], pr[1]); pkp[(g8 & 1) * 4 + 1] = pack_bf16x2(pr[2], pr[3]);
              pkp[(g8 & 1) * 4 + 2] = pack_bf16x2(pr[4], pr[5]); pkp[(g8 & 1) * 4 + 3] = pack_bf16x2(pr[6], pr[7]);
            }
            pkd[(g8 & 1) * 4 + 0] = pack_bf16x2(ds[0], ds[1]); pkd[(g8 & 1) * 4 + 1] = pack_bf16x2(ds[2], ds[3]);
            pkd[(g8 & 1) * 4 + 2] = pack_bf16x2(ds[4], ds[5]); pkd[(g8 & 1) * 4 + 3] = pack_bf16x2(ds[6], ds[7]);
            if (g8 & 1) {     // 16 values = 8 packed columns ready: over this thread's own (already read) S / dP columns
              if (MODE == 0) tmem_st_32x32b_x8(tS + (g8 >> 1) * 8, pkp);
              tmem_st_32x32b_x8(tS + 64 + (g8 >> 1) * 8, pkd);
            }
          } else {
            const int off = r * 128 + (((half * 4 + g8) ^ (r & 7)) * 16);
            if (MODE == 0)
              *reinterpret_cast<uint4*>(myT1 + off) = make_uint4(pack_bf16x2(pr[0], pr[1]), pack_bf16x2(pr[2], pr[3]),
                                                                 pack_bf16x2(pr[4], pr[5]), pack_bf16x2(pr[6], pr[7]));
            *reinterpret_cast<uint4*>(myT2 + off) = make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]),
                                                               pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7]));
          }
        }
        };
        const bool fast = MODE == 0 ? (!edge && __all_sync(0xffffffffu, row_ok)) : !edge;      // warp-uniform
        if (fast) tile_math(std::false_type{});
        else tile_math(std::true_type{});
        if (TS) b2_tmem_st_wait();
        else fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[grp]);
      }
      // ---- epilogue of this outer tile: accumulators -> bf16 -> global.  Four threads share a TMEM lane (2 groups x 2
      // halves): thread e takes columns [16e, 16e+16) and [64+16e, 64+16e+16) — the two halves a rotary pair lives in.
      mbar_wait(acc_full, ti & 1);
      tc_fence_after();
      const bool store = orow < S;
      const int e4 = grp * 2 + half;
      __nv_bfloat16* base = p.dqkv + int64_t(row_base + orow) * p.ld_dqkv + e4 * 16;
      const bool rope = p.rope_cos != nullptr;
      float cs[16], sn[16];
      if (rope && store) {
        const int pos = p.rope_pos ? p.rope_pos[row_base + orow] : orow;
        const uint4* cp = reinterpret_cast<const uint4*>(p.rope_cos + int64_t(pos) * 64 + e4 * 16);
        const uint4* sp = reinterpret_cast<const uint4*>(p.rope_sin + int64_t(pos) * 64 + e4 * 16);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 cq = cp[q], sq = sp[q];
          const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&cq);
          const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&sq);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 cf = __bfloat1622float2(ch[u]), sf = __bfloat1622float2(sh[u]);
            cs[q * 8 + 2 * u] = cf.x; cs[q * 8 + 2 * u + 1] = cf.y;
            sn[q * 8 + 2 * u] = sf.x; sn[q * 8 + 2 * u + 1] = sf.y;
          }
        }
      }
#pragma unroll 1
      for (int a = 0; a < (MODE == 0 ? 2 : 1); ++a) {
        const uint32_t tacc = (a == 0 ? tmem_acc0 : tmem_acc1) + lane_off + e4 * 16;
        // MODE 0: acc0 = dV (v block), acc1 = dK (k block, scaled); MODE 1: acc0 = dQ (q block, scaled)
        const int col = MODE == 0 ? (a == 0 ? col_v : col_k) : col_q;
        const bool is_v = (MODE == 0 && a == 0);
        const float sc = is_v ? 1.f : p.scale;
        uint32_t v1[16], v2[16];
        tmem_ld_32x32b_x16(tacc, v1);
        tmem_ld_32x32b_x16(tacc + 64, v2);
        tmem_ld_wait();
        if (store) {
          float o1[16], o2[16];
          if (rope && !is_v) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              // rope_kernel with sign = -1 on the bf16-rounded gradient: the transpose of the rotation
              const float x1 = bf16_round(__uint_as_float(v1[i]) * sc), x2 = bf16_round(__uint_as_float(v2[i]) * sc);
              const float s = -sn[i];
              o1[i] = bf16_round(x1 * cs[i]) + bf16_round(-x2 * s);
              o2[i] = bf16_round(x2 * cs[i]) + bf16_round(x1 * s);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              o1[i] = __uint_as_float(v1[i]) * sc;
              o2[i] = __uint_as_float(v2[i]) * sc;
            }
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            *reinterpret_cast<uint4*>(base + col + q * 8) =
                make_uint4(pack_bf16x2(o1[q * 8 + 0], o1[q * 8 + 1]), pack_bf16x2(o1[q * 8 + 2], o1[q * 8 + 3]),
                           pack_bf16x2(o1[q * 8 + 4], o1[q * 8 + 5]), pack_bf16x2(o1[q * 8 + 6], o1[q * 8 + 7]));
            *reinterpret_cast<uint4*>(base + col + 64 + q * 8) =
                make_uint4(pack_bf16x2(o2[q * 8 + 0], o2[q * 8 + 1]), pack_bf16x2(o2[q * 8 + 2], o2[q * 8 + 3]),
                           pack_bf16x2(o2[q * 8 + 4], o2[q * 8 + 5]), pack_bf16x2(o2[q * 8 + 6], o2[q * 8 + 7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      it += n;
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// flags[b] = 1 when mask[b, :] contains a zero (one CTA per batch row)
__global__ void attn_mask_flags_kernel(const uint8_t* __restrict__ mask, int S, int32_t* __restrict__ flags) {
  int bad = 0;
  for (int i = threadIdx.x; i < S; i += blockDim.x) bad |= (mask[int64_t(blockIdx.x) * S + i] == 0);
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flags[blockIdx.x] = bad ? 1 : 0;
}

// defined in attention_bwd_sm100.cu: lse2 = lse * log2(e) and delta = rowsum(dO o O), padded to S_pad per (b, h)
int attn_bwd_prep_launch(const void* o, const void* d_o, int64_t ld_o, const void* lse, float* lse2, float* delta, int B,
                         int S, int H, int S_pad, cudaStream_t s);

}  // namespace mla

using namespace mla;

// Backward for head_dim 128, pipelined generation.  Same contract as mla_attn_bwd_sm100 (workspace from
// mla_attn_bwd_sm100_workspace) plus: rope_cos / rope_sin (bf16 [seq, 64], or null) — when given, the gradients written
// to the q and k column blocks of dqkv are those w.r.t. the PRE-RoPE projections (the transposed rotation is applied in
// the epilogue), i.e. dqkv is ready for the weight-gradient / input-gradient GEMMs of the q|k|v projection.
static int g_bwd2_ts = [] { const char* e = getenv("MLA_ATTN_BWD_TS"); return e ? atoi(e) : 1; }();

// 0: P/dS staged in shared memory; 1 (default): handed over in tensor memory (A operand of the accumulation MMAs in TMEM)
extern "C" int mla_attn_bwd2_set_ts(int32_t on) {
  g_bwd2_ts = on ? 1 : 0;
  return MLA_OK;
}

extern "C" int mla_attn_bwd2_sm100_grouped(const void* qkv, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o,
                                           const void* lse, const void* mask, void* dqkv, int64_t ld_dqkv, void* workspace,
                                           const void* rope_cos, const void* rope_sin, const void* rope_pos,
                                           const void* prefix_len, int32_t group, int32_t batch, int32_t seq,
                                           int32_t heads, float scale, void* stream);

extern "C" int mla_attn_bwd2_sm100(const void* qkv, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o,
                                   const void* lse, const void* mask, void* dqkv, int64_t ld_dqkv, void* workspace,
                                   const void* rope_cos, const void* rope_sin, int32_t batch, int32_t seq, int32_t heads,
                                   float scale, void* stream) {
  return mla_attn_bwd2_sm100_grouped(qkv, ld_qkv, o, d_o, ld_o, lse, mask, dqkv, ld_dqkv, workspace, rope_cos, rope_sin,
                                     nullptr, nullptr, 0, batch, seq, heads, scale, stream);
}

// Shared-prefix variant (see mla_attn_fwd_sm100_grouped): prefix_len int32 [batch], group = rows per suffix group;
// rope_pos int32 [batch*seq] (or NULL: position = row index within the sequence) for the fused RoPE transpose.
extern "C" int mla_attn_bwd2_sm100_grouped(const void* qkv, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o,
                                           const void* lse, const void* mask, void* dqkv, int64_t ld_dqkv, void* workspace,
                                           const void* rope_cos, const void* rope_sin, const void* rope_pos,
                                           const void* prefix_len, int32_t group, int32_t batch, int32_t seq,
                                           int32_t heads, float scale, void* stream) {
  if (int rc = device_check()) return rc;
  if (group < 0 || (group > 0 && prefix_len == nullptr))
    return set_error(MLA_ERR_ARG, "attn_bwd2_sm100: a shared-prefix layout needs prefix_len and group > 0");
  if (batch <= 0 || seq <= 0 || heads <= 0) return set_error(MLA_ERR_ARG, "attn_bwd2_sm100: empty problem");
  if ((ld_qkv & 7) || (ld_o & 7) || (ld_dqkv & 7) || (reinterpret_cast<uintptr_t>(qkv) & 15) ||
      (reinterpret_cast<uintptr_t>(d_o) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15) ||
      (reinterpret_cast<uintptr_t>(dqkv) & 15))
    return set_error(MLA_ERR_ARG, "attn_bwd2_sm100: pitches must be multiples of 8 elements, bases 16-byte aligned");
  if ((rope_cos == nullptr) != (rope_sin == nullptr) || (reinterpret_cast<uintptr_t>(rope_cos) & 15) ||
      (reinterpret_cast<uintptr_t>(rope_sin) & 15))
    return set_error(MLA_ERR_ARG, "attn_bwd2_sm100: rope tables must both be given (16-byte aligned) or both be null");
  auto s = (cudaStream_t)stream;
  const int s_pad = (seq + 127) / 128 * 128;
  float* lse2 = (float*)workspace;
  float* delta = lse2 + size_t(batch) * heads * s_pad;
  if (int rc = attn_bwd_prep_launch(o, d_o, ld_o, lse, lse2, delta, batch, seq, heads, s_pad, s)) return rc;
  int32_t* flags = nullptr;
  if (mask != nullptr) {
    flags = reinterpret_cast<int32_t*>(delta + size_t(batch) * heads * s_pad);
    attn_mask_flags_kernel<<<batch, 256, 0, s>>>((const uint8_t*)mask, seq, flags);
    MLA_CHECK_LAUNCH("attn_mask_flags");
  }
  CUtensorMap m_qkv_fix, m_qkv_str, m_do_fix, m_do_str;
  const uint64_t dims_qkv[2] = {uint64_t(3) * heads * B2_D, uint64_t(batch) * seq};
  const uint64_t dims_do[2] = {uint64_t(heads) * B2_D, uint64_t(batch) * seq};
  const uint64_t st_qkv[1] = {uint64_t(ld_qkv) * 2}, st_do[1] = {uint64_t(ld_o) * 2};
  const uint32_t box_fix[2] = {64, B2_OUT}, box_str[2] = {64, B2_IN};
  if (int rc = encode_tmap_2d_bf16(&m_qkv_fix, qkv, dims_qkv, st_qkv, box_fix)) return rc;
  if (int rc = encode_tmap_2d_bf16(&m_qkv_str, qkv, dims_qkv, st_qkv, box_str)) return rc;
  if (int rc = encode_tmap_2d_bf16(&m_do_fix, d_o, dims_do, st_do, box_fix)) return rc;
  if (int rc = encode_tmap_2d_bf16(&m_do_str, d_o, dims_do, st_do, box_str)) return rc;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd2_sm100_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd2_sm100_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd2_sm100_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd2_sm100_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "attn_bwd2_sm100 smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  Bw2Params p;
  p.B = batch; p.S = seq; p.H = heads; p.S_pad = s_pad; p.scale = scale;
  p.lse2 = lse2; p.delta = delta; p.mask = (const uint8_t*)mask;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv;
  p.rope_cos = (const __nv_bfloat16*)rope_cos; p.rope_sin = (const __nv_bfloat16*)rope_sin;
  p.prefix_len = group > 0 ? (const int32_t*)prefix_len : nullptr; p.group = group;
  p.rope_pos = (const int32_t*)rope_pos;
  p.batch_masked = flags;
  // persistent CTAs, one per SM (an even count: the two halves of a (batch, head) never straddle a round)
  const int items = batch * heads * 2;
  int grid = num_sms() & ~1;
  if (grid > items) grid = items;
  if (g_bwd2_ts) attn_bwd2_sm100_kernel<0, 1><<<grid, B2_THREADS, B2_SMEM, s>>>(m_qkv_fix, m_qkv_str, m_do_fix, m_do_str, p);
  else attn_bwd2_sm100_kernel<0, 0><<<grid, B2_THREADS, B2_SMEM, s>>>(m_qkv_fix, m_qkv_str, m_do_fix, m_do_str, p);
  MLA_CHECK_LAUNCH("attn_bwd2_sm100_dkv");
  if (g_bwd2_ts) attn_bwd2_sm100_kernel<1, 1><<<grid, B2_THREADS, B2_SMEM, s>>>(m_qkv_fix, m_qkv_str, m_do_fix, m_do_str, p);
  else attn_bwd2_sm100_kernel<1, 0><<<grid, B2_THREADS, B2_SMEM, s>>>(m_qkv_fix, m_qkv_str, m_do_fix, m_do_str, p);
  MLA_CHECK_LAUNCH("attn_bwd2_sm100_dq");
  return MLA_OK;
}
