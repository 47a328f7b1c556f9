// Causal multi-head attention over the fused multimodal sequence, forward and backward, flash-style
// (online softmax, nothing of size S x S ever reaches HBM).  Replaces flash_attn_func / flash_attn_varlen_func +
// unpad/pad_input at transformers/models/llama/modeling_llama.py:540-557.
//
// Semantics (identical to the reference's varlen path): key j is visible to query i iff j <= i and mask[b,j] != 0;
// rows whose own mask is 0 produce zeros (pad_input) and receive/propagate no gradient.  P and dS are rounded to
// bf16 before the second GEMM of each pair, accumulators are fp32, exactly as flash-attn does.
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 tiles with cp.async double buffering (attention is ~1 % of
// the step's FLOPs at S=548; the tcgen05/TMEM version is scheduled after the GEMM path).
//   fwd   : CTA = 64 query rows of one (batch, head); 4 warps x 16 rows; K/V streamed in 64-row tiles.
//   bwd dQ: same decomposition, recomputes P, dQ += dS.K
//   bwd dKdV: CTA = 64 key rows; computes S^T = K.Q^T so P^T/dS^T come out in A-fragment layout directly.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Swizzled [rows, D] bf16 tile: 16-byte chunk c of row r lives at chunk (c ^ f(r)) so that the 8 rows of one
// ldmatrix phase hit 8 different bank groups.
template <int D>
__device__ __forceinline__ int swz(int row, int chunk) {
  constexpr int CP = D / 8;
  const int f = CP >= 8 ? (row & 7) : ((row >> 1) & (CP - 1));
  return row * D + ((chunk ^ f) << 3);
}

// Cooperative load of a [ROWS, D] tile (rows row0.. of one head) into swizzled smem; rows >= row_limit are zero.
template <int D, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int64_t ld, int row0, int row_limit) {
  constexpr int CP = D / 8;
  for (int idx = threadIdx.x; idx < ROWS * CP; idx += THREADS) {
    const int r = idx / CP, c = idx % CP;
    const bool ok = row0 + r < row_limit;
    const __nv_bfloat16* src = g + int64_t(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(s + swz<D>(r, c), src, ok);
  }
}

// A fragment (16 rows x 16 k) of a swizzled [rows, D] tile, rows r0.., k-chunk pair kc (k = 16*kc..).
template <int D>
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const __nv_bfloat16* s, int r0, int kc, int lane) {
  const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int chunk = kc * 2 + (lane >> 4);
  ldsm_x4(a, s + swz<D>(row, chunk));
}
// B fragments for two adjacent n-tiles (n = n0..n0+15) at k-step kc where B[k][n] = tile[n][k] (tile rows are n).
// r[0],r[1] = (b0,b1) of n-tile n0 ; r[2],r[3] = (b0,b1) of n-tile n0+8.
template <int D>
__device__ __forceinline__ void load_b_nk(uint32_t (&r)[4], const __nv_bfloat16* s, int n0, int kc, int lane) {
  const int row = n0 + (lane & 7) + (lane >> 4) * 8;
  const int chunk = kc * 2 + ((lane >> 3) & 1);
  ldsm_x4(r, s + swz<D>(row, chunk));
}
// B fragments for two adjacent n-tiles (n = n0..n0+15) at k rows k0..k0+15 where B[k][n] = tile[k][n] (rows are k).
template <int D>
__device__ __forceinline__ void load_b_kn(uint32_t (&r)[4], const __nv_bfloat16* s, int k0, int n0, int lane) {
  const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int chunk = (n0 >> 3) + (lane >> 4);
  ldsm_x4_t(r, s + swz<D>(row, chunk));
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v;  // row t = b*S+s, head hd at column hd*D; row pitch ld_qkv
  int64_t ld_qkv;
  __nv_bfloat16* o;                // [B*S, H*D], pitch ld_o
  int64_t ld_o;
  float* lse;                      // [B, H, S]  natural-log LSE; +inf for masked-out query rows
  const uint8_t* mask;             // [B, S] or null
  int B, S, H;
  float scale;
  // backward
  const __nv_bfloat16* d_o;        // pitch ld_o
  float* delta;                    // [B, H, S]
  __nv_bfloat16 *dq, *dk, *dv;     // same layout/pitch as q,k,v: ld_dqkv
  int64_t ld_dqkv;
};

// ======================================================================================================== forward
template <int D>
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnParams p) {
  constexpr int BM = 64, BN = 64;
  extern __shared__ __align__(128) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + BM * D;       // 2 buffers
  __nv_bfloat16* sV = sK + 2 * BN * D;   // 2 buffers
  __shared__ uint8_t sMask[2][BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * BM, hd = blockIdx.y, b = blockIdx.z;
  const int S = p.S;
  const int64_t tok0 = int64_t(b) * S;
  const __nv_bfloat16* gq = p.q + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gk = p.k + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gv = p.v + tok0 * p.ld_qkv + hd * D;
  const uint8_t* gmask = p.mask ? p.mask + tok0 : nullptr;

  const int q_hi = min(q0 + BM, S);
  const int n_kv = (q_hi + BN - 1) / BN;  // causal: keys < q_hi

  load_tile<D, BM, 128>(sQ, gq, p.ld_qkv, q0, S);
  load_tile<D, BN, 128>(sK, gk, p.ld_qkv, 0, S);
  load_tile<D, BN, 128>(sV, gv, p.ld_qkv, 0, S);
  if (threadIdx.x < BN) sMask[0][threadIdx.x] = (threadIdx.x < S) && (!gmask || gmask[threadIdx.x]);
  cp_async_commit();

  float acc_o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc_o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * LOG2E;
  const int row_lo = q0 + warp * 16 + g;  // this thread's two query rows: row_lo, row_lo + 8

  for (int j = 0; j < n_kv; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_kv) {
      const int nb = buf ^ 1;
      load_tile<D, BN, 128>(sK + nb * BN * D, gk, p.ld_qkv, (j + 1) * BN, S);
      load_tile<D, BN, 128>(sV + nb * BN * D, gv, p.ld_qkv, (j + 1) * BN, S);
      if (threadIdx.x < BN) {
        const int kk = (j + 1) * BN + threadIdx.x;
        sMask[nb][threadIdx.x] = (kk < S) && (!gmask || gmask[kk]);
      }
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* cK = sK + buf * BN * D;
    const __nv_bfloat16* cV = sV + buf * BN * D;

    // S = Q K^T  (16 x 64 per warp)
    float acc_s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc_s[i][jj] = 0.f;
#pragma unroll
    for (int kc = 0; kc < D / 16; ++kc) {
      uint32_t a[4];
      load_a<D>(a, sQ, warp * 16, kc, lane);
#pragma unroll
      for (int np = 0; np < BN / 16; ++np) {
        uint32_t bb[4];
        load_b_nk<D>(bb, cK, np * 16, kc, lane);
        mma_bf16(acc_s[2 * np], a, bb[0], bb[1]);
        mma_bf16(acc_s[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    // mask + online softmax
    const int kv0 = j * BN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + t4 * 2 + (e & 1);
        const int row = row_lo + (e >> 1) * 8;
        const bool vis = (kv0 + col <= row) && sMask[buf][col];
        const float sv = vis ? acc_s[nt][e] : -INFINITY;
        acc_s[nt][e] = sv;
        mx[e >> 1] = fmaxf(mx[e >> 1], sv);
      }
    float corr[2], msub[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = quad_max(mx[r]);
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = exp2f((m_run[r] - m_use) * sl2);  // m_run = -inf -> 0
      msub[r] = m_use * sl2;
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      acc_o[i][0] *= corr[0]; acc_o[i][1] *= corr[0];
      acc_o[i][2] *= corr[1]; acc_o[i][3] *= corr[1];
    }
    uint32_t pa[BN / 16][4];
#pragma unroll
    for (int nt = 0; nt < BN / 8; ++nt) {
      float e0 = exp2f(acc_s[nt][0] * sl2 - msub[0]);
      float e1 = exp2f(acc_s[nt][1] * sl2 - msub[0]);
      float e2 = exp2f(acc_s[nt][2] * sl2 - msub[1]);
      float e3 = exp2f(acc_s[nt][3] * sl2 - msub[1]);
      l_run[0] += e0 + e1;
      l_run[1] += e2 + e3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(e0, e1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(e2, e3);
    }
    // O += P V
#pragma unroll
    for (int kc = 0; kc < BN / 16; ++kc) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t bb[4];
        load_b_kn<D>(bb, cV, kc * 16, np * 16, lane);
        mma_bf16(acc_o[2 * np], pa[kc], bb[0], bb[1]);
        mma_bf16(acc_o[2 * np + 1], pa[kc], bb[2], bb[3]);
      }
    }
    __syncthreads();
  }

  // epilogue: normalise, stage through sQ, coalesced store; LSE
  float inv[2];
  bool qvalid[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_lo + r * 8;
    l_run[r] = quad_sum(l_run[r]);
    qvalid[r] = row < S && (!gmask || gmask[row]);
    const bool live = qvalid[r] && l_run[r] > 0.f;
    inv[r] = live ? 1.f / l_run[r] : 0.f;
    if (t4 == 0 && row < S) {
      const float lse = live ? m_run[r] * p.scale + logf(l_run[r]) : INFINITY;
      p.lse[(int64_t(b) * p.H + hd) * S + row] = lse;
    }
  }
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
    const int rl = warp * 16 + g;
    const int col = nt * 8 + t4 * 2;
    *reinterpret_cast<uint32_t*>(sQ + swz<D>(rl, col >> 3) + (col & 7)) =
        pack_bf16x2(acc_o[nt][0] * inv[0], acc_o[nt][1] * inv[0]);
    *reinterpret_cast<uint32_t*>(sQ + swz<D>(rl + 8, col >> 3) + (col & 7)) =
        pack_bf16x2(acc_o[nt][2] * inv[1], acc_o[nt][3] * inv[1]);
  }
  __syncthreads();
  constexpr int CP = D / 8;
  __nv_bfloat16* go = p.o + tok0 * p.ld_o + hd * D;
  for (int idx = threadIdx.x; idx < BM * CP; idx += 128) {
    const int r = idx / CP, c = idx % CP;
    if (q0 + r < S)
      *reinterpret_cast<uint4*>(go + int64_t(q0 + r) * p.ld_o + c * 8) = *reinterpret_cast<const uint4*>(sQ + swz<D>(r, c));
  }
}

// ======================================================================================================== delta
// delta[b,h,s] = sum_d dO*O   (one warp per (token, head))
template <int D>
__global__ void attn_delta_kernel(AttnParams p) {
  const int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total = int64_t(p.B) * p.S * p.H;
  if (w >= total) return;
  const int hd = int(w % p.H);
  const int64_t tok = w / p.H;
  const __nv_bfloat16* o = p.o + tok * p.ld_o + hd * D;
  const __nv_bfloat16* d = p.d_o + tok * p.ld_o + hd * D;
  float acc = 0.f;
  for (int i = lane * 2; i < D; i += 64) {
    float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o + i));
    float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d + i));
    acc += a.x * c.x + a.y * c.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    const int b = int(tok / p.S), s = int(tok % p.S);
    p.delta[(int64_t(b) * p.H + hd) * p.S + s] = acc;
  }
}

// ======================================================================================================== bwd dQ
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(AttnParams p) {
  constexpr int BM = 64, BN = 64;
  extern __shared__ __align__(128) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sdO = sQ + BM * D;
  __nv_bfloat16* sK = sdO + BM * D;      // 2 buffers
  __nv_bfloat16* sV = sK + 2 * BN * D;   // 2 buffers
  __shared__ uint8_t sMask[2][BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * BM, hd = blockIdx.y, b = blockIdx.z;
  const int S = p.S;
  const int64_t tok0 = int64_t(b) * S;
  const __nv_bfloat16* gq = p.q + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gk = p.k + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gv = p.v + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gdo = p.d_o + tok0 * p.ld_o + hd * D;
  const uint8_t* gmask = p.mask ? p.mask + tok0 : nullptr;
  const int q_hi = min(q0 + BM, S);
  const int n_kv = (q_hi + BN - 1) / BN;

  load_tile<D, BM, 128>(sQ, gq, p.ld_qkv, q0, S);
  load_tile<D, BM, 128>(sdO, gdo, p.ld_o, q0, S);
  load_tile<D, BN, 128>(sK, gk, p.ld_qkv, 0, S);
  load_tile<D, BN, 128>(sV, gv, p.ld_qkv, 0, S);
  if (threadIdx.x < BN) sMask[0][threadIdx.x] = (threadIdx.x < S) && (!gmask || gmask[threadIdx.x]);
  cp_async_commit();

  const float sl2 = p.scale * LOG2E;
  const int row_lo = q0 + warp * 16 + g;
  float lse2[2], dlt[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_lo + r * 8;
    const int64_t o = (int64_t(b) * p.H + hd) * S + row;
    lse2[r] = row < S ? p.lse[o] * LOG2E : INFINITY;
    dlt[r] = row < S ? p.delta[o] : 0.f;
  }
  float acc_dq[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc_dq[i][jj] = 0.f;

  for (int j = 0; j < n_kv; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_kv) {
      const int nb = buf ^ 1;
      load_tile<D, BN, 128>(sK + nb * BN * D, gk, p.ld_qkv, (j + 1) * BN, S);
      load_tile<D, BN, 128>(sV + nb * BN * D, gv, p.ld_qkv, (j + 1) * BN, S);
      if (threadIdx.x < BN) {
        const int kk = (j + 1) * BN + threadIdx.x;
        sMask[nb][threadIdx.x] = (kk < S) && (!gmask || gmask[kk]);
      }
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* cK = sK + buf * BN * D;
    const __nv_bfloat16* cV = sV + buf * BN * D;

    float acc_s[BN / 8][4], acc_dp[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) { acc_s[i][jj] = 0.f; acc_dp[i][jj] = 0.f; }
#pragma unroll
    for (int kc = 0; kc < D / 16; ++kc) {
      uint32_t a[4], ad[4];
      load_a<D>(a, sQ, warp * 16, kc, lane);
      load_a<D>(ad, sdO, warp * 16, kc, lane);
#pragma unroll
      for (int np = 0; np < BN / 16; ++np) {
        uint32_t bb[4];
        load_b_nk<D>(bb, cK, np * 16, kc, lane);
        mma_bf16(acc_s[2 * np], a, bb[0], bb[1]);
        mma_bf16(acc_s[2 * np + 1], a, bb[2], bb[3]);
        load_b_nk<D>(bb, cV, np * 16, kc, lane);
        mma_bf16(acc_dp[2 * np], ad, bb[0], bb[1]);
        mma_bf16(acc_dp[2 * np + 1], ad, bb[2], bb[3]);
      }
    }
    const int kv0 = j * BN;
    uint32_t dsa[BN / 16][4];
#pragma unroll
    for (int nt = 0; nt < BN / 8; ++nt) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + t4 * 2 + (e & 1);
        const int row = row_lo + (e >> 1) * 8;
        const bool vis = (kv0 + col <= row) && sMask[buf][col];
        const float pr = vis ? exp2f(acc_s[nt][e] * sl2 - lse2[e >> 1]) : 0.f;
        ds[e] = pr * (acc_dp[nt][e] - dlt[e >> 1]);
      }
      dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
      dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // dQ += dS K
#pragma unroll
    for (int kc = 0; kc < BN / 16; ++kc) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t bb[4];
        load_b_kn<D>(bb, cK, kc * 16, np * 16, lane);
        mma_bf16(acc_dq[2 * np], dsa[kc], bb[0], bb[1]);
        mma_bf16(acc_dq[2 * np + 1], dsa[kc], bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
  // store dQ * scale through sQ
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
    const int rl = warp * 16 + g;
    const int col = nt * 8 + t4 * 2;
    *reinterpret_cast<uint32_t*>(sQ + swz<D>(rl, col >> 3) + (col & 7)) =
        pack_bf16x2(acc_dq[nt][0] * p.scale, acc_dq[nt][1] * p.scale);
    *reinterpret_cast<uint32_t*>(sQ + swz<D>(rl + 8, col >> 3) + (col & 7)) =
        pack_bf16x2(acc_dq[nt][2] * p.scale, acc_dq[nt][3] * p.scale);
  }
  __syncthreads();
  constexpr int CP = D / 8;
  __nv_bfloat16* gdq = p.dq + tok0 * p.ld_dqkv + hd * D;
  for (int idx = threadIdx.x; idx < BM * CP; idx += 128) {
    const int r = idx / CP, c = idx % CP;
    if (q0 + r < S)
      *reinterpret_cast<uint4*>(gdq + int64_t(q0 + r) * p.ld_dqkv + c * 8) =
          *reinterpret_cast<const uint4*>(sQ + swz<D>(r, c));
  }
}

// ======================================================================================================== bwd dK,dV
template <int D, int BMQ>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(AttnParams p) {
  constexpr int BN = 64;  // key rows per CTA (16 per warp)
  extern __shared__ __align__(128) uint8_t smem_attn[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sV = sK + BN * D;
  __nv_bfloat16* sQ = sV + BN * D;          // 2 buffers of [BMQ, D]
  __nv_bfloat16* sdO = sQ + 2 * BMQ * D;    // 2 buffers
  __shared__ float sLse[2][BMQ];
  __shared__ float sDelta[2][BMQ];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int kv0 = blockIdx.x * BN, hd = blockIdx.y, b = blockIdx.z;
  const int S = p.S;
  const int64_t tok0 = int64_t(b) * S;
  const __nv_bfloat16* gq = p.q + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gk = p.k + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gv = p.v + tok0 * p.ld_qkv + hd * D;
  const __nv_bfloat16* gdo = p.d_o + tok0 * p.ld_o + hd * D;
  const uint8_t* gmask = p.mask ? p.mask + tok0 : nullptr;
  const float* glse = p.lse + (int64_t(b) * p.H + hd) * S;
  const float* gdelta = p.delta + (int64_t(b) * p.H + hd) * S;

  // causal: only query rows >= kv0 see these keys
  const int qb0 = kv0 / BMQ;
  const int n_q = (S + BMQ - 1) / BMQ;

  load_tile<D, BN, 128>(sK, gk, p.ld_qkv, kv0, S);
  load_tile<D, BN, 128>(sV, gv, p.ld_qkv, kv0, S);
  load_tile<D, BMQ, 128>(sQ, gq, p.ld_qkv, qb0 * BMQ, S);
  load_tile<D, BMQ, 128>(sdO, gdo, p.ld_o, qb0 * BMQ, S);
  if (threadIdx.x < BMQ) {
    const int qq = qb0 * BMQ + threadIdx.x;
    sLse[0][threadIdx.x] = qq < S ? glse[qq] * LOG2E : INFINITY;
    sDelta[0][threadIdx.x] = qq < S ? gdelta[qq] : 0.f;
  }
  cp_async_commit();

  const float sl2 = p.scale * LOG2E;
  const int krow_lo = kv0 + warp * 16 + g;  // this thread's two key rows
  bool kvalid[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kr = krow_lo + r * 8;
    kvalid[r] = kr < S && (!gmask || gmask[kr]);
  }
  float acc_dk[D / 8][4], acc_dv[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { acc_dk[i][jj] = 0.f; acc_dv[i][jj] = 0.f; }

  for (int qb = qb0; qb < n_q; ++qb) {
    const int buf = (qb - qb0) & 1;
    if (qb + 1 < n_q) {
      const int nb = buf ^ 1;
      load_tile<D, BMQ, 128>(sQ + nb * BMQ * D, gq, p.ld_qkv, (qb + 1) * BMQ, S);
      load_tile<D, BMQ, 128>(sdO + nb * BMQ * D, gdo, p.ld_o, (qb + 1) * BMQ, S);
      if (threadIdx.x < BMQ) {
        const int qq = (qb + 1) * BMQ + threadIdx.x;
        sLse[nb][threadIdx.x] = qq < S ? glse[qq] * LOG2E : INFINITY;
        sDelta[nb][threadIdx.x] = qq < S ? gdelta[qq] : 0.f;
      }
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* cQ = sQ + buf * BMQ * D;
    const __nv_bfloat16* cdO = sdO + buf * BMQ * D;
    const int qq0 = qb * BMQ;

    // S^T = K Q^T   (16 key rows x BMQ query columns per warp)
    float acc_st[BMQ / 8][4];
#pragma unroll
    for (int i = 0; i < BMQ / 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc_st[i][jj] = 0.f;
#pragma unroll
    for (int kc = 0; kc < D / 16; ++kc) {
      uint32_t a[4];
      load_a<D>(a, sK, warp * 16, kc, lane);
#pragma unroll
      for (int np = 0; np < BMQ / 16; ++np) {
        uint32_t bb[4];
        load_b_nk<D>(bb, cQ, np * 16, kc, lane);
        mma_bf16(acc_st[2 * np], a, bb[0], bb[1]);
        mma_bf16(acc_st[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    // P^T (bf16 A fragments)
    uint32_t pta[BMQ / 16][4];
    float pt[BMQ / 8][4];
#pragma unroll
    for (int nt = 0; nt < BMQ / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qcol = nt * 8 + t4 * 2 + (e & 1);
        const int krow = krow_lo + (e >> 1) * 8;
        const bool vis = (krow <= qq0 + qcol) && kvalid[e >> 1];
        pt[nt][e] = vis ? exp2f(acc_st[nt][e] * sl2 - sLse[buf][qcol]) : 0.f;
      }
      pta[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pt[nt][0], pt[nt][1]);
      pta[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pt[nt][2], pt[nt][3]);
    }
    // dV += P^T dO
#pragma unroll
    for (int kc = 0; kc < BMQ / 16; ++kc) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t bb[4];
        load_b_kn<D>(bb, cdO, kc * 16, np * 16, lane);
        mma_bf16(acc_dv[2 * np], pta[kc], bb[0], bb[1]);
        mma_bf16(acc_dv[2 * np + 1], pta[kc], bb[2], bb[3]);
      }
    }
    // dP^T = V dO^T
    float acc_dpt[BMQ / 8][4];
#pragma unroll
    for (int i = 0; i < BMQ / 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc_dpt[i][jj] = 0.f;
#pragma unroll
    for (int kc = 0; kc < D / 16; ++kc) {
      uint32_t a[4];
      load_a<D>(a, sV, warp * 16, kc, lane);
#pragma unroll
      for (int np = 0; np < BMQ / 16; ++np) {
        uint32_t bb[4];
        load_b_nk<D>(bb, cdO, np * 16, kc, lane);
        mma_bf16(acc_dpt[2 * np], a, bb[0], bb[1]);
        mma_bf16(acc_dpt[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    // dS^T = P^T o (dP^T - delta[q])
    uint32_t dsa[BMQ / 16][4];
#pragma unroll
    for (int nt = 0; nt < BMQ / 8; ++nt) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qcol = nt * 8 + t4 * 2 + (e & 1);
        ds[e] = pt[nt][e] * (acc_dpt[nt][e] - sDelta[buf][qcol]);
      }
      dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
      dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // dK += dS^T Q
#pragma unroll
    for (int kc = 0; kc < BMQ / 16; ++kc) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t bb[4];
        load_b_kn<D>(bb, cQ, kc * 16, np * 16, lane);
        mma_bf16(acc_dk[2 * np], dsa[kc], bb[0], bb[1]);
        mma_bf16(acc_dk[2 * np + 1], dsa[kc], bb[2], bb[3]);
      }
    }
    __syncthreads();
  }

  // store dK*scale, dV through sK / sV (free now)
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
    const int rl = warp * 16 + g;
    const int col = nt * 8 + t4 * 2;
    *reinterpret_cast<uint32_t*>(sK + swz<D>(rl, col >> 3) + (col & 7)) =
        pack_bf16x2(acc_dk[nt][0] * p.scale, acc_dk[nt][1] * p.scale);
    *reinterpret_cast<uint32_t*>(sK + swz<D>(rl + 8, col >> 3) + (col & 7)) =
        pack_bf16x2(acc_dk[nt][2] * p.scale, acc_dk[nt][3] * p.scale);
    *reinterpret_cast<uint32_t*>(sV + swz<D>(rl, col >> 3) + (col & 7)) = pack_bf16x2(acc_dv[nt][0], acc_dv[nt][1]);
    *reinterpret_cast<uint32_t*>(sV + swz<D>(rl + 8, col >> 3) + (col & 7)) = pack_bf16x2(acc_dv[nt][2], acc_dv[nt][3]);
  }
  __syncthreads();
  constexpr int CP = D / 8;
  __nv_bfloat16* gdk = p.dk + tok0 * p.ld_dqkv + hd * D;
  __nv_bfloat16* gdv = p.dv + tok0 * p.ld_dqkv + hd * D;
  for (int idx = threadIdx.x; idx < BN * CP; idx += 128) {
    const int r = idx / CP, c = idx % CP;
    if (kv0 + r < S) {
      *reinterpret_cast<uint4*>(gdk + int64_t(kv0 + r) * p.ld_dqkv + c * 8) = *reinterpret_cast<const uint4*>(sK + swz<D>(r, c));
      *reinterpret_cast<uint4*>(gdv + int64_t(kv0 + r) * p.ld_dqkv + c * 8) = *reinterpret_cast<const uint4*>(sV + swz<D>(r, c));
    }
  }
}

template <int D>
static int launch_fwd(const AttnParams& p, cudaStream_t s) {
  const int smem = (64 + 4 * 64) * D * 2;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "attn_fwd smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  dim3 grid((p.S + 63) / 64, p.H, p.B);
  attn_fwd_kernel<D><<<grid, 128, smem, s>>>(p);
  MLA_CHECK_LAUNCH("attn_fwd");
  return MLA_OK;
}

template <int D>
static int launch_bwd(const AttnParams& p, cudaStream_t s) {
  constexpr int BMQ = D >= 128 ? 32 : 64;
  const int smem_dq = (2 * 64 + 4 * 64) * D * 2;
  const int smem_dkv = (2 * 64 + 4 * BMQ) * D * 2;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<D, BMQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "attn_bwd smem attr: %s", cudaGetErrorString(e));
    done = true;
  }
  const int64_t rows = int64_t(p.B) * p.S * p.H;
  attn_delta_kernel<D><<<unsigned((rows * 32 + 255) / 256), 256, 0, s>>>(p);
  MLA_CHECK_LAUNCH("attn_delta");
  dim3 grid((p.S + 63) / 64, p.H, p.B);
  attn_bwd_dkv_kernel<D, BMQ><<<grid, 128, smem_dkv, s>>>(p);
  MLA_CHECK_LAUNCH("attn_bwd_dkv");
  attn_bwd_dq_kernel<D><<<grid, 128, smem_dq, s>>>(p);
  MLA_CHECK_LAUNCH("attn_bwd_dq");
  return MLA_OK;
}

}  // namespace mla

using namespace mla;

static int attn_check(const mla_attn_args* a, bool bwd) {
  if (!a) return set_error(MLA_ERR_ARG, "attention: null args");
  if (a->batch <= 0 || a->seq <= 0 || a->heads <= 0) return set_error(MLA_ERR_ARG, "attention: empty problem");
  if (a->head_dim != 32 && a->head_dim != 64 && a->head_dim != 128)
    return set_error(MLA_ERR_ARG, "attention: head_dim %d unsupported (32, 64, 128)", a->head_dim);
  if ((a->ld_qkv & 7) || (a->ld_o & 7) || (bwd && (a->ld_dqkv & 7)))
    return set_error(MLA_ERR_ARG, "attention: row pitches must be multiples of 8 elements");
  if (!a->q || !a->k || !a->v || !a->o || !a->lse) return set_error(MLA_ERR_ARG, "attention: null tensor");
  if (bwd && (!a->d_o || !a->delta || !a->dq || !a->dk || !a->dv)) return set_error(MLA_ERR_ARG, "attention bwd: null tensor");
  return MLA_OK;
}

static AttnParams to_params(const mla_attn_args* a) {
  AttnParams p;
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.ld_qkv = a->ld_qkv; p.o = (__nv_bfloat16*)a->o; p.ld_o = a->ld_o; p.lse = (float*)a->lse;
  p.mask = (const uint8_t*)a->mask; p.B = a->batch; p.S = a->seq; p.H = a->heads; p.scale = a->scale;
  p.d_o = (const __nv_bfloat16*)a->d_o; p.delta = (float*)a->delta;
  p.dq = (__nv_bfloat16*)a->dq; p.dk = (__nv_bfloat16*)a->dk; p.dv = (__nv_bfloat16*)a->dv; p.ld_dqkv = a->ld_dqkv;
  return p;
}

extern "C" int mla_attn_fwd(const mla_attn_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  if (int rc = attn_check(a, false)) return rc;
  AttnParams p = to_params(a);
  auto s = (cudaStream_t)stream;
  switch (a->head_dim) {
    case 32: return launch_fwd<32>(p, s);
    case 64: return launch_fwd<64>(p, s);
    default: return launch_fwd<128>(p, s);
  }
}

extern "C" int mla_attn_bwd(const mla_attn_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  if (int rc = attn_check(a, true)) return rc;
  AttnParams p = to_params(a);
  auto s = (cudaStream_t)stream;
  switch (a->head_dim) {
    case 32: return launch_bwd<32>(p, s);
    case 64: return launch_bwd<64>(p, s);
    default: return launch_bwd<128>(p, s);
  }
}
