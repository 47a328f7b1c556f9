// Epilogue shared by the 1-CTA and the 2-CTA (cta_group::2) GEMM kernels: one thread owns one accumulator row (TMEM
// lane) and walks the tile's 256 columns in 32-column chunks (tcgen05.ld 32x32b.x32), applying the fused
// bias / activation / residual / rounding or the fp32 (+=) store.
#pragma once
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int GEMM_BN = 256;

struct GemmEpilogue {
  void* c;
  int64_t ldc;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  int64_t ldr;
  __nv_bfloat16* pre_act;
  int64_t ldp;
  float alpha;
  int c_dtype;      // 0 bf16, 1 fp32
  int accumulate;   // fp32 only: C += value
  int activation;   // MLA_ACT_*
  // fused RoPE (head_dim 128) on the leading rope_cols columns of a bf16 output: position = row % rope_seq
  const __nv_bfloat16* rope_cos;   // bf16 [rope_seq, 64]
  const __nv_bfloat16* rope_sin;
  int rope_seq;
  int rope_cols;                   // multiple of 256; 0 = off
  const int32_t* rope_pos;         // int32 [M] row -> position, or null (position = row % rope_seq)
  float* sumsq;                    // fp32 outputs only: += sum of squares of the values written (the global gradient norm
                                   // of clip_grad_norm_ without re-reading the gradients), or null
  // fused SwiGLU (CTA-pair kernel only): the tile holds 128 gate columns | the matching 128 up columns
  __nv_bfloat16* swiglu_out;       // bf16 [M, swiglu_f] = bf16(bf16(silu(gate)) * up)
  int64_t ld_swiglu;
  int swiglu_f;                    // columns of gate (= of up); 0 = off
  // fused SwiGLU BACKWARD (the input-gradient GEMM of the down projection, d_act = dy . Wd, N = f): the epilogue reads
  // gate|up, rounds d_act to bf16 and writes d(gate|up) and the re-materialised act instead of d_act itself
  const __nv_bfloat16* sb_gu;      // bf16 [M, 2f] (gate | up); null = off
  int64_t ld_sb_gu;
  __nv_bfloat16* sb_dgu;           // bf16 [M, 2f]
  int64_t ld_sb_dgu;
  __nv_bfloat16* sb_act;           // bf16 [M, f] or null
  int64_t ld_sb_act;
  int stream_stores;               // 1: C / fused outputs are written with the evict-first hint
  int l2_hints;                    // CTA-pair kernel: 1 = A loads evict_last / B loads evict_first, 2 = the other way round
};

// Output stores with the evict-first hint (st.global.cs): a GEMM's C tile is not re-read by this kernel, and at 0.4-0.8 GB
// per launch it would otherwise push the operand panels the other CTAs are still sharing out of the 126 MB L2.
__device__ __forceinline__ void ep_st16(void* p, const uint4& v, int cs) {
  if (cs) __stcs(reinterpret_cast<uint4*>(p), v); else *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void ep_st16f(void* p, const float4& v, int cs) {
  if (cs) __stcs(reinterpret_cast<float4*>(p), v); else *reinterpret_cast<float4*>(p) = v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MLA_ACT_RELU: return v > 0.f ? v : 0.f;
    case MLA_ACT_GELU_ERF: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case MLA_ACT_GELU_TANH: {
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      float u = k0 * (v + k1 * v * v * v);
      return 0.5f * v * (1.f + tanhf(u));
    }
    case MLA_ACT_SILU: return v / (1.f + __expf(-v));
    default: return v;
  }
}

// RoPE epilogue (modeling_llama.py:184-208 fused into the q|k|v projection): the 256-column tile holds two heads of
// 128; column j < 64 of a head pairs with column j + 64, so the chunks are walked in pairs (c, c + 2).  Rounding points
// of the bf16 path: linear -> bf16, x*cos -> bf16, rotate_half(x)*sin -> bf16, sum -> bf16.
__device__ __forceinline__ void gemm_store_tile_rope(const GemmEpilogue& ep, uint32_t taddr, int64_t row, int n0, int M) {
  const bool row_ok = row < M;
  const int pos = ep.rope_pos ? (row_ok ? ep.rope_pos[row] : 0) : int(row % ep.rope_seq);
#pragma unroll 1
  for (int pair = 0; pair < 4; ++pair) {
    const int c1 = (pair >> 1) * 4 + (pair & 1);     // chunks 0,1 (head 0) and 4,5 (head 1): first halves
    const int c2 = c1 + 2;                           // matching second halves
    uint32_t r1[32], r2[32];
    tmem_ld_32x32b_x32(taddr + c1 * 32, r1);
    tmem_ld_32x32b_x32(taddr + c2 * 32, r2);
    tmem_ld_wait();
    if (!row_ok) continue;
    const __nv_bfloat16* cs = ep.rope_cos + int64_t(pos) * 64 + (pair & 1) * 32;
    const __nv_bfloat16* sn = ep.rope_sin + int64_t(pos) * 64 + (pair & 1) * 32;
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.ldc + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const uint4 cq = *reinterpret_cast<const uint4*>(cs + j);
      const uint4 sq = *reinterpret_cast<const uint4*>(sn + j);
      const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&cq);
      const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&sq);
      float o1[8], o2[8];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 cf = __bfloat1622float2(ch[t]), sf = __bfloat1622float2(sh[t]);
        const float cc[2] = {cf.x, cf.y}, ss[2] = {sf.x, sf.y};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float x1 = bf16_round(__uint_as_float(r1[j + 2 * t + u]) * ep.alpha);
          const float x2 = bf16_round(__uint_as_float(r2[j + 2 * t + u]) * ep.alpha);
          o1[2 * t + u] = bf16_round(x1 * cc[u]) + bf16_round(-x2 * ss[u]);
          o2[2 * t + u] = bf16_round(x2 * cc[u]) + bf16_round(x1 * ss[u]);
        }
      }
      ep_st16(crow + c1 * 32 + j, make_uint4(pack_bf16x2(o1[0], o1[1]), pack_bf16x2(o1[2], o1[3]),
                                                                 pack_bf16x2(o1[4], o1[5]), pack_bf16x2(o1[6], o1[7])), ep.stream_stores);
      ep_st16(crow + c2 * 32 + j, make_uint4(pack_bf16x2(o2[0], o2[1]), pack_bf16x2(o2[2], o2[3]),
                                                                 pack_bf16x2(o2[4], o2[5]), pack_bf16x2(o2[6], o2[7])), ep.stream_stores);
    }
  }
}

// SwiGLU epilogue (LlamaMLP, modeling_llama.py:240, fused into the gate|up projection): tile tn of the CTA-pair kernel
// holds gate columns [128 tn, 128 tn + 128) in TMEM columns 0..127 and the up columns f + the same range in 128..255
// (the peer CTA's half of the B panel comes from the `up` rows), so chunk c pairs with chunk c + 4.  Rounding points of
// swiglu_fwd_kernel: projection -> bf16, silu -> bf16, product -> bf16.  c (the gate|up matrix) is optional.
__device__ __forceinline__ void gemm_store_tile_swiglu(const GemmEpilogue& ep, uint32_t taddr, int64_t row, int tn, int M) {
  const bool row_ok = row < M;
  const int g0 = tn * 128;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t rg[32], ru[32];
    tmem_ld_32x32b_x32(taddr + c * 32, rg);
    tmem_ld_32x32b_x32(taddr + (c + 4) * 32, ru);
    tmem_ld_wait();
    if (!row_ok) continue;
    float g[32], u[32], a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      g[j] = bf16_round(__uint_as_float(rg[j]) * ep.alpha);
      u[j] = bf16_round(__uint_as_float(ru[j]) * ep.alpha);
      a[j] = bf16_round(g[j] * (1.f / (1.f + __expf(-g[j])))) * u[j];
    }
    __nv_bfloat16* arow = ep.swiglu_out + row * ep.ld_swiglu + g0 + c * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8)
      ep_st16(arow + j, make_uint4(pack_bf16x2(a[j], a[j + 1]), pack_bf16x2(a[j + 2], a[j + 3]),
                                                      pack_bf16x2(a[j + 4], a[j + 5]), pack_bf16x2(a[j + 6], a[j + 7])), ep.stream_stores);
    if (ep.c != nullptr) {
      __nv_bfloat16* grow = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.ldc + g0 + c * 32;
      __nv_bfloat16* urow = grow + ep.swiglu_f;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        ep_st16(grow + j, make_uint4(pack_bf16x2(g[j], g[j + 1]), pack_bf16x2(g[j + 2], g[j + 3]),
                                                        pack_bf16x2(g[j + 4], g[j + 5]), pack_bf16x2(g[j + 6], g[j + 7])), ep.stream_stores);
        ep_st16(urow + j, make_uint4(pack_bf16x2(u[j], u[j + 1]), pack_bf16x2(u[j + 2], u[j + 3]),
                                                        pack_bf16x2(u[j + 4], u[j + 5]), pack_bf16x2(u[j + 6], u[j + 7])), ep.stream_stores);
      }
    }
  }
}

// SwiGLU-backward epilogue (the input-gradient GEMM of the down projection; N = f, a multiple of 32): per 32-column chunk
// the thread needs 64 bytes of gate and 64 bytes of up from global memory.  Those loads are issued ONE CHUNK AHEAD (for
// chunk c + 1 before the accumulator of chunk c is read and worked on), otherwise eight dependent DRAM round trips per
// tile make the epilogue slower than the tile's MMAs (measured: the un-prefetched version cost 10 ms per step).
// Same rounding points as d_act -> bf16 -> swiglu_bwd_kernel.
__device__ __forceinline__ void gemm_store_tile_swiglu_bwd(const GemmEpilogue& ep, uint32_t taddr, int64_t row, int n0, int M,
                                                           int N) {
  const bool row_ok = row < M;
  const int64_t rr = row_ok ? row : 0;                 // rows beyond M read row 0 (never stored)
  const __nv_bfloat16* gbase = ep.sb_gu + rr * ep.ld_sb_gu;
  uint4 gq[4], uq[4], gn[4], un[4];
  auto fetch = [&](int col0, uint4 (&g)[4], uint4 (&u)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      g[q] = *reinterpret_cast<const uint4*>(gbase + col0 + q * 8);
      u[q] = *reinterpret_cast<const uint4*>(gbase + N + col0 + q * 8);
    }
  };
  if (n0 < N) fetch(n0, gq, uq);
#pragma unroll 1
  for (int c = 0; c < GEMM_BN / 32; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 >= N) break;  // warp-uniform
    const bool more = (c + 1 < GEMM_BN / 32) && (col0 + 32 < N);
    if (more) fetch(col0 + 32, gn, un);
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + c * 32, r);
    tmem_ld_wait();
    if (row_ok) {
      __nv_bfloat16* dgrow = ep.sb_dgu + row * ep.ld_sb_dgu + col0;
      __nv_bfloat16* arow = ep.sb_act ? ep.sb_act + row * ep.ld_sb_act + col0 : nullptr;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gq[q]);
        const __nv_bfloat162* uh = reinterpret_cast<const __nv_bfloat162*>(&uq[q]);
        float dg[8], du[8], ac[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 gf = __bfloat1622float2(gh[t]), uf = __bfloat1622float2(uh[t]);
          swiglu_bwd_elem(gf.x, uf.x, bf16_round(__uint_as_float(r[q * 8 + 2 * t]) * ep.alpha), dg[2 * t], du[2 * t], ac[2 * t]);
          swiglu_bwd_elem(gf.y, uf.y, bf16_round(__uint_as_float(r[q * 8 + 2 * t + 1]) * ep.alpha), dg[2 * t + 1],
                          du[2 * t + 1], ac[2 * t + 1]);
        }
        ep_st16(dgrow + q * 8, make_uint4(pack_bf16x2(dg[0], dg[1]), pack_bf16x2(dg[2], dg[3]),
                                                             pack_bf16x2(dg[4], dg[5]), pack_bf16x2(dg[6], dg[7])), ep.stream_stores);
        ep_st16(dgrow + N + q * 8, make_uint4(pack_bf16x2(du[0], du[1]), pack_bf16x2(du[2], du[3]),
                                                                 pack_bf16x2(du[4], du[5]), pack_bf16x2(du[6], du[7])), ep.stream_stores);
        if (arow)
          ep_st16(arow + q * 8, make_uint4(pack_bf16x2(ac[0], ac[1]), pack_bf16x2(ac[2], ac[3]),
                                                              pack_bf16x2(ac[4], ac[5]), pack_bf16x2(ac[6], ac[7])), ep.stream_stores);
      }
    }
    if (more) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { gq[q] = gn[q]; uq[q] = un[q]; }
    }
  }
}

// taddr: TMEM address of this thread's warp-quarter and accumulator buffer; row: global output row of this thread.
__device__ __forceinline__ void gemm_store_tile(const GemmEpilogue& ep, uint32_t taddr, int64_t row, int n0, int M, int N) {
  constexpr int BN = GEMM_BN;
  if (ep.sb_gu != nullptr) {    // kernel-uniform
    gemm_store_tile_swiglu_bwd(ep, taddr, row, n0, M, N);
    return;
  }
  if (n0 < ep.rope_cols) {      // warp-uniform (whole tile inside the rotated q|k column block)
    gemm_store_tile_rope(ep, taddr, row, n0, M);
    return;
  }
  const bool row_ok = row < M;
  float ss = 0.f;                  // sum of squares of this thread's fp32 outputs (ep.sumsq)
  // kernel-uniform: plain bf16 output (alpha 1, optional bias), 16-byte aligned rows and bias
  const bool plain = ep.c_dtype == 0 && ep.alpha == 1.f && ep.pre_act == nullptr && ep.activation == MLA_ACT_NONE &&
                     ep.residual == nullptr && (ep.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.c) & 15) == 0 &&
                     (ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) && (n0 & 7) == 0;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 >= N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + c * 32, r);
    tmem_ld_wait();
    if (!row_ok) continue;
    const bool full = (col0 + 32 <= N);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
    if (ep.c_dtype == 1) {
      // fp32 output (weight gradients): optional accumulate, no activation path.
      float* crow = reinterpret_cast<float*>(ep.c) + row * ep.ldc + col0;
      if (full && (ep.ldc & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (ep.accumulate) {
            float4 p = *reinterpret_cast<const float4*>(crow + j);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          ep_st16f(crow + j, o, ep.stream_stores);
          ss += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
        }
      } else {
        #pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) {
            const float o = ep.accumulate ? crow[j] + v[j] : v[j];
            crow[j] = o;
            ss += o * o;
          }
      }
      continue;
    }
    // bf16 output: replicate the reference's rounding points (linear -> bf16, act -> bf16, +residual -> bf16).
    if (plain && full) {
      // no activation / residual / pre-activation copy: acc (+ bias) is rounded ONCE, straight into the packed store
      // (same bits as round-then-pack).  The four epilogue warps are one per scheduler, so their instruction count per
      // tile bounds the tall-skinny (HBM-bound) GEMMs: ~320 instructions per 32-column chunk on the general path.
      __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = __uint_as_float(r[j + t]);
        if (ep.bias != nullptr) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(ep.bias + col0 + j));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(h[t]);
            o[2 * t] += f.x;
            o[2 * t + 1] += f.y;
          }
        }
        ep_st16(crow + j, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                     pack_bf16x2(o[6], o[7])), ep.stream_stores);
      }
      continue;
    }
    if (ep.bias != nullptr) {
      if (full && (reinterpret_cast<uintptr_t>(ep.bias + col0) & 15) == 0) {
        // four 16-byte loads instead of 32 scalar ones: the tall-skinny biased GEMMs of the tokenizers are HBM-bound and
        // spent most of their time here (0.32-0.73 ms against cuBLAS's 0.12-0.14 ms at 1.3 M x 96..384)
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(ep.bias + col0 + j));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(h[t]);
            v[j + 2 * t] += f.x;
            v[j + 2 * t + 1] += f.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || col0 + j < N) v[j] += __bfloat162float(ep.bias[col0 + j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
    if (ep.pre_act != nullptr) {
      __nv_bfloat16* prow = ep.pre_act + row * ep.ldp + col0;
      if (full && (ep.ldp & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          ep_st16(prow + j, make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                                          pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7])), ep.stream_stores);
      } else {
        #pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) prow[j] = __float2bfloat16_rn(v[j]);
      }
    }
    if (ep.activation != MLA_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = bf16_round(apply_act(v[j], ep.activation));
    }
    if (ep.residual != nullptr) {
      const __nv_bfloat16* rrow = ep.residual + row * ep.ldr + col0;
      if (full && (ep.ldr & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 q = *reinterpret_cast<const uint4*>(rrow + j);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 f = __bfloat1622float2(h[t]);
            v[j + 2 * t] += f.x;
            v[j + 2 * t + 1] += f.y;
          }
        }
      } else {
        #pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) v[j] += __bfloat162float(rrow[j]);
      }
    }
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.ldc + col0;
    if (full && (ep.ldc & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        ep_st16(crow + j, make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                                        pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7])), ep.stream_stores);
    } else {
      #pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) crow[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if (ep.sumsq != nullptr) {       // kernel-uniform; every lane of the (converged) epilogue warp takes part
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0 && ss != 0.f) atomicAdd(ep.sumsq, ss);
  }
}

}  // namespace mla
