// Inference denoise loop of the diffusion action head (MLA.predict_action_diff, models/mla/model_mla.py:592-775).
//
// The reference re-runs the whole 548-token sequence through the decoder at each of the 8 DDIM steps although only
// the [t | x_0..x_T] rows change between steps (under causal attention everything in front of them is identical).
// Here the prefix runs once through the training-path kernels (its post-RoPE K/V stay in the per-layer q|k|v buffers)
// and each DDIM step pushes only the few suffix rows through the 32 layers.  At 2..34 rows every linear is a pure
// weight-streaming problem (13.5 GB of bf16 weights per step), so the step is HBM-bound, not tensor-bound:
//   * gemv_bf16_kernel    — skinny GEMM: one warp per output column streams that weight row once with 16-byte loads,
//                           fp32 accumulators for up to 8 activation rows at a time (activations come from L1/L2)
//   * decode_attn_kernel  — the suffix queries against the cached K/V (bottom-right aligned causal mask, as
//                           flash-attn / modeling_llama.py:540-557), keys split over the warps of a CTA, online softmax
//   * ddim_step_kernel    — x_{t-1} of ddim_sample with eta = 0 (gaussian_diffusion.py:342-352,:522-571), in the
//                           reference's fp32 op order (no FMA contraction) so the update is bit-exact.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

__device__ __forceinline__ float d_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// ---------------------------------------------------------------------------------------------- skinny GEMM
// out[m, n] = bf16( bf16(sum_k x[m,k] w[n,k]) + residual[m,n] ),  m < M (any M; processed MB rows at a time), w [N,K]
// row-major (nn.Linear layout).  K % 8 == 0, all row pitches multiples of 8 elements, 16-byte aligned bases.
constexpr int GEMV_MB = 8;
constexpr int GEMV_WARPS = 8;
__global__ void __launch_bounds__(GEMV_WARPS * 32) gemv_bf16_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out,
    const __nv_bfloat16* __restrict__ res, int M, int N, int K, int64_t ldx, int64_t ldw, int64_t ldo, int64_t ldr) {
  const int n = blockIdx.x * GEMV_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const __nv_bfloat16* wr = w + int64_t(n) * ldw;
  const int chunks = K >> 3;
  for (int m0 = 0; m0 < M; m0 += GEMV_MB) {
    const int mb = M - m0 < GEMV_MB ? M - m0 : GEMV_MB;
    float acc[GEMV_MB];
#pragma unroll
    for (int r = 0; r < GEMV_MB; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int c = lane; c < chunks; c += 32) {
      float wf[8];
      // the first pass streams the weight row from HBM; later row blocks find it in L2
      unpack8(ld_stream16(wr + 8 * c), wf);
#pragma unroll
      for (int r = 0; r < GEMV_MB; ++r) {
        if (r < mb) {
          float xf[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(x + int64_t(m0 + r) * ldx + 8 * c)), xf);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[r] = fmaf(xf[e], wf[e], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < GEMV_MB; ++r) {
      if (r < mb) {
        const float s = d_wsum(acc[r]);
        if (lane == r) {
          float v = bf16_round(s);
          if (res) v += __bfloat162float(res[int64_t(m0 + r) * ldr + n]);
          out[int64_t(m0 + r) * ldo + n] = __float2bfloat16_rn(v);
        }
      }
    }
  }
}

// EPL consecutive bf16 (2*EPL bytes, naturally aligned) -> fp32
template <int EPL>
__device__ __forceinline__ void ld_epl(const __nv_bfloat16* p, float* f) {
  if constexpr (EPL == 8) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
  } else if constexpr (EPL == 4) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  } else if constexpr (EPL == 2) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
    f[0] = a.x; f[1] = a.y;
  } else {
    f[0] = __bfloat162float(p[0]);
  }
}

// ---------------------------------------------------------------------------------------------- decode attention
// q rows (b, i), i < Lq, are the LAST Lq positions of a length-Lk sequence whose K/V rows live at
// k/v + (b*Lk + j)*ldkv + h*D.  Query i sees keys j <= Lk - Lq + i.  One CTA per (i, h, b); warp w takes keys
// w, w+NW, ...; partial (max, sum, acc[D]) are merged through shared memory.  D = 32 * EPL.
constexpr int DEC_WARPS = 8;
template <int EPL>
__global__ void __launch_bounds__(DEC_WARPS * 32) decode_attn_kernel(
    const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
    const __nv_bfloat16* __restrict__ v, int64_t ldkv, __nv_bfloat16* __restrict__ o, int64_t ldo, int H, int Lq,
    int Lk, float scale) {
  constexpr int D = 32 * EPL;
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS];
  __shared__ float s_acc[DEC_WARPS][D];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkeys = Lk - Lq + i + 1;
  float qf[EPL], acc[EPL];
  const __nv_bfloat16* qr = q + (int64_t(b) * Lq + i) * ldq + int64_t(h) * D + lane * EPL;
#pragma unroll
  for (int e = 0; e < EPL; ++e) { qf[e] = __bfloat162float(qr[e]) * scale; acc[e] = 0.f; }   // scale folded into q
  float m = -INFINITY, l = 0.f;
  const int64_t base = int64_t(b) * Lk * ldkv + int64_t(h) * D + lane * EPL;
  constexpr int UN = 4;
  for (int j0 = warp; j0 < nkeys; j0 += DEC_WARPS * UN) {
    float s[UN], vf[UN][EPL];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * DEC_WARPS;
      s[u] = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) vf[u][e] = 0.f;
      if (j < nkeys) {
        float kf[EPL];
        ld_epl<EPL>(k + base + int64_t(j) * ldkv, kf);
        ld_epl<EPL>(v + base + int64_t(j) * ldkv, vf[u]);
#pragma unroll
        for (int e = 0; e < EPL; ++e) s[u] = fmaf(qf[e], kf[e], s[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) s[u] = d_wsum(s[u]);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (j0 + u * DEC_WARPS < nkeys) {
        const float mn = fmaxf(m, s[u]);
        const float corr = __expf(m - mn), p = __expf(s[u] - mn);
        l = l * corr + p;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = acc[e] * corr + p * vf[u][e];
        m = mn;
      }
    }
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[warp][lane * EPL + e] = acc[e];
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float mx = -INFINITY;
#pragma unroll
    for (int w2 = 0; w2 < DEC_WARPS; ++w2) mx = fmaxf(mx, s_m[w2]);
    float L = 0.f, a = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < DEC_WARPS; ++w2) {
      const float c = s_m[w2] == -INFINITY ? 0.f : __expf(s_m[w2] - mx);
      L += s_l[w2] * c;
      a += s_acc[w2][d] * c;
    }
    o[(int64_t(b) * Lq + i) * ldo + int64_t(h) * D + d] = __float2bfloat16_rn(L > 0.f ? a / L : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------- DDIM update
// coef f32 [4] = sqrt(1/ac_t), sqrt(1/ac_t - 1), sqrt(ac_prev), sqrt(1 - ac_prev) of the current respaced step.
//   pred_xstart = c0*x - c1*eps ; eps' = (c0*x - pred_xstart)/c1 ; out = pred_xstart*c2 + c3*eps'
// (the reference re-derives eps from pred_xstart, gaussian_diffusion.py:549; every op rounded on its own like ATen's).
template <typename EpsT>
__global__ void ddim_step_kernel(const float* __restrict__ x, const EpsT* __restrict__ eps,
                                 const float* __restrict__ coef, float* __restrict__ out, int64_t n) {
  const float c0 = coef[0], c1 = coef[1], c2 = coef[2], c3 = coef[3];
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    float e;
    if constexpr (sizeof(EpsT) == 4) e = eps[i]; else e = __bfloat162float(eps[i]);
    const float a = __fmul_rn(c0, x[i]);
    const float px = __fsub_rn(a, __fmul_rn(c1, e));
    const float e2 = __fdiv_rn(__fsub_rn(a, px), c1);
    out[i] = __fadd_rn(__fmul_rn(px, c2), __fmul_rn(c3, e2));
  }
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))

extern "C" int mla_gemv_bf16(const void* x, const void* w, void* out, const void* residual, int32_t m, int32_t n,
                             int32_t k, int64_t ldx, int64_t ldw, int64_t ldo, int64_t ldr, void* stream) {
  if (int rc = device_check()) return rc;
  if (m <= 0 || n <= 0) return MLA_OK;
  if (k <= 0 || (k & 7) || (ldx & 7) || (ldw & 7))
    return set_error(MLA_ERR_ARG, "gemv: k and the row pitches of x and w must be multiples of 8 (k=%d)", k);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15)
    return set_error(MLA_ERR_ARG, "gemv: x and w must be 16-byte aligned");
  if (m > 64) return set_error(MLA_ERR_ARG, "gemv: m=%d rows is a GEMM, use mla_gemm_bf16", m);
  gemv_bf16_kernel<<<(n + GEMV_WARPS - 1) / GEMV_WARPS, GEMV_WARPS * 32, 0, S_(stream)>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)out, (const __nv_bfloat16*)residual, m, n, k,
      ldx, ldw, ldo, ldr);
  MLA_CHECK_LAUNCH("gemv_bf16");
  return MLA_OK;
}

extern "C" int mla_decode_attn(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* o,
                               int64_t ldo, int32_t batch, int32_t heads, int32_t len_q, int32_t len_k,
                               int32_t head_dim, float scale, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || heads <= 0 || len_q <= 0) return MLA_OK;
  if (len_k < len_q) return set_error(MLA_ERR_ARG, "decode_attn: len_k=%d < len_q=%d", len_k, len_q);
  if ((ldkv & 7) || ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15))
    return set_error(MLA_ERR_ARG, "decode_attn: K/V must be 16-byte aligned with a row pitch that is a multiple of 8");
  dim3 grid(len_q, heads, batch);
#define MLA_DEC(EPL)                                                                                            \
  decode_attn_kernel<EPL><<<grid, DEC_WARPS * 32, 0, S_(stream)>>>(                                             \
      (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ldkv, (__nv_bfloat16*)o, ldo, \
      heads, len_q, len_k, scale)
  switch (head_dim) {
    case 32: MLA_DEC(1); break;
    case 64: MLA_DEC(2); break;
    case 128: MLA_DEC(4); break;
    case 256: MLA_DEC(8); break;
    default: return set_error(MLA_ERR_ARG, "decode_attn: head_dim %d not in {32, 64, 128, 256}", head_dim);
  }
#undef MLA_DEC
  MLA_CHECK_LAUNCH("decode_attn");
  return MLA_OK;
}

extern "C" int mla_ddim_step(const void* x, const void* eps, int32_t eps_is_f32, const void* coef, void* out, int64_t n,
                             void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  const int grid = int((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  if (eps_is_f32)
    ddim_step_kernel<float><<<grid, 256, 0, S_(stream)>>>((const float*)x, (const float*)eps, (const float*)coef,
                                                          (float*)out, n);
  else
    ddim_step_kernel<__nv_bfloat16><<<grid, 256, 0, S_(stream)>>>((const float*)x, (const __nv_bfloat16*)eps,
                                                                  (const float*)coef, (float*)out, n);
  MLA_CHECK_LAUNCH("ddim_step");
  return MLA_OK;
}
