// Generic (non-causal, any head_dim, Lq != Lk) multi-head attention core, forward and backward, for the
// post-training generation heads (reference: models/mla/generation/models.py — nn.TransformerDecoder self/cross
// attention at d_model 4096 / 8 heads = head_dim 512, nn.MultiheadAttention at 1024 / 8 = 128, tactile 4096 / 4 = 1024).
// These attentions are <0.1 % of the step's FLOPs (a few hundred query rows), so this is a plain SIMT flash-style
// kernel: bf16 in / out, fp32 math, online softmax, K/V tiles staged in shared memory, one warp per query row
// (forward, dQ) or per key row (dK/dV) with the 32 lanes spread over the rows of the staged tile.
//
//   S = scale * Q K^T ;  P = softmax(S) ;  O = (keep ? P / (1-p_drop) : 0) V        (keep: optional dropout mask)
//
// Row addressing: q row (b,i) = q + (b*Lq + i)*ldq + h*D ; k/v row (b,j) = k + (b*Lk + j)*ldk + h*D — so packed
// in_proj outputs ([rows, 3d] or [rows, 2d]) are consumed in place.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

struct MhaParams {
  const __nv_bfloat16 *q, *k, *v;
  int64_t ldq, ldk, ldv;
  __nv_bfloat16* o;
  int64_t ldo;
  float* lse;               // [B,H,Lq]
  const uint8_t* keep;      // [B,H,Lq,Lk] or null
  float keep_scale;         // 1/(1-p_drop)
  int B, H, Lq, Lk, D;
  float scale;
  // backward
  const __nv_bfloat16* d_o;
  int64_t ld_do;
  float* delta;             // [B,H,Lq]
  __nv_bfloat16 *dq, *dk, *dv;
  int64_t lddq, lddk, lddv;
};

constexpr int MHA_ROWS = 8;    // rows (queries or keys) per CTA = warps per CTA
constexpr int MHA_TILE = 32;   // staged rows per tile = lanes
constexpr int MHA_MAXP = 16;   // bf16 pairs per lane: head_dim <= 1024

// Stage `n` rows x D of a strided bf16 matrix into smem with a padded pitch of (D+2) elements (odd number of 4-byte
// words per row for D % 4 == 0, so the 32 lanes reading one column of 32 rows hit 32 different banks); rows >= n are
// zero filled.
__device__ __forceinline__ void stage_rows(uint32_t* dst, const __nv_bfloat16* src, int64_t ld, int n, int D) {
  const int wpr = D >> 1;            // words per row
  const int pitch = wpr + 1;
  for (int idx = threadIdx.x; idx < MHA_TILE * wpr; idx += blockDim.x) {
    const int r = idx / wpr, c = idx - r * wpr;
    uint32_t val = 0u;
    if (r < n) val = *reinterpret_cast<const uint32_t*>(src + int64_t(r) * ld + 2 * c);
    dst[r * pitch + c] = val;
  }
}

__device__ __forceinline__ float2 bf2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// lane-per-row dot product: sum_d vec[d] * tile[lane][d]   (vec fp32 in smem, broadcast reads)
__device__ __forceinline__ float row_dot(const float* vec, const uint32_t* tile, int lane, int D) {
  const int wpr = D >> 1, pitch = wpr + 1;
  const uint32_t* row = tile + lane * pitch;
  float acc = 0.f;
#pragma unroll 4
  for (int c = 0; c < wpr; ++c) {
    const float2 kv = bf2(row[c]);
    const float2 qv = *reinterpret_cast<const float2*>(vec + 2 * c);
    acc = fmaf(qv.x, kv.x, acc);
    acc = fmaf(qv.y, kv.y, acc);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------- forward
// MODE 0: forward (writes O, LSE).  MODE 1: backward dQ (reads O/dO/LSE, writes dQ and delta).
template <int MODE>
__global__ void __launch_bounds__(MHA_ROWS * 32) mha_q_kernel(MhaParams p) {
  extern __shared__ __align__(16) uint8_t mha_smem[];
  const int D = p.D, wpr = D >> 1, pitch = wpr + 1;
  uint32_t* sK = reinterpret_cast<uint32_t*>(mha_smem);
  uint32_t* sV = sK + MHA_TILE * pitch;
  float* sQ = reinterpret_cast<float*>(sV + MHA_TILE * pitch);     // [MHA_ROWS][D]
  float* sDO = sQ + MHA_ROWS * D;                                  // [MHA_ROWS][D]  (MODE 1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * MHA_ROWS + warp;
  const bool q_ok = i < p.Lq;
  const int npair = (wpr + 31) >> 5;      // pairs per lane
  const __nv_bfloat16* qrow = p.q + (int64_t(b) * p.Lq + i) * p.ldq + int64_t(h) * D;
  float* myQ = sQ + warp * D;
  float* myDO = sDO + warp * D;
  float delta = 0.f, lse = 0.f;
  for (int c = lane; c < wpr; c += 32) {
    float2 v = make_float2(0.f, 0.f);
    if (q_ok) v = bf2(*reinterpret_cast<const uint32_t*>(qrow + 2 * c));
    *reinterpret_cast<float2*>(myQ + 2 * c) = v;
    if (MODE == 1) {
      float2 d = make_float2(0.f, 0.f), o = make_float2(0.f, 0.f);
      if (q_ok) {
        d = bf2(*reinterpret_cast<const uint32_t*>(p.d_o + (int64_t(b) * p.Lq + i) * p.ld_do + int64_t(h) * D + 2 * c));
        o = bf2(*reinterpret_cast<const uint32_t*>(p.o + (int64_t(b) * p.Lq + i) * p.ldo + int64_t(h) * D + 2 * c));
      }
      *reinterpret_cast<float2*>(myDO + 2 * c) = d;
      delta += d.x * o.x + d.y * o.y;
    }
  }
  const int64_t stat = (int64_t(b) * p.H + h) * p.Lq + i;
  if (MODE == 1) {
    delta = wsum(delta);
    if (q_ok) {
      lse = p.lse[stat];
      if (lane == 0) p.delta[stat] = delta;
    }
  }
  float m = -INFINITY, l = 0.f;
  float2 acc[MHA_MAXP];
#pragma unroll
  for (int t = 0; t < MHA_MAXP; ++t) acc[t] = make_float2(0.f, 0.f);
  const uint8_t* keep = p.keep ? p.keep + stat * int64_t(p.Lk) : nullptr;

  for (int j0 = 0; j0 < p.Lk; j0 += MHA_TILE) {
    const int n = min(MHA_TILE, p.Lk - j0);
    __syncthreads();     // previous tile fully consumed (also orders the sQ/sDO writes on the first pass)
    stage_rows(sK, p.k + (int64_t(b) * p.Lk + j0) * p.ldk + int64_t(h) * D, p.ldk, n, D);
    stage_rows(sV, p.v + (int64_t(b) * p.Lk + j0) * p.ldv + int64_t(h) * D, p.ldv, n, D);
    __syncthreads();
    const bool j_ok = lane < n;
    float s = row_dot(myQ, sK, lane, D) * p.scale;
    float kp = 1.f;     // dropout multiplier of this (query, key)
    if (keep != nullptr && j_ok && q_ok) kp = keep[j0 + lane] ? p.keep_scale : 0.f;
    float w;            // what multiplies the staged row in the accumulation
    if (MODE == 0) {
      if (!j_ok) s = -INFINITY;
      const float m_new = fmaxf(m, wmax(s));
      const float corr = __expf(m - m_new);         // m = -inf on the first tile -> 0
      const float pj = j_ok ? __expf(s - m_new) : 0.f;
      l = l * corr + wsum(pj);
      m = m_new;
#pragma unroll
      for (int t = 0; t < MHA_MAXP; ++t) { acc[t].x *= corr; acc[t].y *= corr; }
      w = pj * kp;
    } else {
      const float pj = j_ok ? __expf(s - lse) : 0.f;
      const float dp = row_dot(myDO, sV, lane, D) * kp;
      w = pj * (dp - delta) * p.scale;              // dS (scaled): dQ += dS K
    }
    const uint32_t* src = MODE == 0 ? sV : sK;
    for (int j = 0; j < n; ++j) {
      const float wj = __shfl_sync(0xffffffffu, w, j);
      const uint32_t* row = src + j * pitch;
#pragma unroll
      for (int t = 0; t < MHA_MAXP; ++t) {
        if (t < npair) {
          const int c = lane + 32 * t;
          if (c < wpr) {
            const float2 vv = bf2(row[c]);
            acc[t].x = fmaf(wj, vv.x, acc[t].x);
            acc[t].y = fmaf(wj, vv.y, acc[t].y);
          }
        }
      }
    }
  }
  if (!q_ok) return;
  if (MODE == 0) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* orow = p.o + (int64_t(b) * p.Lq + i) * p.ldo + int64_t(h) * D;
#pragma unroll
    for (int t = 0; t < MHA_MAXP; ++t) {
      const int c = lane + 32 * t;
      if (t < npair && c < wpr) *reinterpret_cast<uint32_t*>(orow + 2 * c) = pack_bf16x2(acc[t].x * inv, acc[t].y * inv);
    }
    if (lane == 0) p.lse[stat] = m + __logf(l);
  } else {
    __nv_bfloat16* drow = p.dq + (int64_t(b) * p.Lq + i) * p.lddq + int64_t(h) * D;
#pragma unroll
    for (int t = 0; t < MHA_MAXP; ++t) {
      const int c = lane + 32 * t;
      if (t < npair && c < wpr) *reinterpret_cast<uint32_t*>(drow + 2 * c) = pack_bf16x2(acc[t].x, acc[t].y);
    }
  }
}

// ---------------------------------------------------------------------------------------------- dK / dV
// One warp per key row j; query rows (Q, dO) staged 32 at a time, lane i owns query j0+i.
__global__ void __launch_bounds__(MHA_ROWS * 32) mha_kv_bwd_kernel(MhaParams p) {
  extern __shared__ __align__(16) uint8_t mha_smem[];
  const int D = p.D, wpr = D >> 1, pitch = wpr + 1;
  uint32_t* sQ = reinterpret_cast<uint32_t*>(mha_smem);
  uint32_t* sDO = sQ + MHA_TILE * pitch;
  float* sK = reinterpret_cast<float*>(sDO + MHA_TILE * pitch);    // [MHA_ROWS][D]
  float* sV = sK + MHA_ROWS * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int j = blockIdx.x * MHA_ROWS + warp;
  const bool k_ok = j < p.Lk;
  const int npair = (wpr + 31) >> 5;
  float* myK = sK + warp * D;
  float* myV = sV + warp * D;
  for (int c = lane; c < wpr; c += 32) {
    float2 kk = make_float2(0.f, 0.f), vv = make_float2(0.f, 0.f);
    if (k_ok) {
      kk = bf2(*reinterpret_cast<const uint32_t*>(p.k + (int64_t(b) * p.Lk + j) * p.ldk + int64_t(h) * D + 2 * c));
      vv = bf2(*reinterpret_cast<const uint32_t*>(p.v + (int64_t(b) * p.Lk + j) * p.ldv + int64_t(h) * D + 2 * c));
    }
    *reinterpret_cast<float2*>(myK + 2 * c) = kk;
    *reinterpret_cast<float2*>(myV + 2 * c) = vv;
  }
  float2 dk[MHA_MAXP], dv[MHA_MAXP];
#pragma unroll
  for (int t = 0; t < MHA_MAXP; ++t) { dk[t] = make_float2(0.f, 0.f); dv[t] = make_float2(0.f, 0.f); }
  const int64_t stat0 = (int64_t(b) * p.H + h) * p.Lq;

  for (int i0 = 0; i0 < p.Lq; i0 += MHA_TILE) {
    const int n = min(MHA_TILE, p.Lq - i0);
    __syncthreads();
    stage_rows(sQ, p.q + (int64_t(b) * p.Lq + i0) * p.ldq + int64_t(h) * D, p.ldq, n, D);
    stage_rows(sDO, p.d_o + (int64_t(b) * p.Lq + i0) * p.ld_do + int64_t(h) * D, p.ld_do, n, D);
    __syncthreads();
    const bool i_ok = lane < n;
    float pi = 0.f, ds = 0.f;
    if (i_ok && k_ok) {
      const float s = row_dot(myK, sQ, lane, D) * p.scale;
      const float pr = __expf(s - p.lse[stat0 + i0 + lane]);
      float kp = 1.f;
      if (p.keep != nullptr) kp = p.keep[(stat0 + i0 + lane) * int64_t(p.Lk) + j] ? p.keep_scale : 0.f;
      const float dp = row_dot(myV, sDO, lane, D) * kp;
      pi = pr * kp;
      ds = pr * (dp - p.delta[stat0 + i0 + lane]) * p.scale;
    }
    for (int i = 0; i < n; ++i) {
      const float pw = __shfl_sync(0xffffffffu, pi, i);
      const float dw = __shfl_sync(0xffffffffu, ds, i);
      const uint32_t* qr = sQ + i * pitch;
      const uint32_t* dr = sDO + i * pitch;
#pragma unroll
      for (int t = 0; t < MHA_MAXP; ++t) {
        if (t < npair) {
          const int c = lane + 32 * t;
          if (c < wpr) {
            const float2 qq = bf2(qr[c]);
            const float2 dd = bf2(dr[c]);
            dk[t].x = fmaf(dw, qq.x, dk[t].x);
            dk[t].y = fmaf(dw, qq.y, dk[t].y);
            dv[t].x = fmaf(pw, dd.x, dv[t].x);
            dv[t].y = fmaf(pw, dd.y, dv[t].y);
          }
        }
      }
    }
  }
  if (!k_ok) return;
  __nv_bfloat16* dkr = p.dk + (int64_t(b) * p.Lk + j) * p.lddk + int64_t(h) * D;
  __nv_bfloat16* dvr = p.dv + (int64_t(b) * p.Lk + j) * p.lddv + int64_t(h) * D;
#pragma unroll
  for (int t = 0; t < MHA_MAXP; ++t) {
    const int c = lane + 32 * t;
    if (t < npair && c < wpr) {
      *reinterpret_cast<uint32_t*>(dkr + 2 * c) = pack_bf16x2(dk[t].x, dk[t].y);
      *reinterpret_cast<uint32_t*>(dvr + 2 * c) = pack_bf16x2(dv[t].x, dv[t].y);
    }
  }
}

static size_t mha_smem_bytes(int D) {
  const size_t pitch = size_t(D / 2 + 1);
  return 2 * MHA_TILE * pitch * 4 + 2 * size_t(MHA_ROWS) * D * 4;
}

static int mha_check(const mla_mha_args* a, bool bwd) {
  if (a == nullptr) return set_error(MLA_ERR_ARG, "mha: null args");
  if (a->batch <= 0 || a->heads <= 0 || a->len_q <= 0 || a->len_k <= 0)
    return set_error(MLA_ERR_ARG, "mha: empty problem");
  if (a->head_dim <= 0 || (a->head_dim & 3) || a->head_dim > 64 * MHA_MAXP)
    return set_error(MLA_ERR_ARG, "mha: head_dim must be a multiple of 4 and <= %d", 64 * MHA_MAXP);
  if ((a->ldq | a->ldk | a->ldv | a->ldo) & 1) return set_error(MLA_ERR_ARG, "mha: row strides must be even");
  if (!a->q || !a->k || !a->v || !a->o || !a->lse) return set_error(MLA_ERR_ARG, "mha: null tensor");
  if (bwd && (!a->d_o || !a->delta || !a->dq || !a->dk || !a->dv)) return set_error(MLA_ERR_ARG, "mha bwd: null tensor");
  return MLA_OK;
}

static MhaParams mha_params(const mla_mha_args* a) {
  MhaParams p;
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
  p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo; p.lse = (float*)a->lse;
  p.keep = (const uint8_t*)a->keep_mask; p.keep_scale = a->keep_scale;
  p.B = a->batch; p.H = a->heads; p.Lq = a->len_q; p.Lk = a->len_k; p.D = a->head_dim; p.scale = a->scale;
  p.d_o = (const __nv_bfloat16*)a->d_o; p.ld_do = a->ld_do; p.delta = (float*)a->delta;
  p.dq = (__nv_bfloat16*)a->dq; p.dk = (__nv_bfloat16*)a->dk; p.dv = (__nv_bfloat16*)a->dv;
  p.lddq = a->ld_dq; p.lddk = a->ld_dk; p.lddv = a->ld_dv;
  return p;
}

template <typename K>
static int mha_set_smem(K kern, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "mha smem attribute (%zu B): %s", bytes, cudaGetErrorString(e));
  return MLA_OK;
}

}  // namespace mla

using namespace mla;

extern "C" int mla_mha_fwd(const mla_mha_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  if (int rc = mha_check(a, false)) return rc;
  MhaParams p = mha_params(a);
  const size_t smem = mha_smem_bytes(p.D);
  if (int rc = mha_set_smem(mha_q_kernel<0>, smem)) return rc;
  dim3 grid(ceil_div(p.Lq, MHA_ROWS), p.H, p.B);
  mha_q_kernel<0><<<grid, MHA_ROWS * 32, smem, (cudaStream_t)stream>>>(p);
  MLA_CHECK_LAUNCH("mha_fwd");
  return MLA_OK;
}

extern "C" int mla_mha_bwd(const mla_mha_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  if (int rc = mha_check(a, true)) return rc;
  MhaParams p = mha_params(a);
  const size_t smem = mha_smem_bytes(p.D);
  if (int rc = mha_set_smem(mha_q_kernel<1>, smem)) return rc;
  if (int rc = mha_set_smem(mha_kv_bwd_kernel, smem)) return rc;
  dim3 gq(ceil_div(p.Lq, MHA_ROWS), p.H, p.B), gk(ceil_div(p.Lk, MHA_ROWS), p.H, p.B);
  mha_q_kernel<1><<<gq, MHA_ROWS * 32, smem, (cudaStream_t)stream>>>(p);      // dQ + delta
  MLA_CHECK_LAUNCH("mha_bwd_dq");
  mha_kv_bwd_kernel<<<gk, MHA_ROWS * 32, smem, (cudaStream_t)stream>>>(p);    // dK, dV (needs delta)
  MLA_CHECK_LAUNCH("mha_bwd_dkv");
  return MLA_OK;
}
