// Backward of the two encoder-free tokenizers.  Stage "pretrain" of the reference trains vision_tower_2d and
// vision_tower_3d (models/vlm/prismatic.py:427-434); the finetune / post-training stages freeze them, so none of
// these kernels is on the headline step.  The GEMMs in between (dgrad / wgrad of every 1x1 conv and linear) run on
// gemm_sm100.cu; what is here is the bandwidth-bound glue, each one pass over its operands:
//   * LocalAttention core backward (vision_tokenizer.py:40-45)
//   * LayerNorm backward on bf16 rows with the two residual gradients of the pooling path folded in
//   * train-mode BatchNorm backward for the row layouts of pointcloud.cu (reduce + apply), ReLU / residual masks fused
//   * max-pool-over-neighbours backward, neighbour-gather (index_points) backward.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

__device__ __forceinline__ float t_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result in every thread (blockDim.x <= 1024)
__device__ __forceinline__ float t_bsum(float v, float* red) {
  v = t_wsum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  return t_wsum(lane < nw ? red[lane] : 0.f);
}

// ------------------------------------------------------------------------------------------- LocalAttention bwd
// Forward (vision_splice.cu: local_attn_kernel): a_n = bf16(sum_d bf16(bf16(q_d*scale) * k_nd)), p = softmax_n(a)
// (fp32), out_d = sum_n p_n v_nd.  Backward, all in fp32 from the recomputed p, results rounded once to bf16:
//   dv_nd = p_n dout_d ; dp_n = sum_d dout_d v_nd ; ds_n = p_n (dp_n - sum_m p_m dp_m)
//   dq_d = scale * sum_n ds_n k_nd ; dk_nd = ds_n * bf16(q_d*scale).
// One warp per (window g, head).  win <= 16.
__global__ void local_attn_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                                      const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dq,
                                      __nv_bfloat16* __restrict__ dkv, int64_t G, int C, int heads, int win,
                                      float scale) {
  const int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= G * heads) return;
  const int hh = int(w % heads);
  const int64_t g = w / heads;
  const int dh = C / heads;
  const __nv_bfloat16* qh = q + g * C + hh * dh;
  const __nv_bfloat16* doh = dout + g * C + hh * dh;
  float a[16], dp[16];
  for (int n = 0; n < win; ++n) {
    const __nv_bfloat16* kr = kv + (g * win + n) * 2 * int64_t(C) + hh * dh;
    float acc = 0.f, accp = 0.f;
    for (int d = lane; d < dh; d += 32) {
      const float qs = bf16_round(__bfloat162float(qh[d]) * scale);
      acc += bf16_round(qs * __bfloat162float(kr[d]));
      accp += __bfloat162float(doh[d]) * __bfloat162float(kr[C + d]);
    }
    a[n] = bf16_round(t_wsum(acc));
    dp[n] = t_wsum(accp);
  }
  float mx = -INFINITY;
  for (int n = 0; n < win; ++n) mx = fmaxf(mx, a[n]);
  float den = 0.f;
  for (int n = 0; n < win; ++n) { a[n] = __expf(a[n] - mx); den += a[n]; }
  float dsum = 0.f;
  for (int n = 0; n < win; ++n) { a[n] /= den; dsum += a[n] * dp[n]; }
  for (int d = lane; d < dh; d += 32) {
    const float qs = bf16_round(__bfloat162float(qh[d]) * scale);
    const float go = __bfloat162float(doh[d]);
    float accq = 0.f;
    for (int n = 0; n < win; ++n) {
      const int64_t o = (g * win + n) * 2 * int64_t(C) + hh * dh + d;
      const float ds = a[n] * (dp[n] - dsum);
      accq += ds * __bfloat162float(kv[o]);
      dkv[o] = __float2bfloat16_rn(ds * qs);
      dkv[o + C] = __float2bfloat16_rn(a[n] * go);
    }
    dq[g * C + hh * dh + d] = __float2bfloat16_rn(accq * scale);
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm bwd (bf16 rows)
// x bf16 [rows, h] (the forward's input; statistics are recomputed exactly as layernorm_kernel does), dy bf16,
// w f32.  dx = bf16( rstd (g - mean(g) - xhat mean(g xhat)) [+ dres[row]] [+ dgrp[row / win] / win] ), g = dy w.
// dw += sum_rows dy xhat ; db += sum_rows dy  (fp32 atomics, one per column per CTA).
// The two optional addends are the other gradient paths into the same tensor of LocalAttention.forward: the
// `reduced_features +` residual (:46) and avg_pool2d's broadcast (:28).
constexpr int TLN_CPT = 16;
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
    const __nv_bfloat16* __restrict__ dres, const __nv_bfloat16* __restrict__ dgrp, int win,
    __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int64_t rows, int h, float eps) {
  __shared__ float red[32];
  float aw[TLN_CPT], ab[TLN_CPT];
#pragma unroll
  for (int k = 0; k < TLN_CPT; ++k) { aw[k] = 0.f; ab[k] = 0.f; }
  const float inv_win = dgrp ? 1.f / float(win) : 0.f;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const __nv_bfloat16* xr = x + row * h;
    const __nv_bfloat16* dr = dy + row * h;
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int k = 0; k < TLN_CPT; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < h) { const float v = __bfloat162float(xr[i]); s += v; ss += v * v; }
    }
    const float mean = t_bsum(s, red) / float(h);
    const float var = fmaxf(t_bsum(ss, red) / float(h) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < TLN_CPT; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < h) {
        const float gg = __bfloat162float(dr[i]) * w[i];
        const float xh = (__bfloat162float(xr[i]) - mean) * rstd;
        s1 += gg; s2 += gg * xh;
      }
    }
    s1 = t_bsum(s1, red) / float(h);
    s2 = t_bsum(s2, red) / float(h);
#pragma unroll
    for (int k = 0; k < TLN_CPT; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < h) {
        const float d = __bfloat162float(dr[i]);
        const float xh = (__bfloat162float(xr[i]) - mean) * rstd;
        float o = rstd * (d * w[i] - s1 - xh * s2);
        if (dres) o += __bfloat162float(dres[row * h + i]);
        if (dgrp) o += __bfloat162float(dgrp[(row / win) * h + i]) * inv_win;
        dx[row * h + i] = __float2bfloat16_rn(o);
        aw[k] += d * xh;
        ab[k] += d;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < TLN_CPT; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < h) { atomicAdd(dw + i, aw[k]); atomicAdd(db + i, ab[k]); }
  }
}

// ------------------------------------------------------------------------------------------- BatchNorm bwd (train mode)
// y bf16 [rows, C] = the conv output the forward normalised with coef = (mean | invstd) (pointcloud.cu).
// Upstream gradient `up` is f32 or bf16; it is first masked:
//   mode 0: g = up
//   mode 1: g = up * [bn(y) > 0]                      — BN followed by ReLU (Linear1Layer, Linear2Layer.net1)
//   mode 2: g = up * [xnew > 0], rounded to bf16      — Linear2Layer: relu(bn(y) + x), xnew f32 = the block output;
//           the unrounded g is also the gradient of the residual input x and is written to dres (f32).
// reduce: sums[c] += sum_r g ; sums[C+c] += sum_r g * xhat.     apply: dy = w inv (g - s1/R - xhat s2/R) (bf16).
// dw = sums[C..2C), db = sums[0..C) — read by the caller once the reduce is done.
template <typename UpT>
__device__ __forceinline__ float bn_masked_grad(const UpT* __restrict__ up, const __nv_bfloat16* __restrict__ y,
                                                const float* __restrict__ xnew, int mode, int64_t o, float mean,
                                                float inv, float ww, float bb, float& xhat) {
  xhat = (__bfloat162float(y[o]) - mean) * inv;
  float g;
  if constexpr (sizeof(UpT) == 4) g = up[o]; else g = __bfloat162float(up[o]);
  if (mode == 1) {
    if (!(xhat * ww + bb > 0.f)) g = 0.f;
  } else if (mode == 2) {
    if (!(xnew[o] > 0.f)) g = 0.f;
  }
  return g;
}

template <typename UpT>
__global__ void bn_bwd_reduce_kernel(const UpT* __restrict__ up, const __nv_bfloat16* __restrict__ y,
                                     const float* __restrict__ coef, const float* __restrict__ w,
                                     const float* __restrict__ bias, const float* __restrict__ xnew, int mode,
                                     float* __restrict__ sums, int64_t rows, int C, int rows_per_block) {
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = coef[c], inv = coef[C + c], ww = w[c], bb = bias[c];
    float s1 = 0.f, s2 = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      float xh;
      float g = bn_masked_grad<UpT>(up, y, xnew, mode, r * C + c, mean, inv, ww, bb, xh);
      if (mode == 2) g = bf16_round(g);
      s1 += g; s2 += g * xh;
    }
    atomicAdd(sums + c, s1);
    atomicAdd(sums + C + c, s2);
  }
}

template <typename UpT>
__global__ void bn_bwd_apply_kernel(const UpT* __restrict__ up, const __nv_bfloat16* __restrict__ y,
                                    const float* __restrict__ coef, const float* __restrict__ w,
                                    const float* __restrict__ bias, const float* __restrict__ xnew, int mode,
                                    const float* __restrict__ sums, __nv_bfloat16* __restrict__ dy,
                                    float* __restrict__ dres, int64_t rows, int C) {
  const int64_t total = rows * C;
  const float inv_n = 1.f / float(rows);
  for (int64_t o = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; o < total; o += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(o % C);
    const float mean = coef[c], inv = coef[C + c], ww = w[c], bb = bias[c];
    float xh;
    float g = bn_masked_grad<UpT>(up, y, xnew, mode, o, mean, inv, ww, bb, xh);
    if (mode == 2) {
      if (dres) dres[o] = g;
      g = bf16_round(g);
    }
    dy[o] = __float2bfloat16_rn(ww * inv * (g - sums[c] * inv_n - xh * sums[C + c] * inv_n));
  }
}

// ------------------------------------------------------------------------------------------- max-pool bwd
// pooled[g, c] = max_k xnew[g, k, c] (Point_PN.py:157, x.max(-1)[0]).  d_xnew[g, k, c] = d_pooled[g, c] at the first k
// attaining the maximum, 0 elsewhere.  One CTA per group.
__global__ void maxpool_bwd_kernel(const float* __restrict__ xnew, const float* __restrict__ dpooled,
                                   float* __restrict__ dxnew, int K, int C) {
  const int64_t g = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mx = -INFINITY;
    int km = 0;
    for (int k = 0; k < K; ++k) {
      const float v = xnew[(g * K + k) * C + c];
      if (v > mx) { mx = v; km = k; }
    }
    const float d = dpooled[g * C + c];
    for (int k = 0; k < K; ++k) dxnew[(g * K + k) * C + c] = k == km ? d : 0.f;
  }
}

// ------------------------------------------------------------------------------------------- grouping bwd
// Forward (pointcloud.cu: group_pose_kernel): X[(b,g,k), 0..C) = feat[b, knn_idx[b,g,k]], X[(b,g,k), C..2C) =
// feat[b, fps_idx[b,g]], plus a parameter-free positional term.  d_feat f32 [B*N, C] (zero-filled by the caller):
//   d_feat[b, knn_idx[b,g,k], c] += dX[(b,g,k), c] ;  d_feat[b, fps_idx[b,g], c] += sum_k dX[(b,g,k), C + c].
__global__ void group_pose_bwd_kernel(const float* __restrict__ dx, const int32_t* __restrict__ fps_idx,
                                      const int32_t* __restrict__ knn_idx, float* __restrict__ dfeat, int N, int G,
                                      int K, int C) {
  const int bg = blockIdx.x, b = bg / G;
  const int center = fps_idx[bg];
  const int out_dim = 2 * C;
  float* base = dfeat + int64_t(b) * N * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += dx[(int64_t(bg) * K + k) * out_dim + C + c];
    atomicAdd(base + int64_t(center) * C + c, acc);
  }
  for (int e = threadIdx.x; e < K * C; e += blockDim.x) {
    const int k = e / C, c = e % C;
    const int src = knn_idx[int64_t(bg) * K + k];
    atomicAdd(base + int64_t(src) * C + c, dx[(int64_t(bg) * K + k) * out_dim + c]);
  }
}

// out[r, c] = sum_s x[s*rows + r, ...]: folds the `parts` diagonal blocks of a block-expanded wgrad back together.
// x f32 [parts*m, parts*n] row-major; out f32 [m, n] = sum_s x[s*m + i, s*n + j].
__global__ void diag_block_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int m, int n, int parts) {
  const int64_t total = int64_t(m) * n;
  const int64_t ld = int64_t(parts) * n;
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int i = int(e / n), j = int(e % n);
    float acc = 0.f;
    for (int s = 0; s < parts; ++s) acc += x[(int64_t(s) * m + i) * ld + int64_t(s) * n + j];
    out[e] = acc;
  }
}

static inline int ew_grid(int64_t total, int block = 256) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = int64_t(num_sms()) * 16;
  return int(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))

extern "C" int mla_local_attn_bwd(const void* q, const void* kv, const void* dout, void* dq, void* dkv, int64_t groups,
                                  int32_t c, int32_t heads, int32_t win, float scale, void* stream) {
  if (int rc = device_check()) return rc;
  if (groups <= 0) return MLA_OK;
  if (win > 16 || win < 1 || heads < 1 || c % heads)
    return set_error(MLA_ERR_ARG, "local_attn_bwd: window must be in [1,16] and channels divisible by heads");
  const int64_t warps = groups * heads;
  local_attn_bwd_kernel<<<unsigned((warps * 32 + 255) / 256), 256, 0, S_(stream)>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)kv, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dq,
      (__nv_bfloat16*)dkv, groups, c, heads, win, scale);
  MLA_CHECK_LAUNCH("local_attn_bwd");
  return MLA_OK;
}

extern "C" int mla_layernorm_bwd(const void* dy, const void* x, const void* w, const void* dres, const void* dgrp,
                                 int32_t win, void* dx, void* dw, void* db, int64_t rows, int32_t h, float eps,
                                 void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (h <= 0 || h > 256 * TLN_CPT) return set_error(MLA_ERR_ARG, "layernorm_bwd: h must be in (0, %d]", 256 * TLN_CPT);
  if (dgrp && (win < 1 || rows % win)) return set_error(MLA_ERR_ARG, "layernorm_bwd: rows must be a multiple of win");
  if (!dw || !db) return set_error(MLA_ERR_ARG, "layernorm_bwd: dw and db are required");
  const int grid = int(rows < int64_t(num_sms()) * 4 ? rows : int64_t(num_sms()) * 4);
  layernorm_bwd_kernel<<<grid, 256, 0, S_(stream)>>>(
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const float*)w, (const __nv_bfloat16*)dres,
      (const __nv_bfloat16*)dgrp, win, (__nv_bfloat16*)dx, (float*)dw, (float*)db, rows, h, eps);
  MLA_CHECK_LAUNCH("layernorm_bwd");
  return MLA_OK;
}

extern "C" int mla_bn_bwd(const void* up, int32_t up_is_f32, const void* y, const void* coef, const void* w,
                          const void* bias, const void* xnew, int32_t mode, void* sums, void* dy, void* dres,
                          int64_t rows, int32_t c, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (mode < 0 || mode > 2) return set_error(MLA_ERR_ARG, "bn_bwd: mode must be 0, 1 or 2");
  if (mode == 2 && !xnew) return set_error(MLA_ERR_ARG, "bn_bwd: mode 2 needs the block output");
  if (c <= 0 || c > 1024) return set_error(MLA_ERR_ARG, "bn_bwd: channels must be in (0, 1024]");
  cudaStream_t s = S_(stream);
  if (cudaMemsetAsync(sums, 0, 2 * size_t(c) * sizeof(float), s) != cudaSuccess)
    return set_error(MLA_ERR_CUDA, "bn_bwd: memset failed");
  const int target_blocks = num_sms() * 8;
  int rpb = int((rows + target_blocks - 1) / target_blocks);
  if (rpb < 8) rpb = 8;
  const int blocks = int((rows + rpb - 1) / rpb);
  const int threads = c >= 256 ? 256 : (c + 31) / 32 * 32;
  const int64_t total = rows * c;
  if (up_is_f32) {
    bn_bwd_reduce_kernel<float><<<blocks, threads, 0, s>>>((const float*)up, (const __nv_bfloat16*)y, (const float*)coef,
                                                           (const float*)w, (const float*)bias, (const float*)xnew, mode,
                                                           (float*)sums, rows, c, rpb);
    MLA_CHECK_LAUNCH("bn_bwd_reduce");
    bn_bwd_apply_kernel<float><<<ew_grid(total), 256, 0, s>>>((const float*)up, (const __nv_bfloat16*)y,
                                                              (const float*)coef, (const float*)w, (const float*)bias,
                                                              (const float*)xnew, mode, (const float*)sums,
                                                              (__nv_bfloat16*)dy, (float*)dres, rows, c);
    MLA_CHECK_LAUNCH("bn_bwd_apply");
  } else {
    bn_bwd_reduce_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(
        (const __nv_bfloat16*)up, (const __nv_bfloat16*)y, (const float*)coef, (const float*)w, (const float*)bias,
        (const float*)xnew, mode, (float*)sums, rows, c, rpb);
    MLA_CHECK_LAUNCH("bn_bwd_reduce");
    bn_bwd_apply_kernel<__nv_bfloat16><<<ew_grid(total), 256, 0, s>>>(
        (const __nv_bfloat16*)up, (const __nv_bfloat16*)y, (const float*)coef, (const float*)w, (const float*)bias,
        (const float*)xnew, mode, (const float*)sums, (__nv_bfloat16*)dy, (float*)dres, rows, c);
    MLA_CHECK_LAUNCH("bn_bwd_apply");
  }
  return MLA_OK;
}

extern "C" int mla_maxpool_bwd(const void* xnew, const void* dpooled, void* dxnew, int64_t groups, int32_t k, int32_t c,
                               void* stream) {
  if (int rc = device_check()) return rc;
  if (groups <= 0) return MLA_OK;
  if (k < 1 || c < 1) return set_error(MLA_ERR_ARG, "maxpool_bwd: empty neighbourhood");
  maxpool_bwd_kernel<<<(unsigned)groups, c >= 256 ? 256 : (c + 31) / 32 * 32, 0, S_(stream)>>>(
      (const float*)xnew, (const float*)dpooled, (float*)dxnew, k, c);
  MLA_CHECK_LAUNCH("maxpool_bwd");
  return MLA_OK;
}

extern "C" int mla_group_pose_bwd(const void* dx, const void* fps_idx, const void* knn_idx, void* dfeat, int32_t batch,
                                  int32_t n, int32_t groups, int32_t k, int32_t c, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || groups <= 0) return MLA_OK;
  group_pose_bwd_kernel<<<batch * groups, 256, 0, S_(stream)>>>((const float*)dx, (const int32_t*)fps_idx,
                                                               (const int32_t*)knn_idx, (float*)dfeat, n, groups, k, c);
  MLA_CHECK_LAUNCH("group_pose_bwd");
  return MLA_OK;
}

extern "C" int mla_diag_block_sum(const void* x, void* out, int32_t m, int32_t n, int32_t parts, void* stream) {
  if (int rc = device_check()) return rc;
  if (m <= 0 || n <= 0 || parts <= 0) return set_error(MLA_ERR_ARG, "diag_block_sum: empty problem");
  diag_block_sum_kernel<<<ew_grid(int64_t(m) * n), 256, 0, S_(stream)>>>((const float*)x, (float*)out, m, n, parts);
  MLA_CHECK_LAUNCH("diag_block_sum");
  return MLA_OK;
}
