// Kernels of the post-training generation heads (reference: models/mla/generation/models.py,
// gen_loss.py, utils.py and PrismaticVLM.compute_generation_losses, models/vlm/prismatic.py:771-838).
//
//   * fp32 LayerNorm forward / backward (CUDA autocast keeps layer_norm in fp32; the decoders' residual stream is fp32)
//   * small fp32 <-> bf16 helpers (add, cast, dropout-mask multiply, mean over the sequence)
//   * BatchNorm1d (train-mode batch statistics) + ReLU over rows, forward / backward
//   * ROI mask (scatter of the projected point-cloud patches + 3x3 dilation)
//   * the image head tail fused into ONE pass per patch: tanh/sigmoid heads, translation warp (affine_grid +
//     grid_sample bilinear/border/align_corners), ROI / non-ROI prediction, alpha blend and the three image losses;
//     and its backward (gradients w.r.t. the three raw head outputs)
//   * Chamfer-L2 (Euclidean cdist, min both ways) forward / backward
// All are bandwidth/latency-bound helpers around the tcgen05 GEMMs that do the heads' linear layers.
#include <cfloat>

#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

__device__ __forceinline__ float g_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result in every thread (blockDim.x <= 1024)
__device__ __forceinline__ float g_bsum(float v, float* red) {
  v = g_wsum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  return g_wsum(lane < nw ? red[lane] : 0.f);
}

// ------------------------------------------------------------------------------------------------ LayerNorm (fp32)
// y = (x - mean) * rstd * w + b ; x fp32 [rows, h]; optional bf16 copy of y for the next GEMM; saves mean / rstd.
__global__ void ln_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                  float* __restrict__ y, __nv_bfloat16* __restrict__ y16, float* __restrict__ mean_out,
                                  float* __restrict__ rstd_out, int h, float eps) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* xr = x + row * h;
  float s = 0.f;
  for (int i = threadIdx.x; i < h; i += blockDim.x) s += xr[i];
  const float mean = g_bsum(s, red) / float(h);
  float v = 0.f;
  for (int i = threadIdx.x; i < h; i += blockDim.x) { const float d = xr[i] - mean; v += d * d; }
  const float rstd = rsqrtf(g_bsum(v, red) / float(h) + eps);
  if (threadIdx.x == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  for (int i = threadIdx.x; i < h; i += blockDim.x) {
    const float o = (xr[i] - mean) * rstd * w[i] + b[i];
    y[row * h + i] = o;
    if (y16) y16[row * h + i] = __float2bfloat16_rn(o);
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w ; dw += sum_rows dy * xhat ; db += sum_rows dy.
// A CTA walks rows blockIdx.x, +gridDim.x, ...; thread t owns columns t, t+256, ... (h <= 256*LN_CPT).
constexpr int LN_CPT = 16;
__global__ void __launch_bounds__(256) ln_f32_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, float* __restrict__ dx,
                                                         float* __restrict__ dw, float* __restrict__ db, int64_t rows,
                                                         int h) {
  __shared__ float red[32];
  float aw[LN_CPT], ab[LN_CPT];
#pragma unroll
  for (int k = 0; k < LN_CPT; ++k) { aw[k] = 0.f; ab[k] = 0.f; }
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_CPT; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < h) {
        const float g = dy[row * h + i] * w[i];
        const float xh = (x[row * h + i] - mu) * rs;
        s1 += g;
        s2 += g * xh;
      }
    }
    s1 = g_bsum(s1, red) / float(h);
    s2 = g_bsum(s2, red) / float(h);
#pragma unroll
    for (int k = 0; k < LN_CPT; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < h) {
        const float d = dy[row * h + i];
        const float xh = (x[row * h + i] - mu) * rs;
        dx[row * h + i] = rs * (d * w[i] - s1 - xh * s2);
        aw[k] += d * xh;
        ab[k] += d;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LN_CPT; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < h) { atomicAdd(dw + i, aw[k]); atomicAdd(db + i, ab[k]); }
  }
}

// ------------------------------------------------------------------------------------------------ small helpers
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}
// out = a (fp32) + b (bf16)   [+ rounding of the sum to bf16 when the reference's operands are both bf16]
__global__ void add_f32_bf16_kernel(const float* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                    float* __restrict__ out, int64_t n, int round_bf16) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float s = a[i] + __bfloat162float(b[i]);
    out[i] = round_bf16 ? bf16_round(s) : s;
  }
}
// y = keep ? x * scale : 0  (bf16) — the dropout masks are drawn by torch's generator on the host side
__global__ void mask_scale_bf16_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ keep,
                                       __nv_bfloat16* __restrict__ y, int64_t n, int64_t per_mask, float scale) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16_rn(keep[i / per_mask] ? __bfloat162float(x[i]) * scale : 0.f);
}
// mean over the sequence: x bf16 [B, S, C] -> bf16 [B, C] (fp32 accumulation); backward broadcasts dy / S.
__global__ void seq_mean_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int S, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const __nv_bfloat16* p = x + int64_t(b) * S * C + c;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += __bfloat162float(p[int64_t(s) * C]);
  y[int64_t(b) * C + c] = __float2bfloat16_rn(acc / float(S));
}
__global__ void seq_mean_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int S, int C,
                                    int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    const int64_t b = i / (int64_t(S) * C);
    dx[i] = __float2bfloat16_rn(__bfloat162float(dy[b * C + c]) / float(S));
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm1d + ReLU
// x bf16 [R, C] (rows = batch*length, the reference's [B, C, L] transposed); train-mode statistics over the R rows.
// One CTA = 32 channels x 8 row lanes.  y = relu((x - mean) * rstd * w + b) in bf16; running stats updated with the
// unbiased variance (torch semantics).  Backward recomputes xhat from the saved mean / rstd.
__global__ void __launch_bounds__(256) bn_rows_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                          float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                          float* __restrict__ run_mean, float* __restrict__ run_var,
                                                          int R, int C, float eps, float momentum, int relu) {
  __shared__ float s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a = 0.f, q = 0.f;
  if (c < C)
    for (int r = ry; r < R; r += 8) { const float v = __bfloat162float(x[int64_t(r) * C + c]); a += v; q += v * v; }
  s1[ry][cx] = a; s2[ry][cx] = q;
  __syncthreads();
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sum += s1[k][cx]; sq += s2[k][cx]; }
  const float mean = sum / float(R);
  const float var = fmaxf(sq / float(R) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (c < C && ry == 0) {
    mean_out[c] = mean; rstd_out[c] = rstd;
    if (run_mean) {
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * var * (R > 1 ? float(R) / float(R - 1) : 1.f);
    }
  }
  if (c < C) {
    const float ww = w[c], bb = b[c];
    for (int r = ry; r < R; r += 8) {
      float o = bf16_round((__bfloat162float(x[int64_t(r) * C + c]) - mean) * rstd * ww + bb);
      if (relu) o = fmaxf(o, 0.f);
      y[int64_t(r) * C + c] = __float2bfloat16_rn(o);
    }
  }
}
__global__ void __launch_bounds__(256) bn_rows_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                          const __nv_bfloat16* __restrict__ y, const float* __restrict__ w,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
                                                          float* __restrict__ db, int R, int C, int relu) {
  __shared__ float s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const float mu = c < C ? mean[c] : 0.f, rs = c < C ? rstd[c] : 0.f, ww = c < C ? w[c] : 0.f;
  float a = 0.f, q = 0.f;
  if (c < C)
    for (int r = ry; r < R; r += 8) {
      float g = __bfloat162float(dy[int64_t(r) * C + c]);
      if (relu && !(__bfloat162float(y[int64_t(r) * C + c]) > 0.f)) g = 0.f;
      const float xh = (__bfloat162float(x[int64_t(r) * C + c]) - mu) * rs;
      a += g; q += g * xh;
    }
  s1[ry][cx] = a; s2[ry][cx] = q;
  __syncthreads();
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sg += s1[k][cx]; sgx += s2[k][cx]; }
  if (c < C && ry == 0) { dw[c] = sgx; db[c] = sg; }
  if (c < C)
    for (int r = ry; r < R; r += 8) {
      float g = __bfloat162float(dy[int64_t(r) * C + c]);
      if (relu && !(__bfloat162float(y[int64_t(r) * C + c]) > 0.f)) g = 0.f;
      const float xh = (__bfloat162float(x[int64_t(r) * C + c]) - mu) * rs;
      dx[int64_t(r) * C + c] = __float2bfloat16_rn(ww * rs * (g - sg / float(R) - xh * sgx / float(R)));
    }
}

// ------------------------------------------------------------------------------------------------ ROI mask
// create_roi_mask_from_indices + dilate_mask (generation/utils.py:41-70): scatter (row, col) of every projected point
// (valid or not — the reference does not apply the validity mask here) into a GxG grid, then a KxK max-pool.
__global__ void roi_mask_kernel(const int64_t* __restrict__ patch_idx, uint8_t* __restrict__ mask, int n_pts, int G,
                                int ksize) {
  extern __shared__ uint8_t grid_s[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < G * G; i += blockDim.x) grid_s[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
    const int64_t r = patch_idx[(int64_t(b) * n_pts + i) * 2], c = patch_idx[(int64_t(b) * n_pts + i) * 2 + 1];
    if (r >= 0 && r < G && c >= 0 && c < G) grid_s[r * G + c] = 1;
  }
  __syncthreads();
  const int pad = (ksize - 1) / 2;
  for (int i = threadIdx.x; i < G * G; i += blockDim.x) {
    const int r = i / G, c = i % G;
    uint8_t v = 0;
    for (int dr = -pad; dr <= pad; ++dr)
      for (int dc = -pad; dc <= pad; ++dc) {
        const int rr = r + dr, cc = c + dc;
        if (rr >= 0 && rr < G && cc >= 0 && cc < G) v |= grid_s[rr * G + cc];
      }
    mask[int64_t(b) * G * G + i] = v;
  }
}

// ------------------------------------------------------------------------------------------------ image head tail
// One CTA per patch n = (sample, gy, gx).  Pixel e = (c, y, x) of the patch lives at image[(b % n_img), c, gy*ps+y,
// gx*ps+x] (images_to_patches, generation/utils.py:7-20, folded into the addressing).
//
// Rounding points follow ImageGenerationModule.forward / _generate_generated_patches (generation/models.py:193-286)
// under bf16 autocast: the three heads produce bf16; tanh / sigmoid and the *clip products round to bf16; the patch
// arithmetic is fp32 except 0.95*delta (bf16) and (1 - alpha) (bf16).
struct GenImg {
  const float* cur;  const float* nxt;          // current / next image, fp32
  int64_t cur_sb, cur_sc, nxt_sb, nxt_sc;        // batch and channel strides (elements); row stride = W
  int n_img, W, ps, G;                           // distinct images, image width, patch size, patches per side
  const __nv_bfloat16* delta_raw; int64_t ld_delta;   // [N, >=3*ps*ps]
  const __nv_bfloat16* ao_raw; int64_t ld_ao;         // [N, >=3]: alpha, offset x, offset y (one fused GEMM)
  const uint8_t* roi;                            // [N]
  float delta_clip, max_shift, gen_weight;
  float* blended;                                // [N, 3*ps*ps] fp32 (image_generation)
  __nv_bfloat16* delta_all; __nv_bfloat16* alpha_all; __nv_bfloat16* offset_all;   // outputs dict (bf16)
  float* sums;                                   // [5]: sq_roi, abs_roi, abs_bg, abs_delta, n_roi_patches
  // backward
  const float* coef;                             // [4]: d/d(sq_roi), d/d(abs_roi), d/d(abs_bg), d/d(abs_delta)
  const float* gscale;                           // [1]: upstream gradient of the image loss
  __nv_bfloat16* d_delta_raw; __nv_bfloat16* d_ao_raw;
};

struct PatchGeom { float a, om, tx, ty, d_o_x, d_o_y, sig; bool roi; };

__device__ __forceinline__ PatchGeom patch_geom(const GenImg& g, int n) {
  PatchGeom p;
  const float ar = __bfloat162float(g.ao_raw[int64_t(n) * g.ld_ao]);
  const float oxr = __bfloat162float(g.ao_raw[int64_t(n) * g.ld_ao + 1]);
  const float oyr = __bfloat162float(g.ao_raw[int64_t(n) * g.ld_ao + 2]);
  p.roi = g.roi[n] != 0;
  p.sig = bf16_round(1.f / (1.f + __expf(-ar)));               // torch.sigmoid on bf16
  p.a = p.roi ? 1.f : p.sig;
  p.om = bf16_round(1.f - p.a);                                // (1.0 - alpha) in bf16
  p.d_o_x = bf16_round(tanhf(oxr));
  p.d_o_y = bf16_round(tanhf(oyr));
  const float offx = bf16_round(p.d_o_x * g.max_shift), offy = bf16_round(p.d_o_y * g.max_shift);
  // tx_norm = 2.0 * tx / (ps - 1) in bf16 (two roundings), then fp32 grid arithmetic
  const float txn = bf16_round(bf16_round(2.f * offx) / float(g.ps - 1));
  const float tyn = bf16_round(bf16_round(2.f * offy) / float(g.ps - 1));
  p.tx = txn * 0.5f * float(g.ps - 1);      // shift in pixels: ((x_base + txn + 1) / 2) * (ps-1) = x + tx
  p.ty = tyn * 0.5f * float(g.ps - 1);
  return p;
}

// bilinear sample of the patch at (x + tx, y + ty) with border clamping; returns value and d/d(ix), d/d(iy)
__device__ __forceinline__ float warp_sample(const float* base, int W, int ps, float fx, float fy, float& dvx, float& dvy) {
  const float hi = float(ps - 1);
  const bool inx = fx > 0.f && fx < hi, iny = fy > 0.f && fy < hi;     // clip_coordinates_set_grad: zero outside
  const float cx = fminf(fmaxf(fx, 0.f), hi), cy = fminf(fmaxf(fy, 0.f), hi);
  const int x0 = int(floorf(cx)), y0 = int(floorf(cy));
  const int x1 = min(x0 + 1, ps - 1), y1 = min(y0 + 1, ps - 1);
  const float wx = cx - float(x0), wy = cy - float(y0);
  const float v00 = base[int64_t(y0) * W + x0], v01 = base[int64_t(y0) * W + x1];
  const float v10 = base[int64_t(y1) * W + x0], v11 = base[int64_t(y1) * W + x1];
  dvx = inx ? ((v01 - v00) * (1.f - wy) + (v11 - v10) * wy) : 0.f;
  dvy = iny ? ((v10 - v00) * (1.f - wx) + (v11 - v01) * wx) : 0.f;
  return v00 * (1.f - wx) * (1.f - wy) + v01 * wx * (1.f - wy) + v10 * (1.f - wx) * wy + v11 * wx * wy;
}

template <int BWD>
__global__ void __launch_bounds__(256) gen_image_kernel(GenImg g) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const int per = g.G * g.G;
  const int b = n / per, pidx = n % per, gy = pidx / g.G, gx = pidx % g.G;
  const int ps = g.ps, pp = ps * ps, E = 3 * pp;
  const PatchGeom pg = patch_geom(g, n);
  const float* cur0 = g.cur + int64_t(b % g.n_img) * g.cur_sb + (int64_t(gy) * ps) * g.W + int64_t(gx) * ps;
  const float* nxt0 = g.nxt + int64_t(b % g.n_img) * g.nxt_sb + (int64_t(gy) * ps) * g.W + int64_t(gx) * ps;
  float sq = 0.f, ab = 0.f, ad = 0.f;           // forward partial sums
  float dA = 0.f, dtx = 0.f, dty = 0.f;         // backward per-patch reductions
  float c_sq = 0.f, c_abs = 0.f, c_bg = 0.f, c_del = 0.f;
  if (BWD) {
    const float up = g.gscale[0];
    c_sq = g.coef[0] * up; c_abs = g.coef[1] * up; c_bg = g.coef[2] * up; c_del = g.coef[3] * up;
  }
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int c = e / pp, rem = e - c * pp, y = rem / ps, x = rem - y * ps;
    const float* cb = cur0 + int64_t(c) * g.cur_sc;
    const float curv = cb[int64_t(y) * g.W + x];
    const float gt = nxt0[int64_t(c) * g.nxt_sc + int64_t(y) * g.W + x];
    const float dr = __bfloat162float(g.delta_raw[int64_t(n) * g.ld_delta + e]);
    const float th = bf16_round(tanhf(dr));
    const float delta = bf16_round(th * g.delta_clip);
    float pred, dvx = 0.f, dvy = 0.f;
    if (pg.roi) {
      pred = (1.f - g.gen_weight) * (curv + delta) + bf16_round(g.gen_weight * delta);
    } else {
      pred = warp_sample(cb, g.W, ps, float(x) + pg.tx, float(y) + pg.ty, dvx, dvy) + delta;
    }
    const float bl = pg.a * pred + pg.om * curv;
    const float diff = bl - gt;
    if (!BWD) {
      g.blended[int64_t(n) * E + e] = bl;
      g.delta_all[int64_t(n) * E + e] = __float2bfloat16_rn(delta);
      if (pg.roi) { sq += diff * diff; ab += fabsf(diff); } else { ab += fabsf(diff); }
      ad += fabsf(delta);
    } else {
      const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      const float gbl = pg.roi ? (c_sq * 2.f * diff + c_abs * sgn) : (c_bg * sgn);
      const float dpred = gbl * pg.a;
      if (!pg.roi) {
        dA += gbl * (pred - curv);
        dtx += dpred * dvx;
        dty += dpred * dvy;
      }
      const float sd = delta > 0.f ? 1.f : (delta < 0.f ? -1.f : 0.f);
      const float ddelta = dpred + c_del * sd;            // roi: (1-w) + w = 1 ; non-roi: 1
      g.d_delta_raw[int64_t(n) * g.ld_delta + e] = __float2bfloat16_rn(ddelta * g.delta_clip * (1.f - th * th));
    }
  }
  if (!BWD) {
    sq = g_bsum(sq, red); ab = g_bsum(ab, red); ad = g_bsum(ad, red);
    if (threadIdx.x == 0) {
      if (pg.roi) { atomicAdd(g.sums + 0, sq); atomicAdd(g.sums + 1, ab); atomicAdd(g.sums + 4, 1.f); }
      else atomicAdd(g.sums + 2, ab);
      atomicAdd(g.sums + 3, ad);
      g.alpha_all[n] = __float2bfloat16_rn(pg.sig);
      g.offset_all[2 * n] = __float2bfloat16_rn(bf16_round(pg.d_o_x * g.max_shift));
      g.offset_all[2 * n + 1] = __float2bfloat16_rn(bf16_round(pg.d_o_y * g.max_shift));
    }
  } else {
    dA = g_bsum(dA, red); dtx = g_bsum(dtx, red); dty = g_bsum(dty, red);
    if (threadIdx.x == 0) {
      // alpha = sigmoid(raw) (non-ROI only); offsets: pixels = max_shift * tanh(raw) (through the two bf16 roundings)
      g.d_ao_raw[int64_t(n) * g.ld_ao] = __float2bfloat16_rn(pg.roi ? 0.f : dA * pg.sig * (1.f - pg.sig));
      g.d_ao_raw[int64_t(n) * g.ld_ao + 1] = __float2bfloat16_rn(pg.roi ? 0.f : dtx * g.max_shift * (1.f - pg.d_o_x * pg.d_o_x));
      g.d_ao_raw[int64_t(n) * g.ld_ao + 2] = __float2bfloat16_rn(pg.roi ? 0.f : dty * g.max_shift * (1.f - pg.d_o_y * pg.d_o_y));
    }
  }
}

// compute_generation_losses (prismatic.py:779-816): roi = mse + 0.5*l1 over the ROI patches' pixels, bg = 0.01*l1 over
// the rest, delta reward = -0.1 * mean|delta_all| (a bf16 mean in the reference); terms with no elements are skipped.
// losses[4] = {image_gen_loss, roi, bg, delta}; coef[4] = the partial derivatives the backward kernel needs.
__global__ void gen_image_finalize_kernel(const float* __restrict__ sums, int n_patches, int E, float* __restrict__ losses,
                                          float* __restrict__ coef) {
  const float n_roi = sums[4];
  const float cnt_roi = n_roi * float(E), cnt_bg = (float(n_patches) - n_roi) * float(E);
  const float cnt_all = float(n_patches) * float(E);
  const float roi = cnt_roi > 0.f ? (sums[0] / cnt_roi + 0.5f * (sums[1] / cnt_roi)) : 0.f;
  const float bg = cnt_bg > 0.f ? 0.01f * (sums[2] / cnt_bg) : 0.f;
  const float dl = bf16_round(-0.1f * bf16_round(sums[3] / cnt_all));
  losses[0] = roi + bg + dl; losses[1] = roi; losses[2] = bg; losses[3] = dl;
  coef[0] = cnt_roi > 0.f ? 1.f / cnt_roi : 0.f;
  coef[1] = cnt_roi > 0.f ? 0.5f / cnt_roi : 0.f;
  coef[2] = cnt_bg > 0.f ? 0.01f / cnt_bg : 0.f;
  coef[3] = -0.1f / cnt_all;
}

// ------------------------------------------------------------------------------------------------ Chamfer-L2
// chamfer_distance_l2 (generation/gen_loss.py:12-18): d = cdist(pred, gt) (Euclidean); mean_b( mean_i min_j d +
// mean_j min_i d ).  pred bf16 [B, N1, 3]; gt fp32 [n_gt_batch, N2, 3] (sample b uses gt[b % n_gt_batch]).
// dir 0: one thread per pred point (min over gt); dir 1: one thread per gt point (min over pred).
__global__ void chamfer_min_kernel(const __nv_bfloat16* __restrict__ pred, const float* __restrict__ gt, int N1, int N2,
                                   int n_gt, int* __restrict__ idx_out, float* __restrict__ dist_out, int dir) {
  extern __shared__ float pts[];       // the "other" cloud of this sample, xyz interleaved
  const int b = blockIdx.y;
  const int n_self = dir == 0 ? N1 : N2, n_other = dir == 0 ? N2 : N1;
  for (int i = threadIdx.x; i < n_other * 3; i += blockDim.x)
    pts[i] = dir == 0 ? gt[(int64_t(b % n_gt) * N2) * 3 + i] : __bfloat162float(pred[(int64_t(b) * N1) * 3 + i]);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_self) return;
  float px, py, pz;
  if (dir == 0) {
    const __nv_bfloat16* p = pred + (int64_t(b) * N1 + i) * 3;
    px = __bfloat162float(p[0]); py = __bfloat162float(p[1]); pz = __bfloat162float(p[2]);
  } else {
    const float* p = gt + (int64_t(b % n_gt) * N2 + i) * 3;
    px = p[0]; py = p[1]; pz = p[2];
  }
  float best = FLT_MAX;
  int bi = 0;
  for (int j = 0; j < n_other; ++j) {
    const float dx = px - pts[3 * j], dy = py - pts[3 * j + 1], dz = pz - pts[3 * j + 2];
    const float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < best) { best = d2; bi = j; }
  }
  idx_out[int64_t(b) * n_self + i] = bi;
  dist_out[int64_t(b) * n_self + i] = sqrtf(best);
}
// loss = sum(d_fwd) / (B*N1) + sum(d_bwd) / (B*N2)
__global__ void chamfer_reduce_kernel(const float* __restrict__ d1, const float* __restrict__ d2, int64_t n1, int64_t n2,
                                      float* __restrict__ loss) {
  __shared__ float red[32];
  float a = 0.f, b = 0.f;
  for (int64_t i = threadIdx.x; i < n1; i += blockDim.x) a += d1[i];
  for (int64_t i = threadIdx.x; i < n2; i += blockDim.x) b += d2[i];
  a = g_bsum(a, red);
  b = g_bsum(b, red);
  if (threadIdx.x == 0) loss[0] = a / float(n1) + b / float(n2);
}
// d pred (fp32 scratch, zeroed by the caller): forward direction writes, backward direction scatter-adds.
__global__ void chamfer_bwd_kernel(const __nv_bfloat16* __restrict__ pred, const float* __restrict__ gt, int B, int N1,
                                   int N2, int n_gt, const int* __restrict__ idx1, const float* __restrict__ dist1,
                                   const int* __restrict__ idx2, const float* __restrict__ dist2,
                                   const float* __restrict__ gscale, float* __restrict__ dpred) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t n1 = int64_t(B) * N1, n2 = int64_t(B) * N2;
  const float gs = gscale[0];
  if (t < n1) {
    const int b = int(t / N1);
    const float d = dist1[t];
    if (d > 0.f) {
      const float* q = gt + (int64_t(b % n_gt) * N2 + idx1[t]) * 3;
      const float s = gs / (float(n1) * d);
      for (int k = 0; k < 3; ++k) atomicAdd(dpred + t * 3 + k, s * (__bfloat162float(pred[t * 3 + k]) - q[k]));
    }
  } else if (t < n1 + n2) {
    const int64_t u = t - n1;
    const int b = int(u / N2);
    const float d = dist2[u];
    if (d > 0.f) {
      const int64_t pi = int64_t(b) * N1 + idx2[u];
      const float* q = gt + (int64_t(b % n_gt) * N2 + (u % N2)) * 3;
      const float s = gs / (float(n2) * d);
      for (int k = 0; k < 3; ++k) atomicAdd(dpred + pi * 3 + k, s * (__bfloat162float(pred[pi * 3 + k]) - q[k]));
    }
  }
}

// ------------------------------------------------------------------------------------------------ token plumbing
// tile_rows: a learned [P, h] table (queries / positional embedding, fp32 master, used as bf16 by the reference)
// repeated for every sample -> f32 [B*P, h] holding the bf16 values; backward sums over the samples.
__global__ void tile_rows_fwd_kernel(const float* __restrict__ p, float* __restrict__ out, int64_t ph, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = bf16_round(p[i % ph]);
}
__global__ void tile_rows_bwd_kernel(const float* __restrict__ d, float* __restrict__ dp, int64_t ph, int B) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < ph; i += int64_t(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += d[int64_t(b) * ph + i];
    dp[i] = acc;
  }
}
// MAE decoder input (generation/models.py:183-187): tokens = image features with the ROI rows replaced by the mask
// token, plus the positional embedding — all bf16 in the reference; written as f32 holding those bf16 values.
__global__ void mask_tokens_fwd_kernel(const __nv_bfloat16* __restrict__ feat, const uint8_t* __restrict__ roi,
                                       const float* __restrict__ mtok, const float* __restrict__ pos,
                                       float* __restrict__ out, int P, int h, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = i / h;
    const int c = int(i - row * h);
    const float v = roi[row] ? bf16_round(mtok[c]) : __bfloat162float(feat[i]);
    out[i] = bf16_round(v + bf16_round(pos[(row % P) * int64_t(h) + c]));
  }
}
__global__ void mask_tokens_bwd_kernel(const float* __restrict__ d, const uint8_t* __restrict__ roi,
                                       __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dmtok,
                                       float* __restrict__ dpos, int B, int P, int h) {
  const int64_t ph = int64_t(P) * h;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < ph; i += int64_t(gridDim.x) * blockDim.x) {
    const int pp = int(i / h), c = int(i - int64_t(pp) * h);
    float ap = 0.f, am = 0.f;
    for (int b = 0; b < B; ++b) {
      const int64_t row = int64_t(b) * P + pp;
      const float g = d[row * h + c];
      ap += g;
      if (roi[row]) { am += g; dfeat[row * h + c] = __float2bfloat16_rn(0.f); }
      else dfeat[row * h + c] = __float2bfloat16_rn(g);
    }
    dpos[i] = ap;
    if (am != 0.f) atomicAdd(dmtok + c, am);
  }
}

static inline int ew_blocks(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = int64_t(num_sms()) * 16;
  return int(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace mla

using namespace mla;

extern "C" int mla_ln_f32_fwd(const void* x, const void* w, const void* b, void* y, void* y_bf16, void* mean, void* rstd,
                              int64_t rows, int32_t h, float eps, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (h <= 0) return set_error(MLA_ERR_ARG, "ln_f32_fwd: h must be positive");
  ln_f32_fwd_kernel<<<unsigned(rows), 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)w, (const float*)b,
                                                                     (float*)y, (__nv_bfloat16*)y_bf16, (float*)mean,
                                                                     (float*)rstd, h, eps);
  MLA_CHECK_LAUNCH("ln_f32_fwd");
  return MLA_OK;
}

extern "C" int mla_ln_f32_bwd(const void* dy, const void* x, const void* w, const void* mean, const void* rstd, void* dx,
                              void* dw, void* db, int64_t rows, int32_t h, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (h <= 0 || h > 256 * LN_CPT) return set_error(MLA_ERR_ARG, "ln_f32_bwd: h must be in (0, %d]", 256 * LN_CPT);
  const int grid = int(rows < int64_t(num_sms()) * 2 ? rows : int64_t(num_sms()) * 2);
  ln_f32_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dy, (const float*)x, (const float*)w,
                                                            (const float*)mean, (const float*)rstd, (float*)dx,
                                                            (float*)dw, (float*)db, rows, h);
  MLA_CHECK_LAUNCH("ln_f32_bwd");
  return MLA_OK;
}

extern "C" int mla_cast_bf16_f32(const void* x, void* y, int64_t n, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  cast_bf16_f32_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (float*)y, n);
  MLA_CHECK_LAUNCH("cast_bf16_f32");
  return MLA_OK;
}

extern "C" int mla_add_f32_bf16(const void* a, const void* b, void* out, int64_t n, int32_t round_bf16, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  add_f32_bf16_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)a, (const __nv_bfloat16*)b,
                                                                          (float*)out, n, round_bf16);
  MLA_CHECK_LAUNCH("add_f32_bf16");
  return MLA_OK;
}

extern "C" int mla_mask_scale_bf16(const void* x, const void* keep, void* y, int64_t n, int64_t per_mask, float scale,
                                   void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  if (per_mask <= 0) return set_error(MLA_ERR_ARG, "mask_scale: per_mask must be positive");
  mask_scale_bf16_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const uint8_t*)keep, (__nv_bfloat16*)y, n, per_mask, scale);
  MLA_CHECK_LAUNCH("mask_scale_bf16");
  return MLA_OK;
}

extern "C" int mla_seq_mean_fwd(const void* x, void* y, int32_t B, int32_t S, int32_t C, void* stream) {
  if (int rc = device_check()) return rc;
  if (B <= 0 || S <= 0 || C <= 0) return set_error(MLA_ERR_ARG, "seq_mean: empty problem");
  dim3 grid((C + 127) / 128, B);
  seq_mean_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, S, C);
  MLA_CHECK_LAUNCH("seq_mean_fwd");
  return MLA_OK;
}

extern "C" int mla_seq_mean_bwd(const void* dy, void* dx, int32_t B, int32_t S, int32_t C, void* stream) {
  if (int rc = device_check()) return rc;
  if (B <= 0 || S <= 0 || C <= 0) return set_error(MLA_ERR_ARG, "seq_mean: empty problem");
  const int64_t n = int64_t(B) * S * C;
  seq_mean_bwd_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx,
                                                                          S, C, n);
  MLA_CHECK_LAUNCH("seq_mean_bwd");
  return MLA_OK;
}

extern "C" int mla_bn_rows_fwd(const void* x, const void* w, const void* b, void* y, void* mean, void* rstd,
                               void* running_mean, void* running_var, int32_t R, int32_t C, float eps, float momentum,
                               int32_t relu, void* stream) {
  if (int rc = device_check()) return rc;
  if (R <= 0 || C <= 0) return set_error(MLA_ERR_ARG, "bn_rows: empty problem");
  bn_rows_fwd_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const float*)w, (const float*)b, (__nv_bfloat16*)y, (float*)mean, (float*)rstd,
      (float*)running_mean, (float*)running_var, R, C, eps, momentum, relu);
  MLA_CHECK_LAUNCH("bn_rows_fwd");
  return MLA_OK;
}

extern "C" int mla_bn_rows_bwd(const void* dy, const void* x, const void* y, const void* w, const void* mean,
                               const void* rstd, void* dx, void* dw, void* db, int32_t R, int32_t C, int32_t relu,
                               void* stream) {
  if (int rc = device_check()) return rc;
  if (R <= 0 || C <= 0) return set_error(MLA_ERR_ARG, "bn_rows: empty problem");
  bn_rows_bwd_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (const float*)w, (const float*)mean,
      (const float*)rstd, (__nv_bfloat16*)dx, (float*)dw, (float*)db, R, C, relu);
  MLA_CHECK_LAUNCH("bn_rows_bwd");
  return MLA_OK;
}

extern "C" int mla_roi_mask(const void* patch_idx, void* mask, int32_t B, int32_t n_pts, int32_t G, int32_t ksize,
                            void* stream) {
  if (int rc = device_check()) return rc;
  if (B <= 0 || G <= 0 || ksize < 1) return set_error(MLA_ERR_ARG, "roi_mask: bad arguments");
  roi_mask_kernel<<<B, 256, G * G, (cudaStream_t)stream>>>((const int64_t*)patch_idx, (uint8_t*)mask, n_pts, G, ksize);
  MLA_CHECK_LAUNCH("roi_mask");
  return MLA_OK;
}

static int gen_img_fill(GenImg& g, const mla_gen_image_args* a) {
  if (a == nullptr) return set_error(MLA_ERR_ARG, "gen_image: null args");
  if (a->n_patches <= 0 || a->patch <= 1 || a->grid <= 0 || a->n_images <= 0)
    return set_error(MLA_ERR_ARG, "gen_image: bad geometry");
  g.cur = (const float*)a->cur; g.nxt = (const float*)a->nxt;
  g.cur_sb = a->cur_stride_b; g.cur_sc = a->cur_stride_c; g.nxt_sb = a->nxt_stride_b; g.nxt_sc = a->nxt_stride_c;
  g.n_img = a->n_images; g.W = a->width; g.ps = a->patch; g.G = a->grid;
  g.delta_raw = (const __nv_bfloat16*)a->delta_raw; g.ld_delta = a->ld_delta;
  g.ao_raw = (const __nv_bfloat16*)a->ao_raw; g.ld_ao = a->ld_ao;
  g.roi = (const uint8_t*)a->roi;
  g.delta_clip = a->delta_clip; g.max_shift = a->max_shift; g.gen_weight = a->gen_weight;
  g.blended = (float*)a->blended; g.delta_all = (__nv_bfloat16*)a->delta_all;
  g.alpha_all = (__nv_bfloat16*)a->alpha_all; g.offset_all = (__nv_bfloat16*)a->offset_all;
  g.sums = (float*)a->sums; g.coef = (const float*)a->coef; g.gscale = (const float*)a->grad_scale;
  g.d_delta_raw = (__nv_bfloat16*)a->d_delta_raw; g.d_ao_raw = (__nv_bfloat16*)a->d_ao_raw;
  return MLA_OK;
}

extern "C" int mla_gen_image_fwd(const mla_gen_image_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  GenImg g;
  if (int rc = gen_img_fill(g, a)) return rc;
  if (!g.blended || !g.delta_all || !g.alpha_all || !g.offset_all || !g.sums)
    return set_error(MLA_ERR_ARG, "gen_image_fwd: null output");
  if (!a->losses || !a->coef) return set_error(MLA_ERR_ARG, "gen_image_fwd: null losses / coef");
  cudaMemsetAsync(g.sums, 0, 5 * sizeof(float), (cudaStream_t)stream);
  gen_image_kernel<0><<<a->n_patches, 256, 0, (cudaStream_t)stream>>>(g);
  MLA_CHECK_LAUNCH("gen_image_fwd");
  gen_image_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(g.sums, a->n_patches, 3 * a->patch * a->patch,
                                                              (float*)a->losses, (float*)a->coef);
  MLA_CHECK_LAUNCH("gen_image_finalize");
  return MLA_OK;
}

extern "C" int mla_gen_image_bwd(const mla_gen_image_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  GenImg g;
  if (int rc = gen_img_fill(g, a)) return rc;
  if (!g.coef || !g.gscale || !g.d_delta_raw || !g.d_ao_raw) return set_error(MLA_ERR_ARG, "gen_image_bwd: null tensor");
  gen_image_kernel<1><<<a->n_patches, 256, 0, (cudaStream_t)stream>>>(g);
  MLA_CHECK_LAUNCH("gen_image_bwd");
  return MLA_OK;
}

extern "C" int mla_chamfer_fwd(const void* pred, const void* gt, int32_t B, int32_t N1, int32_t N2, int32_t n_gt,
                               void* idx1, void* dist1, void* idx2, void* dist2, void* loss, void* stream) {
  if (int rc = device_check()) return rc;
  if (B <= 0 || N1 <= 0 || N2 <= 0 || n_gt <= 0) return set_error(MLA_ERR_ARG, "chamfer: empty problem");
  if (size_t(N1 > N2 ? N1 : N2) * 12 > 200 * 1024) return set_error(MLA_ERR_ARG, "chamfer: clouds above 17k points unsupported");
  auto s = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(chamfer_min_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  chamfer_min_kernel<<<dim3((N1 + 127) / 128, B), 128, size_t(N2) * 12, s>>>(
      (const __nv_bfloat16*)pred, (const float*)gt, N1, N2, n_gt, (int*)idx1, (float*)dist1, 0);
  MLA_CHECK_LAUNCH("chamfer_min(pred)");
  chamfer_min_kernel<<<dim3((N2 + 127) / 128, B), 128, size_t(N1) * 12, s>>>(
      (const __nv_bfloat16*)pred, (const float*)gt, N1, N2, n_gt, (int*)idx2, (float*)dist2, 1);
  MLA_CHECK_LAUNCH("chamfer_min(gt)");
  chamfer_reduce_kernel<<<1, 1024, 0, s>>>((const float*)dist1, (const float*)dist2, int64_t(B) * N1, int64_t(B) * N2,
                                           (float*)loss);
  MLA_CHECK_LAUNCH("chamfer_reduce");
  return MLA_OK;
}

extern "C" int mla_chamfer_bwd(const void* pred, const void* gt, int32_t B, int32_t N1, int32_t N2, int32_t n_gt,
                               const void* idx1, const void* dist1, const void* idx2, const void* dist2,
                               const void* grad_scale, void* dpred_f32, void* stream) {
  if (int rc = device_check()) return rc;
  if (B <= 0 || N1 <= 0 || N2 <= 0 || n_gt <= 0) return set_error(MLA_ERR_ARG, "chamfer: empty problem");
  const int64_t n = int64_t(B) * (N1 + N2);
  chamfer_bwd_kernel<<<int((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)pred, (const float*)gt, B, N1, N2, n_gt, (const int*)idx1, (const float*)dist1,
      (const int*)idx2, (const float*)dist2, (const float*)grad_scale, (float*)dpred_f32);
  MLA_CHECK_LAUNCH("chamfer_bwd");
  return MLA_OK;
}

extern "C" int mla_tile_rows_fwd(const void* table, void* out, int64_t table_elems, int32_t B, void* stream) {
  if (int rc = device_check()) return rc;
  if (table_elems <= 0 || B <= 0) return set_error(MLA_ERR_ARG, "tile_rows: empty problem");
  const int64_t n = table_elems * B;
  tile_rows_fwd_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)table, (float*)out, table_elems, n);
  MLA_CHECK_LAUNCH("tile_rows_fwd");
  return MLA_OK;
}

extern "C" int mla_tile_rows_bwd(const void* d_out, void* d_table, int64_t table_elems, int32_t B, void* stream) {
  if (int rc = device_check()) return rc;
  if (table_elems <= 0 || B <= 0) return set_error(MLA_ERR_ARG, "tile_rows: empty problem");
  tile_rows_bwd_kernel<<<ew_blocks(table_elems, 256), 256, 0, (cudaStream_t)stream>>>((const float*)d_out, (float*)d_table,
                                                                                      table_elems, B);
  MLA_CHECK_LAUNCH("tile_rows_bwd");
  return MLA_OK;
}

extern "C" int mla_mask_tokens_fwd(const void* feat, const void* roi, const void* mask_token, const void* pos, void* out,
                                   int32_t B, int32_t P, int32_t h, void* stream) {
  if (int rc = device_check()) return rc;
  if (B <= 0 || P <= 0 || h <= 0) return set_error(MLA_ERR_ARG, "mask_tokens: empty problem");
  const int64_t n = int64_t(B) * P * h;
  mask_tokens_fwd_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)feat, (const uint8_t*)roi, (const float*)mask_token, (const float*)pos, (float*)out, P, h, n);
  MLA_CHECK_LAUNCH("mask_tokens_fwd");
  return MLA_OK;
}

extern "C" int mla_mask_tokens_bwd(const void* d_out, const void* roi, void* d_feat, void* d_mask_token, void* d_pos,
                                   int32_t B, int32_t P, int32_t h, void* stream) {
  if (int rc = device_check()) return rc;
  if (B <= 0 || P <= 0 || h <= 0) return set_error(MLA_ERR_ARG, "mask_tokens: empty problem");
  mask_tokens_bwd_kernel<<<ew_blocks(int64_t(P) * h, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float*)d_out, (const uint8_t*)roi, (__nv_bfloat16*)d_feat, (float*)d_mask_token, (float*)d_pos, B, P, h);
  MLA_CHECK_LAUNCH("mask_tokens_bwd");
  return MLA_OK;
}
