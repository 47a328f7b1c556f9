// InfoNCE alignment losses (reference: models/mla/fuser/contrastive.py:185-215 and :241-258) and the 3D->2D
// patch-correspondence projection (:5-45).
//
// Arithmetic under the reference's bf16 autocast: F.normalize runs in fp32 (autocast's fp32 list covers `norm`),
// the similarity matmul/bmm rounds to bf16, `/ temperature` rounds to bf16 again, cross_entropy runs in fp32.
//
// The reference compacts the valid rows with a boolean mask (host sync, :203-206).  Here invalid rows/columns stay in
// place and are masked out of both softmax directions and of the mean; M = number of valid rows lives on the device.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// y = bf16(x / max(||x||_2, eps)) per row; norm saved (fp32).  One warp per row.
__global__ void l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                  float* __restrict__ norms, int64_t rows, int d, int64_t ldx, float eps) {
  const int64_t r = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) { float v = __bfloat162float(x[r * ldx + i]); ss += v * v; }
  const float nrm = fmaxf(sqrtf(warp_sum_f(ss)), eps);
  if (lane == 0) norms[r] = nrm;
  for (int i = lane; i < d; i += 32) y[r * d + i] = __float2bfloat16_rn(__bfloat162float(x[r * ldx + i]) / nrm);
}
// dx = (dy - y*(y.dy)) / norm,  y recomputed from x in fp32.
__global__ void l2norm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ norms,
                                  const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int64_t rows,
                                  int d, int64_t ldx) {
  const int64_t r = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float inv = 1.f / norms[r];
  float dot = 0.f;
  for (int i = lane; i < d; i += 32)
    dot += __bfloat162float(x[r * ldx + i]) * inv * __bfloat162float(dy[r * d + i]);
  dot = warp_sum_f(dot);
  for (int i = lane; i < d; i += 32) {
    const float y = __bfloat162float(x[r * ldx + i]) * inv;
    dx[r * d + i] = __float2bfloat16_rn((__bfloat162float(dy[r * d + i]) - y * dot) * inv);
  }
}

// ---- symmetric InfoNCE over a materialised bf16 similarity matrix sim [N,N] (= bf16(a_i.b_j)) ------------------
// l_ij = bf16(sim_ij / T).  Pass 1: per-row LSE over valid columns (one warp per row) and per-column partial
// (max, sumexp) over row chunks, merged by pass 2.  Diagonal picked up on the way.
__global__ void nce_row_kernel(const __nv_bfloat16* __restrict__ sim, const uint8_t* __restrict__ valid,
                               float* __restrict__ row_lse, float* __restrict__ diag, int N, float temp) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= N) return;
  if (!valid[r]) { if (lane == 0) { row_lse[r] = 0.f; diag[r] = 0.f; } return; }
  const __nv_bfloat16* s = sim + int64_t(r) * N;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32)
    if (valid[j]) mx = fmaxf(mx, bf16_round(__bfloat162float(s[j]) / temp));
  mx = warp_max_f(mx);
  float se = 0.f;
  for (int j = lane; j < N; j += 32)
    if (valid[j]) se += __expf(bf16_round(__bfloat162float(s[j]) / temp) - mx);
  se = warp_sum_f(se);
  if (lane == 0) {
    row_lse[r] = mx + logf(se);
    diag[r] = bf16_round(__bfloat162float(s[r]) / temp);
  }
}
// Column LSE: each thread owns one column and walks rows [r0, r1) of its chunk (coalesced across the warp);
// chunk partials (max, sum) are merged with atomics-free two-level reduction by the finalize kernel.
__global__ void nce_col_partial_kernel(const __nv_bfloat16* __restrict__ sim, const uint8_t* __restrict__ valid,
                                       float* __restrict__ pmax, float* __restrict__ psum, int N, int chunk,
                                       float temp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * chunk, r1 = min(N, r0 + chunk);
  if (c >= N) return;
  float mx = -INFINITY, se = 0.f;
  if (valid[c]) {
    for (int r = r0; r < r1; ++r) {
      if (!valid[r]) continue;
      const float l = bf16_round(__bfloat162float(sim[int64_t(r) * N + c]) / temp);
      if (l > mx) { se = se * __expf(mx - l) + 1.f; mx = l; } else { se += __expf(l - mx); }
    }
  }
  pmax[int64_t(blockIdx.y) * N + c] = mx;
  psum[int64_t(blockIdx.y) * N + c] = se;
}
// Single CTA: merge column partials, count M, produce the loss.  out[0] = loss, out[1] = M.
__global__ void nce_finalize_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, int nchunks,
                                    const uint8_t* __restrict__ valid, const float* __restrict__ row_lse,
                                    const float* __restrict__ diag, float* __restrict__ col_lse,
                                    float* __restrict__ out, int N) {
  __shared__ float red[2][32];
  float acc = 0.f, cnt = 0.f;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float lse = 0.f;
    if (valid[c]) {
      float mx = -INFINITY;
      for (int k = 0; k < nchunks; ++k) mx = fmaxf(mx, pmax[int64_t(k) * N + c]);
      float se = 0.f;
      for (int k = 0; k < nchunks; ++k) {
        const float m = pmax[int64_t(k) * N + c];
        if (m > -INFINITY) se += psum[int64_t(k) * N + c] * __expf(m - mx);
      }
      lse = mx + logf(se);
      acc += (row_lse[c] - diag[c]) + (lse - diag[c]);
      cnt += 1.f;
    }
    col_lse[c] = lse;
  }
  acc = warp_sum_f(acc); cnt = warp_sum_f(cnt);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc; red[1][threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    acc = threadIdx.x < nw ? red[0][threadIdx.x] : 0.f;
    cnt = threadIdx.x < nw ? red[1][threadIdx.x] : 0.f;
    acc = warp_sum_f(acc); cnt = warp_sum_f(cnt);
    if (threadIdx.x == 0) { out[0] = cnt > 0.f ? acc / (2.f * cnt) : 0.f; out[1] = cnt; }
  }
}
// dsim_ij = g/(2M) * inv_t * (softmax_row_i(j) + softmax_col_j(i) - 2*delta_ij)   (valid i, j; else 0), bf16 in place
__global__ void nce_dsim_kernel(__nv_bfloat16* __restrict__ sim, const uint8_t* __restrict__ valid,
                                const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                                const float* __restrict__ out, const float* __restrict__ gscale, int N, float temp) {
  const float inv_t = 1.f / temp;
  const float M = out[1];
  const float coef = M > 0.f ? gscale[0] * inv_t / (2.f * M) : 0.f;
  const int64_t total = int64_t(N) * N;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int r = int(i / N), c = int(i % N);
    float d = 0.f;
    if (valid[r] && valid[c]) {
      const float l = bf16_round(__bfloat162float(sim[i]) / temp);
      d = coef * (__expf(l - row_lse[r]) + __expf(l - col_lse[c]) - (r == c ? 2.f : 0.f));
    }
    sim[i] = __float2bfloat16_rn(d);
  }
}

// ---- tactile InfoNCE: query [B,A,D] vs keys [B,K,D] (both L2-normalised bf16), positive index per (b,a) -------
// loss_sum += CE(bf16(bf16(q.k)/T), pos).  One CTA per (b,a); thread j owns key j (K <= 1024).
__global__ void tac_nce_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ keys,
                                   const int64_t* __restrict__ pos, float* __restrict__ probs, float* __restrict__ loss_rows,
                                   int A, int K, int D, float temp) {
  extern __shared__ float sh[];
  float* sq = sh;           // D
  float* red = sh + D;      // 32
  const int ba = blockIdx.x, b = ba / A;
  for (int i = threadIdx.x; i < D; i += blockDim.x) sq[i] = __bfloat162float(q[int64_t(ba) * D + i]);
  __syncthreads();
  const int j = threadIdx.x;
  float l = -INFINITY;
  if (j < K) {
    float acc = 0.f;
    const __nv_bfloat16* kr = keys + (int64_t(b) * K + j) * D;
    for (int i = 0; i < D; ++i) acc += sq[i] * __bfloat162float(kr[i]);
    l = bf16_round(bf16_round(acc) / temp);
  }
  float mx = warp_max_f(l);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int w = 0; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float e = j < K ? __expf(l - mx) : 0.f;
  float se = warp_sum_f(e);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
  __syncthreads();
  se = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) se += red[w];
  if (j < K) probs[int64_t(ba) * K + j] = e / se;
  if (j == int(pos[ba])) loss_rows[ba] = (mx + logf(se)) - l;
}
// dq[ba,:] = sum_j dl_j k_j ; dkeys[b,j,:] += dl_j q   with dl_j = coef*(p_j - [j==pos])*inv_t
__global__ void tac_nce_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ keys,
                                   const int64_t* __restrict__ pos, const float* __restrict__ probs,
                                   const float* __restrict__ gscale, float* __restrict__ dq, float* __restrict__ dkeys,
                                   int A, int K, int D, float inv_t, float inv_rows) {
  const int ba = blockIdx.x, b = ba / A;
  const float coef = gscale[0] * inv_rows * inv_t;
  const int p = int(pos[ba]);
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < K; ++j) {
      const float dl = coef * (probs[int64_t(ba) * K + j] - (j == p ? 1.f : 0.f));
      acc += dl * __bfloat162float(keys[(int64_t(b) * K + j) * D + i]);
    }
    dq[int64_t(ba) * D + i] = acc;
  }
  for (int idx = threadIdx.x; idx < K * D; idx += blockDim.x) {
    const int j = idx / D, i = idx % D;
    const float dl = coef * (probs[int64_t(ba) * K + j] - (j == p ? 1.f : 0.f));
    atomicAdd(dkeys + (int64_t(b) * K + j) * D + i, dl * __bfloat162float(q[int64_t(ba) * D + i]));
  }
}

// ---- 3D -> 2D patch correspondence (contrastive.py:5-45), op order of the reference kept so floor() agrees ------
// cam[0..8] = R (row-major), cam[9..11] = t, cam[12..20] = K (row-major, unscaled); sx, sy = resize/orig ratios.
__global__ void project_kernel(const float* __restrict__ xyz, const float* __restrict__ cam, int64_t n, float sx,
                               float sy, float total_stride, int patch_h, int patch_w, float img_w, float img_h,
                               int64_t* __restrict__ patch_idx, uint8_t* __restrict__ valid) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float* R = cam;
  const float* t = cam + 9;
  const float* Km = cam + 12;
  // R_world_to_cam = R^T ; t_w2c = -(R^T @ t)
  float tw[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc = __fmaf_rn(R[k * 3 + r], t[k], acc);
    tw[r] = -acc;
  }
  const float p[3] = {xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]};
  // xyz_cam = xyz @ (R^T)^T + t_w2c = xyz @ R + t_w2c
  float pc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc = __fmaf_rn(p[k], R[k * 3 + c], acc);
    pc[c] = acc + tw[c];
  }
  float Ks[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Ks[k] = Km[k];
  Ks[0] *= sx; Ks[4] *= sy; Ks[2] *= sx; Ks[5] *= sy;
  // uvw = xyz_cam @ K_scaled^T
  float uvw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc = __fmaf_rn(pc[k], Ks[c * 3 + k], acc);
    uvw[c] = acc;
  }
  const float z = uvw[2];
  const float x = uvw[0] / (z + 1e-6f), y = uvw[1] / (z + 1e-6f);
  float row = floorf(y / total_stride), col = floorf(x / total_stride);
  const bool ok = (z > 0.f) && (x >= 0.f) && (x < img_w) && (y >= 0.f) && (y < img_h);
  row = fminf(fmaxf(row, 0.f), float(patch_h - 1));
  col = fminf(fmaxf(col, 0.f), float(patch_w - 1));
  // NaN -> the reference's floor().long() is undefined; map to 0 like a clamp would
  patch_idx[i * 2] = (row == row) ? int64_t(row) : 0;
  patch_idx[i * 2 + 1] = (col == col) ? int64_t(col) : 0;
  valid[i] = ok ? 1 : 0;
}


// Tactile positives (models/vlm/prismatic.py:742-749): for every gripper position the nearest point-cloud centre
// (torch.cdist + topk(k=1, largest=False)) and the linear index row*patch_w + col of the image patch that centre
// projects to.  One warp per (sample, arm); fp32 squared distances, lowest index wins ties.
__global__ void nearest_center_kernel(const float* __restrict__ grip, const float* __restrict__ centers,
                                      const int64_t* __restrict__ patch_idx, int n_query, int arms, int G, int patch_w,
                                      int64_t* __restrict__ pos_pc, int64_t* __restrict__ lin_img) {
  const int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wq >= n_query) return;
  const int b = wq / arms;
  const float gx = grip[wq * 3], gy = grip[wq * 3 + 1], gz = grip[wq * 3 + 2];
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int g = lane; g < G; g += 32) {
    const float* c = centers + (int64_t(b) * G + g) * 3;
    const float dx = gx - c[0], dy = gy - c[1], dz = gz - c[2];
    const float d = dx * dx + dy * dy + dz * dz;
    if (d < best) { best = d; bi = g; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    if (bi >= G) bi = 0;   // all distances NaN/inf: torch.topk would still return an index
    pos_pc[wq] = bi;
    const int64_t* pi = patch_idx + (int64_t(b) * G + bi) * 2;
    lin_img[wq] = pi[0] * patch_w + pi[1];
  }
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))

extern "C" int mla_l2norm_fwd(const void* x, void* y, void* norms, int64_t rows, int32_t d, int64_t ldx, float eps,
                              void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  l2norm_fwd_kernel<<<unsigned((rows * 32 + 255) / 256), 256, 0, S_(stream)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                                                               (float*)norms, rows, d, ldx, eps);
  MLA_CHECK_LAUNCH("l2norm_fwd");
  return MLA_OK;
}
extern "C" int mla_l2norm_bwd(const void* x, const void* norms, const void* dy, void* dx, int64_t rows, int32_t d,
                              int64_t ldx, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  l2norm_bwd_kernel<<<unsigned((rows * 32 + 255) / 256), 256, 0, S_(stream)>>>(
      (const __nv_bfloat16*)x, (const float*)norms, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, rows, d, ldx);
  MLA_CHECK_LAUNCH("l2norm_bwd");
  return MLA_OK;
}

extern "C" size_t mla_infonce_workspace(int32_t n) {
  const int nchunks = (n + 127) / 128;
  return sizeof(float) * (size_t(3) * n + size_t(2) * nchunks * n);
}

extern "C" int mla_infonce_fwd(const void* sim, const void* valid, void* workspace, void* out, int32_t n,
                               float temperature, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return set_error(MLA_ERR_ARG, "infonce: empty");
  const int chunk = 128, nchunks = (n + chunk - 1) / chunk;
  float* ws = (float*)workspace;
  float *row_lse = ws, *col_lse = ws + n, *diag = ws + 2 * size_t(n), *pmax = ws + 3 * size_t(n), *psum = pmax + size_t(nchunks) * n;
  nce_row_kernel<<<(n * 32 + 255) / 256, 256, 0, S_(stream)>>>((const __nv_bfloat16*)sim, (const uint8_t*)valid, row_lse, diag, n, temperature);
  MLA_CHECK_LAUNCH("nce_row");
  dim3 g((n + 255) / 256, nchunks);
  nce_col_partial_kernel<<<g, 256, 0, S_(stream)>>>((const __nv_bfloat16*)sim, (const uint8_t*)valid, pmax, psum, n, chunk, temperature);
  MLA_CHECK_LAUNCH("nce_col_partial");
  nce_finalize_kernel<<<1, 1024, 0, S_(stream)>>>(pmax, psum, nchunks, (const uint8_t*)valid, row_lse, diag, col_lse, (float*)out, n);
  MLA_CHECK_LAUNCH("nce_finalize");
  return MLA_OK;
}

extern "C" int mla_infonce_bwd(void* sim_inout, const void* valid, const void* workspace, const void* out,
                               const void* gscale, int32_t n, float temperature, void* stream) {
  if (int rc = device_check()) return rc;
  const float* ws = (const float*)workspace;
  const int64_t total = int64_t(n) * n;
  int grid = int((total + 255) / 256 < int64_t(num_sms()) * 32 ? (total + 255) / 256 : int64_t(num_sms()) * 32);
  nce_dsim_kernel<<<grid, 256, 0, S_(stream)>>>((__nv_bfloat16*)sim_inout, (const uint8_t*)valid, ws, ws + n,
                                                (const float*)out, (const float*)gscale, n, temperature);
  MLA_CHECK_LAUNCH("nce_dsim");
  return MLA_OK;
}

extern "C" int mla_tac_nce_fwd(const void* q, const void* keys, const void* pos, void* probs, void* loss_rows,
                               int32_t batch, int32_t arms, int32_t k, int32_t d, float temperature, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch * arms <= 0) return MLA_OK;
  if (k > 1024) return set_error(MLA_ERR_ARG, "tac_nce: more than 1024 keys");
  const int block = (k + 31) / 32 * 32;
  tac_nce_fwd_kernel<<<batch * arms, block, (d + 32) * sizeof(float), S_(stream)>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)keys, (const int64_t*)pos, (float*)probs, (float*)loss_rows, arms,
      k, d, temperature);
  MLA_CHECK_LAUNCH("tac_nce_fwd");
  return MLA_OK;
}
extern "C" int mla_tac_nce_bwd(const void* q, const void* keys, const void* pos, const void* probs, const void* gscale,
                               void* dq, void* dkeys, int32_t batch, int32_t arms, int32_t k, int32_t d,
                               float temperature, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch * arms <= 0) return MLA_OK;
  tac_nce_bwd_kernel<<<batch * arms, 256, 0, S_(stream)>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)keys,
                                                           (const int64_t*)pos, (const float*)probs, (const float*)gscale,
                                                           (float*)dq, (float*)dkeys, arms, k, d, 1.f / temperature,
                                                           1.f / float(batch * arms));
  MLA_CHECK_LAUNCH("tac_nce_bwd");
  return MLA_OK;
}

extern "C" int mla_project_points(const void* xyz, const void* cam, int64_t n, float sx, float sy, float total_stride,
                                  int32_t patch_h, int32_t patch_w, float img_w, float img_h, void* patch_idx,
                                  void* valid, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  project_kernel<<<unsigned((n + 255) / 256), 256, 0, S_(stream)>>>((const float*)xyz, (const float*)cam, n, sx, sy,
                                                                    total_stride, patch_h, patch_w, img_w, img_h,
                                                                    (int64_t*)patch_idx, (uint8_t*)valid);
  MLA_CHECK_LAUNCH("project_points");
  return MLA_OK;
}

extern "C" int mla_nearest_center(const void* gripper_xyz, const void* centers, const void* patch_idx, int32_t batch,
                                  int32_t arms, int32_t groups, int32_t patch_w, void* pos_pc, void* lin_img,
                                  void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || arms <= 0) return MLA_OK;
  if (groups <= 0) return set_error(MLA_ERR_ARG, "nearest_center: no centres");
  const int nq = batch * arms;
  nearest_center_kernel<<<(nq * 32 + 127) / 128, 128, 0, S_(stream)>>>(
      (const float*)gripper_xyz, (const float*)centers, (const int64_t*)patch_idx, nq, arms, groups, patch_w,
      (int64_t*)pos_pc, (int64_t*)lin_img);
  MLA_CHECK_LAUNCH("nearest_center");
  return MLA_OK;
}
