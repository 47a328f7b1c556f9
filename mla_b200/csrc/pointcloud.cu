// Point-cloud tokenizer kernels (reference: models/mla/pointcloud/backbone/Point_PN.py).
//   fps            furthest_point_sample (:6-21): the reference runs 512+256 sequential iterations of ~10 ATen
//                  launches each; here one CTA per sample keeps xyz and the running distance in shared memory.
//   knn            knn_point (:62-73) on the reference's distance arithmetic (bf16 dot under autocast, in-place bf16
//                  accumulation of the norms), exact k-smallest by a two-level radix select on the 16-bit keys;
//                  ties resolved by lowest index (torch.topk leaves them implementation-defined).
//   group_pose     LGA 'scan' normalisation + feature concat + PosE_Geo (:125-150,:228-249) in one pass.
//   bn_stats/...   train-mode BatchNorm statistics and the fused normalise(+residual)+ReLU(+max-pool over K).
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

// ---------------------------------------------------------------------------------------------------- FPS
// dist_i = min(dist_i, ((x-cx)^2 + (y-cy)^2) + (z-cz)^2) without FMA contraction (bit-exact with the ATen
// elementwise ops), argmax with lowest-index tie-break.
__global__ void __launch_bounds__(1024) fps_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start,
                                                   int32_t* __restrict__ idx_out, float* __restrict__ centers, int N,
                                                   int npoint) {
  extern __shared__ float sm[];
  float* sx = sm;
  float* sy = sm + N;
  float* sz = sm + 2 * N;
  float* sd = sm + 3 * N;
  __shared__ float rv[32];
  __shared__ int ri[32];
  __shared__ int s_far;
  const int b = blockIdx.x;
  const float* p = xyz + int64_t(b) * N * 3;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sx[i] = p[i * 3]; sy[i] = p[i * 3 + 1]; sz[i] = p[i * 3 + 2];
    sd[i] = 1e10f;
  }
  if (threadIdx.x == 0) s_far = int(start[b]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int it = 0; it < npoint; ++it) {
    const int far = s_far;
    const float cx = sx[far], cy = sy[far], cz = sz[far];
    if (threadIdx.x == 0) {
      idx_out[int64_t(b) * npoint + it] = far;
      float* c = centers + (int64_t(b) * npoint + it) * 3;
      c[0] = cx; c[1] = cy; c[2] = cz;
    }
    float best = -1.f;
    int besti = 0x7fffffff;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const float dx = __fsub_rn(sx[i], cx), dy = __fsub_rn(sy[i], cy), dz = __fsub_rn(sz[i], cz);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      float cur = sd[i];
      if (d < cur) { cur = d; sd[i] = d; }
      if (cur > best) { best = cur; besti = i; }  // ascending i: first maximum kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (lane == 0) { rv[warp] = best; ri[warp] = besti; }
    __syncthreads();
    if (warp == 0) {
      best = lane < nw ? rv[lane] : -1.f;
      besti = lane < nw ? ri[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (lane == 0) s_far = besti;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------- kNN
__device__ __forceinline__ uint32_t bf16_sort_key(float v) {
  // bf16 bit pattern -> unsigned key with the same order as the float
  uint32_t u = __float_as_uint(v) >> 16;
  return (u & 0x8000u) ? (~u & 0xFFFFu) : (u | 0x8000u);
}
// One CTA (128 threads) per query point.  bf16_dist != 0: distances as the reference's autocast path computes them.
__global__ void __launch_bounds__(128) knn_kernel(const float* __restrict__ xyz, const float* __restrict__ query,
                                                  int32_t* __restrict__ knn_idx, int N, int G, int K, int bf16_dist) {
  extern __shared__ uint32_t skey[];  // N keys
  __shared__ int hist[256];
  __shared__ int s_bin, s_less, s_cnt_tie, s_cnt_less;
  const int bg = blockIdx.x;
  const int b = bg / G;
  const float* p = xyz + int64_t(b) * N * 3;
  const float qx = query[int64_t(bg) * 3], qy = query[int64_t(bg) * 3 + 1], qz = query[int64_t(bg) * 3 + 2];
  const float qn = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
  const float qbx = bf16_round(qx), qby = bf16_round(qy), qbz = bf16_round(qz);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float x = p[i * 3], y = p[i * 3 + 1], z = p[i * 3 + 2];
    const float pn = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    float d;
    if (bf16_dist) {
      float dot = __fmaf_rn(qbz, bf16_round(z), __fmaf_rn(qby, bf16_round(y), __fmul_rn(qbx, bf16_round(x))));
      d = bf16_round(-2.f * bf16_round(dot));
      d = bf16_round(__fadd_rn(d, qn));
      d = bf16_round(__fadd_rn(d, pn));
      skey[i] = bf16_sort_key(d);
    } else {
      float dot = __fmaf_rn(qz, z, __fmaf_rn(qy, y, __fmul_rn(qx, x)));
      d = __fadd_rn(__fadd_rn(-2.f * dot, qn), pn);
      uint32_t u = __float_as_uint(d);
      skey[i] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // full 32-bit order key
    }
  }
  // radix select of the K-th smallest key, 8 bits per round from the top
  const int rounds = bf16_dist ? 2 : 4;
  uint32_t prefix = 0, prefix_mask = 0;
  int need = K;  // rank (1-based) still to find inside the current prefix class
  for (int r = 0; r < rounds; ++r) {
    const int shift = (bf16_dist ? 8 : 24) - 8 * r;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const uint32_t k = skey[i];
      if ((k & prefix_mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, bin = 0;
      for (; bin < 256; ++bin) {
        if (acc + hist[bin] >= need) break;
        acc += hist[bin];
      }
      s_bin = bin;
      s_less = acc;
    }
    __syncthreads();
    prefix |= uint32_t(s_bin) << shift;
    prefix_mask |= 255u << shift;
    need -= s_less;
    __syncthreads();
  }
  // prefix == exact key of the K-th smallest; `need` of the elements equal to it are taken, lowest index first
  if (threadIdx.x == 0) { s_cnt_less = 0; s_cnt_tie = 0; }
  __syncthreads();
  int32_t* out = knn_idx + int64_t(bg) * K;
  const int n_less = K - need;
  // deterministic order: thread-strided scan would reorder ties, so ties are resolved by a serial index scan per warp 0
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (skey[i] < prefix) out[atomicAdd(&s_cnt_less, 1)] = i;  // order inside the set is irrelevant downstream
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // ballot-based ordered compaction of the tie class
    int taken = 0;
    for (int base = 0; base < N && taken < need; base += 32) {
      const int i = base + threadIdx.x;
      const bool is_tie = i < N && skey[i] == prefix;
      const uint32_t m = __ballot_sync(0xffffffffu, is_tie);
      const int rank = taken + __popc(m & ((1u << threadIdx.x) - 1u));
      if (is_tie && rank < need) out[n_less + rank] = i;
      taken += __popc(m);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- grouping
// One CTA per group (b,g).  feat: [B, N, C] (bf16 if feat_is_bf16 else f32).  Output rows (b,g,k), 2C + PosE
// channels: X f32 [rows, out_dim] and a bf16 copy for the GEMM.  dim_embed: f32 [fd] = alpha^(j/fd).
__global__ void group_pose_kernel(const float* __restrict__ xyz, const void* __restrict__ feat, int feat_is_bf16,
                                  const int32_t* __restrict__ fps_idx, const int32_t* __restrict__ knn_idx,
                                  const float* __restrict__ dim_embed, float* __restrict__ xf,
                                  __nv_bfloat16* __restrict__ xb, int N, int G, int K, int C, float beta) {
  extern __shared__ float sh[];
  float* nrm = sh;             // K*3 normalised offsets
  __shared__ float smax[3];
  const int bg = blockIdx.x, b = bg / G;
  const int out_dim = 2 * C, fd = out_dim / 6;
  const int center = fps_idx[bg];
  const float* P = xyz + int64_t(b) * N * 3;
  const float cx = P[center * 3], cy = P[center * 3 + 1], cz = P[center * 3 + 2];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int j = knn_idx[int64_t(bg) * K + k];
    nrm[k * 3] = __fsub_rn(P[j * 3], cx);
    nrm[k * 3 + 1] = __fsub_rn(P[j * 3 + 1], cy);
    nrm[k * 3 + 2] = __fsub_rn(P[j * 3 + 2], cz);
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float m = 0.f;
    for (int k = 0; k < K; ++k) m = fmaxf(m, fabsf(nrm[k * 3 + threadIdx.x]));
    smax[threadIdx.x] = fmaxf(m, 1e-6f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 3; i += blockDim.x) nrm[i] = nrm[i] / smax[i % 3];
  __syncthreads();
  const __nv_bfloat16* fb = static_cast<const __nv_bfloat16*>(feat);
  const float* ff = static_cast<const float*>(feat);
  // channel ch = c3 * 2fd + j carries sin(arg) for j < fd and cos(arg) of the SAME arg at j + fd: one sincosf per pair
  // (the kernel is bound by the two accurate transcendentals per element, not by its 6 bytes of output)
  const int half = out_dim / 2;                    // = 3 * fd pairs per neighbour
  for (int e = threadIdx.x; e < K * half; e += blockDim.x) {
    const int k = e / half, pr = e % half;
    const int c3 = pr / fd, jj = pr % fd;
    const float arg = (beta * nrm[k * 3 + c3]) / dim_embed[jj];
    float sn, cs;
    sincosf(arg, &sn, &cs);
    const int nb = knn_idx[int64_t(bg) * K + k];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ch = c3 * 2 * fd + jj + q * fd;
      const int src = ch < C ? nb : center;
      const int cc = ch < C ? ch : ch - C;
      const int64_t fo = (int64_t(b) * N + src) * C + cc;
      const float fv = feat_is_bf16 ? __bfloat162float(fb[fo]) : ff[fo];
      const float v = fv + (q == 0 ? sn : cs);
      const int64_t o = (int64_t(bg) * K + k) * out_dim + ch;
      xf[o] = v;
      xb[o] = __float2bfloat16_rn(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- BatchNorm
// sums[0..C) += sum_r y[r,c] ; sums[C..2C) += sum_r y[r,c]^2     (y bf16 [rows, C]); 256 threads, C <= 512
__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ sums, int64_t rows, int C,
                                int rows_per_block) {
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      const float v = __bfloat162float(y[r * C + c]);
      s += v; ss += v * v;
    }
    atomicAdd(sums + c, s);
    atomicAdd(sums + C + c, ss);
  }
}
// scale/shift + running-stat update (momentum, unbiased variance), nn.BatchNorm train-mode semantics.
// coef[0..C) = mean, coef[C..2C) = invstd
__global__ void bn_finalize_kernel(const float* __restrict__ sums, float* __restrict__ coef,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int64_t rows,
                                   int C, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float n = float(rows);
  const float mean = sums[c] / n;
  const float var = fmaxf(sums[C + c] / n - mean * mean, 0.f);
  coef[c] = mean;
  coef[C + c] = rsqrtf(var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (n / fmaxf(n - 1.f, 1.f));
}
// 16-byte vectors (C % 8 == 0): the scalar kernels above/below move 2 bytes per thread and pay a 64-bit modulo per element
// (0.22 / 0.32 ms per call at 1.3 M x 96..192 against ~0.05 / 0.08 ms of HBM time).
__global__ void __launch_bounds__(256)
bn_stats_vec_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ sums, int64_t rows, int C, int rows_per_block) {
  __shared__ float red[256 * 16];
  const int cpr = C >> 3;                         // 16-byte chunks per row
  const int lanes = blockDim.x / cpr;             // rows processed per step
  const int cc = threadIdx.x % cpr, rr = threadIdx.x / cpr;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (rr < lanes) {
    auto add = [&](const uint4& u) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(h[t]);
        s[2 * t] += f.x; q[2 * t] += f.x * f.x;
        s[2 * t + 1] += f.y; q[2 * t + 1] += f.y * f.y;
      }
    };
    int64_t r = r0 + rr;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {       // four loads in flight per thread
      const uint4 u0 = *reinterpret_cast<const uint4*>(y + r * C + cc * 8);
      const uint4 u1 = *reinterpret_cast<const uint4*>(y + (r + lanes) * C + cc * 8);
      const uint4 u2 = *reinterpret_cast<const uint4*>(y + (r + 2 * lanes) * C + cc * 8);
      const uint4 u3 = *reinterpret_cast<const uint4*>(y + (r + 3 * lanes) * C + cc * 8);
      add(u0); add(u1); add(u2); add(u3);
    }
    for (; r < r1; r += lanes) add(*reinterpret_cast<const uint4*>(y + r * C + cc * 8));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = s[j]; red[threadIdx.x * 16 + 8 + j] = q[j]; }
  __syncthreads();
  // thread (cc, j) with rr == 0 folds the row lanes of its column
  for (int idx = threadIdx.x; idx < cpr * 16; idx += blockDim.x) {
    const int c2 = idx / 16, j = idx % 16;
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[(l * cpr + c2) * 16 + j];
    atomicAdd(sums + (j < 8 ? 0 : C) + c2 * 8 + (j & 7), t);
  }
}
__global__ void __launch_bounds__(256)
bn_relu_vec_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ coef, const float* __restrict__ w,
                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int64_t total8, int C) {
  // the launcher makes gridDim * blockDim a multiple of the chunks per row, so a thread meets the same 8 columns in every
  // iteration and keeps their coefficients in registers
  const int cpr = C >> 3;
  const int64_t i0 = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int c0 = int(i0 % cpr) * 8;
  float mean[8], inv[8], ww[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mean[j] = coef[c0 + j]; inv[j] = coef[C + c0 + j]; ww[j] = w[c0 + j]; bb[j] = bias[c0 + j]; }
  for (int64_t i = i0; i < total8; i += int64_t(gridDim.x) * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(y + i * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    float o[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __bfloat1622float2(h[t]);
      o[2 * t] = f.x; o[2 * t + 1] = f.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf((o[j] - mean[j]) * inv[j] * ww[j] + bb[j], 0.f);
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                        pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  }
}
// out = bf16(relu(bn(y)))   (in place allowed)
__global__ void bn_relu_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ coef,
                               const float* __restrict__ w, const float* __restrict__ bias,
                               __nv_bfloat16* __restrict__ out, int64_t total, int C) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    const float v = (__bfloat162float(y[i]) - coef[c]) * coef[C + c] * w[c] + bias[c];
    out[i] = __float2bfloat16_rn(fmaxf(v, 0.f));
  }
}
// v = relu(bf16(bn(y)) + x);  writes whichever of xf (f32), xb (bf16 copy for the next GEMM) and
// pooled[g, c] = max_k v (fused max-pool over the K rows of each group) is non-null.  One CTA per group.
__global__ void bn_res_relu_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ coef,
                                   const float* __restrict__ w, const float* __restrict__ bias,
                                   const float* __restrict__ x, float* __restrict__ xf_out,
                                   __nv_bfloat16* __restrict__ xb_out, float* __restrict__ pooled, int K, int C) {
  const int64_t g = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = coef[c], inv = coef[C + c], ww = w[c], bb = bias[c];
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const int64_t o = (g * K + k) * C + c;
      const float bn = bf16_round((__bfloat162float(y[o]) - mean) * inv * ww + bb);
      const float v = fmaxf(bn + x[o], 0.f);
      mx = fmaxf(mx, v);
      if (xf_out) xf_out[o] = v;
      if (xb_out) xb_out[o] = __float2bfloat16_rn(v);
    }
    if (pooled) pooled[g * C + c] = mx;
  }
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))

extern "C" int mla_fps(const void* xyz, const void* start, void* idx_out, void* centers, int32_t batch, int32_t n,
                       int32_t npoint, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || npoint <= 0) return MLA_OK;
  if (n > 12000) return set_error(MLA_ERR_ARG, "fps: more than 12000 points per cloud does not fit shared memory");
  const int smem = 4 * n * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    done = true;
  }
  int block = n >= 1024 ? 1024 : (n + 31) / 32 * 32;
  fps_kernel<<<batch, block, smem, S_(stream)>>>((const float*)xyz, (const int64_t*)start, (int32_t*)idx_out,
                                                 (float*)centers, n, npoint);
  MLA_CHECK_LAUNCH("fps");
  return MLA_OK;
}

extern "C" int mla_knn(const void* xyz, const void* query, void* knn_idx, int32_t batch, int32_t n, int32_t groups,
                       int32_t k, int32_t bf16_dist, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || groups <= 0) return MLA_OK;
  if (k > n) return set_error(MLA_ERR_ARG, "knn: k=%d exceeds the number of points %d", k, n);
  if (n > 40000) return set_error(MLA_ERR_ARG, "knn: more than 40000 points per cloud does not fit shared memory");
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    done = true;
  }
  knn_kernel<<<batch * groups, 128, n * sizeof(uint32_t), S_(stream)>>>((const float*)xyz, (const float*)query,
                                                                        (int32_t*)knn_idx, n, groups, k, bf16_dist);
  MLA_CHECK_LAUNCH("knn");
  return MLA_OK;
}

extern "C" int mla_group_pose(const void* xyz, const void* feat, int32_t feat_is_bf16, const void* fps_idx,
                              const void* knn_idx, const void* dim_embed, void* x_f32, void* x_bf16, int32_t batch,
                              int32_t n, int32_t groups, int32_t k, int32_t c, float beta, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0) return MLA_OK;
  if ((2 * c) % 6) return set_error(MLA_ERR_ARG, "group_pose: 2*c must be divisible by 6");
  group_pose_kernel<<<batch * groups, 256, k * 3 * sizeof(float), S_(stream)>>>(
      (const float*)xyz, feat, feat_is_bf16, (const int32_t*)fps_idx, (const int32_t*)knn_idx, (const float*)dim_embed,
      (float*)x_f32, (__nv_bfloat16*)x_bf16, n, groups, k, c, beta);
  MLA_CHECK_LAUNCH("group_pose");
  return MLA_OK;
}

extern "C" int mla_bn_stats(const void* y, void* sums, int64_t rows, int32_t c, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  const int target_blocks = num_sms() * 8;
  int rpb = int((rows + target_blocks - 1) / target_blocks);
  if (rpb < 8) rpb = 8;
  const int blocks = int((rows + rpb - 1) / rpb);
  if ((c & 7) == 0 && c <= 2048 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && rows >= 4096) {
    int rpb2 = rpb < 64 ? 64 : rpb;
    const int blocks2 = int((rows + rpb2 - 1) / rpb2);
    bn_stats_vec_kernel<<<blocks2, 256, 0, S_(stream)>>>((const __nv_bfloat16*)y, (float*)sums, rows, c, rpb2);
    MLA_CHECK_LAUNCH("bn_stats_vec");
    return MLA_OK;
  }
  bn_stats_kernel<<<blocks, c >= 256 ? 256 : (c + 31) / 32 * 32, 0, S_(stream)>>>((const __nv_bfloat16*)y, (float*)sums, rows, c, rpb);
  MLA_CHECK_LAUNCH("bn_stats");
  return MLA_OK;
}

extern "C" int mla_bn_finalize(const void* sums, void* coef, void* running_mean, void* running_var, int64_t rows,
                               int32_t c, float eps, float momentum, void* stream) {
  if (int rc = device_check()) return rc;
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, S_(stream)>>>((const float*)sums, (float*)coef, (float*)running_mean,
                                                              (float*)running_var, rows, c, eps, momentum);
  MLA_CHECK_LAUNCH("bn_finalize");
  return MLA_OK;
}

extern "C" int mla_bn_relu(const void* y, const void* coef, const void* w, const void* bias, void* out, int64_t rows,
                           int32_t c, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  const int64_t total = rows * c;
  if ((c & 7) == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int64_t total8 = total >> 3;
    int grid8 = int((total8 + 255) / 256 < int64_t(num_sms()) * 16 ? (total8 + 255) / 256 : int64_t(num_sms()) * 16);
    {   // gridDim * 256 must be a multiple of the chunks per row (c / 8)
      const int cpr = c >> 3;
      int a = cpr, b = 256;
      while (b) { const int t = a % b; a = b; b = t; }
      const int need = cpr / a;
      grid8 = (grid8 + need - 1) / need * need;
    }
    bn_relu_vec_kernel<<<grid8, 256, 0, S_(stream)>>>((const __nv_bfloat16*)y, (const float*)coef, (const float*)w,
                                                      (const float*)bias, (__nv_bfloat16*)out, total8, c);
    MLA_CHECK_LAUNCH("bn_relu_vec");
    return MLA_OK;
  }
  int grid = int((total + 255) / 256 < int64_t(num_sms()) * 16 ? (total + 255) / 256 : int64_t(num_sms()) * 16);
  bn_relu_kernel<<<grid, 256, 0, S_(stream)>>>((const __nv_bfloat16*)y, (const float*)coef, (const float*)w,
                                               (const float*)bias, (__nv_bfloat16*)out, total, c);
  MLA_CHECK_LAUNCH("bn_relu");
  return MLA_OK;
}

extern "C" int mla_bn_res_relu(const void* y, const void* coef, const void* w, const void* bias, const void* x,
                               void* x_f32_out, void* x_bf16_out, void* pooled, int64_t groups, int32_t k, int32_t c,
                               void* stream) {
  if (int rc = device_check()) return rc;
  if (groups <= 0) return MLA_OK;
  bn_res_relu_kernel<<<(unsigned)groups, c >= 256 ? 256 : (c + 31) / 32 * 32, 0, S_(stream)>>>(
      (const __nv_bfloat16*)y, (const float*)coef, (const float*)w, (const float*)bias, (const float*)x,
      (float*)x_f32_out, (__nv_bfloat16*)x_bf16_out, (float*)pooled, k, c);
  MLA_CHECK_LAUNCH("bn_res_relu");
  return MLA_OK;
}
