// Small bandwidth-bound helpers around the GEMMs: dtype casts, activation backward, bias gradients, row
// gather/scatter (sequence splice and embedding lookup), LayerNorm, the diffusion-loss pieces.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

#define GRID_STRIDE(i, n) \
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < (n); i += int64_t(gridDim.x) * blockDim.x)

static inline int grid_for(int64_t total, int block = 256) {
  int64_t g = (total + block - 1) / block;
  int64_t cap = int64_t(num_sms()) * 16;
  return int(g < 1 ? 1 : (g < cap ? g : cap));
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, int64_t n) {
  const int64_t n4 = n >> 2;
  GRID_STRIDE(i, n4) {
    float4 v = reinterpret_cast<const float4*>(s)[i];
    uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    reinterpret_cast<uint2*>(d)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) d[n4 * 4 + threadIdx.x] = __float2bfloat16_rn(s[n4 * 4 + threadIdx.x]);
}

// Strided 2-D cast with zero padding: dst [rows, ldd] <- src [rows_src, cols_src] (pitch lds); outside -> 0.
__global__ void cast_pad_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, int64_t rows, int64_t cols,
                                int64_t rows_src, int64_t cols_src, int64_t lds) {
  GRID_STRIDE(i, rows * cols) {
    const int64_t r = i / cols, c = i % cols;
    d[i] = __float2bfloat16_rn((r < rows_src && c < cols_src) ? s[r * lds + c] : 0.f);
  }
}

__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case MLA_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case MLA_ACT_GELU_ERF: {
      const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case MLA_ACT_GELU_TANH: {
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      const float u = k0 * (x + k1 * x * x * x);
      const float t = tanhf(u);
      return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x * x);
    }
    case MLA_ACT_SILU: {
      const float s = 1.f / (1.f + __expf(-x));
      return s * (1.f + x * (1.f - s));
    }
    default: return 1.f;
  }
}

__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pre,
                               __nv_bfloat16* __restrict__ dx, int64_t n, int act) {
  GRID_STRIDE(i, n) {
    dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * act_grad(__bfloat162float(pre[i]), act));
  }
}

// out[c] (+)= sum_r x[r, c]   (bias gradient).  grid.x covers columns in chunks of 32, grid.y splits rows.
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int64_t rows, int cols,
                              int64_t ld) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;  // 8 row lanes
  float acc = 0.f;
  if (c < cols)
    for (int64_t r = blockIdx.y * 8 + ry; r < rows; r += int64_t(gridDim.y) * 8) acc += __bfloat162float(x[r * ld + c]);
  part[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// dst[i, :] = idx[i] >= 0 ? src[idx[i], :] : 0      (h % 8 == 0)
template <typename IdxT>
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const IdxT* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dst, int64_t n, int h, int64_t lds) {
  const int cpr = h >> 3;
  GRID_STRIDE(i, n * cpr) {
    const int64_t r = i / cpr;
    const int c = int(i % cpr);
    const int64_t s = int64_t(idx[r]);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (s >= 0) v = *reinterpret_cast<const uint4*>(src + s * lds + c * 8);
    *reinterpret_cast<uint4*>(dst + r * h + c * 8) = v;
  }
}
// dsrc[idx[i], :] = ddst[i, :]  for idx[i] >= 0  (idx injective: a permutation scatter, no atomics)
__global__ void scatter_rows_kernel(__nv_bfloat16* __restrict__ dsrc, const int32_t* __restrict__ idx,
                                    const __nv_bfloat16* __restrict__ ddst, int64_t n, int h, int64_t lds) {
  const int cpr = h >> 3;
  GRID_STRIDE(i, n * cpr) {
    const int64_t r = i / cpr;
    const int c = int(i % cpr);
    const int64_t s = idx[r];
    if (s >= 0) *reinterpret_cast<uint4*>(dsrc + s * lds + c * 8) = *reinterpret_cast<const uint4*>(ddst + r * h + c * 8);
  }
}
// grad[ids[i], :] += dout[i, :]   (embedding backward; fp32 atomics, duplicates allowed)
__global__ void embedding_bwd_kernel(float* __restrict__ grad, const int64_t* __restrict__ ids,
                                     const __nv_bfloat16* __restrict__ dout, int64_t n, int h, int64_t pad_id) {
  GRID_STRIDE(i, n * h) {
    const int64_t r = i / h;
    const int c = int(i % h);
    const int64_t id = ids[r];
    if (id != pad_id) atomicAdd(grad + id * h + c, __bfloat162float(dout[i]));
  }
}

// LayerNorm over the last dim (fp32 statistics, nn.LayerNorm semantics), bf16 in / bf16 out, fp32 affine.
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ b, __nv_bfloat16* __restrict__ y, int h, float eps) {
  __shared__ float red[2][32];
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * h;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < h; i += blockDim.x) {
    const float v = __bfloat162float(xr[i]);
    s += v; ss += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  s = lane < nw ? red[0][lane] : 0.f;
  ss = lane < nw ? red[1][lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  const float mean = s / h;
  const float var = fmaxf(ss / h - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x; i < h; i += blockDim.x)
    y[row * h + i] = __float2bfloat16_rn((__bfloat162float(xr[i]) - mean) * rstd * w[i] + b[i]);
}

// q_sample: x = sqrt_ac[t]*a + sqrt_1mac[t]*noise  (gaussian_diffusion.py:214-229), all fp32.
__global__ void q_sample_kernel(const float* __restrict__ a, const float* __restrict__ noise,
                                const int64_t* __restrict__ t, const float* __restrict__ sa, const float* __restrict__ sb,
                                float* __restrict__ out, int64_t n, int per_sample) {
  GRID_STRIDE(i, n) {
    const int64_t ts = t[i / per_sample];
    out[i] = sa[ts] * a[i] + sb[ts] * noise[i];
  }
}

// loss = mean((pred - target)^2), pred bf16, target f32 (model_mla.py:215); single block, deterministic.
__global__ void mse_fwd_kernel(const __nv_bfloat16* __restrict__ pred, const float* __restrict__ target,
                               float* __restrict__ loss, int64_t n) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = __bfloat162float(pred[i]) - target[i];
    acc += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) loss[0] = acc / float(n);
  }
}
// dpred = bf16( gscale[0] * 2*(pred - target)/n )
__global__ void mse_bwd_kernel(const __nv_bfloat16* __restrict__ pred, const float* __restrict__ target,
                               const float* __restrict__ gscale, __nv_bfloat16* __restrict__ dpred, int64_t n) {
  const float g = gscale[0] * 2.f / float(n);
  GRID_STRIDE(i, n) dpred[i] = __float2bfloat16_rn(g * (__bfloat162float(pred[i]) - target[i]));
}

// out = bf16(a + b)
__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, int64_t n) {
  GRID_STRIDE(i, n) out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))

extern "C" int mla_cast_f32_bf16(const void* src, void* dst, int64_t n, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 7))
    return set_error(MLA_ERR_ARG, "cast: src must be 16-byte and dst 8-byte aligned");
  cast_f32_bf16_kernel<<<grid_for(n / 4 + 1), 256, 0, S_(stream)>>>((const float*)src, (__nv_bfloat16*)dst, n);
  MLA_CHECK_LAUNCH("cast_f32_bf16");
  return MLA_OK;
}

extern "C" int mla_cast_pad_f32_bf16(const void* src, void* dst, int64_t rows, int64_t cols, int64_t rows_src,
                                     int64_t cols_src, int64_t lds, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows * cols <= 0) return MLA_OK;
  cast_pad_kernel<<<grid_for(rows * cols), 256, 0, S_(stream)>>>((const float*)src, (__nv_bfloat16*)dst, rows, cols,
                                                                 rows_src, cols_src, lds);
  MLA_CHECK_LAUNCH("cast_pad");
  return MLA_OK;
}

extern "C" int mla_act_bwd(const void* dy, const void* pre, void* dx, int64_t n, int32_t act, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  act_bwd_kernel<<<grid_for(n), 256, 0, S_(stream)>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)pre,
                                                      (__nv_bfloat16*)dx, n, act);
  MLA_CHECK_LAUNCH("act_bwd");
  return MLA_OK;
}

extern "C" int mla_colsum_bf16(const void* x, void* out, int64_t rows, int32_t cols, int64_t ld, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0 || cols <= 0) return MLA_OK;
  dim3 grid((cols + 31) / 32, (unsigned)(rows < 8 * 64 ? (rows + 7) / 8 : 64));
  colsum_kernel<<<grid, 256, 0, S_(stream)>>>((const __nv_bfloat16*)x, (float*)out, rows, cols, ld);
  MLA_CHECK_LAUNCH("colsum");
  return MLA_OK;
}

extern "C" int mla_gather_rows(const void* src, const void* idx, void* dst, int64_t n, int32_t h, int64_t lds,
                               int32_t idx_is_i64, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  if ((h & 7) || (lds & 7)) return set_error(MLA_ERR_ARG, "gather_rows: h and lds must be multiples of 8");
  if (idx_is_i64)
    gather_rows_kernel<int64_t><<<grid_for(n * (h / 8)), 256, 0, S_(stream)>>>((const __nv_bfloat16*)src, (const int64_t*)idx,
                                                                               (__nv_bfloat16*)dst, n, h, lds);
  else
    gather_rows_kernel<int32_t><<<grid_for(n * (h / 8)), 256, 0, S_(stream)>>>((const __nv_bfloat16*)src, (const int32_t*)idx,
                                                                               (__nv_bfloat16*)dst, n, h, lds);
  MLA_CHECK_LAUNCH("gather_rows");
  return MLA_OK;
}

extern "C" int mla_scatter_rows(void* dsrc, const void* idx, const void* ddst, int64_t n, int32_t h, int64_t lds,
                                void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  if ((h & 7) || (lds & 7)) return set_error(MLA_ERR_ARG, "scatter_rows: h and lds must be multiples of 8");
  scatter_rows_kernel<<<grid_for(n * (h / 8)), 256, 0, S_(stream)>>>((__nv_bfloat16*)dsrc, (const int32_t*)idx,
                                                                     (const __nv_bfloat16*)ddst, n, h, lds);
  MLA_CHECK_LAUNCH("scatter_rows");
  return MLA_OK;
}

extern "C" int mla_embedding_bwd(void* grad, const void* ids, const void* dout, int64_t n, int32_t h, int64_t pad_id,
                                 void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  embedding_bwd_kernel<<<grid_for(n * h), 256, 0, S_(stream)>>>((float*)grad, (const int64_t*)ids,
                                                                (const __nv_bfloat16*)dout, n, h, pad_id);
  MLA_CHECK_LAUNCH("embedding_bwd");
  return MLA_OK;
}

extern "C" int mla_layernorm_fwd(const void* x, const void* w, const void* b, void* y, int64_t rows, int32_t h,
                                 float eps, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  layernorm_kernel<<<(unsigned)rows, 256, 0, S_(stream)>>>((const __nv_bfloat16*)x, (const float*)w, (const float*)b,
                                                           (__nv_bfloat16*)y, h, eps);
  MLA_CHECK_LAUNCH("layernorm");
  return MLA_OK;
}

extern "C" int mla_q_sample(const void* a, const void* noise, const void* t, const void* sqrt_ac,
                            const void* sqrt_1mac, void* out, int64_t n, int32_t per_sample, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  q_sample_kernel<<<grid_for(n), 256, 0, S_(stream)>>>((const float*)a, (const float*)noise, (const int64_t*)t,
                                                       (const float*)sqrt_ac, (const float*)sqrt_1mac, (float*)out, n,
                                                       per_sample);
  MLA_CHECK_LAUNCH("q_sample");
  return MLA_OK;
}

extern "C" int mla_mse_fwd(const void* pred, const void* target, void* loss, int64_t n, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return set_error(MLA_ERR_ARG, "mse: empty input");
  mse_fwd_kernel<<<1, 1024, 0, S_(stream)>>>((const __nv_bfloat16*)pred, (const float*)target, (float*)loss, n);
  MLA_CHECK_LAUNCH("mse_fwd");
  return MLA_OK;
}

extern "C" int mla_mse_bwd(const void* pred, const void* target, const void* gscale, void* dpred, int64_t n,
                           void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  mse_bwd_kernel<<<grid_for(n), 256, 0, S_(stream)>>>((const __nv_bfloat16*)pred, (const float*)target,
                                                      (const float*)gscale, (__nv_bfloat16*)dpred, n);
  MLA_CHECK_LAUNCH("mse_bwd");
  return MLA_OK;
}

extern "C" int mla_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  add_bf16_kernel<<<grid_for(n), 256, 0, S_(stream)>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                       (__nv_bfloat16*)out, n);
  MLA_CHECK_LAUNCH("add_bf16");
  return MLA_OK;
}
