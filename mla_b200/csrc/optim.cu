// Optimizer-side kernels of the data-parallel training step: global gradient norm, clip coefficient, fused AdamW
// (fp32 master update + bf16 compute-copy refresh in one pass).  Mirrors torch.optim.AdamW + clip_grad_norm_ as the
// reference's strategy uses them (training/strategies/fsdp.py:242-257,:310), minus the host round trips.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = x[n4 * 4 + threadIdx.x]; acc += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

// scale[0] = min(1, max_norm / (||g_mean|| + 1e-6)) * inv_world ; scale[1] = ||g_mean||   (g_mean = g_sum * inv_world)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float inv_world,
                                 float* __restrict__ scale) {
  const float norm = sqrtf(sumsq[0]) * inv_world;
  float c = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
  c = c < 1.f ? c : 1.f;
  scale[0] = c * inv_world;
  scale[1] = norm;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, __nv_bfloat16* __restrict__ p_bf16, int64_t n, float lr, float beta1,
                             float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                             const float* __restrict__ gscale) {
  const float gs = gscale ? gscale[0] : 1.f;
  const float step = lr / bc1;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}

// 16-byte vector form (all pointers 16-byte aligned, n4 = n/4 groups): 4 loads + 3.5 stores of 16 B per group keep
// more bytes in flight per thread than the scalar loop (30 B of HBM traffic per parameter either way).
__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, float gs, float lr, float wd,
                                           float beta1, float beta2, float eps, float step, float bc2_sqrt) {
  const float gi = g * gs;
  float pi = p * (1.f - lr * wd);
  const float mi = beta1 * m + (1.f - beta1) * gi;
  const float vi = beta2 * v + (1.f - beta2) * gi * gi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= step * (mi / denom);
  p = pi; m = mi; v = vi;
  return pi;
}

__global__ void __launch_bounds__(256)
adamw_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                  uint2* __restrict__ p_bf16, int64_t n4, float lr, float beta1, float beta2, float eps, float wd,
                  float bc1, float bc2_sqrt, const float* __restrict__ gscale) {
  const float gs = gscale ? gscale[0] : 1.f;
  const float step = lr / bc1;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pv = p[i], mv = m[i], vv = v[i];
    const float4 gv = g[i];
    const float a = adamw_one(pv.x, gv.x, mv.x, vv.x, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float b = adamw_one(pv.y, gv.y, mv.y, vv.y, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float c = adamw_one(pv.z, gv.z, mv.z, vv.z, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float d = adamw_one(pv.w, gv.w, mv.w, vv.w, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    p[i] = pv; m[i] = mv; v[i] = vv;
    if (p_bf16) p_bf16[i] = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
  }
}

// Small-footprint variant for running NEXT TO the forward GEMMs of the following step (trainer: MLA_ADAM_STREAM=1): the
// persistent GEMM CTAs take ~55 K of an SM's 64 K registers and all of its shared memory, so the update may only use
// what is left — one CTA of 128 threads per SM, two float4 per array and thread in flight (32 KB of loads per SM; the
// forward lasts ~190 ms, the update needs ~200 GB, i.e. ~1.1 TB/s of the bandwidth the tensor-bound GEMMs leave idle).
__global__ void __launch_bounds__(128)
adamw_vec4_lean_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                       uint2* __restrict__ p_bf16, int64_t n4, float lr, float beta1, float beta2, float eps, float wd,
                       float bc1, float bc2_sqrt, const float* __restrict__ gscale) {
  const float gs = gscale ? gscale[0] : 1.f;
  const float step = lr / bc1;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    const int64_t j = i + stride;
    float4 pv = p[i], mv = m[i], vv = v[i], pw = p[j], mw = m[j], vw = v[j];
    const float4 gv = g[i], gw = g[j];
    const float a = adamw_one(pv.x, gv.x, mv.x, vv.x, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float b = adamw_one(pv.y, gv.y, mv.y, vv.y, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float c = adamw_one(pv.z, gv.z, mv.z, vv.z, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float d = adamw_one(pv.w, gv.w, mv.w, vv.w, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float a2 = adamw_one(pw.x, gw.x, mw.x, vw.x, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float b2 = adamw_one(pw.y, gw.y, mw.y, vw.y, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float c2 = adamw_one(pw.z, gw.z, mw.z, vw.z, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float d2 = adamw_one(pw.w, gw.w, mw.w, vw.w, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    p[i] = pv; m[i] = mv; v[i] = vv;
    p[j] = pw; m[j] = mw; v[j] = vw;
    if (p_bf16) {
      p_bf16[i] = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
      p_bf16[j] = make_uint2(pack_bf16x2(a2, b2), pack_bf16x2(c2, d2));
    }
  }
  if (i < n4) {
    float4 pv = p[i], mv = m[i], vv = v[i];
    const float4 gv = g[i];
    const float a = adamw_one(pv.x, gv.x, mv.x, vv.x, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float b = adamw_one(pv.y, gv.y, mv.y, vv.y, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float c = adamw_one(pv.z, gv.z, mv.z, vv.z, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    const float d = adamw_one(pv.w, gv.w, mv.w, vv.w, gs, lr, wd, beta1, beta2, eps, step, bc2_sqrt);
    p[i] = pv; m[i] = mv; v[i] = vv;
    if (p_bf16) p_bf16[i] = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
  }
}

}  // namespace mla

using namespace mla;

static int g_adamw_lean = 0;
extern "C" int mla_adamw_set_lean(int32_t on) {
  g_adamw_lean = on;
  return MLA_OK;
}

extern "C" int mla_sumsq_f32(const void* x, int64_t n, void* out, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  int64_t cap = int64_t(num_sms()) * 8;
  sumsq_kernel<<<int(blocks < 1 ? 1 : (blocks < cap ? blocks : cap)), 256, 0, (cudaStream_t)stream>>>((const float*)x, n, (float*)out);
  MLA_CHECK_LAUNCH("sumsq");
  return MLA_OK;
}

extern "C" int mla_clip_coef(const void* sumsq, float max_norm, float inv_world, void* scale, void* stream) {
  if (int rc = device_check()) return rc;
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const float*)sumsq, max_norm, inv_world, (float*)scale);
  MLA_CHECK_LAUNCH("clip_coef");
  return MLA_OK;
}

extern "C" int mla_adamw_f32(void* p, const void* g, void* m, void* v, void* p_bf16, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int64_t step, const void* grad_scale,
                             void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  if (step < 1) return set_error(MLA_ERR_ARG, "adamw: step must be >= 1");
  const float bc1 = 1.f - powf(beta1, float(step));
  const float bc2 = 1.f - powf(beta2, float(step));
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v);
  if ((n & 3) == 0 && (al & 15) == 0 && (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0) {
    const int64_t n4 = n >> 2;
    if (g_adamw_lean) {
      int64_t need = (n4 + 127) / 128;
      const int64_t cap = int64_t(num_sms()) * g_adamw_lean;
      adamw_vec4_lean_kernel<<<int(need < cap ? need : cap), 128, 0, (cudaStream_t)stream>>>(
          (float4*)p, (const float4*)g, (float4*)m, (float4*)v, (uint2*)p_bf16, n4, lr, beta1, beta2, eps, weight_decay,
          bc1, sqrtf(bc2), (const float*)grad_scale);
      MLA_CHECK_LAUNCH("adamw_lean");
      return MLA_OK;
    }
    int64_t blocks4 = (n4 + 255) / 256;
    int64_t cap4 = int64_t(num_sms()) * 8;
    adamw_vec4_kernel<<<int(blocks4 < cap4 ? blocks4 : cap4), 256, 0, (cudaStream_t)stream>>>(
        (float4*)p, (const float4*)g, (float4*)m, (float4*)v, (uint2*)p_bf16, n4, lr, beta1, beta2, eps, weight_decay,
        bc1, sqrtf(bc2), (const float*)grad_scale);
    MLA_CHECK_LAUNCH("adamw");
    return MLA_OK;
  }
  int64_t blocks = (n + 255) / 256;
  int64_t cap = int64_t(num_sms()) * 16;
  adamw_kernel<<<int(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
      (float*)p, (const float*)g, (float*)m, (float*)v, (__nv_bfloat16*)p_bf16, n, lr, beta1, beta2, eps, weight_decay,
      bc1, sqrtf(bc2), (const float*)grad_scale);
  MLA_CHECK_LAUNCH("adamw");
  return MLA_OK;
}
