// Bandwidth-bound row kernels of the decoder layer: RMSNorm fwd/bwd, RoPE (in place, fwd and transposed for bwd),
// SwiGLU fwd/bwd.  One pass over HBM each, 16-byte vector accesses, fp32 math, and the reference's bf16 rounding
// points reproduced (each PyTorch op on the bf16 autocast path rounds its result to bf16).
#include <cstdlib>

#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024; result broadcast to every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();  // protect red[] reuse across consecutive calls
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < nw ? red[lane] : 0.f;
  return warp_sum(t);
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float2 p = __bfloat1622float2(h[t]);
    f[2 * t] = p.x;
    f[2 * t + 1] = p.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------ RMSNorm fwd
// modeling_llama.py:85-90:  xf = x.float(); n = xf * rsqrt(mean(xf^2) + eps); y = w * n.to(bf16)
// mode 1 ("var", timm 0.9.x RmsNorm hazard): divide by the unbiased variance instead of the mean square.
__global__ void rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                   __nv_bfloat16* __restrict__ y, float* __restrict__ rstd_out, int64_t rows, int h,
                                   int64_t ldx, int64_t ldy, float eps, int mode) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * ldx;
  const int nvec = h >> 3;
  float ss = 0.f, s1 = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + i * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { ss += f[j] * f[j]; s1 += f[j]; }
  }
  ss = block_sum(ss, red);
  float rstd;
  if (mode == 0) {
    rstd = rsqrtf(ss / float(h) + eps);
  } else {
    s1 = block_sum(s1, red);
    float mean = s1 / float(h);
    float var = (ss - float(h) * mean * mean) / float(h - 1);
    rstd = rsqrtf(var + eps);
  }
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
  __nv_bfloat16* yr = y + row * ldy;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float f[8], g[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + i * 8), f);
    unpack8(*reinterpret_cast<const uint4*>(w + i * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = g[j] * bf16_round(f[j] * rstd);
    *reinterpret_cast<uint4*>(yr + i * 8) = pack8(f);
  }
}

// Register-resident variant for h <= 256*8*VPT: every element is read from HBM exactly once (the generic kernel above
// re-reads the row for the scaling pass) and one CTA walks rows blockIdx.x, +gridDim.x, ... with the next row's loads in
// flight across the reduction.
template <int VPT>
__global__ void __launch_bounds__(256) rmsnorm_fwd_reg_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const __nv_bfloat16* __restrict__ w,
                                                              __nv_bfloat16* __restrict__ y, float* __restrict__ rstd_out,
                                                              int64_t rows, int h, int64_t ldx, int64_t ldy, float eps) {
  __shared__ float red[32];
  const int nvec = h >> 3;
  uint4 wp[VPT], nx[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int i = threadIdx.x + v * blockDim.x;
    wp[v] = i < nvec ? *reinterpret_cast<const uint4*>(w + i * 8) : make_uint4(0, 0, 0, 0);
  }
  auto prefetch = [&](int64_t r) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      nx[v] = (r < rows && i < nvec) ? *reinterpret_cast<const uint4*>(x + r * ldx + i * 8) : make_uint4(0, 0, 0, 0);
    }
  };
  int64_t row = blockIdx.x;
  prefetch(row);
  for (; row < rows; row += gridDim.x) {
    float xf[VPT][8];
    float ss = 0.f;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      unpack8(nx[v], xf[v]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += xf[v][j] * xf[v][j];
    }
    prefetch(row + gridDim.x);
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / float(h) + eps);
    if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      if (i < nvec) {
        float g[8], o[8];
        unpack8(wp[v], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j] * bf16_round(xf[v][j] * rstd);
        *reinterpret_cast<uint4*>(y + row * ldy + i * 8) = pack8(o);
      }
    }
  }
}

// Warp-per-row variant (h a multiple of 256, h <= 4096): a warp keeps its whole row in registers (NV 16-byte vectors per
// lane), so there is no block barrier at all, and with 16 warps resident per SM ~128 KB of loads are in flight per SM —
// the block-per-row kernels above top out at 0.57 of the HBM peak because only ~48 KB are.
template <int NV>
__global__ void __launch_bounds__(256, 2) rmsnorm_fwd_warp_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ y, float* __restrict__ rstd_out,
                                                               int64_t rows, int h, int64_t ldx, int64_t ldy, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const __nv_bfloat16* xr = x + row * ldx;
    uint4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const uint4*>(xr + (i * 32 + lane) * 8);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / float(h) + eps);
    if (lane == 0 && rstd_out) rstd_out[row] = rstd;
    __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float f[8], g[8];
      // opaque to the optimiser: unpack again here instead of keeping the 8*NV floats of the first pass alive (spills)
      asm volatile("" : "+r"(v[i].x), "+r"(v[i].y), "+r"(v[i].z), "+r"(v[i].w));
      unpack8(v[i], f);
      // weights: 8 KB, L1-resident; the volatile asm keeps the compiler from hoisting all NV weight loads above the loop
      // (that would double the live registers and spill)
      uint4 wq;
      asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wq.x), "=r"(wq.y), "=r"(wq.z), "=r"(wq.w)
                   : "l"(w + (i * 32 + lane) * 8));
      unpack8(wq, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * bf16_round(f[j] * rstd);
      *reinterpret_cast<uint4*>(yr + (i * 32 + lane) * 8) = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------ RMSNorm bwd
// dn = dy*w ; dx = rstd*(dn - n*mean(dn*n)) (+ dres) ; dw += sum_rows dy*n.   Each CTA walks rows blockIdx.x,
// +gridDim.x, ... keeping its dw partial in registers, then does one fp32 atomicAdd per column.
template <int VPT>  // uint4 vectors per thread (h <= blockDim*8*VPT)
__global__ void __launch_bounds__(256, 2)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ dres,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, int64_t rows, int h, float eps) {
  // One CTA walks rows blockIdx.x, +gridDim.x, ...; a thread owns the same columns of every row, so the weight
  // gradient accumulates in registers.  Both row statistics (sum x^2 and sum dn*x) come out of ONE block reduction,
  // and the next row's x / dy are prefetched (kept packed as bf16) before the reduction's barriers.
  __shared__ float red[2][32];
  const int nvec = h >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  uint4 wp[VPT];           // weights stay packed (bf16x8) to keep two CTAs per SM
  float dwacc[VPT][8];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int i = threadIdx.x + v * blockDim.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[v][j] = 0.f;
    wp[v] = make_uint4(0, 0, 0, 0);
    if (i < nvec) wp[v] = *reinterpret_cast<const uint4*>(w + i * 8);
  }
  uint4 nx[VPT], ndy[VPT], nres[VPT];
  int64_t row = blockIdx.x;
  auto prefetch = [&](int64_t r) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      nx[v] = make_uint4(0, 0, 0, 0);
      ndy[v] = make_uint4(0, 0, 0, 0);
      nres[v] = make_uint4(0, 0, 0, 0);
      if (r < rows && i < nvec) {
        nx[v] = *reinterpret_cast<const uint4*>(x + r * h + i * 8);
        ndy[v] = *reinterpret_cast<const uint4*>(dy + r * h + i * 8);
        if (dres) nres[v] = *reinterpret_cast<const uint4*>(dres + r * h + i * 8);
      }
    }
  };
  prefetch(row);
  for (; row < rows; row += gridDim.x) {
    float xf[VPT][8], dyf[VPT][8];
    uint4 res[VPT];
    float ss = 0.f, dp = 0.f;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      float wv[8];
      unpack8(nx[v], xf[v]);
      unpack8(ndy[v], dyf[v]);
      unpack8(wp[v], wv);
      res[v] = nres[v];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ss += xf[v][j] * xf[v][j];
        dp += (dyf[v][j] * wv[j]) * xf[v][j];
      }
    }
    prefetch(row + gridDim.x);   // next row's x / dy / dres stay in flight across the reduction and the stores
    // one reduction for both sums
    ss = warp_sum(ss);
    dp = warp_sum(dp);
    __syncthreads();
    if (lane == 0) { red[0][warp] = ss; red[1][warp] = dp; }
    __syncthreads();
    ss = warp_sum(lane < nw ? red[0][lane] : 0.f);
    dp = warp_sum(lane < nw ? red[1][lane] : 0.f);
    const float rstd = rsqrtf(ss / float(h) + eps);
    const float dot = rstd * dp / float(h);      // mean(dn * n), n = x * rstd
    __nv_bfloat16* dxr = dx + row * h;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      if (i < nvec) {
        float o[8], wv[8];
        unpack8(wp[v], wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float n = xf[v][j] * rstd;
          dwacc[v][j] += dyf[v][j] * bf16_round(n);   // dy * n (n is what forward multiplied by w)
          o[j] = bf16_round(rstd * (dyf[v][j] * wv[j] - n * dot));
        }
        if (dres) {
          float r[8];
          unpack8(res[v], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        *reinterpret_cast<uint4*>(dxr + i * 8) = pack8(o);
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      if (i < nvec)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw + i * 8 + j, dwacc[v][j]);
    }
  }
}

// Same arithmetic, loads staged through shared memory with cp.async: a thread copies the 16-byte chunks IT will consume
// (x, dy, dres of rows r + k*grid) RB_STAGES - 1 rows ahead, so ~70 KB per CTA (two CTAs per SM) are in flight
// regardless of the register budget — the register-prefetch kernel above keeps one row (24 KB) per CTA in flight and
// sits at 0.69 of the HBM peak.  No barrier guards the pipeline: every thread only ever reads what it copied itself.
constexpr int RB_STAGES = 4;
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VPT>
__global__ void __launch_bounds__(256, 2)
rmsnorm_bwd_pipe_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                        const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ dres,
                        __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, int64_t rows, int h, float eps) {
  extern __shared__ __align__(16) uint8_t rb_smem[];      // [RB_STAGES][3][h] bf16
  __shared__ float red[2][32];
  const int nvec = h >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const uint32_t row_bytes = uint32_t(h) * 2u;
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(rb_smem));
  uint4 wp[VPT];
  float dwacc[VPT][8];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int i = threadIdx.x + v * blockDim.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[v][j] = 0.f;
    wp[v] = make_uint4(0, 0, 0, 0);
    if (i < nvec) wp[v] = *reinterpret_cast<const uint4*>(w + i * 8);
  }
  auto issue = [&](int64_t r, int stage) {
    if (r < rows) {
      const uint32_t sb = sbase + uint32_t(stage) * 3u * row_bytes;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const int i = threadIdx.x + v * blockDim.x;
        if (i < nvec) {
          cp_async16(sb + 16u * i, x + r * h + i * 8);
          cp_async16(sb + row_bytes + 16u * i, dy + r * h + i * 8);
          if (dres) cp_async16(sb + 2u * row_bytes + 16u * i, dres + r * h + i * 8);
        }
      }
    }
    cp_async_commit();
  };
  int64_t row = blockIdx.x;
#pragma unroll
  for (int k = 0; k < RB_STAGES - 1; ++k) issue(row + int64_t(k) * gridDim.x, k);
  int stage = 0;
  for (; row < rows; row += gridDim.x) {
    issue(row + int64_t(RB_STAGES - 1) * gridDim.x, (stage + RB_STAGES - 1) % RB_STAGES);
    cp_async_wait<RB_STAGES - 1>();
    const uint8_t* sb = rb_smem + size_t(stage) * 3u * row_bytes;
    float xf[VPT][8], dyf[VPT][8];
    uint4 res[VPT];
    float ss = 0.f, dp = 0.f;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      float wv[8];
      uint4 qx = make_uint4(0, 0, 0, 0), qd = qx;
      res[v] = qx;
      if (i < nvec) {
        qx = *reinterpret_cast<const uint4*>(sb + 16u * i);
        qd = *reinterpret_cast<const uint4*>(sb + row_bytes + 16u * i);
        if (dres) res[v] = *reinterpret_cast<const uint4*>(sb + 2u * row_bytes + 16u * i);
      }
      unpack8(qx, xf[v]);
      unpack8(qd, dyf[v]);
      unpack8(wp[v], wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ss += xf[v][j] * xf[v][j];
        dp += (dyf[v][j] * wv[j]) * xf[v][j];
      }
    }
    ss = warp_sum(ss);
    dp = warp_sum(dp);
    __syncthreads();
    if (lane == 0) { red[0][warp] = ss; red[1][warp] = dp; }
    __syncthreads();
    ss = warp_sum(lane < nw ? red[0][lane] : 0.f);
    dp = warp_sum(lane < nw ? red[1][lane] : 0.f);
    const float rstd = rsqrtf(ss / float(h) + eps);
    const float dot = rstd * dp / float(h);      // mean(dn * n), n = x * rstd
    __nv_bfloat16* dxr = dx + row * h;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      if (i < nvec) {
        float o[8], wv[8];
        unpack8(wp[v], wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float n = xf[v][j] * rstd;
          dwacc[v][j] += dyf[v][j] * bf16_round(n);   // dy * n (n is what forward multiplied by w)
          o[j] = bf16_round(rstd * (dyf[v][j] * wv[j] - n * dot));
        }
        if (dres) {
          float r[8];
          unpack8(res[v], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        *reinterpret_cast<uint4*>(dxr + i * 8) = pack8(o);
      }
    }
    stage = (stage + 1) % RB_STAGES;
  }
  cp_async_wait<0>();
  if (dw) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * blockDim.x;
      if (i < nvec)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw + i * 8 + j, dwacc[v][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ RoPE
// modeling_llama.py:184-208 on the bf16 path: out = bf16(bf16(x*cos) + bf16(rotate_half(x)*sin)), cos/sin bf16.
// Applied in place to the q and k column blocks of the fused [T, ld] projection buffer.  sign=-1 is the transpose
// (the backward of the rotation).  One thread per (token, head, 8-wide chunk of the first half).
__global__ void rope_kernel(__nv_bfloat16* __restrict__ base, const __nv_bfloat16* __restrict__ cos_t,
                            const __nv_bfloat16* __restrict__ sin_t, int64_t tokens, int seq, int heads, int d,
                            int64_t ld, float sign) {
  const int half = d >> 1;
  const int cpr = half >> 3;  // chunks per (token, head)
  const int64_t total = tokens * heads * cpr;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % cpr);
    const int hd = int((idx / cpr) % heads);
    const int64_t t = idx / (int64_t(cpr) * heads);
    const int pos = int(t % seq);
    __nv_bfloat16* p = base + t * ld + int64_t(hd) * d + c * 8;
    float x1[8], x2[8], cs[8], sn[8];
    unpack8(*reinterpret_cast<const uint4*>(p), x1);
    unpack8(*reinterpret_cast<const uint4*>(p + half), x2);
    unpack8(*reinterpret_cast<const uint4*>(cos_t + int64_t(pos) * half + c * 8), cs);
    unpack8(*reinterpret_cast<const uint4*>(sin_t + int64_t(pos) * half + c * 8), sn);
    float o1[8], o2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = sign * sn[j];
      o1[j] = bf16_round(x1[j] * cs[j]) + bf16_round(-x2[j] * s);
      o2[j] = bf16_round(x2[j] * cs[j]) + bf16_round(x1[j] * s);
    }
    *reinterpret_cast<uint4*>(p) = pack8(o1);
    *reinterpret_cast<uint4*>(p + half) = pack8(o2);
  }
}

// ------------------------------------------------------------------------------------------------ SwiGLU
// modeling_llama.py:240: act_fn(gate_proj(x)) * up_proj(x) with bf16 rounding after silu and after the product.
// gu is the fused [T, 2f] projection (gate columns [0,f), up columns [f,2f)).
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ out,
                                  int64_t rows, int f) {
  const int cpr = f >> 3;
  const int64_t total = rows * cpr;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / cpr;
    const int c = int(idx % cpr);
    float g[8], u[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(gu + r * 2 * f + c * 8), g);
    unpack8(*reinterpret_cast<const uint4*>(gu + r * 2 * f + f + c * 8), u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bf16_round(g[j] * sigmoidf_(g[j])) * u[j];
    *reinterpret_cast<uint4*>(out + r * f + c * 8) = pack8(o);
  }
}

// d_gu[:, :f] = d_act*u*silu'(g) ; d_gu[:, f:] = d_act*silu(g)
// With act_out the kernel also re-materialises act = bf16(silu(g)) * u (what swiglu_fwd produced), which backward needs
// as the operand of the down-projection weight gradient: one pass over gate|up instead of two.
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ dact, const __nv_bfloat16* __restrict__ gu,
                                  __nv_bfloat16* __restrict__ dgu, __nv_bfloat16* __restrict__ act_out, int64_t rows,
                                  int f) {
  const int cpr = f >> 3;
  const int64_t total = rows * cpr;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / cpr;
    const int c = int(idx % cpr);
    float g[8], u[8], d[8], dg[8], du[8], ac[8];
    unpack8(*reinterpret_cast<const uint4*>(gu + r * 2 * f + c * 8), g);
    unpack8(*reinterpret_cast<const uint4*>(gu + r * 2 * f + f + c * 8), u);
    unpack8(*reinterpret_cast<const uint4*>(dact + r * f + c * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) swiglu_bwd_elem(g[j], u[j], d[j], dg[j], du[j], ac[j]);
    *reinterpret_cast<uint4*>(dgu + r * 2 * f + c * 8) = pack8(dg);
    *reinterpret_cast<uint4*>(dgu + r * 2 * f + f + c * 8) = pack8(du);
    if (act_out) *reinterpret_cast<uint4*>(act_out + r * f + c * 8) = pack8(ac);
  }
}

static inline int ew_grid(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  int64_t cap = int64_t(num_sms()) * 16;
  return int(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace mla

using namespace mla;

extern "C" int mla_rmsnorm_fwd(const void* x, const void* w, void* y, void* rstd, int64_t rows, int32_t h, int64_t ldx,
                               int64_t ldy, float eps, int32_t mode, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (h <= 0 || (h & 7) || (ldx & 7) || (ldy & 7)) return set_error(MLA_ERR_ARG, "rmsnorm_fwd: h, ldx, ldy must be multiples of 8");
  if (mode != 0 && mode != 1) return set_error(MLA_ERR_ARG, "rmsnorm_fwd: mode must be 0 (mean-square) or 1 (variance)");
  if (mode == 0 && (h & 255) == 0 && h <= 4096 && rows >= 4096) {
    // large token counts (the decoder's norms): warp-per-row
    auto s_ = (cudaStream_t)stream;
    const int64_t want = (rows + 7) / 8;
    const int grid = int(want < int64_t(num_sms()) * 2 ? want : int64_t(num_sms()) * 2);
#define LAUNCH_RW(NV_)                                                                                            \
  rmsnorm_fwd_warp_kernel<NV_><<<grid, 256, 0, s_>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w,            \
                                                     (__nv_bfloat16*)y, (float*)rstd, rows, h, ldx, ldy, eps)
    switch (h / 256) {
      case 1: LAUNCH_RW(1); break;
      case 2: LAUNCH_RW(2); break;
      case 4: LAUNCH_RW(4); break;
      case 8: LAUNCH_RW(8); break;
      case 16: LAUNCH_RW(16); break;
      default: goto generic;
    }
#undef LAUNCH_RW
    MLA_CHECK_LAUNCH("rmsnorm_fwd");
    return MLA_OK;
  }
generic:
  int block = (h / 8 + 31) / 32 * 32;
  block = block > 256 ? 256 : block;
  const int vpt = (h / 8 + block - 1) / block;
  if (mode == 0 && vpt <= 4) {
    const int grid = int(rows < int64_t(num_sms()) * 6 ? rows : int64_t(num_sms()) * 6);
    auto s_ = (cudaStream_t)stream;
#define LAUNCH_RF(V)                                                                                             \
  rmsnorm_fwd_reg_kernel<V><<<grid, block, 0, s_>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w,            \
                                                    (__nv_bfloat16*)y, (float*)rstd, rows, h, ldx, ldy, eps)
    if (vpt == 1) LAUNCH_RF(1);
    else if (vpt == 2) LAUNCH_RF(2);
    else if (vpt == 3) LAUNCH_RF(3);
    else LAUNCH_RF(4);
#undef LAUNCH_RF
    MLA_CHECK_LAUNCH("rmsnorm_fwd");
    return MLA_OK;
  }
  rmsnorm_fwd_kernel<<<(unsigned)rows, block, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, (float*)rstd, rows, h, ldx, ldy, eps, mode);
  MLA_CHECK_LAUNCH("rmsnorm_fwd");
  return MLA_OK;
}

extern "C" int mla_rmsnorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, void* dw,
                               int64_t rows, int32_t h, float eps, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (h <= 0 || (h & 7)) return set_error(MLA_ERR_ARG, "rmsnorm_bwd: h must be a multiple of 8");
  if (h > 256 * 8 * 4) return set_error(MLA_ERR_ARG, "rmsnorm_bwd: h > 8192 unsupported");
  const int nvec = h / 8;
  int block = (nvec + 31) / 32 * 32;
  block = block > 256 ? 256 : block;
  const int vpt = (nvec + block - 1) / block;
  int grid = int(rows < int64_t(num_sms()) * 2 ? rows : int64_t(num_sms()) * 2);
  auto s = (cudaStream_t)stream;
  {
    // cp.async-staged variant: needs 4 stages x 3 rows in shared memory twice per SM, and enough rows to fill the pipe
    static int pipe = -1;
    if (pipe < 0) {
      const char* e = getenv("MLA_RMSNORM_BWD_PIPE");
      pipe = (e && e[0] == '0') ? 0 : 1;
    }
    const size_t smem = size_t(RB_STAGES) * 3 * size_t(h) * 2;
    if (pipe && vpt <= 2 && block == 256 && smem <= 100 * 1024 && rows >= int64_t(grid) * RB_STAGES &&
        ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dres)) & 15) == 0) {
      auto kern = vpt == 1 ? rmsnorm_bwd_pipe_kernel<1> : rmsnorm_bwd_pipe_kernel<2>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "rmsnorm_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      kern<<<grid, 256, smem, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                   (const __nv_bfloat16*)dres, (__nv_bfloat16*)dx, (float*)dw, rows, h, eps);
      MLA_CHECK_LAUNCH("rmsnorm_bwd_pipe");
      return MLA_OK;
    }
  }
#define LAUNCH_RB(V)                                                                                              \
  rmsnorm_bwd_kernel<V><<<grid, block, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,                 \
                                                (const __nv_bfloat16*)w, (const __nv_bfloat16*)dres,              \
                                                (__nv_bfloat16*)dx, (float*)dw, rows, h, eps)
  if (vpt == 1) LAUNCH_RB(1);
  else if (vpt == 2) LAUNCH_RB(2);
  else if (vpt == 3) LAUNCH_RB(3);
  else LAUNCH_RB(4);
#undef LAUNCH_RB
  MLA_CHECK_LAUNCH("rmsnorm_bwd");
  return MLA_OK;
}

extern "C" int mla_rope_inplace(void* base, const void* cos_t, const void* sin_t, int64_t tokens, int32_t seq,
                                int32_t heads, int32_t d, int64_t ld, int32_t transpose, void* stream) {
  if (int rc = device_check()) return rc;
  if (tokens <= 0 || heads <= 0) return MLA_OK;
  if (d <= 0 || (d & 15) || (ld & 7)) return set_error(MLA_ERR_ARG, "rope: head dim must be a multiple of 16 and ld of 8");
  if (seq <= 0) return set_error(MLA_ERR_ARG, "rope: seq must be positive");
  const int64_t total = tokens * heads * (d / 16);
  rope_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)base, (const __nv_bfloat16*)cos_t,
                                                                     (const __nv_bfloat16*)sin_t, tokens, seq, heads, d,
                                                                     ld, transpose ? -1.f : 1.f);
  MLA_CHECK_LAUNCH("rope");
  return MLA_OK;
}

extern "C" int mla_swiglu_fwd(const void* gu, void* out, int64_t rows, int32_t f, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (f <= 0 || (f & 7)) return set_error(MLA_ERR_ARG, "swiglu_fwd: f must be a multiple of 8");
  swiglu_fwd_kernel<<<ew_grid(rows * (f / 8), 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)gu,
                                                                                    (__nv_bfloat16*)out, rows, f);
  MLA_CHECK_LAUNCH("swiglu_fwd");
  return MLA_OK;
}

extern "C" int mla_swiglu_bwd(const void* dact, const void* gu, void* dgu, int64_t rows, int32_t f, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (f <= 0 || (f & 7)) return set_error(MLA_ERR_ARG, "swiglu_bwd: f must be a multiple of 8");
  swiglu_bwd_kernel<<<ew_grid(rows * (f / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dact, (const __nv_bfloat16*)gu, (__nv_bfloat16*)dgu, nullptr, rows, f);
  MLA_CHECK_LAUNCH("swiglu_bwd");
  return MLA_OK;
}

extern "C" int mla_swiglu_bwd_act(const void* dact, const void* gu, void* dgu, void* act_out, int64_t rows, int32_t f,
                                  void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  if (f <= 0 || (f & 7)) return set_error(MLA_ERR_ARG, "swiglu_bwd_act: f must be a multiple of 8");
  if (act_out == nullptr) return set_error(MLA_ERR_ARG, "swiglu_bwd_act: null act_out");
  swiglu_bwd_kernel<<<ew_grid(rows * (f / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dact, (const __nv_bfloat16*)gu, (__nv_bfloat16*)dgu, (__nv_bfloat16*)act_out, rows, f);
  MLA_CHECK_LAUNCH("swiglu_bwd_act");
  return MLA_OK;
}
