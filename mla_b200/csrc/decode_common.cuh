// Helpers shared by the per-op decode kernels (decode.cu) and the whole-stack persistent kernel (decode_stack.cu).
#pragma once
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

constexpr int GV_CWARPS = 16;
constexpr int GV_CONSUMERS = GV_CWARPS * 32;
constexpr int GV_THREADS = GV_CONSUMERS + 32;
constexpr int GV_MAX_STAGES = 8;
constexpr int GV_RING_BYTES = 192 * 1024;
enum { GV_PRO_NONE = 0, GV_PRO_RMSNORM = 1, GV_PRO_SWIGLU = 2 };

__device__ __forceinline__ void bulk_load_row(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// programmatic dependent launch: no-ops unless the launch carried the attribute / has a programmatic dependent
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GV_CONSUMERS) : "memory"); }
__device__ __forceinline__ uint4 pack8f(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, float acc) {
  float fa[8], fb[8];
  unpack8(a, fa);
  unpack8(b, fb);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc = fmaf(fa[e], fb[e], acc);
  return acc;
}

// Sum V per-lane values across the warp with V - 1 + (5 - log2 V) shuffles instead of 5 V: while more than one value
// is left, the two halves of the warp exchange the half of the values the other one keeps.  On return lane l holds in
// v[0] the total of value (l * V) >> 5 (for V >= 32: value l).
template <int V>
__device__ __forceinline__ void warp_reduce_many(float (&v)[V], int lane) {
  static_assert(V >= 1 && V <= 32 && (V & (V - 1)) == 0, "V must be a power of two <= 32");
  int cur = V;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (cur > 1) {
      const int half = cur >> 1;
      const bool hi = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        if (i < half) {
          const float send = hi ? v[i] : v[i + half];
          const float keep = hi ? v[i + half] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      cur = half;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
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
constexpr int DEC_WARPS = 16;

// raw (packed bf16) loads first, conversion later: the loads of all UN keys are in flight before the first use
template <int EPL> struct RawEpl { uint32_t w[(EPL + 1) / 2]; };
template <int EPL>
__device__ __forceinline__ RawEpl<EPL> ld_raw(const __nv_bfloat16* p) {
  RawEpl<EPL> r;
  if constexpr (EPL == 8) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    r.w[0] = u.x; r.w[1] = u.y; r.w[2] = u.z; r.w[3] = u.w;
  } else if constexpr (EPL == 4) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    r.w[0] = u.x; r.w[1] = u.y;
  } else if constexpr (EPL == 2) {
    r.w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
  } else {
    r.w[0] = uint32_t(__ldg(reinterpret_cast<const unsigned short*>(p)));
  }
  return r;
}
template <int EPL>
__device__ __forceinline__ void cvt_raw(const RawEpl<EPL>& r, float* f) {
  if constexpr (EPL == 1) {
    f[0] = __uint_as_float(r.w[0] << 16);
  } else {
#pragma unroll
    for (int i = 0; i < EPL / 2; ++i) {
      f[2 * i] = __uint_as_float(r.w[i] << 16);
      f[2 * i + 1] = __uint_as_float(r.w[i] & 0xffff0000u);
    }
  }
}
// RoPE of one head row spread over the warp (lane holds dims [lane*EPL, lane*EPL+EPL); lanes 0-15 hold the first half,
// their partners lane^16 the second): o1 = bf16(x1 c) + bf16(-x2 s), o2 = bf16(x2 c) + bf16(x1 s), rounded to bf16 —
// the arithmetic of rope_kernel (norm_rope_act.cu; modeling_llama.py:184-208).  cs/sn: table row of this position.
template <int EPL>
__device__ __forceinline__ void rope_lanes(float* x, const __nv_bfloat16* __restrict__ cs,
                                           const __nv_bfloat16* __restrict__ sn, int lane) {
  const int col = (lane & 15) * EPL;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const float other = __shfl_xor_sync(0xffffffffu, x[e], 16);
    const float c = __bfloat162float(cs[col + e]), sgn = __bfloat162float(sn[col + e]);
    const float t = lane < 16 ? bf16_round(x[e] * c) + bf16_round(-other * sgn) : bf16_round(x[e] * c) + bf16_round(other * sgn);
    x[e] = bf16_round(t);
  }
}

}  // namespace mla
