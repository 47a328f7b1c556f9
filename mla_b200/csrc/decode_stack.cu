// One DDIM step's trip through the WHOLE decoder stack as ONE persistent kernel (inference denoise loop,
// models/mla/model_mla.py:592-775; per layer: modeling_llama.py:405-597 on the few suffix rows, prefix K/V cached).
//
// Why: the per-op path (decode.cu, 5 launches per layer) streams 12.95 GB of weights per step at 0.57 of the HBM peak.
// Each of its kernels is 15-38 us long and owns the whole shared memory of an SM, so the next kernel's CTA cannot become
// resident — and start prefetching ITS weights — before the previous one has drained: every launch boundary empties
// the memory pipe (ramp + tail ~4 us of a ~22 us kernel), programmatic dependent launch notwithstanding.
// Here ONE CTA per SM lives for all layers:
//   * the producer thread streams this CTA's share of EVERY weight matrix, layer after layer, through one
//     shared-memory ring (cp.async.bulk per weight row, full/empty mbarriers).  It never waits for anything but a free
//     slot: weights do not depend on activations, so it runs ahead across operator and layer boundaries and the HBM
//     queue stays full while the consumers synchronise;
//   * the 16 consumer warps go through the five phases of a layer — RMSNorm+QKV, attention, O-proj+residual,
//     RMSNorm+gate|up, SwiGLU+down+residual — separated by a grid barrier (sense-reversing counter in global memory:
//     ~1 us, hidden behind the ring's ~3 us of buffered weights);
//   * attention runs split-K over all SMs (item = (sample, head, 128-key split), last arrival merges), the prefix
//     K/V of the layer having been pulled into L2 by a bulk prefetch the producer issues one phase earlier.
// The arithmetic of every phase is the per-op kernels' (same thread -> k-chunk mapping, same reduction trees, same
// rounding points; attention = decode_attn_kernel<.,ROPE=1> with split-K), so the result is BIT-IDENTICAL to
// LlamaDecoderLayer.decode run with split-K attention — that is what tests/test_decode_stack_gpu.py asserts.
// Activations written by other CTAs inside this launch are read with ld.global.cg (L2): L1 is not coherent.
// Limits: B*n <= 2 suffix rows (activations live in registers), h <= 12288, f <= 12288, head_dim in {32, 64, 128}.
#include <cstdlib>

#include "decode_common.cuh"

namespace mla {

constexpr int ST_MAX_STAGES = 8;
constexpr int ST_ATTN_UN = 8;                       // keys per warp and split: 16 warps x 8 = 128 keys per item
constexpr int ST_SPLIT_KEYS = DEC_WARPS * ST_ATTN_UN;

struct StackParams {
  const __nv_bfloat16* const* wqkv;     // [L] device pointers: [3h, h]
  const __nv_bfloat16* const* wo;       // [h, h]
  const __nv_bfloat16* const* wgu;      // [2f, h]
  const __nv_bfloat16* const* wd;       // [h, f]
  const __nv_bfloat16* const* ln1;      // [h]
  const __nv_bfloat16* const* ln2;      // [h]
  const __nv_bfloat16* const* cache;    // [B, 2, H, P, D] rotated prefix keys | values (head-major)
  __nv_bfloat16 *x, *qkv, *ctx, *xmid, *gu;     // [M,h] (in/out) | [M,3h] | [M,h] | [M,h] | [M,2f]
  const __nv_bfloat16 *cos, *sin;       // [n, D/2]: table rows of positions P..P+n-1
  float* attn_ws;                       // [B*H*n, S, D+2] split-K partials
  int* attn_cnt;                        // [B*H] arrival counters (zero on entry, left zero)
  unsigned* bar;                        // [2] grid barrier: arrivals, generation (zero-initialised once)
  int L, B, n, P, H, D, h, f;
  float eps, scale;
  int stages;
  uint32_t slot_bytes;
};

__device__ __forceinline__ uint4 ld_cg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float ld_cg_bf16(const __nv_bfloat16* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return __uint_as_float(uint32_t(v) << 16);
}
template <int EPL>
__device__ __forceinline__ RawEpl<EPL> ld_raw_cg(const __nv_bfloat16* p) {
  RawEpl<EPL> r;
  if constexpr (EPL == 4) {
    const uint2 u = __ldcg(reinterpret_cast<const uint2*>(p));
    r.w[0] = u.x; r.w[1] = u.y;
  } else if constexpr (EPL == 2) {
    r.w[0] = __ldcg(reinterpret_cast<const unsigned int*>(p));
  } else {
    unsigned short v;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
    r.w[0] = uint32_t(v);
  }
  return r;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// All consumer threads of all CTAs.  Sense-reversing: bar[0] counts arrivals and is reset by the last one, bar[1] is the
// generation everybody else spins on — no host-side reset between launches.
__device__ __forceinline__ void grid_barrier(unsigned* bar, int tid) {
  consumers_sync();
  if (tid == 0) {
    const unsigned gen = ld_acquire_u32(bar + 1);
    __threadfence();
    if (atomicAdd(bar, 1u) == gridDim.x - 1) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (ld_acquire_u32(bar + 1) == gen) {
      }
    }
    __threadfence();
  }
  consumers_sync();
}

struct Ring {
  uint8_t* base;
  uint64_t *full_bar, *empty_bar;
  int stages;
  uint32_t slot_bytes;
};

// the linear phases of a layer, in order: which weight, how many output rows, contraction length
struct PhaseW {
  const __nv_bfloat16* w;
  int N, K;
};
__device__ __forceinline__ PhaseW phase_weights(const StackParams& p, int l, int ph) {
  switch (ph) {
    case 0: return {p.wqkv[l], 3 * p.h, p.h};
    case 1: return {p.wo[l], p.h, p.h};
    case 2: return {p.wgu[l], 2 * p.f, p.h};
    default: return {p.wd[l], p.h, p.f};
  }
}
__device__ __forceinline__ int rows_per_slot(int K) { return K <= GV_CONSUMERS * 8 ? 4 : 2; }

// group g of a phase goes to CTA (off + g) % grid, where off continues the round-robin of the previous phase: the CTAs
// that get one group more than the others are different ones in every phase
struct Deal {
  int off;
  __device__ __forceinline__ int first(int cta, int grid) const { int d = cta - off; return d < 0 ? d + grid : d; }
  __device__ __forceinline__ void next(int groups, int grid) { off = (off + groups) % grid; }
};

// ---------------------------------------------------------------------------------------------- producer
__device__ void stack_producer(const StackParams& p, const Ring& r) {
  int it = 0;
  Deal deal{0};
  const int grid = gridDim.x, cta = blockIdx.x;
  for (int l = 0; l < p.L; ++l) {
    for (int ph = 0; ph < 4; ++ph) {
      const PhaseW w = phase_weights(p, l, ph);
      const int rpi = rows_per_slot(w.K);
      const uint32_t row_bytes = uint32_t(w.K) * 2u, pitch = (row_bytes + 127u) & ~127u;
      const int groups = (w.N + rpi - 1) / rpi;
      for (int g = deal.first(cta, grid); g < groups; g += grid, ++it) {
        const int s = it % r.stages;
        const uint32_t par = (it / r.stages) & 1;
        mbar_wait(&r.empty_bar[s], par ^ 1);
        const int n0 = g * rpi;
        const int valid = w.N - n0 < rpi ? w.N - n0 : rpi;
        mbar_arrive_expect_tx(&r.full_bar[s], row_bytes * valid);
        for (int rr = 0; rr < valid; ++rr)
          bulk_load_row(r.base + size_t(s) * r.slot_bytes + size_t(rr) * pitch, w.w + int64_t(n0 + rr) * w.K, row_bytes,
                        &r.full_bar[s]);
      }
      deal.next(groups, grid);
      if (ph == 0) {
        // this layer's prefix K/V -> L2 while the QKV projection is still being consumed: the attention phase that
        // follows then pays L2, not DRAM, latency.  Each CTA asks for its 1/grid of the cache in 16 KB pieces.
        const size_t total = size_t(p.B) * 2 * p.H * p.P * p.D * 2;
        const size_t per = ((total + grid - 1) / grid + 15) & ~size_t(15);
        size_t beg = per * cta;
        const size_t end = beg + per < total ? beg + per : (total & ~size_t(15));
        const uint8_t* c = reinterpret_cast<const uint8_t*>(p.cache[l]);
        for (; beg < end; beg += 16384) prefetch_l2_bulk(c + beg, uint32_t(end - beg < 16384 ? end - beg : 16384));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- linear phase (consumers)
// out[m, n] = bf16( bf16(sum_k x'[m,k] w[n,k]) + res[m,n] ): the fast path of gemv_ring_kernel (XF = 1) with coherent
// activation loads and the ring / deal state carried across phases.
template <int MB, int CPT, int RPI, int PRO>
__device__ __forceinline__ void stack_linear(const Ring& r, int& it, int& pb, const Deal& deal, float* partial /*[2][16][8]*/,
                                             float* red_ss /*[16][2]*/, const __nv_bfloat16* x, int64_t ldx,
                                             const __nv_bfloat16* ln_w, const __nv_bfloat16* res, int64_t ldr,
                                             __nv_bfloat16* out, int64_t ldo, int M, int N, int K, float eps) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunks = K >> 3;
  const uint32_t pitch = (uint32_t(K) * 2u + 127u) & ~127u;
  const int groups = (N + RPI - 1) / RPI;
  float xf[MB][CPT][8];
  {
    uint4 xr[MB][CPT];
    float ss[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) ss[m] = 0.f;
#pragma unroll
    for (int m = 0; m < MB; ++m) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = tid + j * GV_CONSUMERS;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (m < M && c < chunks) {
          if (PRO == GV_PRO_SWIGLU) {
            float g[8], u[8], o[8];
            unpack8(ld_cg16(x + int64_t(m) * ldx + 8 * c), g);
            unpack8(ld_cg16(x + int64_t(m) * ldx + K + 8 * c), u);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = bf16_round(g[e] * (1.f / (1.f + __expf(-g[e])))) * u[e];
            v = pack8f(o);
          } else {
            v = ld_cg16(x + int64_t(m) * ldx + 8 * c);
            if (PRO == GV_PRO_RMSNORM) {
              float f[8];
              unpack8(v, f);
#pragma unroll
              for (int e = 0; e < 8; ++e) ss[m] = fmaf(f[e], f[e], ss[m]);
            }
          }
        }
        xr[m][j] = v;
      }
    }
    if (PRO == GV_PRO_RMSNORM) {
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        const float t = d_wsum(ss[m]);
        if (lane == 0) red_ss[warp * 2 + m] = t;
      }
      consumers_sync();
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < GV_CWARPS; ++w2) t += red_ss[w2 * 2 + m];
        const float rstd = rsqrtf(t / float(K) + eps);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int c = tid + j * GV_CONSUMERS;
          if (c < chunks) {
            float f[8], g[8], o[8];
            unpack8(xr[m][j], f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(ln_w + 8 * c)), g);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g[e] * bf16_round(f[e] * rstd);
            xr[m][j] = pack8f(o);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
      for (int j = 0; j < CPT; ++j) unpack8(xr[m][j], xf[m][j]);
  }

  constexpr int V = RPI * MB;
  for (int g = deal.first(blockIdx.x, gridDim.x); g < groups; g += gridDim.x, ++it) {
    const int s = it % r.stages;
    const uint32_t par = (it / r.stages) & 1;
    mbar_wait(&r.full_bar[s], par);
    const uint8_t* slot = r.base + size_t(s) * r.slot_bytes;
    const int n0 = g * RPI;
    float acc[RPI][MB];
#pragma unroll
    for (int rr = 0; rr < RPI; ++rr)
#pragma unroll
      for (int m = 0; m < MB; ++m) acc[rr][m] = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tid + j * GV_CONSUMERS;
      if (c < chunks) {
#pragma unroll
        for (int rr = 0; rr < RPI; ++rr) {
          const uint4 wv = *reinterpret_cast<const uint4*>(slot + size_t(rr) * pitch + 16 * c);
          float wf[8];
          unpack8(wv, wf);
#pragma unroll
          for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[rr][m] = fmaf(wf[e], xf[m][j][e], acc[rr][m]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.empty_bar[s]);      // this warp no longer reads the slot
    float flat[V];
#pragma unroll
    for (int rr = 0; rr < RPI; ++rr)
#pragma unroll
      for (int m = 0; m < MB; ++m) flat[rr * MB + m] = acc[rr][m];
    warp_reduce_many<V>(flat, lane);
    constexpr int LPV = 32 / V;
    float* part = partial + pb * (GV_CWARPS * 8);
    if (lane % LPV == 0) part[warp * V + lane / LPV] = flat[0];
    consumers_sync();
    if (tid < V) {
      const int rr = tid / MB, m = tid % MB;
      const int n = n0 + rr;
      if (n < N && m < M) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < GV_CWARPS; ++w2) t += part[w2 * V + tid];
        float v = bf16_round(t);
        if (res) v += ld_cg_bf16(res + int64_t(m) * ldr + n);
        out[int64_t(m) * ldo + n] = __float2bfloat16_rn(v);
      }
    }
    pb ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------- attention phase
// decode_attn_kernel<EPL, ROPE = 1> with split-K, every item (b, h, split) serving all n <= NQ query rows of the sample
// from ONE load of the keys / values.  Per query the arithmetic is that kernel's: warp w takes keys jbeg + w + 16 u in
// order, the 16 warps are merged in order, the splits are merged in order by the last CTA to arrive.
template <int EPL, int NQ>
__device__ __forceinline__ void stack_attention(const StackParams& p, const __nv_bfloat16* cache, float* s_m /*[NQ][16]*/,
                                                float* s_l, float* s_acc /*[NQ][16][D]*/, int* s_last) {
  constexpr int D = 32 * EPL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.n, P = p.P, H = p.H, hdim = H * D;
  const int Lk = P + n;
  const int S = (Lk + ST_SPLIT_KEYS - 1) / ST_SPLIT_KEYS;
  const int items = p.B * H * S;
  const int64_t ldq = 3 * int64_t(hdim);
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int sp = item % S, h = (item / S) % H, b = item / (S * H);
    float qf[NQ][EPL], acc[NQ][EPL], m[NQ], l[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      m[i] = -INFINITY;
      l[i] = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) { qf[i][e] = 0.f; acc[i][e] = 0.f; }
      if (i < n) {
        cvt_raw<EPL>(ld_raw_cg<EPL>(p.qkv + (int64_t(b) * n + i) * ldq + int64_t(h) * D + lane * EPL), qf[i]);
        rope_lanes<EPL>(qf[i], p.cos + int64_t(i) * (D / 2), p.sin + int64_t(i) * (D / 2), lane);
#pragma unroll
        for (int e = 0; e < EPL; ++e) qf[i][e] *= p.scale;
      }
    }
    const __nv_bfloat16* kc = cache + (int64_t(b) * 2 * H + h) * int64_t(P) * D + lane * EPL;
    const __nv_bfloat16* vc = kc + int64_t(H) * P * D;
    const int jbeg = sp * ST_SPLIT_KEYS;
    const int jend = Lk < jbeg + ST_SPLIT_KEYS ? Lk : jbeg + ST_SPLIT_KEYS;     // keys any query of the sample may see
    RawEpl<EPL> kr[ST_ATTN_UN], vr[ST_ATTN_UN];
#pragma unroll
    for (int u = 0; u < ST_ATTN_UN; ++u) {
      const int j = jbeg + warp + u * DEC_WARPS;
#pragma unroll
      for (int w2 = 0; w2 < (EPL + 1) / 2; ++w2) { kr[u].w[w2] = 0u; vr[u].w[w2] = 0u; }
      if (j < jend) {
        if (j >= P) {
          const __nv_bfloat16* row = p.qkv + (int64_t(b) * n + (j - P)) * ldq + int64_t(h) * D + lane * EPL;
          kr[u] = ld_raw_cg<EPL>(row + hdim);
          vr[u] = ld_raw_cg<EPL>(row + 2 * hdim);
        } else {
          kr[u] = ld_raw<EPL>(kc + int64_t(j) * D);
          vr[u] = ld_raw<EPL>(vc + int64_t(j) * D);
        }
      }
    }
    float s[NQ][ST_ATTN_UN];
#pragma unroll
    for (int u = 0; u < ST_ATTN_UN; ++u) {
      const int j = jbeg + warp + u * DEC_WARPS;
      float kf[EPL];
      cvt_raw<EPL>(kr[u], kf);
      if (j >= P && j < jend) rope_lanes<EPL>(kf, p.cos + int64_t(j - P) * (D / 2), p.sin + int64_t(j - P) * (D / 2), lane);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        s[i][u] = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) s[i][u] = fmaf(qf[i][e], kf[e], s[i][u]);
      }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
      for (int u = 0; u < ST_ATTN_UN; ++u) s[i][u] = d_wsum(s[i][u]);
#pragma unroll
    for (int u = 0; u < ST_ATTN_UN; ++u) {
      const int j = jbeg + warp + u * DEC_WARPS;
      float vf[EPL];
      cvt_raw<EPL>(vr[u], vf);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        if (i < n && j < jend && j <= P + i) {          // query i sees keys j <= P + i
          const float mn = fmaxf(m[i], s[i][u]);
          const float corr = __expf(m[i] - mn), pj = __expf(s[i][u] - mn);
          l[i] = l[i] * corr + pj;
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[i][e] = acc[i][e] * corr + pj * vf[e];
          m[i] = mn;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      if (lane == 0) { s_m[i * DEC_WARPS + warp] = m[i]; s_l[i * DEC_WARPS + warp] = l[i]; }
#pragma unroll
      for (int e = 0; e < EPL; ++e) s_acc[(i * DEC_WARPS + warp) * D + lane * EPL + e] = acc[i][e];
    }
    consumers_sync();
    for (int t = tid; t < n * D; t += GV_CONSUMERS) {
      const int i = t / D, d = t % D;
      float mx = -INFINITY;
#pragma unroll
      for (int w2 = 0; w2 < DEC_WARPS; ++w2) mx = fmaxf(mx, s_m[i * DEC_WARPS + w2]);
      float Ls = 0.f, a = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < DEC_WARPS; ++w2) {
        const float mw = s_m[i * DEC_WARPS + w2];
        const float c = mw == -INFINITY ? 0.f : __expf(mw - mx);
        Ls += s_l[i * DEC_WARPS + w2] * c;
        a += s_acc[(i * DEC_WARPS + w2) * D + d] * c;
      }
      const int64_t unit = (int64_t(b) * H + h) * n + i;
      float* part = p.attn_ws + (unit * S + sp) * (D + 2);
      part[2 + d] = a;
      if (d == 0) { part[0] = mx; part[1] = Ls; }
    }
    __threadfence();                       // this CTA's partials are visible before its arrival is counted
    consumers_sync();
    if (tid == 0) *s_last = atomicAdd(p.attn_cnt + (b * H + h), 1) == S - 1;
    consumers_sync();
    if (*s_last) {
      __threadfence();
      for (int t = tid; t < n * D; t += GV_CONSUMERS) {
        const int i = t / D, d = t % D;
        const int64_t unit = (int64_t(b) * H + h) * n + i;
        const float* all = p.attn_ws + unit * S * (D + 2);
        float mx = -INFINITY;
        for (int s2 = 0; s2 < S; ++s2) mx = fmaxf(mx, __ldcg(all + s2 * (D + 2)));
        float Ls = 0.f, a = 0.f;
        for (int s2 = 0; s2 < S; ++s2) {
          const float ms = __ldcg(all + s2 * (D + 2));
          const float c = ms == -INFINITY ? 0.f : __expf(ms - mx);
          Ls += __ldcg(all + s2 * (D + 2) + 1) * c;
          a += __ldcg(all + s2 * (D + 2) + 2 + d) * c;
        }
        p.ctx[(int64_t(b) * n + i) * hdim + int64_t(h) * D + d] = __float2bfloat16_rn(Ls > 0.f ? a / Ls : 0.f);
      }
      if (tid == 0) p.attn_cnt[b * H + h] = 0;       // re-armed for the next layer / launch
    }
    consumers_sync();                      // s_m / s_l / s_acc / s_last are re-used by the next item
  }
}

template <int MB, int PRO>
__device__ __forceinline__ void stack_linear_k(const Ring& r, int& it, int& pb, const Deal& deal, float* partial,
                                               float* red_ss, const __nv_bfloat16* x, int64_t ldx,
                                               const __nv_bfloat16* ln_w, const __nv_bfloat16* res, int64_t ldr,
                                               __nv_bfloat16* out, int64_t ldo, int M, int N, int K, float eps) {
  if (K <= GV_CONSUMERS * 8)
    stack_linear<MB, 1, 4, PRO>(r, it, pb, deal, partial, red_ss, x, ldx, ln_w, res, ldr, out, ldo, M, N, K, eps);
  else
    stack_linear<MB, 3, 2, PRO>(r, it, pb, deal, partial, red_ss, x, ldx, ln_w, res, ldr, out, ldo, M, N, K, eps);
}

template <int MB>
__global__ void __launch_bounds__(GV_THREADS, 1) decode_stack_kernel(const StackParams p) {
  extern __shared__ uint8_t st_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[ST_MAX_STAGES], empty_bar[ST_MAX_STAGES];
  __shared__ float partial[2 * GV_CWARPS * 8];
  __shared__ float red_ss[GV_CWARPS * 2];
  __shared__ float s_m[MB * DEC_WARPS], s_l[MB * DEC_WARPS];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5;
  // attention merge buffer [MB][16][D] fp32: behind the ring
  float* s_acc = reinterpret_cast<float*>(ring + size_t(p.stages) * p.slot_bytes);
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], GV_CWARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  Ring r{ring, full_bar, empty_bar, p.stages, p.slot_bytes};
  if (warp == GV_CWARPS) {
    if ((tid & 31) == 0) stack_producer(p, r);
    return;
  }
  const int M = p.B * p.n, h = p.h, f = p.f, grid = gridDim.x;
  int it = 0, pb = 0;
  Deal deal{0};
  for (int l = 0; l < p.L; ++l) {
    // RMSNorm + QKV projection
    stack_linear_k<MB, GV_PRO_RMSNORM>(r, it, pb, deal, partial, red_ss, p.x, h, p.ln1[l], nullptr, 0, p.qkv, 3 * int64_t(h),
                                       M, 3 * h, h, p.eps);
    deal.next((3 * h + rows_per_slot(h) - 1) / rows_per_slot(h), grid);
    grid_barrier(p.bar, tid);
    // attention (RoPE of q and of the new keys on the fly)
    switch (p.D) {
      case 32: stack_attention<1, MB>(p, p.cache[l], s_m, s_l, s_acc, &s_last); break;
      case 64: stack_attention<2, MB>(p, p.cache[l], s_m, s_l, s_acc, &s_last); break;
      default: stack_attention<4, MB>(p, p.cache[l], s_m, s_l, s_acc, &s_last); break;
    }
    grid_barrier(p.bar, tid);
    // output projection + residual
    stack_linear_k<MB, GV_PRO_NONE>(r, it, pb, deal, partial, red_ss, p.ctx, h, nullptr, p.x, h, p.xmid, h, M, h, h, p.eps);
    deal.next((h + rows_per_slot(h) - 1) / rows_per_slot(h), grid);
    grid_barrier(p.bar, tid);
    // RMSNorm + gate | up projection
    stack_linear_k<MB, GV_PRO_RMSNORM>(r, it, pb, deal, partial, red_ss, p.xmid, h, p.ln2[l], nullptr, 0, p.gu, 2 * int64_t(f),
                                       M, 2 * f, h, p.eps);
    deal.next((2 * f + rows_per_slot(h) - 1) / rows_per_slot(h), grid);
    grid_barrier(p.bar, tid);
    // SwiGLU + down projection + residual -> the next layer's input, in place
    stack_linear_k<MB, GV_PRO_SWIGLU>(r, it, pb, deal, partial, red_ss, p.gu, 2 * int64_t(f), nullptr, p.xmid, h, p.x, h, M, h,
                                      f, p.eps);
    deal.next((h + rows_per_slot(f) - 1) / rows_per_slot(f), grid);
    grid_barrier(p.bar, tid);
  }
}

}  // namespace mla

using namespace mla;

static int g_stack_coop = -1;

extern "C" size_t mla_decode_stack_workspace(int32_t batch, int32_t n, int32_t prefix, int32_t heads, int32_t head_dim) {
  // split-K partials | arrival counters | grid barrier (the last two must be zero before the FIRST launch only)
  const size_t splits = size_t((prefix + n + ST_SPLIT_KEYS - 1) / ST_SPLIT_KEYS);
  const size_t ws = size_t(batch) * heads * n * splits * (head_dim + 2) * sizeof(float);
  return ((ws + 255) & ~size_t(255)) + ((size_t(batch) * heads * sizeof(int) + 255) & ~size_t(255)) + 256;
}

extern "C" int mla_decode_stack(const mla_decode_stack_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  if (a == nullptr) return set_error(MLA_ERR_ARG, "decode_stack: null args");
  if (a->layers <= 0 || a->batch <= 0 || a->n <= 0) return MLA_OK;
  const int M = a->batch * a->n;
  if (M > 2) return set_error(MLA_ERR_ARG, "decode_stack: batch*n = %d suffix rows, at most 2 are supported", M);
  const int h = a->heads * a->head_dim;
  if (a->head_dim != 32 && a->head_dim != 64 && a->head_dim != 128)
    return set_error(MLA_ERR_ARG, "decode_stack: head_dim %d not in {32, 64, 128}", a->head_dim);
  if ((h & 7) || (a->ffn & 7) || h > GV_CONSUMERS * 8 * 3 || a->ffn > GV_CONSUMERS * 8 * 3)
    return set_error(MLA_ERR_ARG, "decode_stack: hidden %d / ffn %d must be multiples of 8 and <= %d", h, a->ffn,
                     GV_CONSUMERS * 8 * 3);
  if (a->prefix < 0 || !a->workspace) return set_error(MLA_ERR_ARG, "decode_stack: negative prefix or no workspace");
  StackParams p;
  p.wqkv = (const __nv_bfloat16* const*)a->w_qkv; p.wo = (const __nv_bfloat16* const*)a->w_o;
  p.wgu = (const __nv_bfloat16* const*)a->w_gate_up; p.wd = (const __nv_bfloat16* const*)a->w_down;
  p.ln1 = (const __nv_bfloat16* const*)a->ln1; p.ln2 = (const __nv_bfloat16* const*)a->ln2;
  p.cache = (const __nv_bfloat16* const*)a->kv_cache;
  p.x = (__nv_bfloat16*)a->x; p.qkv = (__nv_bfloat16*)a->qkv; p.ctx = (__nv_bfloat16*)a->ctx;
  p.xmid = (__nv_bfloat16*)a->x_mid; p.gu = (__nv_bfloat16*)a->gate_up;
  p.cos = (const __nv_bfloat16*)a->cos_t; p.sin = (const __nv_bfloat16*)a->sin_t;
  p.L = a->layers; p.B = a->batch; p.n = a->n; p.P = a->prefix; p.H = a->heads; p.D = a->head_dim; p.h = h; p.f = a->ffn;
  p.eps = a->eps; p.scale = a->scale;
  const size_t splits = size_t((a->prefix + a->n + ST_SPLIT_KEYS - 1) / ST_SPLIT_KEYS);
  const size_t ws = (size_t(a->batch) * a->heads * a->n * splits * (a->head_dim + 2) * sizeof(float) + 255) & ~size_t(255);
  const size_t cnt = (size_t(a->batch) * a->heads * sizeof(int) + 255) & ~size_t(255);
  p.attn_ws = (float*)a->workspace;
  p.attn_cnt = (int*)((uint8_t*)a->workspace + ws);
  p.bar = (unsigned*)((uint8_t*)a->workspace + ws + cnt);
  auto pitch_of = [](int K) { return (size_t(K) * 2 + 127) & ~size_t(127); };
  auto rpi_of = [](int K) { return K <= GV_CONSUMERS * 8 ? 4 : 2; };
  size_t slot = pitch_of(h) * rpi_of(h);
  if (pitch_of(a->ffn) * rpi_of(a->ffn) > slot) slot = pitch_of(a->ffn) * rpi_of(a->ffn);
  const size_t attn_smem = size_t(M <= 1 ? 1 : 2) * DEC_WARPS * a->head_dim * sizeof(float);
  const size_t budget = 227 * 1024 - 4096 - attn_smem - 128;       // static shared memory + alignment slack
  int stages = int(budget / slot);
  if (stages < 2) return set_error(MLA_ERR_ARG, "decode_stack: hidden %d / ffn %d do not fit the shared-memory ring", h, a->ffn);
  p.stages = stages > ST_MAX_STAGES ? ST_MAX_STAGES : stages;
  p.slot_bytes = uint32_t(slot);
  const size_t smem = slot * p.stages + attn_smem + 128;
  if (g_stack_coop < 0) {
    const char* e = getenv("MLA_DECODE_STACK_COOP");
    g_stack_coop = (e && e[0] == '0') ? 0 : 1;
  }
  auto kern = M <= 1 ? decode_stack_kernel<1> : decode_stack_kernel<2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "cudaFuncSetAttribute(decode_stack smem %zu): %s", smem, cudaGetErrorString(e));
  // every CTA must be resident at once (grid barrier): one per SM, and the cooperative attribute makes the driver
  // refuse the launch instead of deadlocking if that ever does not hold
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms());
  cfg.blockDim = dim3(GV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_stack_coop ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "decode_stack launch: %s", cudaGetErrorString(e));
  MLA_CHECK_LAUNCH("decode_stack");
  return MLA_OK;
}
