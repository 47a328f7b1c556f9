// Inference denoise loop of the diffusion action head (MLA.predict_action_diff, models/mla/model_mla.py:592-775).
//
// The reference re-runs the whole 548-token sequence through the decoder at each of the 8 DDIM steps although only
// the [t | x_0..x_T] rows change between steps (under causal attention everything in front of them is identical).
// Here the prefix runs once through the training-path kernels (its post-RoPE K/V stay in the per-layer q|k|v buffers)
// and each DDIM step pushes only the few suffix rows through the 32 layers.  At 2..34 rows every linear is a pure
// weight-streaming problem (13.5 GB of bf16 weights per step), so the step is HBM-bound, not tensor-bound:
//   * gemv_bf16_kernel    — skinny GEMM: one warp per output column streams that weight row once with 16-byte loads,
//                           fp32 accumulators for up to 8 activation rows at a time (activations come from L1/L2)
//   * decode_attn_kernel  — the suffix queries against the cached K/V (bottom-right aligned causal mask, as
//                           flash-attn / modeling_llama.py:540-557), keys split over the warps of a CTA, online softmax
//   * ddim_step_kernel    — x_{t-1} of ddim_sample with eta = 0 (gaussian_diffusion.py:342-352,:522-571), in the
//                           reference's fp32 op order (no FMA contraction) so the update is bit-exact.
#include <cstdlib>

#include "decode_common.cuh"

namespace mla {


// ---------------------------------------------------------------------------------------------- skinny GEMM
// out[m, n] = bf16( bf16(sum_k x'[m,k] w[n,k]) + residual[m,n] ),  w [N,K] row-major (nn.Linear layout), m < M <= 16.
//
// Pure weight streaming, organised the way the DMA engine likes it: ONE producer thread per CTA issues a
// cp.async.bulk (TMA, no tensor map) per weight row — 8-22 KB contiguous — into a multi-stage shared-memory ring
// guarded by full/empty mbarriers; 16 consumer warps take the dot products from shared memory.  One persistent CTA per
// SM owns ~190 KB of ring, so ~150 KB per SM is always in flight.  A slot holds RPI consecutive weight rows; per slot
// the consumers do one warp-shuffle + one named-barrier reduction.
//   Launch overlap (programmatic dependent launch): the weights do not depend on the previous kernel — only the
//   activations do — so every gemv is launched with programmaticStreamSerialization: the producer starts streaming as
//   soon as the previous kernel lets dependents launch, and only the consumers execute griddepcontrol.wait before they
//   touch activations / residual / output.  (MLA_DECODE_PDL=0 turns it off.)
//   Measured on B200 (profiles/r01_denoise_T0_v*.json, ms per decode step = 12.95 GB of weights at Llama-2-7B size):
//   warp-per-row direct loads 6.74 -> this ring, 16 consumer warps / 192 KB 4.33 -> 9-shuffle multi-value reduction +
//   latency-tolerant attention 3.76 -> RoPE fused into the attention (5 launches per layer) 3.64 -> PDL 3.50.  A
//   narrower CTA (8 consumer warps, 100 KB ring, two CTAs of consecutive launches per SM) was slower (5.72).

struct GemvParams {
  const __nv_bfloat16 *x, *w, *res, *ln_w;
  __nv_bfloat16* out;
  int M, N, K;
  int64_t ldx, ldw, ldo, ldr;
  float eps;
  int stages;          // ring depth
  uint32_t pitch;      // bytes per weight row in the ring (K*2 rounded up to 128)
  int one_copy;        // the rows of a slot are contiguous in global memory and in the ring: one bulk copy per slot
};


// MB: activation rows held per thread (fast path: M <= MB, registers) — or 8 in the general path (GEN = 1).
// CPT: k-chunks (8 bf16) per consumer thread (fast path).  RPI: weight rows per ring slot.
template <int MB, int CPT, int RPI, int PRO, int GEN, int XF>
__global__ void __launch_bounds__(GV_THREADS, 1) gemv_ring_kernel(GemvParams p) {
  extern __shared__ uint8_t gv_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gv_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[GV_MAX_STAGES], empty_bar[GV_MAX_STAGES];
  __shared__ float partial[2][GV_CWARPS][RPI * MB];
  __shared__ float red_ss[GV_CWARPS][MB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();       // the next kernel may become resident and prefetch ITS weights while we run
  const int chunks = p.K >> 3;
  const uint32_t row_bytes = uint32_t(p.K) * 2u;
  const uint32_t slot_bytes = p.pitch * RPI;
  const int groups = (p.N + RPI - 1) / RPI;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], GV_CWARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == GV_CWARPS) {
    // ===================== producer: one thread streams this CTA's weight rows =====================
    if (lane == 0) {
      int it = 0;
      for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int n0 = g * RPI;
        const int valid = p.N - n0 < RPI ? p.N - n0 : RPI;
        mbar_arrive_expect_tx(&full_bar[s], row_bytes * valid);
        if (p.one_copy) {
          // the rows of a slot are contiguous in global memory (ldw = K) and, with an unpadded pitch, in the ring: ONE bulk
          // copy.  The producer is a single thread — every instruction it spends per slot is serial latency in front of
          // the memory pipe (profiles/r02_decode_stack_trace_*.json: issuing row by row cost ~5 % of the streaming rate)
          bulk_load_row(ring + size_t(s) * slot_bytes, p.w + int64_t(n0) * p.ldw, row_bytes * valid, &full_bar[s]);
        } else {
          for (int r = 0; r < valid; ++r)
            bulk_load_row(ring + size_t(s) * slot_bytes + size_t(r) * p.pitch, p.w + int64_t(n0 + r) * p.ldw, row_bytes,
                          &full_bar[s]);
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  pdl_wait();                    // activations / residual / output buffers belong to the kernels before us
  // activations: thread t always meets the same k-chunks, so they stay in registers — packed bf16 (XF = 0, unpacked
  // at every use) or unpacked once to fp32 (XF = 1)
  float xf[(GEN || !XF) ? 1 : MB][(GEN || !XF) ? 1 : CPT][8];
  uint4 xr[GEN ? 1 : MB][GEN ? 1 : CPT];
  if (!GEN) {
    float ss[MB];
#pragma unroll
    for (int r = 0; r < MB; ++r) ss[r] = 0.f;
#pragma unroll
    for (int r = 0; r < MB; ++r) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = tid + j * GV_CONSUMERS;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < p.M && c < chunks) {
          if (PRO == GV_PRO_SWIGLU) {
            float g[8], u[8], o[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + int64_t(r) * p.ldx + 8 * c)), g);
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + int64_t(r) * p.ldx + p.K + 8 * c)), u);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = bf16_round(g[e] * (1.f / (1.f + __expf(-g[e])))) * u[e];
            v = pack8f(o);
          } else {
            v = __ldg(reinterpret_cast<const uint4*>(p.x + int64_t(r) * p.ldx + 8 * c));
            if (PRO == GV_PRO_RMSNORM) {
              float f[8];
              unpack8(v, f);
#pragma unroll
              for (int e = 0; e < 8; ++e) ss[r] = fmaf(f[e], f[e], ss[r]);
            }
          }
        }
        xr[r][j] = v;
      }
    }
    if (PRO == GV_PRO_RMSNORM) {
#pragma unroll
      for (int r = 0; r < MB; ++r) {
        const float t = d_wsum(ss[r]);
        if (lane == 0) red_ss[warp][r] = t;
      }
      consumers_sync();
#pragma unroll
      for (int r = 0; r < MB; ++r) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < GV_CWARPS; ++w2) t += red_ss[w2][r];
        const float rstd = rsqrtf(t / float(p.K) + p.eps);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int c = tid + j * GV_CONSUMERS;
          if (c < chunks) {
            float f[8], g[8], o[8];
            unpack8(xr[r][j], f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.ln_w + 8 * c)), g);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g[e] * bf16_round(f[e] * rstd);
            xr[r][j] = pack8f(o);
          }
        }
      }
    }
    if (XF) {
#pragma unroll
      for (int r = 0; r < MB; ++r)
#pragma unroll
        for (int j = 0; j < CPT; ++j) unpack8(xr[r][j], xf[XF ? r : 0][XF ? j : 0]);
    }
  }

  int it = 0, pb = 0;
  for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
    const int s = it % p.stages;
    const uint32_t ph = (it / p.stages) & 1;
    const int n0 = g * RPI;
    const int passes = GEN ? (p.M + MB - 1) / MB : 1;
    // fast path: the residual element this thread adds in the slot's epilogue is requested BEFORE the wait for the
    // weights — otherwise its L2 round trip is paid once per slot by all 512 threads at the next named barrier
    float resv = 0.f;
    if (!GEN && p.res && tid < RPI * MB && n0 + tid / MB < p.N && tid % MB < p.M)
      resv = __bfloat162float(p.res[int64_t(tid % MB) * p.ldr + n0 + tid / MB]);
    mbar_wait(&full_bar[s], ph);
    const uint8_t* slot = ring + size_t(s) * slot_bytes;
    for (int pass = 0; pass < passes; ++pass) {
      const int m0 = pass * MB;
      float acc[RPI][MB];
#pragma unroll
      for (int rr = 0; rr < RPI; ++rr)
#pragma unroll
        for (int r = 0; r < MB; ++r) acc[rr][r] = 0.f;
      if (!GEN) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int c = tid + j * GV_CONSUMERS;
          if (c < chunks) {
#pragma unroll
            for (int rr = 0; rr < RPI; ++rr) {
              const uint4 wv = *reinterpret_cast<const uint4*>(slot + size_t(rr) * p.pitch + 16 * c);
              if (XF) {
                float wf[8];
                unpack8(wv, wf);
#pragma unroll
                for (int r = 0; r < MB; ++r)
#pragma unroll
                  for (int e = 0; e < 8; ++e) acc[rr][r] = fmaf(wf[e], xf[XF ? r : 0][XF ? j : 0][e], acc[rr][r]);
              } else {
#pragma unroll
                for (int r = 0; r < MB; ++r) acc[rr][r] = dot8(wv, xr[r][j], acc[rr][r]);
              }
            }
          }
        }
      } else {
        for (int c = tid; c < chunks; c += GV_CONSUMERS) {
          uint4 wv[RPI];
#pragma unroll
          for (int rr = 0; rr < RPI; ++rr) wv[rr] = *reinterpret_cast<const uint4*>(slot + size_t(rr) * p.pitch + 16 * c);
#pragma unroll
          for (int r = 0; r < MB; ++r) {
            if (m0 + r < p.M) {
              const uint4 xv = __ldg(reinterpret_cast<const uint4*>(p.x + int64_t(m0 + r) * p.ldx + 8 * c));
#pragma unroll
              for (int rr = 0; rr < RPI; ++rr) acc[rr][r] = dot8(wv[rr], xv, acc[rr][r]);
            }
          }
        }
      }
      if (pass == passes - 1) {           // this warp no longer reads the slot: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
      }
      {
        constexpr int V = RPI * MB;
        float flat[V];
#pragma unroll
        for (int rr = 0; rr < RPI; ++rr)
#pragma unroll
          for (int r = 0; r < MB; ++r) flat[rr * MB + r] = acc[rr][r];
        warp_reduce_many<V>(flat, lane);
        constexpr int LPV = 32 / V;                       // lanes that end up holding the same value
        if (lane % LPV == 0) partial[pb][warp][lane / LPV] = flat[0];
      }
      consumers_sync();
      if (tid < RPI * MB) {
        const int rr = tid / MB, r = tid % MB;
        const int n = n0 + rr, m = m0 + r;
        if (n < p.N && m < p.M) {
          float t = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < GV_CWARPS; ++w2) t += partial[pb][w2][tid];
          float v = bf16_round(t);
          if (p.res) v += GEN ? __bfloat162float(p.res[int64_t(m) * p.ldr + n]) : resv;
          p.out[int64_t(m) * p.ldo + n] = __float2bfloat16_rn(v);
        }
      }
      pb ^= 1;      // the next reduction writes the other buffer; this one is re-used only after another barrier
    }
  }
}


// ---------------------------------------------------------------------------------------------- decode attention
// q rows (b, i), i < Lq, are the LAST Lq positions of a length-Lk sequence whose K/V rows live at
// k/v + (b*Lk + j)*ldkv + h*D.  Query i sees keys j <= Lk - Lq + i.  One CTA per (i, h, b); warp w takes keys
// w, w+NW, ... (16 warps x 8 keys per pass cover 128 keys at once); partial (max, sum, acc[D]) are merged through
// shared memory.  D = 32 * EPL.

// ROPE = 0: q rows are already rotated and every key/value row (b, j), j < Lk, lives in the cache.
// ROPE = 1: q and the last Lq key rows come un-rotated from the packed projection `qkv` (row b*Lq + i: q | k | v at
//           column offsets 0 / hdim / 2*hdim) and are rotated on the fly with the table rows cs/sn [Lq, D/2] of
//           positions Lk-Lq..Lk-1; the cache holds the rotated prefix rows j < Lk - Lq (rows b*Lk + j).  The new K/V
//           are never written back: within the DDIM loop the next step recomputes them from the next x_t.
template <int EPL, int ROPE>
__global__ void __launch_bounds__(DEC_WARPS * 32) decode_attn_kernel(
    const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
    const __nv_bfloat16* __restrict__ v, int64_t kv_sb, int64_t kv_sh, int64_t kv_sj, __nv_bfloat16* __restrict__ o,
    int64_t ldo, int H, int Lq, int Lk, float scale, const __nv_bfloat16* __restrict__ cs,
    const __nv_bfloat16* __restrict__ sn, float* __restrict__ ws, int* __restrict__ counters, int S) {
  // Split-K: with S > 1 splits, CTA (i, sp) covers keys [sp*128, sp*128+128) — ONE pass of 16 warps x 8 keys, i.e. a
  // single DRAM round trip instead of ceil(keys/128) dependent ones — and leaves its un-normalised partial
  // (max, sum, acc[D]) in ws; the last CTA of a (b, h, i) to arrive (atomic counter, self-resetting) merges them.
  // cached key/value row (b, h, j) at k|v + b*kv_sb + h*kv_sh + j*kv_sj.  Token-major caches ([token][head][d]:
  // kv_sj = row pitch, kv_sh = D) make every head's stream a 256-byte gather at a 16 KB stride; the head-major layout
  // the denoise loop uses ([head][token][d]: kv_sj = D) lets the 16 warps of a CTA sweep 32 KB contiguous per pass.
  constexpr int D = 32 * EPL;
  pdl_launch_dependents();       // the output projection's gemv may start prefetching its weights
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS];
  __shared__ float s_acc[DEC_WARPS][D];
  const int i = blockIdx.x / S, sp = blockIdx.x % S, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = Lk - Lq;
  const int nall = P + i + 1;                              // keys visible to query i
  const int hdim = H * D;
  float qf[EPL], acc[EPL];
  cvt_raw<EPL>(ld_raw<EPL>(q + (int64_t(b) * Lq + i) * ldq + int64_t(h) * D + lane * EPL), qf);
  if (ROPE) rope_lanes<EPL>(qf, cs + int64_t(i) * (D / 2), sn + int64_t(i) * (D / 2), lane);
#pragma unroll
  for (int e = 0; e < EPL; ++e) { qf[e] *= scale; acc[e] = 0.f; }   // scale folded into q
  float m = -INFINITY, l = 0.f;
  const int64_t base = int64_t(b) * kv_sb + int64_t(h) * kv_sh + lane * EPL;
  constexpr int UN = 8;          // 16 independent loads per lane in flight: the kernel is DRAM-latency bound
  const int jbeg = S > 1 ? sp * (DEC_WARPS * UN) : 0;
  const int nkeys = S > 1 ? (nall < jbeg + DEC_WARPS * UN ? nall : jbeg + DEC_WARPS * UN) : nall;
  for (int j0 = jbeg + warp; j0 < nkeys; j0 += DEC_WARPS * UN) {
    RawEpl<EPL> kr[UN], vr[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * DEC_WARPS;
#pragma unroll
      for (int w2 = 0; w2 < (EPL + 1) / 2; ++w2) { kr[u].w[w2] = 0u; vr[u].w[w2] = 0u; }
      if (j < nkeys) {
        if (ROPE && j >= P) {
          const __nv_bfloat16* row = q + (int64_t(b) * Lq + (j - P)) * ldq + int64_t(h) * D + lane * EPL;
          kr[u] = ld_raw<EPL>(row + hdim);
          vr[u] = ld_raw<EPL>(row + 2 * hdim);
        } else {
          kr[u] = ld_raw<EPL>(k + base + int64_t(j) * kv_sj);
          vr[u] = ld_raw<EPL>(v + base + int64_t(j) * kv_sj);
        }
      }
    }
    float s[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * DEC_WARPS;
      float kf[EPL];
      cvt_raw<EPL>(kr[u], kf);
      if (ROPE && j >= P && j < nkeys)       // warp-uniform: j depends on the warp index only
        rope_lanes<EPL>(kf, cs + int64_t(j - P) * (D / 2), sn + int64_t(j - P) * (D / 2), lane);
      s[u] = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) s[u] = fmaf(qf[e], kf[e], s[u]);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) s[u] = d_wsum(s[u]);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (j0 + u * DEC_WARPS < nkeys) {
        float vf[EPL];
        cvt_raw<EPL>(vr[u], vf);
        const float mn = fmaxf(m, s[u]);
        const float corr = __expf(m - mn), pj = __expf(s[u] - mn);
        l = l * corr + pj;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = acc[e] * corr + pj * vf[e];
        m = mn;
      }
    }
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[warp][lane * EPL + e] = acc[e];
  __syncthreads();
  __nv_bfloat16* orow = o + (int64_t(b) * Lq + i) * ldo + int64_t(h) * D;
  const int64_t unit = (int64_t(b) * H + h) * Lq + i;
  float* part = ws ? ws + (unit * S + sp) * (D + 2) : nullptr;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float mx = -INFINITY;
#pragma unroll
    for (int w2 = 0; w2 < DEC_WARPS; ++w2) mx = fmaxf(mx, s_m[w2]);
    float L = 0.f, a = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < DEC_WARPS; ++w2) {
      const float c = s_m[w2] == -INFINITY ? 0.f : __expf(s_m[w2] - mx);
      L += s_l[w2] * c;
      a += s_acc[w2][d] * c;
    }
    if (S == 1) {
      orow[d] = __float2bfloat16_rn(L > 0.f ? a / L : 0.f);
    } else {
      part[2 + d] = a;
      if (d == 0) { part[0] = mx; part[1] = L; }
    }
  }
  if (S == 1) return;
  __shared__ int s_last;
  __threadfence();                       // this CTA's partial is visible before its arrival is counted
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counters + unit, 1) == S - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* all = ws + unit * S * (D + 2);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float mx = -INFINITY;
    for (int s2 = 0; s2 < S; ++s2) mx = fmaxf(mx, __ldcg(all + s2 * (D + 2)));
    float L = 0.f, a = 0.f;
    for (int s2 = 0; s2 < S; ++s2) {
      const float ms = __ldcg(all + s2 * (D + 2));
      const float c = ms == -INFINITY ? 0.f : __expf(ms - mx);
      L += __ldcg(all + s2 * (D + 2) + 1) * c;
      a += __ldcg(all + s2 * (D + 2) + 2 + d) * c;
    }
    orow[d] = __float2bfloat16_rn(L > 0.f ? a / L : 0.f);
  }
  if (threadIdx.x == 0) counters[unit] = 0;          // re-armed for the next launch
}

// ---------------------------------------------------------------------------------------------- RoPE + cache append
// The n new rows per sample of a packed q|k|v projection [B*n, 3h] at positions P..P+n-1: q is rotated in place (the
// attention kernel reads it there), k is rotated into the cache row (b*Lk + P + i) columns [0,h), v is copied to
// columns [h,2h).  Same arithmetic and rounding points as rope_kernel (norm_rope_act.cu; modeling_llama.py:184-208).
// cos/sin bf16 [n, d/2] = the table rows of positions P..P+n-1.  One thread per (row, head, 8-element chunk of d/2).
__global__ void rope_cache_kernel(__nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ cache,
                                  const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                                  int B, int n, int P, int heads, int d) {
  const int half = d >> 1, cpr = half >> 3, h = heads * d;
  const int64_t total = int64_t(B) * n * heads * cpr;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % cpr);
    const int hd = int((idx / cpr) % heads);
    const int64_t t = idx / (int64_t(cpr) * heads);         // row of qkv: b*n + i
    const int i = int(t % n), b = int(t / n);
    float cs[8], sn[8];
    unpack8(*reinterpret_cast<const uint4*>(cos_t + int64_t(i) * half + c * 8), cs);
    unpack8(*reinterpret_cast<const uint4*>(sin_t + int64_t(i) * half + c * 8), sn);
    __nv_bfloat16* row = qkv + t * 3 * int64_t(h);
    __nv_bfloat16* crow = cache + (int64_t(b) * (P + n) + P + i) * 2 * int64_t(h);
#pragma unroll
    for (int part = 0; part < 2; ++part) {                 // 0: q (in place), 1: k (into the cache)
      const __nv_bfloat16* src = row + part * h + int64_t(hd) * d + c * 8;
      __nv_bfloat16* dst = part == 0 ? row + int64_t(hd) * d + c * 8 : crow + int64_t(hd) * d + c * 8;
      float x1[8], x2[8], o1[8], o2[8];
      unpack8(*reinterpret_cast<const uint4*>(src), x1);
      unpack8(*reinterpret_cast<const uint4*>(src + half), x2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o1[j] = bf16_round(x1[j] * cs[j]) + bf16_round(-x2[j] * sn[j]);
        o2[j] = bf16_round(x2[j] * cs[j]) + bf16_round(x1[j] * sn[j]);
      }
      *reinterpret_cast<uint4*>(dst) = pack8f(o1);
      *reinterpret_cast<uint4*>(dst + half) = pack8f(o2);
    }
    const __nv_bfloat16* vs = row + 2 * h + int64_t(hd) * d + c * 8;
    __nv_bfloat16* vd = crow + h + int64_t(hd) * d + c * 8;
    *reinterpret_cast<uint4*>(vd) = *reinterpret_cast<const uint4*>(vs);
    *reinterpret_cast<uint4*>(vd + half) = *reinterpret_cast<const uint4*>(vs + half);
  }
}

// ---------------------------------------------------------------------------------------------- DDIM update
// coef f32 [4] = sqrt(1/ac_t), sqrt(1/ac_t - 1), sqrt(ac_prev), sqrt(1 - ac_prev) of the current respaced step.
//   pred_xstart = c0*x - c1*eps ; eps' = (c0*x - pred_xstart)/c1 ; out = pred_xstart*c2 + c3*eps'
// (the reference re-derives eps from pred_xstart, gaussian_diffusion.py:549; every op rounded on its own like ATen's).
template <typename EpsT>
__global__ void ddim_step_kernel(const float* __restrict__ x, const EpsT* __restrict__ eps,
                                 const float* __restrict__ coef, float* __restrict__ out, int64_t n) {
  const float c0 = coef[0], c1 = coef[1], c2 = coef[2], c3 = coef[3];
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    float e;
    if constexpr (sizeof(EpsT) == 4) e = eps[i]; else e = __bfloat162float(eps[i]);
    const float a = __fmul_rn(c0, x[i]);
    const float px = __fsub_rn(a, __fmul_rn(c1, e));
    const float e2 = __fdiv_rn(__fsub_rn(a, px), c1);
    out[i] = __fadd_rn(__fmul_rn(px, c2), __fmul_rn(c3, e2));
  }
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))
int gemv2_launch(const mla_gemv_args* a, void* stream);      // decode_stack.cu

static int g_gemv_pdl = -1;
static bool gemv_pdl_enabled() {
  if (g_gemv_pdl < 0) {
    const char* e = getenv("MLA_DECODE_PDL");
    g_gemv_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_gemv_pdl == 1;
}
extern "C" int mla_decode_set_pdl(int32_t on) {
  g_gemv_pdl = on ? 1 : 0;
  return MLA_OK;
}

template <int MB, int CPT, int RPI, int PRO, int GEN, int XF>
static int launch_gemv_x(const GemvParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = gemv_ring_kernel<MB, CPT, RPI, PRO, GEN, XF>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GV_RING_BYTES + 128);
    if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "cudaFuncSetAttribute(gemv smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = gemv_pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "gemv launch: %s", cudaGetErrorString(e));
  return MLA_OK;
}

static int g_gemv_xf32 = -1;
template <int MB, int CPT, int RPI, int PRO, int GEN>
static int launch_gemv(const GemvParams& p, int grid, size_t smem, cudaStream_t stream) {
  if (g_gemv_xf32 < 0) {
    const char* e = getenv("MLA_GEMV_XF32");
    g_gemv_xf32 = (e && e[0] == '0') ? 0 : 1;       // fp32 activations in registers by default
  }
  if (!GEN && g_gemv_xf32) return launch_gemv_x<MB, CPT, RPI, PRO, GEN, 1>(p, grid, smem, stream);
  return launch_gemv_x<MB, CPT, RPI, PRO, GEN, 0>(p, grid, smem, stream);
}

template <int PRO>
static int dispatch_gemv_fast(const GemvParams& p, int mb, bool small_k, int grid, size_t smem, cudaStream_t st) {
  // small_k: K <= 4096 -> one k-chunk per thread, 4 weight rows per slot; else up to 3 chunks per thread, 2 rows per slot
  if (small_k) {
    if (mb == 1) return launch_gemv<1, 1, 4, PRO, 0>(p, grid, smem, st);
    if (mb == 2) return launch_gemv<2, 1, 4, PRO, 0>(p, grid, smem, st);
    return launch_gemv<4, 1, 4, PRO, 0>(p, grid, smem, st);
  }
  if (mb == 1) return launch_gemv<1, 3, 2, PRO, 0>(p, grid, smem, st);
  return launch_gemv<2, 3, 2, PRO, 0>(p, grid, smem, st);
}

extern "C" int mla_gemv_fused(const mla_gemv_args* a, void* stream) {
  if (int rc = device_check()) return rc;
  if (a == nullptr) return set_error(MLA_ERR_ARG, "gemv: null args");
  if (a->m <= 0 || a->n <= 0) return MLA_OK;
  const int M = a->m, N = a->n, K = a->k;
  if (K <= 0 || (K & 7) || (a->ldx & 7) || (a->ldw & 7))
    return set_error(MLA_ERR_ARG, "gemv: k and the row pitches of x and w must be multiples of 8 (k=%d)", K);
  if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w)) & 15)
    return set_error(MLA_ERR_ARG, "gemv: x and w must be 16-byte aligned");
  if (M > 16) return set_error(MLA_ERR_ARG, "gemv: m=%d rows is a GEMM, use mla_gemm_bf16", M);
  if (a->prologue < GV_PRO_NONE || a->prologue > GV_PRO_SWIGLU) return set_error(MLA_ERR_ARG, "gemv: unknown prologue");
  if (a->prologue == GV_PRO_RMSNORM && (!a->ln_weight || (reinterpret_cast<uintptr_t>(a->ln_weight) & 15)))
    return set_error(MLA_ERR_ARG, "gemv: the RMSNorm prologue needs a 16-byte aligned weight vector");
  {
    // m <= 2: consumers without a per-slot CTA barrier + FFMA2 + finisher warp (decode_stack.cu: gemv2_kernel), MLA_GEMV2=1
    static int v2 = -1;
    if (v2 < 0) {
      const char* e = getenv("MLA_GEMV2");
      v2 = (e && e[0] == '1') ? 1 : 0;
    }
    if (v2 && M <= 2) {
      const int rc = gemv2_launch(a, stream);
      if (rc <= 0) return rc;
    }
  }
  const bool small_k = K <= GV_CONSUMERS * 8;                           // 4096
  const bool fast = (M <= 4 && small_k) || (M <= 2 && K <= GV_CONSUMERS * 8 * 3);   // activations fit the registers
  if (a->prologue != GV_PRO_NONE && !fast)
    return set_error(MLA_ERR_ARG, "gemv: fused prologues need m <= 4 with k <= %d, or m <= 2 with k <= %d",
                     GV_CONSUMERS * 8, GV_CONSUMERS * 8 * 3);
  GemvParams p;
  p.x = (const __nv_bfloat16*)a->x; p.w = (const __nv_bfloat16*)a->w; p.res = (const __nv_bfloat16*)a->residual;
  p.ln_w = (const __nv_bfloat16*)a->ln_weight; p.out = (__nv_bfloat16*)a->out;
  p.M = M; p.N = N; p.K = K; p.ldx = a->ldx; p.ldw = a->ldw; p.ldo = a->ldo; p.ldr = a->ldr; p.eps = a->eps;
  p.pitch = (uint32_t(K) * 2u + 127u) & ~127u;
  {
    static int one = -1;
    if (one < 0) {
      const char* e = getenv("MLA_GEMV_ONE_COPY");
      one = (e && e[0] == '0') ? 0 : 1;
    }
    p.one_copy = one && a->ldw == K && p.pitch == uint32_t(K) * 2u;
  }
  const int rpi = small_k ? 4 : 2;
  const size_t slot = size_t(p.pitch) * rpi;
  int stages = int(GV_RING_BYTES / slot);
  if (stages < 2) return set_error(MLA_ERR_ARG, "gemv: k=%d does not fit the shared-memory ring, use mla_gemm_bf16", K);
  p.stages = stages > GV_MAX_STAGES ? GV_MAX_STAGES : stages;
  const size_t smem = slot * p.stages + 128;
  const int groups = (N + rpi - 1) / rpi;
  const int grid = groups < num_sms() ? groups : num_sms();
  cudaStream_t st = S_(stream);
  int rc;
  if (fast) {
    const int mb = M <= 1 ? 1 : (M <= 2 ? 2 : 4);
    if (a->prologue == GV_PRO_RMSNORM) rc = dispatch_gemv_fast<GV_PRO_RMSNORM>(p, mb, small_k, grid, smem, st);
    else if (a->prologue == GV_PRO_SWIGLU) rc = dispatch_gemv_fast<GV_PRO_SWIGLU>(p, mb, small_k, grid, smem, st);
    else rc = dispatch_gemv_fast<GV_PRO_NONE>(p, mb, small_k, grid, smem, st);
  } else {
    rc = small_k ? launch_gemv<8, 1, 4, GV_PRO_NONE, 1>(p, grid, smem, st)
                 : launch_gemv<8, 1, 2, GV_PRO_NONE, 1>(p, grid, smem, st);
  }
  if (rc) return rc;
  MLA_CHECK_LAUNCH("gemv_bf16");
  return MLA_OK;
}

extern "C" int mla_gemv_bf16(const void* x, const void* w, void* out, const void* residual, int32_t m, int32_t n,
                             int32_t k, int64_t ldx, int64_t ldw, int64_t ldo, int64_t ldr, void* stream) {
  mla_gemv_args a;
  a.x = x; a.w = w; a.out = out; a.residual = residual; a.ln_weight = nullptr;
  a.m = m; a.n = n; a.k = k; a.ldx = ldx; a.ldw = ldw; a.ldo = ldo; a.ldr = ldr;
  a.prologue = GV_PRO_NONE; a.eps = 0.f;
  return mla_gemv_fused(&a, stream);
}

static int decode_attn_launch(const void* q, int64_t ldq, const void* k, const void* v, int64_t kv_sb, int64_t kv_sh,
                              int64_t kv_sj, void* o, int64_t ldo, int32_t batch, int32_t heads, int32_t len_q,
                              int32_t len_k, int32_t head_dim, float scale, const void* cs, const void* sn,
                              void* ws, void* counters, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || heads <= 0 || len_q <= 0) return MLA_OK;
  if (len_k < len_q) return set_error(MLA_ERR_ARG, "decode_attn: len_k=%d < len_q=%d", len_k, len_q);
  if (((kv_sb | kv_sh | kv_sj) & 7) || ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15))
    return set_error(MLA_ERR_ARG, "decode_attn: K/V must be 16-byte aligned with strides that are multiples of 8");
  if ((ldq & 7) || (reinterpret_cast<uintptr_t>(q) & 15))
    return set_error(MLA_ERR_ARG, "decode_attn: q must be 16-byte aligned with a row pitch that is a multiple of 8");
  const int splits = (ws && counters) ? (len_k + DEC_WARPS * 8 - 1) / (DEC_WARPS * 8) : 1;
  dim3 grid(len_q * splits, heads, batch);
  const bool rope = cs != nullptr;
  float* wsp = splits > 1 ? (float*)ws : nullptr;
  int* cnt = splits > 1 ? (int*)counters : nullptr;
#define MLA_DEC(EPL)                                                                                             \
  do {                                                                                                           \
    if (rope)                                                                                                    \
      decode_attn_kernel<EPL, 1><<<grid, DEC_WARPS * 32, 0, S_(stream)>>>(                                       \
          (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, kv_sb, kv_sh, kv_sj,    \
          (__nv_bfloat16*)o, ldo, heads, len_q, len_k, scale, (const __nv_bfloat16*)cs, (const __nv_bfloat16*)sn, \
          wsp, cnt, splits);                                                                                     \
    else                                                                                                         \
      decode_attn_kernel<EPL, 0><<<grid, DEC_WARPS * 32, 0, S_(stream)>>>(                                       \
          (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, kv_sb, kv_sh, kv_sj,    \
          (__nv_bfloat16*)o, ldo, heads, len_q, len_k, scale, nullptr, nullptr, wsp, cnt, splits);               \
  } while (0)
  switch (head_dim) {
    case 32: MLA_DEC(1); break;
    case 64: MLA_DEC(2); break;
    case 128: MLA_DEC(4); break;
    case 256: MLA_DEC(8); break;
    default: return set_error(MLA_ERR_ARG, "decode_attn: head_dim %d not in {32, 64, 128, 256}", head_dim);
  }
#undef MLA_DEC
  MLA_CHECK_LAUNCH("decode_attn");
  return MLA_OK;
}

extern "C" int mla_decode_attn(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* o,
                               int64_t ldo, int32_t batch, int32_t heads, int32_t len_q, int32_t len_k,
                               int32_t head_dim, float scale, void* stream) {
  return decode_attn_launch(q, ldq, k, v, int64_t(len_k) * ldkv, head_dim, ldkv, o, ldo, batch, heads, len_q, len_k,
                            head_dim, scale, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int mla_decode_attn_rope(const void* qkv, int64_t ldqkv, const void* k_cache, const void* v_cache,
                                    int64_t kv_stride_b, int64_t kv_stride_h, int64_t kv_stride_j, const void* cos_t,
                                    const void* sin_t, void* o, int64_t ldo, int32_t batch, int32_t heads,
                                    int32_t len_q, int32_t len_k, int32_t head_dim, float scale, void* workspace,
                                    void* counters, void* stream) {
  if (!cos_t || !sin_t) return set_error(MLA_ERR_ARG, "decode_attn_rope: RoPE tables are required");
  return decode_attn_launch(qkv, ldqkv, k_cache, v_cache, kv_stride_b, kv_stride_h, kv_stride_j, o, ldo, batch, heads,
                            len_q, len_k, head_dim, scale, cos_t, sin_t, workspace, counters, stream);
}

extern "C" size_t mla_decode_attn_workspace(int32_t batch, int32_t heads, int32_t len_q, int32_t len_k,
                                            int32_t head_dim) {
  const size_t splits = size_t((len_k + DEC_WARPS * 8 - 1) / (DEC_WARPS * 8));
  return size_t(batch) * heads * len_q * splits * (head_dim + 2) * sizeof(float);
}

extern "C" int mla_ddim_step(const void* x, const void* eps, int32_t eps_is_f32, const void* coef, void* out, int64_t n,
                             void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  const int grid = int((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  if (eps_is_f32)
    ddim_step_kernel<float><<<grid, 256, 0, S_(stream)>>>((const float*)x, (const float*)eps, (const float*)coef,
                                                          (float*)out, n);
  else
    ddim_step_kernel<__nv_bfloat16><<<grid, 256, 0, S_(stream)>>>((const float*)x, (const __nv_bfloat16*)eps,
                                                                  (const float*)coef, (float*)out, n);
  MLA_CHECK_LAUNCH("ddim_step");
  return MLA_OK;
}

extern "C" int mla_rope_cache(void* qkv, void* cache, const void* cos_t, const void* sin_t, int32_t batch, int32_t n,
                              int32_t prefix, int32_t heads, int32_t head_dim, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0 || n <= 0) return MLA_OK;
  if (heads <= 0 || head_dim <= 0 || (head_dim & 15)) return set_error(MLA_ERR_ARG, "rope_cache: head_dim must be a multiple of 16");
  if (prefix < 0) return set_error(MLA_ERR_ARG, "rope_cache: negative prefix length");
  const int64_t total = int64_t(batch) * n * heads * (head_dim / 16);
  const int grid = int((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  rope_cache_kernel<<<grid, 256, 0, S_(stream)>>>((__nv_bfloat16*)qkv, (__nv_bfloat16*)cache, (const __nv_bfloat16*)cos_t,
                                                  (const __nv_bfloat16*)sin_t, batch, n, prefix, heads, head_dim);
  MLA_CHECK_LAUNCH("rope_cache");
  return MLA_OK;
}
