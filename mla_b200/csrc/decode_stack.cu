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
//     RMSNorm+gate|up, SwiGLU+down+residual — separated by a grid barrier (arrival counter in global memory: one
//     L2 round trip, hidden behind the ring's ~3 us of buffered weights);
//   * attention: one CTA per (sample, head) serves all query rows from one pass over the keys, the prefix K/V of the
//     layer having been pulled into L2 by a bulk prefetch the producer issues one phase earlier; the other CTAs' rings
//     fill with the output projection's weights meanwhile.
// The arithmetic of every phase is the per-op kernels' (same thread -> k-chunk mapping, same reduction trees, same
// rounding points; attention = decode_attn_kernel<.,ROPE=1>), so the result is BIT-IDENTICAL to
// LlamaDecoderLayer.decode — that is what tests/test_decode_stack_gpu.py asserts.
// Activations written by other CTAs inside this launch are read with ld.global.cg (L2): L1 is not coherent.
// Limits: B*n <= 2 suffix rows (activations live in registers), h <= 12288, f <= 12288, head_dim in {32, 64, 128}.
#include <cstdlib>
#include <type_traits>

#include "decode_common.cuh"

namespace mla {

constexpr int ST_MAX_STAGES = 8;
constexpr int ST_THREADS = GV_CONSUMERS + 64;        // 16 consumer warps | producer warp | finisher warp
constexpr int ST_ATTN_UN = 8;                       // keys per warp in flight: 16 warps x 8 = 128 keys per pass

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
  unsigned* bar;                        // [3] grid barrier: two alternating arrival counters + selector (zeroed once)
  long long* trace;                     // null, or [grid][L][5 phases][3] globaltimer ns: phase entry, work done, barrier passed
  int L, B, n, P, H, D, h, f;
  float eps, scale;
  int stages;
  uint32_t slot_bytes;
  int rpi_big;    // weight rows per ring slot where K > 4096 (1 or 2)
  int ahead;      // weight groups the L2 prefetch cursor runs in front of the ring (0 = no run-ahead)
  int dbg;        // profiling only (mla_decode_stack_set_debug): 1 = consumers skip the math, 2 = no grid barriers, 4 = no attention
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
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Grid barrier for the consumer threads of all CTAs: a monotonic arrival counter — one fire-and-forget red.add per CTA,
// then everybody polls until the count reaches (k+1) * grid for the k-th barrier of the launch: one L2 round trip after
// the last arrival.  Two counters alternate between launches (bar[2] = which one this launch uses): CTA 0 clears the
// idle one at its start and flips the selector after the first barrier, when every CTA is known to have read it — so
// nothing has to be reset from the host.
struct GridBar {
  unsigned* ctr;
  unsigned target;
};
// consumers + the finisher warp (whose output stores must be ordered before the arrival)
__device__ __forceinline__ void cta_workers_sync() { asm volatile("bar.sync 2, %0;" ::"n"(GV_CONSUMERS + 32) : "memory"); }
__device__ __forceinline__ void grid_barrier(GridBar& gb, int tid, int dbg = 0) {
  if (dbg & 2) { cta_workers_sync(); return; }
  cta_workers_sync();
  if (tid == 0) {
    gb.target += gridDim.x;
    __threadfence();                                  // this CTA's writes (seen through the bar.sync) before the arrival
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
    while (ld_acquire_u32(gb.ctr) < gb.target) {
    }
  }
  cta_workers_sync();
}

// Fixed-size ring slots (the largest group of any phase: 2 rows x 22 KB for the down projection at 7B width; the
// K = 4096 phases use 32 KB of each).  A byte-granular ring that packed six 32 KB slots into the same space was tried
// and streamed no faster: what bounds the rate is not the bytes in flight but the work per slot on both sides — above
// all the PRODUCER thread's own instruction path (one thread, ~1000 cycles per slot with the generic cursor logic),
// which is why the per-slot bookkeeping below is a handful of adds.
struct Ring {
  uint8_t* base;
  uint64_t *full_bar, *empty_bar;
  int stages;
  uint32_t slot_bytes;
};
struct RingPos {        // a role's position in the slot sequence: stage index and the parity of its current round
  int s;
  uint32_t par;
  __device__ __forceinline__ void next(int stages) {
    if (++s == stages) { s = 0; par ^= 1u; }
  }
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
__device__ __forceinline__ int rows_per_slot(int K, int rpi_big) { return K <= GV_CONSUMERS * 8 ? 4 : rpi_big; }

// group g of a phase goes to CTA (off + g) % grid, where off continues the round-robin of the previous phase: the CTAs
// that get one group more than the others are different ones in every phase
struct Deal {
  int off;
  __device__ __forceinline__ int first(int cta, int grid) const { int d = cta - off; return d < 0 ? d + grid : d; }
  __device__ __forceinline__ void next(int groups, int grid) { off = (off + groups) % grid; }
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait_a(bar, parity)) {
  }
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}


// ---------------------------------------------------------------------------------------------- producer
// One thread.  Everything it does per slot is serial latency in front of the memory pipe, so the inner loop is: wait for
// the stage to be free, arm its barrier, ONE bulk copy (the rows of a group are contiguous in global memory and, with an
// unpadded row pitch, in the ring too), a few adds.
__device__ void stack_producer(const StackParams& p, const Ring& r) {
  const int grid = gridDim.x, cta = blockIdx.x;
  RingPos pos{0, 0u};
  Deal deal{0};
  long long wait_cycles = 0, slots = 0;
  const long long t_begin = p.trace ? clock64() : 0;
  for (int l = 0; l < p.L; ++l) {
    for (int ph = 0; ph < 4; ++ph) {
      const PhaseW w = phase_weights(p, l, ph);
      const int rpi = rows_per_slot(w.K, p.rpi_big);
      const uint32_t row_bytes = uint32_t(w.K) * 2u, pitch = (row_bytes + 127u) & ~127u;
      const int groups = (w.N + rpi - 1) / rpi;
      const uint32_t group_bytes = row_bytes * uint32_t(rpi);
      const int g0 = deal.first(cta, grid);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(w.w) + size_t(g0) * group_bytes;
      const size_t stride = size_t(grid) * group_bytes;
      const int full_groups = w.N / rpi;                   // groups with all rpi rows
      for (int g = g0; g < groups; g += grid, src += stride) {
        const uint32_t bar_e = smem_u32(&r.empty_bar[pos.s]), bar_f = smem_u32(&r.full_bar[pos.s]);
        if (p.trace) {
          const long long t0 = clock64();
          mbar_wait_a(bar_e, pos.par ^ 1u);
          wait_cycles += clock64() - t0;
          ++slots;
        } else {
          mbar_wait_a(bar_e, pos.par ^ 1u);
        }
        uint8_t* dst = r.base + size_t(pos.s) * r.slot_bytes;
        const uint32_t bytes = g < full_groups ? group_bytes : row_bytes * uint32_t(w.N - g * rpi);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_f), "r"(bytes) : "memory");
        if (p.dbg & 32) {               // profiling: no memory traffic at all — how fast are the consumers on their own?
          asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar_f), "r"(bytes) : "memory");
        } else if (pitch == row_bytes) {
          bulk_load_row(dst, src, bytes, &r.full_bar[pos.s]);
        } else {
          for (uint32_t o = 0, q = 0; o < bytes; o += row_bytes, q += pitch) bulk_load_row(dst + q, src + o, row_bytes, &r.full_bar[pos.s]);
        }
        pos.next(r.stages);
      }
      deal.next(groups, grid);
      if (ph == 0) {
        // the QKV weights of this layer are all requested: pull the layer's prefix K/V into L2, so the attention phase
        // that follows pays L2, not DRAM, latency.  Each CTA asks for its 1/grid of the cache in 16 KB pieces.
        const size_t total = size_t(p.B) * 2 * p.H * p.P * p.D * 2;
        const size_t per = ((total + grid - 1) / grid + 15) & ~size_t(15);
        size_t beg = per * cta;
        const size_t end = beg + per < total ? beg + per : (total & ~size_t(15));
        const uint8_t* c = reinterpret_cast<const uint8_t*>(p.cache[l]);
        for (; beg < end; beg += 16384) prefetch_l2_bulk(c + beg, uint32_t(end - beg < 16384 ? end - beg : 16384));
      }
    }
  }
  if (p.trace) {          // behind the phase stamps: [grid][4] = producer wait / total cycles, slots, (consumer wait, written there)
    long long* st = p.trace + size_t(gridDim.x) * p.L * 15 + size_t(blockIdx.x) * 4;
    st[0] = wait_cycles;
    st[1] = clock64() - t_begin;
    st[2] = slots;
  }
}

// ---------------------------------------------------------------------------------------------- linear phase (consumers)
// out[m, n] = bf16( bf16(sum_k x'[m,k] w[n,k]) + res[m,n] ): the arithmetic of gemv_ring_kernel's fast path (XF = 1;
// same thread -> k-chunk mapping, same reduction trees), with coherent activation loads and the ring / deal state
// carried across phases.
//
// The 16 consumer warps are NOT synchronised per slot.  Each one takes its k-slice of the slot's rows, reduces its V
// partial sums inside the warp, leaves them in shared memory and arrives on the slot's `part` mbarrier (arrive /
// try_wait are release / acquire at CTA scope — no fence, no atomic counter).  A dedicated FINISHER warp waits for that
// barrier, adds the 16 partials in warp order (so the result does not depend on timing), rounds, adds the residual,
// stores, and hands the slot back to the producer.  Consumer warps drift apart by up to a ring's worth of slots, so one
// slot's reduction latency overlaps the next slots' dot products.
// Partials are indexed by slot mod 16 = 2 * ST_MAX_STAGES: the ring never holds more than ST_MAX_STAGES slots between
// the finisher and the fastest consumer.
struct LinArgs {          // by value: each instantiation is a real function with its own register allocation
  uint32_t ring_base;     // shared-space byte address of the ring
  uint32_t slot_bytes;
  int stages;
  uint32_t full_bar, empty_bar, part_bar;   // shared-space addresses of the barrier arrays
  uint32_t partial;       // shared-space address: float [16 slots][16 warps][8]
  float* red_ss;
  const __nv_bfloat16 *x, *ln_w, *res;
  __nv_bfloat16* out;
  int64_t ldx, ldr, ldo;
  int M, N, K;
  float eps;
  int dbg;
  int deal_off;
  RingPos pos;
  int it;
  long long* wait_acc;    // profiling: cycles thread 0 spent waiting for weights
};
// two fp32 FMAs per instruction (FFMA2, sm_100): the dot products are issue-bound on the CUDA cores
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long bf16x2_to_f32x2(uint32_t v) {      // (lo, hi) bf16 -> (lo, hi) fp32
  unsigned long long r;
  asm("{\n\t.reg .b32 l, h;\n\tshl.b32 l, %1, 16;\n\tand.b32 h, %1, 0xffff0000;\n\tmov.b64 %0, {l, h};\n\t}" : "=l"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ float sum_f32x2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

// MB activation rows (registers), CPT k-chunks (8 bf16) per thread, RPI weight rows per slot.
struct LinRet {
  RingPos pos;
  int it;
};
template <int MB, int CPT, int RPI, int PRO>
__device__ __noinline__ LinRet stack_linear(const LinArgs a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.M, N = a.N, K = a.K;
  const int chunks = K >> 3;
  const uint32_t pitch = (uint32_t(K) * 2u + 127u) & ~127u;
  const int groups = (N + RPI - 1) / RPI;
  // activations of this thread's k-chunks as fp32 pairs (element 2i, 2i+1): thread t always meets the same chunks
  unsigned long long xp[MB][CPT][4];
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
            unpack8(ld_cg16(a.x + int64_t(m) * a.ldx + 8 * c), g);
            unpack8(ld_cg16(a.x + int64_t(m) * a.ldx + K + 8 * c), u);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = bf16_round(g[e] * (1.f / (1.f + __expf(-g[e])))) * u[e];
            v = pack8f(o);
          } else {
            v = ld_cg16(a.x + int64_t(m) * a.ldx + 8 * c);
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
        if (lane == 0) a.red_ss[warp * 2 + m] = t;
      }
      consumers_sync();
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < GV_CWARPS; ++w2) t += a.red_ss[w2 * 2 + m];
        const float rstd = rsqrtf(t / float(K) + a.eps);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int c = tid + j * GV_CONSUMERS;
          if (c < chunks) {
            float f[8], g[8], o[8];
            unpack8(xr[m][j], f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(a.ln_w + 8 * c)), g);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g[e] * bf16_round(f[e] * rstd);
            xr[m][j] = pack8f(o);
          }
        }
      }
      consumers_sync();                 // red_ss is re-used by the next RMSNorm phase
    }
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        xp[m][j][0] = bf16x2_to_f32x2(xr[m][j].x);
        xp[m][j][1] = bf16x2_to_f32x2(xr[m][j].y);
        xp[m][j][2] = bf16x2_to_f32x2(xr[m][j].z);
        xp[m][j][3] = bf16x2_to_f32x2(xr[m][j].w);
        // opaque: keep the fp32 pairs in registers instead of re-deriving them from the packed bf16 in every slot
        asm volatile("" : "+l"(xp[m][j][0]), "+l"(xp[m][j][1]), "+l"(xp[m][j][2]), "+l"(xp[m][j][3]));
      }
  }

  constexpr int V = RPI * MB;
  constexpr int LPV = 32 / V;
  constexpr int NBUF = 2 * ST_MAX_STAGES;
  const uint32_t lane_off = 16u * uint32_t(tid);
  RingPos pos = a.pos;
  int it = a.it;                        // slots so far: picks the finisher warp and the partial buffer
  int g = int(blockIdx.x) - a.deal_off;
  if (g < 0) g += gridDim.x;
  for (; g < groups; g += gridDim.x, ++it, pos.next(a.stages)) {
    const int s = pos.s;
    const uint32_t par = pos.par;
    {
      const long long t0 = (a.wait_acc && tid == 0) ? clock64() : 0;
      mbar_wait_a(a.full_bar + 8u * s, par);
      if (a.wait_acc && tid == 0) *a.wait_acc += clock64() - t0;
    }
    const uint32_t part = a.partial + uint32_t(it % NBUF) * (GV_CWARPS * 8 * 4);
    if (!(a.dbg & 1)) {
      const uint32_t slot = a.ring_base + uint32_t(s) * a.slot_bytes + lane_off;
      unsigned long long acc[RPI][MB];          // (sum over even elements, sum over odd elements)
#pragma unroll
      for (int rr = 0; rr < RPI; ++rr)
#pragma unroll
        for (int m = 0; m < MB; ++m) acc[rr][m] = 0ull;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        if (tid + j * GV_CONSUMERS < chunks) {
          uint4 wv[RPI];
#pragma unroll
          for (int rr = 0; rr < RPI; ++rr) wv[rr] = lds128(slot + uint32_t(rr) * pitch + uint32_t(j) * (16u * GV_CONSUMERS));
#pragma unroll
          for (int rr = 0; rr < RPI; ++rr) {
            const unsigned long long w0 = bf16x2_to_f32x2(wv[rr].x), w1 = bf16x2_to_f32x2(wv[rr].y),
                                     w2 = bf16x2_to_f32x2(wv[rr].z), w3 = bf16x2_to_f32x2(wv[rr].w);
#pragma unroll
            for (int m = 0; m < MB; ++m) {
              ffma2(acc[rr][m], w0, xp[m][j][0]);
              ffma2(acc[rr][m], w1, xp[m][j][1]);
              ffma2(acc[rr][m], w2, xp[m][j][2]);
              ffma2(acc[rr][m], w3, xp[m][j][3]);
            }
          }
        }
      }
      float flat[V];
#pragma unroll
      for (int rr = 0; rr < RPI; ++rr)
#pragma unroll
        for (int m = 0; m < MB; ++m) flat[rr * MB + m] = sum_f32x2(acc[rr][m]);
      warp_reduce_many<V>(flat, lane);
      if (lane % LPV == 0) sts32(part + uint32_t(warp * V + lane / LPV) * 4u, flat[0]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_a(a.part_bar + 8u * s);        // this warp is done with the slot and its partials are published
  }
  LinRet ret{pos, it};
  return ret;
}

// The finisher warp's side of a linear phase: for every slot, once the 16 consumer warps have published their partial
// sums (part_bar), add them in warp order, round, add the residual (requested while waiting), store, and only then hand
// the slot back to the producer (empty_bar, one arrival).  The consumer warps therefore never wait for each other: with
// a consumer warp doing this job in turn (first decoupled version) every slot's last step sat on the critical path of
// the next slot, and the consumers ran at 2000 cycles per slot with an IPC of 0.3.
template <int MB, int RPI>
__device__ __forceinline__ LinRet stack_finish(const LinArgs& a) {
  const int lane = threadIdx.x & 31;
  const int M = a.M, N = a.N;
  const int groups = (N + RPI - 1) / RPI;
  constexpr int V = RPI * MB;
  constexpr int NBUF = 2 * ST_MAX_STAGES;
  RingPos pos = a.pos;
  int it = a.it;
  int g = int(blockIdx.x) - a.deal_off;
  if (g < 0) g += gridDim.x;
  const int rr = lane / MB, m = lane % MB;
  for (; g < groups; g += gridDim.x, ++it, pos.next(a.stages)) {
    const int n = g * RPI + rr;
    const bool live = lane < V && n < N && m < M;
    float resv = 0.f;
    if (live && a.res) resv = ld_cg_bf16(a.res + int64_t(m) * a.ldr + n);
    mbar_wait_a(a.part_bar + 8u * pos.s, pos.par);
    if (live && !(a.dbg & 1)) {
      const uint32_t part = a.partial + uint32_t(it % NBUF) * (GV_CWARPS * 8 * 4);
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < GV_CWARPS; ++w2) t += lds32(part + uint32_t(w2 * V + lane) * 4u);
      float v = bf16_round(t);
      if (a.res) v += resv;
      a.out[int64_t(m) * a.ldo + n] = __float2bfloat16_rn(v);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_a(a.empty_bar + 8u * pos.s);
  }
  LinRet ret{pos, it};
  return ret;
}
template <int MB>
__device__ __forceinline__ LinRet stack_finish_k(const LinArgs& a, int rpi_big) {
  if (a.K <= GV_CONSUMERS * 8) return stack_finish<MB, 4>(a);
  if (rpi_big == 2) return stack_finish<MB, 2>(a);
  return stack_finish<MB, 1>(a);
}

// ---------------------------------------------------------------------------------------------- attention phase
// decode_attn_kernel<EPL, ROPE = 1> (un-split), one CTA per (sample, head) serving all n <= NQ query rows from ONE load
// of the keys / values.  Per query the arithmetic is that kernel's: warp w folds keys w, w + 16, ... in ascending order
// into its online softmax, the 16 warps are merged in order.  No partials in global memory, no atomics: the phase is
// five L2 round trips (the producer pulled this layer's K/V into L2 during the QKV phase) plus a shared-memory merge,
// while the other CTAs' rings fill with the output projection's weights.
struct AttnArgsDev {      // by value: a reference to the kernel parameters would force a local-memory copy of them
  const __nv_bfloat16 *qkv, *cos, *sin;
  __nv_bfloat16* ctx;
  int B, n, P, H;
  float scale;
};
template <int EPL, int NQ>
__device__ __noinline__ void stack_attention(const AttnArgsDev p, const __nv_bfloat16* cache, float* s_m /*[NQ][16]*/,
                                             float* s_l, float* s_acc /*[NQ][16][D]*/) {
  constexpr int D = 32 * EPL;
  constexpr int UN = ST_ATTN_UN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.n, P = p.P, H = p.H, hdim = H * D;
  const int Lk = P + n;
  const int items = p.B * H;
  const int64_t ldq = 3 * int64_t(hdim);
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int h = item % H, b = item / H;
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
    for (int j0 = warp; j0 < Lk; j0 += DEC_WARPS * UN) {
      RawEpl<EPL> kr[UN], vr[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int j = j0 + u * DEC_WARPS;
#pragma unroll
        for (int w2 = 0; w2 < (EPL + 1) / 2; ++w2) { kr[u].w[w2] = 0u; vr[u].w[w2] = 0u; }
        if (j < Lk) {
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
      float s[NQ][UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int j = j0 + u * DEC_WARPS;
        float kf[EPL];
        cvt_raw<EPL>(kr[u], kf);
        if (j >= P && j < Lk) rope_lanes<EPL>(kf, p.cos + int64_t(j - P) * (D / 2), p.sin + int64_t(j - P) * (D / 2), lane);
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
        for (int u = 0; u < UN; ++u) s[i][u] = d_wsum(s[i][u]);
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int j = j0 + u * DEC_WARPS;
        float vf[EPL];
        cvt_raw<EPL>(vr[u], vf);
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          if (i < n && j <= P + i) {                    // query i sees keys j <= P + i (< Lk)
            const float mn = fmaxf(m[i], s[i][u]);
            const float corr = __expf(m[i] - mn), pj = __expf(s[i][u] - mn);
            l[i] = l[i] * corr + pj;
#pragma unroll
            for (int e = 0; e < EPL; ++e) acc[i][e] = acc[i][e] * corr + pj * vf[e];
            m[i] = mn;
          }
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
      p.ctx[(int64_t(b) * n + i) * hdim + int64_t(h) * D + d] = __float2bfloat16_rn(Ls > 0.f ? a / Ls : 0.f);
    }
    consumers_sync();                      // s_m / s_l / s_acc are re-used by the next item
  }
}

template <int MB, int PRO>
__device__ __forceinline__ LinRet stack_linear_k(const LinArgs& a, int rpi_big) {
  if (a.K <= GV_CONSUMERS * 8) return stack_linear<MB, 1, 4, PRO>(a);
  if (rpi_big == 2) return stack_linear<MB, 3, 2, PRO>(a);
  return stack_linear<MB, 3, 1, PRO>(a);
}

template <int MB>
__global__ void __launch_bounds__(ST_THREADS, 1) decode_stack_kernel(const StackParams p) {
  extern __shared__ uint8_t st_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[ST_MAX_STAGES], empty_bar[ST_MAX_STAGES], part_bar[ST_MAX_STAGES];
  __shared__ float partial[2 * ST_MAX_STAGES * GV_CWARPS * 8];
  __shared__ float red_ss[GV_CWARPS * 2];
  __shared__ float s_m[MB * DEC_WARPS], s_l[MB * DEC_WARPS];
  const int tid = threadIdx.x, warp = tid >> 5;
  // attention merge buffer [MB][16][D] fp32: behind the ring
  float* s_acc = reinterpret_cast<float*>(ring + size_t(p.stages) * p.slot_bytes);
  if (tid == 0) {
    for (int s = 0; s < ST_MAX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);              // the finisher warp hands a slot back
      mbar_init(&part_bar[s], GV_CWARPS);       // the 16 consumer warps have published their partial sums
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
  Deal deal{0};
  GridBar gb{p.bar, 0u};
  if (warp == GV_CWARPS + 1) {
    // ---- finisher warp: the same walk over layers and phases as the consumers, one final sum per slot
    LinArgs fa;
    fa.empty_bar = smem_u32(empty_bar); fa.part_bar = smem_u32(part_bar); fa.partial = smem_u32(partial);
    fa.stages = p.stages; fa.M = M; fa.dbg = p.dbg; fa.ldr = h;
    fa.pos = RingPos{0, 0u};
    fa.it = 0;
    auto finish = [&](const __nv_bfloat16* res, __nv_bfloat16* out, int64_t ldo, int N, int K) {
      fa.res = res; fa.out = out; fa.ldo = ldo; fa.N = N; fa.K = K; fa.deal_off = deal.off;
      const LinRet lr = stack_finish_k<MB>(fa, p.rpi_big);
      fa.pos = lr.pos;
      fa.it = lr.it;
      const int rpi = rows_per_slot(K, p.rpi_big);
      deal.next((N + rpi - 1) / rpi, grid);
      grid_barrier(gb, tid, p.dbg);             // tid != 0: only the two CTA-level syncs
    };
    for (int l = 0; l < p.L; ++l) {
      finish(nullptr, p.qkv, 3 * int64_t(h), 3 * h, h);
      grid_barrier(gb, tid, p.dbg);             // attention phase
      finish(p.x, p.xmid, h, h, h);
      finish(nullptr, p.gu, 2 * int64_t(f), 2 * f, h);
      finish(p.xmid, p.x, h, h, f);
    }
    return;
  }
  const AttnArgsDev aa{p.qkv, p.cos, p.sin, p.ctx, p.B, p.n, p.P, p.H, p.scale};
  LinArgs la;
  la.ring_base = smem_u32(ring); la.slot_bytes = p.slot_bytes; la.stages = p.stages;
  la.full_bar = smem_u32(full_bar); la.empty_bar = smem_u32(empty_bar); la.part_bar = smem_u32(part_bar);
  la.partial = smem_u32(partial); la.red_ss = red_ss;
  la.M = M; la.eps = p.eps; la.dbg = p.dbg;
  la.pos = RingPos{0, 0u};
  la.it = 0;
  la.wait_acc = p.trace ? p.trace + size_t(gridDim.x) * p.L * 15 + size_t(blockIdx.x) * 4 + 3 : nullptr;
  if (p.trace && tid == 0) *la.wait_acc = 0;
  if (tid == 0) {
    const unsigned sel = ld_acquire_u32(p.bar + 2) & 1u;
    gb.ctr = p.bar + sel;
    if (blockIdx.x == 0) p.bar[sel ^ 1u] = 0u;          // the counter the NEXT launch will use
  }
  long long* tr = (p.trace && tid == 0) ? p.trace + size_t(blockIdx.x) * p.L * 15 : nullptr;
#define ST_TRACE(ph, k) do { if (tr) tr[(l * 5 + (ph)) * 3 + (k)] = global_ns(); } while (0)
  auto linear = [&](auto pro, const __nv_bfloat16* x, int64_t ldx, const __nv_bfloat16* ln_w, const __nv_bfloat16* res,
                    __nv_bfloat16* out, int64_t ldo, int N, int K) {
    la.x = x; la.ldx = ldx; la.ln_w = ln_w; la.res = res; la.ldr = h; la.out = out; la.ldo = ldo; la.N = N; la.K = K;
    la.deal_off = deal.off;
    const LinRet lr = stack_linear_k<MB, decltype(pro)::value>(la, p.rpi_big);
    la.pos = lr.pos;
    la.it = lr.it;
    const int rpi = rows_per_slot(K, p.rpi_big);
    deal.next((N + rpi - 1) / rpi, grid);
  };
  for (int l = 0; l < p.L; ++l) {
    // RMSNorm + QKV projection
    ST_TRACE(0, 0);
    linear(std::integral_constant<int, GV_PRO_RMSNORM>{}, p.x, h, p.ln1[l], nullptr, p.qkv, 3 * int64_t(h), 3 * h, h);
    ST_TRACE(0, 1);
    grid_barrier(gb, tid, p.dbg);
    if (l == 0 && tid == 0 && blockIdx.x == 0) p.bar[2] = unsigned(gb.ctr - p.bar) ^ 1u;   // every CTA has read it by now
    ST_TRACE(0, 2);
    // attention (RoPE of q and of the new keys on the fly)
    ST_TRACE(1, 0);
    if (!(p.dbg & 4)) switch (p.D) {
      case 32: stack_attention<1, MB>(aa, p.cache[l], s_m, s_l, s_acc); break;
      case 64: stack_attention<2, MB>(aa, p.cache[l], s_m, s_l, s_acc); break;
      default: stack_attention<4, MB>(aa, p.cache[l], s_m, s_l, s_acc); break;
    }
    ST_TRACE(1, 1);
    grid_barrier(gb, tid, p.dbg);
    ST_TRACE(1, 2);
    // output projection + residual
    ST_TRACE(2, 0);
    linear(std::integral_constant<int, GV_PRO_NONE>{}, p.ctx, h, nullptr, p.x, p.xmid, h, h, h);
    ST_TRACE(2, 1);
    grid_barrier(gb, tid, p.dbg);
    ST_TRACE(2, 2);
    // RMSNorm + gate | up projection
    ST_TRACE(3, 0);
    linear(std::integral_constant<int, GV_PRO_RMSNORM>{}, p.xmid, h, p.ln2[l], nullptr, p.gu, 2 * int64_t(f), 2 * f, h);
    ST_TRACE(3, 1);
    grid_barrier(gb, tid, p.dbg);
    ST_TRACE(3, 2);
    // SwiGLU + down projection + residual -> the next layer's input, in place
    ST_TRACE(4, 0);
    linear(std::integral_constant<int, GV_PRO_SWIGLU>{}, p.gu, 2 * int64_t(f), nullptr, p.xmid, p.x, h, h, f);
    ST_TRACE(4, 1);
    grid_barrier(gb, tid, p.dbg);
    ST_TRACE(4, 2);
  }
#undef ST_TRACE
}

// ---------------------------------------------------------------------------------------------- one linear, per-op form
// The same consumer / finisher / producer roles for ONE skinny linear (mla_gemv_fused's fast path, MLA_GEMV2=1): the 16
// consumer warps are not synchronised per slot (partials published through an mbarrier, a dedicated finisher warp adds
// them), dot products with FFMA2 — the two changes that took the consumers from ~3.0 to ~2.1 ms-equivalent per step inside
// the stack kernel.  Programmatic dependent launch as in gemv_ring_kernel: the producer streams weights right away, the
// consumers and the finisher wait for the previous kernel before touching activations / residual / output.
struct Gemv2Params {
  const __nv_bfloat16 *x, *w, *res, *ln_w;
  __nv_bfloat16* out;
  int M, N, K;
  int64_t ldx, ldo, ldr;
  float eps;
  int prologue;
  int stages;
  uint32_t slot_bytes;
  int rpi;
};

template <int MB>
__global__ void __launch_bounds__(ST_THREADS, 1) gemv2_kernel(const Gemv2Params p) {
  extern __shared__ uint8_t g2_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(g2_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[ST_MAX_STAGES], empty_bar[ST_MAX_STAGES], part_bar[ST_MAX_STAGES];
  __shared__ float partial[2 * ST_MAX_STAGES * GV_CWARPS * 8];
  __shared__ float red_ss[GV_CWARPS * 2];
  const int tid = threadIdx.x, warp = tid >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    for (int s = 0; s < ST_MAX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&part_bar[s], GV_CWARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int rpi = p.rpi;
  if (warp == GV_CWARPS) {
    if ((tid & 31) == 0) {
      // ---- producer: this CTA's groups, one bulk copy per slot where the pitch needs no padding
      const uint32_t row_bytes = uint32_t(p.K) * 2u, pitch = (row_bytes + 127u) & ~127u;
      const int groups = (p.N + rpi - 1) / rpi, full_groups = p.N / rpi;
      const uint32_t group_bytes = row_bytes * uint32_t(rpi);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w) + size_t(blockIdx.x) * group_bytes;
      const size_t stride = size_t(gridDim.x) * group_bytes;
      RingPos pos{0, 0u};
      for (int g = blockIdx.x; g < groups; g += gridDim.x, src += stride) {
        mbar_wait_a(smem_u32(&empty_bar[pos.s]), pos.par ^ 1u);
        const uint32_t bytes = g < full_groups ? group_bytes : row_bytes * uint32_t(p.N - g * rpi);
        mbar_arrive_expect_tx(&full_bar[pos.s], bytes);
        uint8_t* dst = ring + size_t(pos.s) * p.slot_bytes;
        if (pitch == row_bytes) {
          bulk_load_row(dst, src, bytes, &full_bar[pos.s]);
        } else {
          for (uint32_t o = 0, q = 0; o < bytes; o += row_bytes, q += pitch) bulk_load_row(dst + q, src + o, row_bytes, &full_bar[pos.s]);
        }
        pos.next(p.stages);
      }
    }
    return;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  LinArgs la;
  la.ring_base = smem_u32(ring); la.slot_bytes = p.slot_bytes; la.stages = p.stages;
  la.full_bar = smem_u32(full_bar); la.empty_bar = smem_u32(empty_bar); la.part_bar = smem_u32(part_bar);
  la.partial = smem_u32(partial); la.red_ss = red_ss;
  la.x = p.x; la.ln_w = p.ln_w; la.res = p.res; la.out = p.out;
  la.ldx = p.ldx; la.ldr = p.ldr; la.ldo = p.ldo;
  la.M = p.M; la.N = p.N; la.K = p.K; la.eps = p.eps; la.dbg = 0; la.deal_off = 0;
  la.pos = RingPos{0, 0u};
  la.it = 0;
  la.wait_acc = nullptr;
  if (warp == GV_CWARPS + 1) {
    stack_finish_k<MB>(la, rpi);
    return;
  }
  if (p.prologue == GV_PRO_RMSNORM) stack_linear_k<MB, GV_PRO_RMSNORM>(la, rpi);
  else if (p.prologue == GV_PRO_SWIGLU) stack_linear_k<MB, GV_PRO_SWIGLU>(la, rpi);
  else stack_linear_k<MB, GV_PRO_NONE>(la, rpi);
}

}  // namespace mla

using namespace mla;

// mla_gemv_fused's fast path (m <= 2) through gemv2_kernel; returns MLA_ERR_ARG-free 1 when the shape is not taken
int gemv2_launch(const mla_gemv_args* a, void* stream) {
  const int M = a->m, N = a->n, K = a->k;
  if (M > 2 || K > GV_CONSUMERS * 8 * 3 || (K & 7) || a->ldw != K) return 1;
  Gemv2Params p;
  p.x = (const __nv_bfloat16*)a->x; p.w = (const __nv_bfloat16*)a->w; p.res = (const __nv_bfloat16*)a->residual;
  p.ln_w = (const __nv_bfloat16*)a->ln_weight; p.out = (__nv_bfloat16*)a->out;
  p.M = M; p.N = N; p.K = K; p.ldx = a->ldx; p.ldo = a->ldo; p.ldr = a->ldr; p.eps = a->eps; p.prologue = a->prologue;
  p.rpi = K <= GV_CONSUMERS * 8 ? 4 : 2;
  const size_t pitch = (size_t(K) * 2 + 127) & ~size_t(127);
  const size_t slot = pitch * p.rpi;
  const size_t ring_bytes = 227 * 1024 - 10240 - 128;
  int stages = int(ring_bytes / slot);
  if (stages < 2) return 1;
  p.stages = stages > ST_MAX_STAGES ? ST_MAX_STAGES : stages;
  p.slot_bytes = uint32_t(slot);
  const size_t smem = slot * p.stages + 128;
  auto kern = M <= 1 ? gemv2_kernel<1> : gemv2_kernel<2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "cudaFuncSetAttribute(gemv2 smem %zu): %s", smem, cudaGetErrorString(e));
  const int groups = (N + p.rpi - 1) / p.rpi;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups < num_sms() ? groups : num_sms());
  cfg.blockDim = dim3(ST_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const char* env = getenv("MLA_DECODE_PDL");
  cfg.numAttrs = (env && env[0] == '0') ? 0 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "gemv2 launch: %s", cudaGetErrorString(e));
  MLA_CHECK_LAUNCH("gemv2");
  return MLA_OK;
}

static int g_stack_coop = -1;
static int g_stack_dbg = 0;
static int g_stack_rpi_big = 2;
extern "C" int mla_decode_stack_set_rows_per_slot_big(int32_t rows) {
  g_stack_rpi_big = rows == 1 ? 1 : 2;
  return MLA_OK;
}
static int g_stack_ring_kb = 0;
extern "C" int mla_decode_stack_set_ring_kb(int32_t kb) {
  g_stack_ring_kb = kb;
  return MLA_OK;
}
static int g_stack_ahead = -1;
extern "C" int mla_decode_stack_set_ahead(int32_t groups) {
  g_stack_ahead = groups < 0 ? 0 : groups;
  return MLA_OK;
}
extern "C" int mla_decode_stack_set_debug(int32_t flags) {
  g_stack_dbg = flags;
  return MLA_OK;
}

extern "C" size_t mla_decode_stack_workspace(int32_t, int32_t, int32_t, int32_t, int32_t) {
  return 256;       // the grid barrier's words; must be zero before the FIRST launch only
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
  p.bar = (unsigned*)a->workspace;
  p.trace = (long long*)a->trace;
  p.dbg = g_stack_dbg;
  if (g_stack_ahead < 0) {
    const char* e = getenv("MLA_DECODE_STACK_AHEAD");
    g_stack_ahead = e ? atoi(e) : 0;
  }
  p.ahead = g_stack_ahead;
  p.rpi_big = g_stack_rpi_big;
  auto pitch_of = [](int K) { return (size_t(K) * 2 + 127) & ~size_t(127); };
  auto rpi_of = [](int K) { return K <= GV_CONSUMERS * 8 ? 4 : g_stack_rpi_big; };
  size_t slot = pitch_of(h) * rpi_of(h);
  if (pitch_of(a->ffn) * rpi_of(a->ffn) > slot) slot = pitch_of(a->ffn) * rpi_of(a->ffn);
  const size_t attn_smem = size_t(M <= 1 ? 1 : 2) * DEC_WARPS * a->head_dim * sizeof(float);
  size_t ring_bytes = 227 * 1024 - 10240 - attn_smem - 128;          // minus static shared memory + alignment slack
  if (g_stack_ring_kb > 0 && size_t(g_stack_ring_kb) * 1024 < ring_bytes) ring_bytes = size_t(g_stack_ring_kb) * 1024;
  int stages = int(ring_bytes / slot);
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
  cfg.blockDim = dim3(ST_THREADS);
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
