// Skinny linear for the inference denoise step on the 5th-gen tensor cores, "swap-AB":
//
//   out[m, n] = bf16( bf16( sum_k x'[m,k] w[n,k] ) + residual[m,n] ),   m < M <= 32 rows, w [N,K] an nn.Linear weight
//
// computed as out^T = W . x'^T: the WEIGHTS are the 128-row A operand of tcgen05.mma (streamed by TMA into a 128B-swizzled
// shared-memory ring, exactly like the training GEMM's operands), the few activation rows are its 16- or 32-column B
// operand, the accumulator is a 128 x 16/32 fp32 tile in TMEM.  The weights never pass through the CUDA cores: the
// per-op gemv kernels (decode.cu) and the one-launch stack (decode_stack.cu) were both bounded by the consumer warps'
// instruction stream (bf16 unpack + FMA + warp reduction per weight element), not by HBM
// (profiles/r02_decode_stack_findings.md); here a k-block of 128 x 64 weights costs the issuing thread four MMAs.
//
//   warp 0      TMA producer: W tiles {64 k, 128 rows} -> 8-stage ring (16 KB per stage); starts before the previous
//               kernel has finished (programmatic dependent launch: weights do not depend on it)
//   warp 1      MMA issuer (owns the TMEM allocation: two accumulators, so the epilogue of one unit overlaps the next)
//   warps 2-5   prologue: x' (RMSNorm / SwiGLU of the activations, modeling_llama.py:85-90,:240) written straight into
//               the B operand's swizzled layout for this CTA's k-range; then epilogue: TMEM -> split-K partial ->
//               (last arrival per row tile) fixed-order sum, bf16 round, residual, store
// Work split: unit = (128-row tile, k-split); the k-splits divide the grid so a CTA keeps ONE k-range (one B operand)
// and walks row tiles.  Split-K partials meet in a small fp32 workspace; the last CTA to arrive for a row tile (atomic
// counter, self re-arming) adds them in split order, so the result does not depend on timing.
#include <cstdlib>

#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int SK_BM = 128, SK_BK = 64, SK_UK = 16;
constexpr int SK_MAX_STAGES = 8;
constexpr int SK_A_BYTES = SK_BM * SK_BK * 2;          // 16 KB
constexpr int SK_THREADS = 192;
constexpr int SK_EPI_THREADS = 128;
enum { SK_PRO_NONE = 0, SK_PRO_RMSNORM = 1, SK_PRO_SWIGLU = 2 };

struct SkinnyParams {
  const __nv_bfloat16 *x, *ln_w, *res;
  __nv_bfloat16* out;
  int M, N, K;
  int64_t ldx, ldr, ldo;
  float eps;
  int prologue;
  int ksplit;            // divides gridDim.x
  int num_kb;            // ceil(K / 64)
  int max_kb;            // k-blocks of the largest split (sizes the B area)
  int stages;            // ring depth (16 KB each): sized so that TWO CTAs fit an SM — two of this kernel's, or one of
                         // this and one of the next launch, which then streams its weights while this one drains
  float* ws;             // [row tiles][ksplit][NT][128] fp32 partials (ksplit > 1)
  int* counters;         // [row tiles], zero on entry, left zero
};

__device__ __forceinline__ void sk_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sk_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void sk_epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SK_EPI_THREADS) : "memory"); }
__device__ __forceinline__ void sk_unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 sk_pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void sk_tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <int NT>       // token columns of the MMA: 16 or 32
__global__ void __launch_bounds__(SK_THREADS, 2)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const SkinnyParams p) {
  extern __shared__ uint8_t sk_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sk_smem_raw) + 1023) & ~uintptr_t(1023));
  const int STG = p.stages;
  uint8_t* sB = smem + STG * SK_A_BYTES;                       // [max_kb][NT rows][128 B], 128B-swizzled
  constexpr int B_TILE = NT * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + size_t(p.max_kb) * B_TILE);
  uint64_t *full_bar = bars, *empty_bar = bars + SK_MAX_STAGES, *acc_full = bars + 2 * SK_MAX_STAGES, *acc_empty = acc_full + 2,
           *b_ready = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_ready + 1);
  __shared__ float red[2 * 4];            // per-warp partial sums of the row being reduced (double-buffered by row parity)
  __shared__ float s_rstd[32];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int ks = blockIdx.x % p.ksplit;
  const int rt0 = blockIdx.x / p.ksplit, rt_step = gridDim.x / p.ksplit;
  const int row_tiles = (p.N + SK_BM - 1) / SK_BM;
  const int kb_begin = int(int64_t(ks) * p.num_kb / p.ksplit), kb_end = int(int64_t(ks + 1) * p.num_kb / p.ksplit);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < STG; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
    mbar_init(b_ready, SK_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);            // two accumulators of NT <= 32 columns
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  sk_pdl_launch_dependents();             // the next kernel may become resident and start streaming ITS weights

  if (warp == 0) {
    // ===================== TMA producer: weights only, no dependence on the previous kernel =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int rt = rt0; rt < row_tiles; rt += rt_step) {
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], SK_A_BYTES);
          tma_load_2d(smem + stage * SK_A_BYTES, &map_w, &full_bar[stage], kb * SK_BK, rt * SK_BM);
          if (++stage == STG) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(SK_BM, NT, 0, 0);
      mbar_wait(b_ready, 0);              // x' is in shared memory (written through the generic proxy + proxy fence)
      tc_fence_after();
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int rt = rt0; rt < row_tiles; rt += rt_step) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 32;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SK_A_BYTES);
          const uint32_t sb = smem_u32(sB + size_t(kb) * B_TILE);
#pragma unroll
          for (int k = 0; k < SK_BK / SK_UK; ++k)
            umma_f16_ss(tmem_d, umma_smem_desc_sw128(sa + k * 32, 16, 1024), umma_smem_desc_sw128(sb + k * 32, 16, 1024),
                        idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STG) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== prologue (x' -> B operand), then epilogue =====================
    const int et = threadIdx.x - 64;          // 0..127
    const int ew = et >> 5;                   // warp within the group = TMEM lane quarter (warp index % 4 == ew + 2 ... see below)
    sk_pdl_wait();                            // activations / residual / output belong to the kernels before us
    const int M = p.M, K = p.K;
    const int chunks_row = K >> 3;
    // ---- RMSNorm statistics over the FULL row (every CTA needs them for its k-range)
    if (p.prologue == SK_PRO_RMSNORM) {
      for (int m = 0; m < M; ++m) {
        float ss = 0.f;
        for (int c = et; c < chunks_row; c += SK_EPI_THREADS) {
          float f[8];
          sk_unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + int64_t(m) * p.ldx + 8 * c)), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[(m & 1) * 4 + ew] = ss;
        sk_epi_sync();
        if (et == 0) {
          const float t = (red[(m & 1) * 4 + 0] + red[(m & 1) * 4 + 1]) + (red[(m & 1) * 4 + 2] + red[(m & 1) * 4 + 3]);
          s_rstd[m] = rsqrtf(t / float(K) + p.eps);
        }
      }
      sk_epi_sync();
    }
    // ---- this CTA's k-range of x' into the swizzled B tiles: row r = token, 128 B per row and k-block, 16-byte chunk j
    //      of row r sits at chunk j ^ (r & 7) of its 8-row group (what TMA SWIZZLE_128B would have produced)
    const int cpb = SK_BK / 8;                                     // 16-byte chunks per row and k-block
    for (int idx = et; idx < nkb * NT * cpb; idx += SK_EPI_THREADS) {
      const int j = idx % cpb, r = (idx / cpb) % NT, kbl = idx / (cpb * NT);
      const int k0 = (kb_begin + kbl) * SK_BK + j * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < M && k0 < K) {
        if (p.prologue == SK_PRO_SWIGLU) {
          float g[8], u[8], o[8];
          sk_unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + int64_t(r) * p.ldx + k0)), g);
          sk_unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + int64_t(r) * p.ldx + K + k0)), u);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = bf16_round(g[e] * (1.f / (1.f + __expf(-g[e])))) * u[e];
          v = sk_pack8(o);
        } else {
          v = __ldg(reinterpret_cast<const uint4*>(p.x + int64_t(r) * p.ldx + k0));
          if (p.prologue == SK_PRO_RMSNORM) {
            float f[8], g[8], o[8];
            sk_unpack8(v, f);
            sk_unpack8(__ldg(reinterpret_cast<const uint4*>(p.ln_w + k0)), g);
            const float rs = s_rstd[r];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = g[e] * bf16_round(f[e] * rs);
            v = sk_pack8(o);
          }
        }
      }
      *reinterpret_cast<uint4*>(sB + size_t(kbl) * B_TILE + (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core's async-proxy reads
    mbar_arrive(b_ready);

    // ---- epilogue: TMEM lane = weight row of the tile.  A warp may only touch the lane quarter (warp index % 4).
    const int quarter = warp & 3;
    const int n_loc = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int rt = rt0; rt < row_tiles; rt += rt_step) {
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      float v[NT];
      {
        uint32_t r[16];
        sk_tmem_ld_x16(tmem_base + (uint32_t(quarter * 32) << 16) + acc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (NT == 32) {
          sk_tmem_ld_x16(tmem_base + (uint32_t(quarter * 32) << 16) + acc * 32 + 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[(NT == 32 ? 16 : 0) + i] = __uint_as_float(r[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);          // the accumulator can take the next unit
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      const int n = rt * SK_BM + n_loc;
      bool finish = true;
      if (p.ksplit > 1) {
        float* part = p.ws + (size_t(rt) * p.ksplit + ks) * (NT * SK_BM);
#pragma unroll
        for (int m = 0; m < NT; ++m)
          if (m < M) part[m * SK_BM + n_loc] = v[m];
        __threadfence();                       // this CTA's partial is visible before its arrival is counted
        sk_epi_sync();
        if (et == 0) s_last = atomicAdd(p.counters + rt, 1) == p.ksplit - 1;
        sk_epi_sync();
        finish = s_last != 0;
        if (finish) {
          __threadfence();
          const float* all = p.ws + size_t(rt) * p.ksplit * (NT * SK_BM);
#pragma unroll
          for (int m = 0; m < NT; ++m) {
            if (m < M) {
              float t = 0.f;
              for (int s2 = 0; s2 < p.ksplit; ++s2) t += __ldcg(all + size_t(s2) * (NT * SK_BM) + m * SK_BM + n_loc);
              v[m] = t;
            }
          }
          if (et == 0) p.counters[rt] = 0;     // re-armed for the next launch
        }
        sk_epi_sync();                         // s_last is re-used by the next row tile
      }
      if (finish && n < p.N) {
#pragma unroll
        for (int m = 0; m < NT; ++m) {
          if (m < M) {
            float o = bf16_round(v[m]);
            if (p.res) o += __bfloat162float(p.res[int64_t(m) * p.ldr + n]);
            p.out[int64_t(m) * p.ldo + n] = __float2bfloat16_rn(o);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace mla

using namespace mla;

static int g_skinny_pdl = -1;

extern "C" size_t mla_skinny_gemm_workspace(int32_t n, int32_t m) {
  // split-K partials for up to 32 splits + one arrival counter per 128-row tile (zero before the FIRST launch only)
  const size_t rt = size_t((n + SK_BM - 1) / SK_BM);
  const size_t nt = m <= 16 ? 16 : 32;
  return rt * 32 * nt * SK_BM * sizeof(float) + ((rt * sizeof(int) + 255) & ~size_t(255));
}

extern "C" int mla_skinny_gemm(const mla_gemv_args* a, void* workspace, void* stream) {
  if (int rc = device_check()) return rc;
  if (a == nullptr) return set_error(MLA_ERR_ARG, "skinny_gemm: null args");
  if (a->m <= 0 || a->n <= 0) return MLA_OK;
  const int M = a->m, N = a->n, K = a->k;
  if (M > 32) return set_error(MLA_ERR_ARG, "skinny_gemm: m=%d rows, at most 32 (use mla_gemm_bf16)", M);
  if (K <= 0 || (K & 7) || (a->ldx & 7) || (a->ldw & 7))
    return set_error(MLA_ERR_ARG, "skinny_gemm: k and the row pitches of x and w must be multiples of 8 (k=%d)", K);
  if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w)) & 15)
    return set_error(MLA_ERR_ARG, "skinny_gemm: x and w must be 16-byte aligned");
  if (a->prologue < SK_PRO_NONE || a->prologue > SK_PRO_SWIGLU) return set_error(MLA_ERR_ARG, "skinny_gemm: unknown prologue");
  if (a->prologue == SK_PRO_RMSNORM && (!a->ln_weight || (reinterpret_cast<uintptr_t>(a->ln_weight) & 15)))
    return set_error(MLA_ERR_ARG, "skinny_gemm: the RMSNorm prologue needs a 16-byte aligned weight vector");
  if (!workspace) return set_error(MLA_ERR_ARG, "skinny_gemm: no workspace");
  const int NT = M <= 16 ? 16 : 32;
  const int row_tiles = (N + SK_BM - 1) / SK_BM;
  const int num_kb = (K + SK_BK - 1) / SK_BK;
  // grid and k-split: the k-split divides the grid (a CTA keeps one k-range), every split has >= 2 k-blocks and its B
  // operand fits beside the ring; among those, the best-balanced split (units / (rounds * grid)), then the smallest
  // Two CTAs per SM (~113 KB each: ring + B operand + barriers).  Grid and k-split: the k-split divides the grid (a CTA
  // keeps one k-range), every split has >= 2 k-blocks and its B operand leaves room for >= 3 ring stages; among those,
  // the best-balanced one (units / (rounds * slots)), then the smallest.
  const int slots = 2 * num_sms();
  const size_t cta_cap = 113 * 1024 - 2048;
  int best_ks = 0, best_grid = 0, best_stages = 0;
  double best_eff = -1.0;
  for (int ks = 1; ks <= 32; ++ks) {
    if (ks > 1 && num_kb / ks < 2) break;
    const int max_kb = (num_kb + ks - 1) / ks;
    const size_t bbytes = size_t(max_kb) * NT * 128;
    if (bbytes + 3 * size_t(SK_A_BYTES) > cta_cap) continue;
    int stages = int((cta_cap - bbytes) / SK_A_BYTES);
    if (stages > SK_MAX_STAGES) stages = SK_MAX_STAGES;
    int grid = (slots / ks) * ks;
    const int units = row_tiles * ks;
    if (units < grid) grid = units;
    if (grid <= 0) continue;
    const int rounds = (units + grid - 1) / grid;
    const double eff = double(units) / (double(rounds) * slots);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_ks = ks; best_grid = grid; best_stages = stages; }
  }
  if (best_ks == 0) return set_error(MLA_ERR_ARG, "skinny_gemm: k=%d does not fit the shared-memory operand area", K);
  SkinnyParams p;
  p.x = (const __nv_bfloat16*)a->x; p.ln_w = (const __nv_bfloat16*)a->ln_weight; p.res = (const __nv_bfloat16*)a->residual;
  p.out = (__nv_bfloat16*)a->out;
  p.M = M; p.N = N; p.K = K; p.ldx = a->ldx; p.ldr = a->ldr; p.ldo = a->ldo; p.eps = a->eps; p.prologue = a->prologue;
  p.ksplit = best_ks; p.num_kb = num_kb; p.max_kb = (num_kb + best_ks - 1) / best_ks; p.stages = best_stages;
  p.ws = (float*)workspace;
  p.counters = (int*)((uint8_t*)workspace + size_t(row_tiles) * 32 * NT * SK_BM * sizeof(float));
  CUtensorMap map_w;
  const uint64_t dims[2] = {uint64_t(K), uint64_t(N)};
  const uint64_t strides[1] = {uint64_t(a->ldw) * 2};
  const uint32_t box[2] = {SK_BK, SK_BM};
  if (int rc = encode_tmap_2d_bf16(&map_w, a->w, dims, strides, box)) return rc;
  const size_t smem = size_t(best_stages) * SK_A_BYTES + size_t(p.max_kb) * NT * 128 + 256 + 1024;
  auto kern = NT == 16 ? skinny_gemm_kernel<16> : skinny_gemm_kernel<32>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "cudaFuncSetAttribute(skinny smem %zu): %s", smem, cudaGetErrorString(e));
  if (g_skinny_pdl < 0) {
    const char* env = getenv("MLA_DECODE_PDL");
    g_skinny_pdl = (env && env[0] == '0') ? 0 : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(best_grid);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_skinny_pdl ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, map_w, p);
  if (e != cudaSuccess) return set_error(MLA_ERR_CUDA, "skinny_gemm launch: %s", cudaGetErrorString(e));
  MLA_CHECK_LAUNCH("skinny_gemm");
  return MLA_OK;
}
