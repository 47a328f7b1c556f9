// (1) Shifted cross-entropy over the vocabulary (LlamaForCausalLM.forward, modeling_llama.py:1254-1269): logits come
//     out of the lm_head GEMM as bf16, the reference upcasts them (`logits.float()`) and runs CrossEntropyLoss
//     (ignore_index -100, mean over the non-ignored rows) on logits[:, :-1] vs labels[:, 1:].
// (2) ActionTokenizer (vla/action_tokenizer.py:43-71): np.clip + np.digitize against the float64 bin edges, id =
//     vocab_size - bin; decode through the bin centres.  Integer results are bit-exact: the comparison is done in
//     double against the host-computed np.linspace table.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

__device__ __forceinline__ float blk_reduce(float v, float* red, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = (threadIdx.x & 31) < nw ? red[threadIdx.x & 31] : (is_max ? -INFINITY : 0.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, r, o);
    r = is_max ? fmaxf(r, t) : r + t;
  }
  return r;
}

// One CTA per row r = b*S+s.  Target = labels[b, s+1] (ignored for s = S-1 or label == -100).
__global__ void ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                              int S, int V, float* __restrict__ lse_out, float* __restrict__ acc /* [2]: sum, count */) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int s = int(row % S);
  const int64_t tgt = (s == S - 1) ? -100 : labels[row + 1];
  if (tgt < 0) { if (threadIdx.x == 0) lse_out[row] = 0.f; return; }   // uniform per CTA
  const __nv_bfloat16* x = logits + row * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, __bfloat162float(x[i]));
  mx = blk_reduce(mx, red, true);
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) se += __expf(__bfloat162float(x[i]) - mx);
  se = blk_reduce(se, red, false);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(se);
    lse_out[row] = lse;
    atomicAdd(acc, lse - __bfloat162float(x[tgt]));
    atomicAdd(acc + 1, 1.f);
  }
}
__global__ void ce_finalize_kernel(const float* __restrict__ acc, float* __restrict__ loss) {
  loss[0] = acc[1] > 0.f ? acc[0] / acc[1] : 0.f / 0.f;   // torch returns nan when every target is ignored
}
// logits -> dlogits in place: g/count * (softmax - onehot) for rows with a target, 0 otherwise
__global__ void ce_bwd_kernel(__nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                              int S, int V, const float* __restrict__ lse, const float* __restrict__ acc,
                              const float* __restrict__ gscale) {
  const int64_t row = blockIdx.x;
  const int s = int(row % S);
  const int64_t tgt = (s == S - 1) ? -100 : labels[row + 1];
  __nv_bfloat16* x = logits + row * ld;
  if (tgt < 0) {
    for (int i = threadIdx.x; i < V; i += blockDim.x) x[i] = __float2bfloat16_rn(0.f);
    return;
  }
  const float coef = gscale[0] / acc[1];
  const float l = lse[row];
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float p = __expf(__bfloat162float(x[i]) - l);
    x[i] = __float2bfloat16_rn(coef * (p - (i == tgt ? 1.f : 0.f)));
  }
}

template <typename T>
__global__ void digitize_kernel(const T* __restrict__ x, int64_t n, const double* __restrict__ edges, int bins,
                                double lo, double hi, int64_t vocab, int64_t* __restrict__ ids) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  T v = x[i];
  v = v < T(lo) ? T(lo) : (v > T(hi) ? T(hi) : v);   // np.clip keeps the input dtype
  const double d = double(v);
  // np.digitize(x, bins) with increasing bins, right=False: number of edges <= x
  int a = 0, b = bins;
  while (a < b) {
    const int m = (a + b) >> 1;
    if (edges[m] <= d) a = m + 1; else b = m;
  }
  ids[i] = vocab - int64_t(a);
}
__global__ void action_decode_kernel(const int64_t* __restrict__ ids, int64_t n, const double* __restrict__ centers,
                                     int n_centers, int64_t vocab, double* __restrict__ out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int64_t d = vocab - ids[i] - 1;
  d = d < 0 ? 0 : (d > n_centers - 1 ? n_centers - 1 : d);
  out[i] = centers[d];
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))

extern "C" int mla_ce_fwd(const void* logits, int64_t ld, const void* labels, int64_t rows, int32_t seq, int32_t vocab,
                          void* lse, void* acc2, void* loss, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return set_error(MLA_ERR_ARG, "ce: empty");
  cudaMemsetAsync(acc2, 0, 2 * sizeof(float), S_(stream));
  ce_fwd_kernel<<<(unsigned)rows, 256, 0, S_(stream)>>>((const __nv_bfloat16*)logits, ld, (const int64_t*)labels, seq, vocab,
                                                       (float*)lse, (float*)acc2);
  MLA_CHECK_LAUNCH("ce_fwd");
  ce_finalize_kernel<<<1, 1, 0, S_(stream)>>>((const float*)acc2, (float*)loss);
  MLA_CHECK_LAUNCH("ce_finalize");
  return MLA_OK;
}
extern "C" int mla_ce_bwd(void* logits_inout, int64_t ld, const void* labels, int64_t rows, int32_t seq, int32_t vocab,
                          const void* lse, const void* acc2, const void* gscale, void* stream) {
  if (int rc = device_check()) return rc;
  if (rows <= 0) return MLA_OK;
  ce_bwd_kernel<<<(unsigned)rows, 256, 0, S_(stream)>>>((__nv_bfloat16*)logits_inout, ld, (const int64_t*)labels, seq, vocab,
                                                       (const float*)lse, (const float*)acc2, (const float*)gscale);
  MLA_CHECK_LAUNCH("ce_bwd");
  return MLA_OK;
}
extern "C" int mla_action_digitize(const void* x, int32_t x_is_f64, int64_t n, const void* edges, int32_t bins,
                                   double lo, double hi, int64_t vocab_size, void* ids_out, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  const unsigned grid = unsigned((n + 255) / 256);
  if (x_is_f64)
    digitize_kernel<double><<<grid, 256, 0, S_(stream)>>>((const double*)x, n, (const double*)edges, bins, lo, hi, vocab_size, (int64_t*)ids_out);
  else
    digitize_kernel<float><<<grid, 256, 0, S_(stream)>>>((const float*)x, n, (const double*)edges, bins, lo, hi, vocab_size, (int64_t*)ids_out);
  MLA_CHECK_LAUNCH("action_digitize");
  return MLA_OK;
}
extern "C" int mla_action_decode(const void* ids, int64_t n, const void* centers, int32_t n_centers, int64_t vocab_size,
                                 void* out, void* stream) {
  if (int rc = device_check()) return rc;
  if (n <= 0) return MLA_OK;
  action_decode_kernel<<<unsigned((n + 255) / 256), 256, 0, S_(stream)>>>((const int64_t*)ids, n, (const double*)centers,
                                                                          n_centers, vocab_size, (double*)out);
  MLA_CHECK_LAUNCH("action_decode");
  return MLA_OK;
}
