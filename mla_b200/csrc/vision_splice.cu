// Image tokenizer glue (patchify im2col, 3x3 window pooling, local attention core) and the multimodal sequence
// splice index.  The GEMMs in between run on gemm_sm100.cu.
//
// Row order trick: patch rows are emitted WINDOW-MAJOR — row = ((b*G + g)*9 + n) with g the 3x3 pooling window and
// n the patch inside it — so the avg-pool, the 9 keys/values of LocalAttention and its output are contiguous
// (the reference unfolds/permutes to the same grouping, models/mla/image/vision_tokenizer.py:35-38).
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

// pixels f32 [B, C_total, Himg, Wimg] (first 3 channels used) -> bf16 [B*G*9, k_pad]; column = c*P*P + ky*P + kx,
// matching nn.Conv2d(3, C, P, P).weight.view(C, -1)  (vision_tokenizer.py:112,:124).  One warp per output row.
__global__ void patchify_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ out, int B, int c_total,
                                int himg, int wimg, int P, int cs, int k_pad) {
  const int gw = wimg / (P * cs), gh = himg / (P * cs);
  const int64_t rows = int64_t(B) * gh * gw * cs * cs;
  const int64_t row = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int n = int(row % (cs * cs));
  const int64_t bg = row / (cs * cs);
  const int g = int(bg % (gh * gw));
  const int b = int(bg / (gh * gw));
  const int py = (g / gw) * cs + n / cs, pxx = (g % gw) * cs + n % cs;
  const int K = 3 * P * P;
  const float* base = px + (int64_t(b) * c_total) * himg * wimg + int64_t(py * P) * wimg + pxx * P;
  __nv_bfloat16* o = out + row * k_pad;
  for (int i = lane; i < k_pad; i += 32) {
    float v = 0.f;
    if (i < K) {
      const int c = i / (P * P), r = i % (P * P);
      v = base[int64_t(c) * himg * wimg + (r / P) * wimg + (r % P)];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}

// mean over each group of `win` consecutive rows: x bf16 [G*win, C] -> bf16 [G, C]  (F.avg_pool2d in bf16)
__global__ void window_mean_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t G,
                                   int C, int win) {
  const int64_t total = G * C;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t g = i / C;
    const int c = int(i % C);
    float acc = 0.f;
    for (int n = 0; n < win; ++n) acc += __bfloat162float(x[(g * win + n) * C + c]);
    out[i] = __float2bfloat16_rn(acc / float(win));
  }
}

// LocalAttention core (vision_tokenizer.py:40-45): per window g and head hh (dh = C/heads):
//   a_n = bf16( sum_d bf16(bf16(q_d*scale) * k_{n,d}) ), w = softmax_n(a) in fp32, out_d = sum_n w_n * v_{n,d}.
// q bf16 [G, C]; kv bf16 [G*win, 2C] (k | v); out bf16 [G, C].  One warp per (g, head).
__global__ void local_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                                  __nv_bfloat16* __restrict__ out, int64_t G, int C, int heads, int win, float scale) {
  const int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= G * heads) return;
  const int hh = int(w % heads);
  const int64_t g = w / heads;
  const int dh = C / heads;
  float a[16];
  for (int n = 0; n < win; ++n) {
    float acc = 0.f;
    for (int d = lane; d < dh; d += 32) {
      const float qs = bf16_round(__bfloat162float(q[g * C + hh * dh + d]) * scale);
      acc += bf16_round(qs * __bfloat162float(kv[(g * win + n) * 2 * C + hh * dh + d]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    a[n] = bf16_round(acc);
  }
  float mx = -INFINITY;
  for (int n = 0; n < win; ++n) mx = fmaxf(mx, a[n]);
  float den = 0.f;
  for (int n = 0; n < win; ++n) { a[n] = __expf(a[n] - mx); den += a[n]; }
  for (int d = lane; d < dh; d += 32) {
    float acc = 0.f;
    for (int n = 0; n < win; ++n) acc += (a[n] / den) * __bfloat162float(kv[(g * win + n) * 2 * C + C + hh * dh + d]);
    out[g * C + hh * dh + d] = __float2bfloat16_rn(acc);
  }
}

// Sequence splice index (models/vlm/prismatic.py:949-1042).  One CTA per sample.
//   z = [text0 | fused(F) | text1..]; with diffusion the n_ins rows [proprio | t | x0..xT] go in before the LAST
//   occurrence of `eos_id` in input_ids (:983).  Sources live in one row table: text rows at text_base + b*Lt + j,
//   fused rows at fused_base + b*F + j, inserted rows at ins_base + b*n_ins + j.
// Outputs: src_idx int32 [B,S], mask uint8 [B,S], labels int64 [B,S] (if labels_in), lti int32 [B] (position of the
// first inserted row), head_rows int32 [B, n_x]: flat rows (b*S + lti + 2 + j) of the noisy-action tokens (:1121-1124).
__global__ void splice_index_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ amask,
                                    const int64_t* __restrict__ labels_in, int Lt, int F, int n_ins, int n_x,
                                    int64_t eos_id, int text_base, int fused_base, int ins_base, int S,
                                    int32_t* __restrict__ src_idx, uint8_t* __restrict__ mask_out,
                                    int64_t* __restrict__ labels_out, int32_t* __restrict__ lti_out,
                                    int32_t* __restrict__ head_rows, int32_t* __restrict__ err_flag) {
  __shared__ int s_last;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_last = -1;
  __syncthreads();
  if (n_ins > 0)
    for (int j = threadIdx.x; j < Lt; j += blockDim.x)
      if (ids[int64_t(b) * Lt + j] == eos_id) atomicMax(&s_last, j);
  __syncthreads();
  int lti = S;  // no insertion
  if (n_ins > 0) {
    if (s_last < 0) {
      if (threadIdx.x == 0) atomicExch(err_flag, 1);  // the reference raises IndexError here
      lti = F + Lt - 1;
    } else {
      lti = s_last + F;
    }
  }
  if (threadIdx.x == 0) lti_out[b] = lti;
  for (int j = threadIdx.x; j < n_x; j += blockDim.x) head_rows[b * n_x + j] = b * S + lti + 2 + j;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    int src, m;
    int64_t lab = -100;
    if (s >= lti && s < lti + n_ins) {
      src = ins_base + b * n_ins + (s - lti);
      m = 1;
    } else {
      const int zi = s < lti ? s : s - n_ins;
      if (zi >= 1 && zi <= F) {
        src = fused_base + b * F + (zi - 1);
        m = 1;
      } else {
        const int j = zi == 0 ? 0 : zi - F;
        src = text_base + b * Lt + j;
        m = amask ? (amask[int64_t(b) * Lt + j] != 0) : 1;
        if (labels_in) lab = labels_in[int64_t(b) * Lt + j];
      }
    }
    src_idx[int64_t(b) * S + s] = src;
    mask_out[int64_t(b) * S + s] = uint8_t(m);
    if (labels_out) labels_out[int64_t(b) * S + s] = lab;
  }
}

// Shared-prefix variant of the splice: MLA.forward repeats every sample R times (model_mla.py:147-176) but
// only the [t | x0..xT] rows differ between the copies.  Sample b becomes ONE packed sequence
//   [ z[:lti] | proprio ]  (prefix, P_b = lti + 1 rows)  ++  R groups [ t_e | x_e,0..x_e,T | EOS ],  e = r*B + b,
// padded with masked filler rows to the common length S' = F + Lt + R*(n_x + 2).  The rows the reference puts behind the
// EOS (padding of a right-padded prompt) are masked keys and masked queries there and are simply left out.
// Row sources: text rows at text_base + b*Lt + j, fused rows at fused_base + b*F + j, proprio rows at pr_base + b,
// timestep rows at t_base + e, noisy-action rows at x_base + e*n_x + j.
// Outputs: src_idx int32 [B,S'], mask uint8 [B,S'], rope_pos int32 [B,S'] (groups repeat the positions P_b ..),
// prefix_len int32 [B], lti int32 [B*R], head_rows int32 [B*R, n_x] (flat packed rows of the x tokens of copy e).
__global__ void splice_index_shared_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ amask, int B, int Lt,
                                           int F, int n_x, int R, int64_t eos_id, int text_base, int fused_base,
                                           int pr_base, int t_base, int x_base, int Sp, int32_t* __restrict__ src_idx,
                                           uint8_t* __restrict__ mask_out, int32_t* __restrict__ rope_pos,
                                           int32_t* __restrict__ prefix_len, int32_t* __restrict__ lti_out,
                                           int32_t* __restrict__ head_rows, int32_t* __restrict__ err_flag) {
  __shared__ int s_last;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_last = -1;
  __syncthreads();
  for (int j = threadIdx.x; j < Lt; j += blockDim.x)
    if (ids[int64_t(b) * Lt + j] == eos_id) atomicMax(&s_last, j);
  __syncthreads();
  int last = s_last;
  if (last < 0) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);  // the reference raises IndexError here
    last = Lt - 1;
  }
  const int lti = last + F;       // position of the proprio row
  const int P = lti + 1, n = n_x + 2;
  if (threadIdx.x == 0) prefix_len[b] = P;
  for (int r = threadIdx.x; r < R; r += blockDim.x) lti_out[r * B + b] = lti;
  for (int i = threadIdx.x; i < R * n_x; i += blockDim.x) {
    const int r = i / n_x, j = i - r * n_x;
    head_rows[(r * B + b) * n_x + j] = b * Sp + P + r * n + 1 + j;
  }
  for (int s = threadIdx.x; s < Sp; s += blockDim.x) {
    int src, m = 1, pos = s;
    if (s < lti) {
      if (s >= 1 && s <= F) {
        src = fused_base + b * F + (s - 1);
      } else {
        const int j = s == 0 ? 0 : s - F;
        src = text_base + b * Lt + j;
        m = amask ? (amask[int64_t(b) * Lt + j] != 0) : 1;
      }
    } else if (s == lti) {
      src = pr_base + b;
    } else if (s < P + R * n) {
      const int q = s - P, r = q / n, k = q - r * n, e = r * B + b;
      pos = P + k;
      if (k == 0) src = t_base + e;
      else if (k <= n_x) src = x_base + e * n_x + (k - 1);
      else src = text_base + b * Lt + last;           // the EOS token's embedding
    } else {
      src = text_base + b * Lt;                       // filler row: masked key, masked query
      m = 0;
      pos = 0;
    }
    src_idx[int64_t(b) * Sp + s] = src;
    mask_out[int64_t(b) * Sp + s] = uint8_t(m);
    rope_pos[int64_t(b) * Sp + s] = pos;
  }
}

}  // namespace mla

using namespace mla;

extern "C" int mla_splice_index_shared(const void* input_ids, const void* attn_mask, int32_t batch, int32_t lt,
                                       int32_t n_fused, int32_t n_x, int32_t repeats, int64_t eos_id, int32_t text_base,
                                       int32_t fused_base, int32_t pr_base, int32_t t_base, int32_t x_base, void* src_idx,
                                       void* mask_out, void* rope_pos, void* prefix_len, void* lti_out, void* head_rows,
                                       void* err_flag, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0) return MLA_OK;
  if (lt < 1 || n_fused < 0 || n_x < 1 || repeats < 1) return set_error(MLA_ERR_ARG, "splice_index_shared: bad sizes");
  const int Sp = n_fused + lt + repeats * (n_x + 2);
  splice_index_shared_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(
      (const int64_t*)input_ids, (const uint8_t*)attn_mask, batch, lt, n_fused, n_x, repeats, eos_id, text_base, fused_base,
      pr_base, t_base, x_base, Sp, (int32_t*)src_idx, (uint8_t*)mask_out, (int32_t*)rope_pos, (int32_t*)prefix_len,
      (int32_t*)lti_out, (int32_t*)head_rows, (int32_t*)err_flag);
  MLA_CHECK_LAUNCH("splice_index_shared");
  return MLA_OK;
}

extern "C" int mla_patchify(const void* pixels, void* out, int32_t batch, int32_t c_total, int32_t himg, int32_t wimg,
                            int32_t patch, int32_t conv_stride, int32_t k_pad, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0) return MLA_OK;
  if (himg % (patch * conv_stride) || wimg % (patch * conv_stride))
    return set_error(MLA_ERR_ARG, "patchify: image %dx%d not divisible by patch*conv_stride=%d", himg, wimg, patch * conv_stride);
  if (c_total < 3 || k_pad < 3 * patch * patch || (k_pad & 7)) return set_error(MLA_ERR_ARG, "patchify: bad channel count or k_pad");
  const int64_t rows = int64_t(batch) * (himg / patch) * (wimg / patch);
  patchify_kernel<<<unsigned((rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float*)pixels, (__nv_bfloat16*)out, batch, c_total, himg, wimg, patch, conv_stride, k_pad);
  MLA_CHECK_LAUNCH("patchify");
  return MLA_OK;
}

extern "C" int mla_window_mean(const void* x, void* out, int64_t groups, int32_t c, int32_t win, void* stream) {
  if (int rc = device_check()) return rc;
  if (groups <= 0) return MLA_OK;
  int64_t total = groups * c;
  int grid = int((total + 255) / 256 < int64_t(num_sms()) * 16 ? (total + 255) / 256 : int64_t(num_sms()) * 16);
  window_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, groups, c, win);
  MLA_CHECK_LAUNCH("window_mean");
  return MLA_OK;
}

extern "C" int mla_local_attn(const void* q, const void* kv, void* out, int64_t groups, int32_t c, int32_t heads,
                              int32_t win, float scale, void* stream) {
  if (int rc = device_check()) return rc;
  if (groups <= 0) return MLA_OK;
  if (win > 16 || c % heads) return set_error(MLA_ERR_ARG, "local_attn: window > 16 or channels not divisible by heads");
  const int64_t warps = groups * heads;
  local_attn_kernel<<<unsigned((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)kv, (__nv_bfloat16*)out, groups, c, heads, win, scale);
  MLA_CHECK_LAUNCH("local_attn");
  return MLA_OK;
}

extern "C" int mla_splice_index(const void* input_ids, const void* attn_mask, const void* labels, int32_t batch,
                                int32_t lt, int32_t n_fused, int32_t n_ins, int32_t n_x, int64_t eos_id,
                                int32_t text_base, int32_t fused_base, int32_t ins_base, void* src_idx, void* mask_out,
                                void* labels_out, void* lti_out, void* head_rows, void* err_flag, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0) return MLA_OK;
  if (lt < 1 || n_fused < 0 || n_ins < 0) return set_error(MLA_ERR_ARG, "splice_index: bad sizes");
  const int S = n_fused + lt + n_ins;
  splice_index_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(
      (const int64_t*)input_ids, (const uint8_t*)attn_mask, (const int64_t*)labels, lt, n_fused, n_ins, n_x, eos_id,
      text_base, fused_base, ins_base, S, (int32_t*)src_idx, (uint8_t*)mask_out, (int64_t*)labels_out,
      (int32_t*)lti_out, (int32_t*)head_rows, (int32_t*)err_flag);
  MLA_CHECK_LAUNCH("splice_index");
  return MLA_OK;
}
