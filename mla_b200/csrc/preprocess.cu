// Data side of the image path: the CLIP preprocessing the reference runs on the CPU for every camera frame
// (vla/datasets/datasets.py:53-69 and models/mla/model_mla.py:661-665: CLIPImageProcessor(672) = PIL bicubic resize
// + rescale 1/255 + CLIP mean/std normalisation, then an all-ones mask channel), as integer work on the GPU.
//
// PIL's resize (libImaging/Resample.c) is a separable convolution on uint8 with 22-bit fixed-point coefficients: a
// horizontal pass then a vertical pass, each accumulating in int32 from 1 << 21, shifting right by 22 and clamping to
// [0, 255].  The coefficient tables are built on the host exactly as PIL does (mla_b200/preprocess.py) and passed in:
//   tab [out, 2 + ksize] int32 = (first input index, tap count, taps...).
// uint8 -> normalised f32 is a 3 x 256 lookup table computed on the host in the reference's arithmetic (float64
// rescale, float32 normalise), so the result is BIT-EXACT with the reference's tensor by construction.
//   * clip_preprocess_kernel: u8 [B,H,W,3] -> f32 [B,4,S,S] (what the model's batch contract carries today)
//   * patchify_u8_kernel:     u8 [B,H,W,3] -> bf16 im2col rows of the 14x14 patch embedding directly (same rows as
//     mla_patchify on the f32 tensor, bit for bit) — the 7.2 MB/sample f32 image never exists.
#include "mla_internal.cuh"
#include "ptx.cuh"

namespace mla {

constexpr int PRE_BITS = 22;

__device__ __forceinline__ int pre_clip8(int v) {
  v >>= PRE_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// resized uint8 value at output (y, x), channel c of one HxWx3 frame
__device__ __forceinline__ int resized_px(const uint8_t* __restrict__ img, int W, const int32_t* __restrict__ tab_h,
                                          const int32_t* __restrict__ tab_v, int ld_h, int ld_v, int y, int x, int c) {
  const int32_t* th = tab_h + int64_t(x) * ld_h;
  const int32_t* tv = tab_v + int64_t(y) * ld_v;
  const int x0 = th[0], nx = th[1], y0 = tv[0], ny = tv[1];
  int acc_v = 1 << (PRE_BITS - 1);
  for (int j = 0; j < ny; ++j) {
    const uint8_t* row = img + (int64_t(y0 + j) * W + x0) * 3 + c;
    int acc_h = 1 << (PRE_BITS - 1);
    for (int i = 0; i < nx; ++i) acc_h += int(row[i * 3]) * th[2 + i];
    acc_v += pre_clip8(acc_h) * tv[2 + j];
  }
  return pre_clip8(acc_v);
}

__global__ void clip_preprocess_kernel(const uint8_t* __restrict__ img, const int32_t* __restrict__ tab_h,
                                       const int32_t* __restrict__ tab_v, const float* __restrict__ lut,
                                       float* __restrict__ out, int B, int H, int W, int S, int ld_h, int ld_v,
                                       int out_channels) {
  const int64_t total = int64_t(B) * out_channels * S * S;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int x = int(idx % S), y = int((idx / S) % S);
    const int c = int((idx / (int64_t(S) * S)) % out_channels);
    const int b = int(idx / (int64_t(S) * S * out_channels));
    float v = 1.f;                                             // mask channel
    if (c < 3) v = lut[c * 256 + resized_px(img + int64_t(b) * H * W * 3, W, tab_h, tab_v, ld_h, ld_v, y, x, c)];
    out[idx] = v;
  }
}

// Same row / column order as patchify_kernel (vision_splice.cu): row = ((b*G + g)*cs^2 + n), column = c*P*P + ky*P + kx.
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, const int32_t* __restrict__ tab_h,
                                   const int32_t* __restrict__ tab_v, const float* __restrict__ lut,
                                   __nv_bfloat16* __restrict__ out, int B, int H, int W, int S, int ld_h, int ld_v, int P,
                                   int cs, int k_pad) {
  const int gw = S / (P * cs), gh = S / (P * cs);
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
  const uint8_t* frame = img + int64_t(b) * H * W * 3;
  __nv_bfloat16* o = out + row * k_pad;
  for (int i = lane; i < k_pad; i += 32) {
    float v = 0.f;
    if (i < K) {
      const int c = i / (P * P), r = i % (P * P);
      v = lut[c * 256 + resized_px(frame, W, tab_h, tab_v, ld_h, ld_v, py * P + r / P, pxx * P + r % P, c)];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace mla

using namespace mla;
#define S_(x) ((cudaStream_t)(x))

static int pre_check(int32_t batch, int32_t h, int32_t w, int32_t size, int32_t taps_h, int32_t taps_v) {
  if (h <= 0 || w <= 0 || size <= 0) return set_error(MLA_ERR_ARG, "clip_preprocess: empty image");
  if (taps_h < 1 || taps_v < 1 || taps_h > 64 || taps_v > 64)
    return set_error(MLA_ERR_ARG, "clip_preprocess: tap counts must be in [1, 64]");
  (void)batch;
  return MLA_OK;
}

extern "C" int mla_clip_preprocess(const void* frames_u8, const void* tab_h, const void* tab_v, const void* lut,
                                   void* out, int32_t batch, int32_t h, int32_t w, int32_t size, int32_t taps_h,
                                   int32_t taps_v, int32_t out_channels, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0) return MLA_OK;
  if (int rc = pre_check(batch, h, w, size, taps_h, taps_v)) return rc;
  if (out_channels != 3 && out_channels != 4) return set_error(MLA_ERR_ARG, "clip_preprocess: out_channels must be 3 or 4");
  const int64_t total = int64_t(batch) * out_channels * size * size;
  const int64_t blocks = (total + 255) / 256;
  const int grid = int(blocks < int64_t(num_sms()) * 32 ? blocks : int64_t(num_sms()) * 32);
  clip_preprocess_kernel<<<grid, 256, 0, S_(stream)>>>((const uint8_t*)frames_u8, (const int32_t*)tab_h,
                                                      (const int32_t*)tab_v, (const float*)lut, (float*)out, batch, h, w,
                                                      size, 2 + taps_h, 2 + taps_v, out_channels);
  MLA_CHECK_LAUNCH("clip_preprocess");
  return MLA_OK;
}

extern "C" int mla_patchify_u8(const void* frames_u8, const void* tab_h, const void* tab_v, const void* lut, void* out,
                               int32_t batch, int32_t h, int32_t w, int32_t size, int32_t taps_h, int32_t taps_v,
                               int32_t patch, int32_t conv_stride, int32_t k_pad, void* stream) {
  if (int rc = device_check()) return rc;
  if (batch <= 0) return MLA_OK;
  if (int rc = pre_check(batch, h, w, size, taps_h, taps_v)) return rc;
  if (size % (patch * conv_stride))
    return set_error(MLA_ERR_ARG, "patchify_u8: size %d not divisible by patch*conv_stride=%d", size, patch * conv_stride);
  if (k_pad < 3 * patch * patch || (k_pad & 7)) return set_error(MLA_ERR_ARG, "patchify_u8: bad k_pad");
  const int64_t rows = int64_t(batch) * (size / patch) * (size / patch);
  patchify_u8_kernel<<<unsigned((rows * 32 + 255) / 256), 256, 0, S_(stream)>>>(
      (const uint8_t*)frames_u8, (const int32_t*)tab_h, (const int32_t*)tab_v, (const float*)lut, (__nv_bfloat16*)out,
      batch, h, w, size, 2 + taps_h, 2 + taps_v, patch, conv_stride, k_pad);
  MLA_CHECK_LAUNCH("patchify_u8");
  return MLA_OK;
}
