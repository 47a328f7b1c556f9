// Shared host-side helpers of libmla_b200: error slot, device check, launch counter, TMA descriptor encoding.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mla_b200.h"

namespace mla {

int set_error(int code, const char* fmt, ...);
int device_check();
int num_sms();
void count_launch(int n = 1);

// 2D bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first; strides[0] = byte pitch of dim 1.
int encode_tmap_2d_bf16(CUtensorMap* map, const void* ptr, const uint64_t dims[2], const uint64_t strides[1],
                        const uint32_t box[2]);
// 3D variant (dims/box innermost first; strides = byte pitches of dims 1 and 2).
int encode_tmap_3d_bf16(CUtensorMap* map, const void* ptr, const uint64_t dims[3], const uint64_t strides[2],
                        const uint32_t box[3]);

#define MLA_CHECK_LAUNCH(what)                                                                   \
  do {                                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) return mla::set_error(MLA_ERR_CUDA, what ": %s", cudaGetErrorString(e__)); \
    mla::count_launch();                                                                         \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return int((a + b - 1) / b); }

}  // namespace mla
