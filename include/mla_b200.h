/*
 * mla_b200.h — C ABI of libmla_b200.so: the sm_100a kernels behind the MLA training-step hot path.
 *
 * The reference (ZhuoyangLiu2005/MLA) has no FFI for this path: its hot ops are PyTorch library calls
 * (cuBLAS / cuDNN / ATen / flash-attn).  Each entry point below names the reference call site it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; row-major, 16-byte aligned;
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue work and return, they never synchronise;
 *   - return value: 0 on success, negative mla_status otherwise; mla_last_error() gives the message
 *     (thread-local); the library never allocates device memory and keeps no global device state;
 *   - bf16 everywhere unless stated; "f32" = IEEE binary32; indices are int32 unless stated;
 *   - rounding points follow the reference's bf16 autocast path (each PyTorch op rounds its result to bf16).
 */
#ifndef MLA_B200_H
#define MLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mla_status {
  MLA_OK = 0,
  MLA_ERR_ARG = -1,     /* bad argument (shape, alignment, null pointer) */
  MLA_ERR_CUDA = -2,    /* a CUDA runtime/driver call failed */
  MLA_ERR_DEVICE = -3,  /* not an sm_100 device */
} mla_status;

enum { MLA_ACT_NONE = 0, MLA_ACT_RELU = 1, MLA_ACT_GELU_ERF = 2, MLA_ACT_GELU_TANH = 3, MLA_ACT_SILU = 4 };

/* ---- library ------------------------------------------------------------------------------------------- */
const char* mla_version(void);
const char* mla_last_error(void);
/* 0 if the current CUDA device is sm_100; MLA_ERR_DEVICE otherwise (also when no device is present). */
int mla_device_check(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches counter). */
int64_t mla_launch_count(void);

/* ---- GEMM (tcgen05 + TMA) --------------------------------------------------------------------------------
 * C[M,N] = epilogue(alpha * A_op[M,K] . B_op[K,N]).
 *   a_mn_major = 0: A stored [M,K], K contiguous (lda = row pitch);  1: A stored [K,M], M contiguous.
 *   b_mn_major = 0: B stored [N,K], K contiguous (an nn.Linear weight); 1: B stored [K,N], N contiguous.
 * Epilogue (bf16 output): v = bf16(acc*alpha + bias); pre_act <- v; v = bf16(act(v)); v = bf16(v + residual).
 * fp32 output: C = acc*alpha (+ C if accumulate) — the weight-gradient path.
 * Replaces: every nn.Linear on the path — modeling_llama.py:240 (gate/up/down), :435-437,:495 (q/k/v/o),
 * models/mla/image/vision_tokenizer.py:21-25,:82-87,:112, util/nn_utils.py:25-31, fuser/contrastive.py:173-182,
 * models/diffusion/models.py:34-38 — and their autograd backward (dgrad / wgrad). */
typedef struct mla_gemm_args {
  const void* a;
  const void* b;
  void* c;
  int64_t m, n, k;
  int64_t lda, ldb, ldc;       /* in elements */
  int32_t a_mn_major, b_mn_major;
  int32_t c_dtype;             /* 0 = bf16, 1 = f32 */
  int32_t accumulate;          /* f32 output only */
  int32_t activation;          /* MLA_ACT_* */
  float alpha;
  const void* bias;            /* bf16 [N] or NULL */
  const void* residual;        /* bf16 [M,N] or NULL */
  int64_t ldr;
  void* pre_act;               /* bf16 [M,N] or NULL: value before the activation (saved for backward) */
  int64_t ldp;
} mla_gemm_args;
int mla_gemm_bf16(const mla_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLA_B200_H */
