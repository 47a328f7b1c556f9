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
/* TMA descriptors (CUtensorMap) are cached per host thread by (pointer, dims, strides, box): hits / driver encodes so far. */
void mla_tmap_cache_stats(int64_t* hits, int64_t* misses);

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
  void* sched_ws;              /* NULL: static tile schedule.  Else 8 bytes of device memory, zero before the first use and
                                  private to one stream: tiles are then claimed dynamically (CTAs that start late because a
                                  concurrent kernel, e.g. an NCCL all-reduce, holds their SM take fewer tiles); the kernel
                                  leaves it zeroed again. */
  /* Fused RoPE for the q|k|v projection (head_dim 128; modeling_llama.py:184-208 folded into the epilogue of :435-437):
   * the leading rope_cols output columns (a multiple of 256 = whole q and k heads) are rotated with position =
   * row % rope_seq; rope_cos / rope_sin: bf16 [rope_seq, 64].  rope_cols = 0 switches it off.  Plain bf16 outputs only. */
  const void* rope_cos;
  const void* rope_sin;
  int32_t rope_seq;
  int32_t rope_cols;
  /* Fused SwiGLU for the gate|up projection (LlamaMLP, modeling_llama.py:240; CTA-pair kernel): B = [gate; up] stored
   * [2f, K], f a multiple of 128.  swiglu_out bf16 [M, f] (pitch ld_swiglu) receives bf16(bf16(silu(gate)) * up); c
   * (the [M, 2f] gate|up matrix backward wants) is then optional — NULL skips its store.  NULL swiglu_out = off.
   * Bit-identical to the projection followed by mla_swiglu_fwd (tests/test_gemm2_gpu.py); the training step uses it by
   * default (MLA_FUSE_SWIGLU=0 restores the separate pass). */
  void* swiglu_out;
  int64_t ld_swiglu;
  /* Fused SwiGLU BACKWARD for the input-gradient GEMM of the down projection (autograd of modeling_llama.py:240):
   * with A = dy [M, h] and B = W_down, N = f (a multiple of 32), the epilogue rounds d_act = dy . W_down to bf16, reads
   * gate|up from swiglu_bwd_gu (bf16 [M, 2f]) and writes d(gate|up) to swiglu_bwd_dgu (bf16 [M, 2f]) and, when
   * swiglu_bwd_act is not NULL, the re-materialised act = bf16(bf16(silu(gate)) * up) (bf16 [M, f], the operand of the
   * down projection's weight gradient).  c is not written.  Bit-identical to the GEMM followed by mla_swiglu_bwd_act.
   * NULL swiglu_bwd_gu = off. */
  /* Optional row -> rotary position table for the fused RoPE (int32 [M]; NULL: position = row % rope_seq): the
   * shared-prefix layout repeats positions across the suffix groups of a sequence. */
  const void* rope_pos;
  const void* swiglu_bwd_gu;
  int64_t ld_swiglu_bwd_gu;
  void* swiglu_bwd_dgu;
  int64_t ld_swiglu_bwd_dgu;
  void* swiglu_bwd_act;
  int64_t ld_swiglu_bwd_act;
  /* fp32 outputs (weight gradients) only: *sumsq (f32, device) += sum of squares of the values WRITTEN (after the optional
   * accumulate) — the global gradient norm of clip_grad_norm_ (training/strategies/fsdp.py:310) comes out of the
   * weight-gradient GEMMs instead of a second pass over 27.8 GB of gradients.  NULL = off. */
  void* sumsq;
} mla_gemm_args;
int mla_gemm_bf16(const mla_gemm_args* args, void* stream);
/* Kernel selection for mla_gemm_bf16: 0 = one CTA per 128x256 tile, 1 = CTA pairs (tcgen05.mma.cta_group::2, 256x256
 * tile per 2-CTA cluster) for problems of at least 1024 rows (default), 2 = CTA pairs always.  Env MLA_GEMM_2CTA sets
 * the initial mode.  Both kernels honour sched_ws (dynamic tile claiming). */
int mla_gemm_set_mode(int32_t mode);
/* Tuning switch: M-tiles per rasterisation group of the CTA-pair kernel (0 = built-in heuristic from the K-panel size;
 * env MLA_GEMM2_GROUP_M sets the initial value).  Results do not depend on it. */
int mla_gemm_set_group_m(int32_t group_m);

/* ---- RMSNorm ----------------------------------------------------------------------------------------------
 * y = w * bf16(x * rsqrt(mean(x^2) + eps))   (modeling_llama.py:85-90, LlamaRMSNorm; x,y,w bf16, stats f32).
 * mode 1 divides by the unbiased variance instead (timm 0.9.x RmsNorm hazard) — used only by
 * FinalLayer.norm_final (models/diffusion/models.py:179) when that arithmetic is selected.
 * rstd (f32 [rows]) may be NULL.  bwd: dx = rmsnorm'(dy) (+ dres if non-NULL), dw (f32 [h]) += sum_rows dy * n. */
int mla_rmsnorm_fwd(const void* x, const void* w, void* y, void* rstd, int64_t rows, int32_t h, int64_t ldx,
                    int64_t ldy, float eps, int32_t mode, void* stream);
int mla_rmsnorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, void* dw,
                    int64_t rows, int32_t h, float eps, void* stream);

/* ---- RoPE ------------------------------------------------------------------------------------------------
 * In-place rotation of `heads` consecutive heads of width d in each row of a [tokens, ld] buffer
 * (modeling_llama.py:184-208, apply_rotary_pos_emb on the bf16 path; position = token index mod seq, as
 * position_ids = arange(S) for every sample, :985-990).  cos_t/sin_t: bf16 [seq, d/2] tables.
 * transpose != 0 applies the inverse rotation (the backward). */
int mla_rope_inplace(void* base, const void* cos_t, const void* sin_t, int64_t tokens, int32_t seq, int32_t heads,
                     int32_t d, int64_t ld, int32_t transpose, void* stream);

/* ---- SwiGLU ----------------------------------------------------------------------------------------------
 * gu = [rows, 2f] (gate | up).  out[rows,f] = bf16(silu(gate)) * up   (modeling_llama.py:240).
 * bwd: dgu[rows,2f] from dact[rows,f] and gu. */
int mla_swiglu_fwd(const void* gu, void* out, int64_t rows, int32_t f, void* stream);
int mla_swiglu_bwd(const void* dact, const void* gu, void* dgu, int64_t rows, int32_t f, void* stream);
/* swiglu_bwd that also re-materialises act = bf16(silu(gate)) * up into act_out [rows, f] (the operand of the
 * down-projection weight gradient), saving the separate swiglu_fwd recompute pass of backward. */
int mla_swiglu_bwd_act(const void* dact, const void* gu, void* dgu, void* act_out, int64_t rows, int32_t f, void* stream);

/* ---- causal attention ------------------------------------------------------------------------------------
 * Replaces flash_attn_func / flash_attn_varlen_func + unpad/pad_input (modeling_llama.py:540-557).
 * q,k,v: row t = b*seq+s, head hd at column hd*head_dim, row pitch ld_qkv (so the fused QKV projection buffer is
 * consumed in place); o/d_o: [batch*seq, heads*head_dim] pitch ld_o; lse/delta: f32 [batch, heads, seq];
 * mask: uint8 [batch, seq] (attention_mask) or NULL.  Key j visible to query i iff j <= i and mask[b,j];
 * rows with mask 0 output zeros and get zero gradient.  head_dim in {32, 64, 128}. */
typedef struct mla_attn_args {
  const void *q, *k, *v;
  int64_t ld_qkv;
  void* o;
  int64_t ld_o;
  void* lse;
  const void* mask;
  int32_t batch, seq, heads, head_dim;
  float scale;
  /* backward only */
  const void* d_o;
  void* delta;
  void *dq, *dk, *dv;
  int64_t ld_dqkv;
} mla_attn_args;
int mla_attn_fwd(const mla_attn_args* args, void* stream);
int mla_attn_bwd(const mla_attn_args* args, void* stream);
/* tcgen05/TMEM/TMA forward for head_dim 128: qkv is the fused [batch*seq, 3*heads*128] projection (q | k | v column
 * blocks, pitch ld_qkv); same outputs and semantics as mla_attn_fwd. */
int mla_attn_fwd_sm100(const void* qkv, int64_t ld_qkv, void* o, int64_t ld_o, void* lse, const void* mask,
                       int32_t batch, int32_t seq, int32_t heads, float scale, void* stream);
/* Shared-prefix attention (models/mla/model_mla.py:147-176 repeats every sample R times although only the
 * [t | x | EOS] rows differ): sequence b = a causal prefix of prefix_len[b] rows followed by groups of `group` rows; a row
 * of a group sees the whole prefix and, causally, its own group.  prefix_len: int32 [batch] on the device. */
int mla_attn_fwd_sm100_grouped(const void* qkv, int64_t ld_qkv, void* o, int64_t ld_o, void* lse, const void* mask,
                               const void* prefix_len, int32_t group, int32_t batch, int32_t seq, int32_t heads,
                               float scale, void* stream);
/* tcgen05/TMEM/TMA backward for head_dim 128 (dK/dV kernel + dQ kernel): dqkv has the layout of qkv (dq | dk | dv);
 * lse is the forward's [batch, heads, seq] output; workspace: mla_attn_bwd_sm100_workspace() bytes, 16-byte aligned. */
size_t mla_attn_bwd_sm100_workspace(int32_t batch, int32_t seq, int32_t heads);
int mla_attn_bwd_sm100(const void* qkv, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o, const void* lse,
                       const void* mask, void* dqkv, int64_t ld_dqkv, void* workspace, int32_t batch, int32_t seq,
                       int32_t heads, float scale, void* stream);
/* Pipelined generation of the same backward (two math-warp groups ping-pong over the streamed tiles, 3-slot TMA ring,
 * double-buffered P/dS staging): same contract and workspace.  rope_cos / rope_sin (bf16 [seq, 64], or both NULL): when
 * given, the transposed rotary embedding — the backward of apply_rotary_pos_emb, modeling_llama.py:184-208 — is applied
 * to the dq and dk column blocks in the epilogue, so dqkv holds the gradient w.r.t. the projection outputs BEFORE RoPE
 * (bit-identical to mla_attn_bwd_sm100 followed by mla_rope_inplace(transpose = 1)). */
int mla_attn_bwd2_sm100(const void* qkv, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o, const void* lse,
                        const void* mask, void* dqkv, int64_t ld_dqkv, void* workspace, const void* rope_cos,
                        const void* rope_sin, int32_t batch, int32_t seq, int32_t heads, float scale, void* stream);
/* Backward of mla_attn_fwd_sm100_grouped; rope_pos int32 [batch*seq] (or NULL) = rotary position of every row for the
 * fused RoPE transpose. */
int mla_attn_bwd2_sm100_grouped(const void* qkv, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o,
                                const void* lse, const void* mask, void* dqkv, int64_t ld_dqkv, void* workspace,
                                const void* rope_cos, const void* rope_sin, const void* rope_pos, const void* prefix_len,
                                int32_t group, int32_t batch, int32_t seq, int32_t heads, float scale, void* stream);
/* How mla_attn_bwd2_sm100 hands P / dS to its accumulation MMAs: 1 (default; env MLA_ATTN_BWD_TS) = packed bf16 in
 * tensor memory, read as the A operand straight from TMEM (no shared-memory staging: the kernel is bound by
 * shared-memory operand bandwidth); 0 = staged in 128B-swizzled shared memory.  Same results bit for bit. */
int mla_attn_bwd2_set_ts(int32_t on);

/* ---- small row/elementwise kernels around the GEMMs ---------------------------------------------------------
 * (ATen glue in the reference: dtype casts under autocast, activation backward, bias gradients, torch.cat / index
 *  splices in models/vlm/prismatic.py:949-1040, nn.Embedding, nn.LayerNorm in vision_tokenizer.py:21-24.) */
int mla_cast_f32_bf16(const void* src, void* dst, int64_t n, void* stream);
/* dst bf16 [rows, cols] <- src f32 [rows_src, cols_src] (pitch lds), zero padded */
int mla_cast_pad_f32_bf16(const void* src, void* dst, int64_t rows, int64_t cols, int64_t rows_src, int64_t cols_src,
                          int64_t lds, void* stream);
/* dx = bf16(dy * act'(pre)) */
int mla_act_bwd(const void* dy, const void* pre, void* dx, int64_t n, int32_t act, void* stream);
/* out f32 [cols] += column sums of x bf16 [rows, cols] (pitch ld) */
int mla_colsum_bf16(const void* x, void* out, int64_t rows, int32_t cols, int64_t ld, void* stream);
/* dst[i,:] = idx[i] >= 0 ? src[idx[i],:] : 0 ; idx int32 (idx_is_i64 = 0) or int64 */
int mla_gather_rows(const void* src, const void* idx, void* dst, int64_t n, int32_t h, int64_t lds,
                    int32_t idx_is_i64, void* stream);
/* dsrc[idx[i],:] = ddst[i,:] for idx[i] >= 0 (idx injective) */
int mla_scatter_rows(void* dsrc, const void* idx, const void* ddst, int64_t n, int32_t h, int64_t lds, void* stream);
/* grad f32 [V,h]: grad[ids[i],:] += dout[i,:] unless ids[i] == pad_id (nn.Embedding backward) */
int mla_embedding_bwd(void* grad, const void* ids, const void* dout, int64_t n, int32_t h, int64_t pad_id,
                      void* stream);
int mla_layernorm_fwd(const void* x, const void* w, const void* b, void* y, int64_t rows, int32_t h, float eps,
                      void* stream);
/* x_t = sqrt_ac[t]*a + sqrt_1mac[t]*noise (models/diffusion/gaussian_diffusion.py:214-229); f32, t int64 [n/per_sample] */
int mla_q_sample(const void* a, const void* noise, const void* t, const void* sqrt_ac, const void* sqrt_1mac,
                 void* out, int64_t n, int32_t per_sample, void* stream);
/* loss[0] = mean((pred - target)^2), pred bf16, target f32 (models/mla/model_mla.py:215); bwd: dpred = g*2(pred-target)/n */
int mla_mse_fwd(const void* pred, const void* target, void* loss, int64_t n, void* stream);
int mla_mse_bwd(const void* pred, const void* target, const void* gscale, void* dpred, int64_t n, void* stream);
int mla_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);

/* ---- image tokenizer (models/mla/image/vision_tokenizer.py) ----------------------------------------------------
 * mla_patchify: im2col of nn.Conv2d(3,C,patch,patch) (:112,:124): pixels f32 [B,c_total,H,W] (first 3 channels) ->
 *   bf16 [B*(H/patch)*(W/patch), k_pad], rows window-major ((b*G+g)*cs^2 + n), column = c*patch^2 + ky*patch + kx.
 * mla_window_mean: F.avg_pool2d(features, cs, cs) (:28) on window-major rows: [groups*win, c] -> [groups, c].
 * mla_local_attn: the 9-way attention of LocalAttention.forward (:40-45): q [groups,c], kv [groups*win, 2c]. */
int mla_patchify(const void* pixels, void* out, int32_t batch, int32_t c_total, int32_t himg, int32_t wimg,
                 int32_t patch, int32_t conv_stride, int32_t k_pad, void* stream);
int mla_window_mean(const void* x, void* out, int64_t groups, int32_t c, int32_t win, void* stream);
int mla_local_attn(const void* q, const void* kv, void* out, int64_t groups, int32_t c, int32_t heads, int32_t win,
                   float scale, void* stream);

/* ---- image preprocessing on uint8 camera frames (data side: vla/datasets/datasets.py:53-69, model_mla.py:661-665) ----
 * CLIPImageProcessor(672) = PIL bicubic resize + rescale 1/255 + CLIP mean/std, + all-ones mask channel, bit-exact.
 * frames u8 [batch, h, w, 3]; tab_h int32 [size, 2+taps_h], tab_v int32 [size, 2+taps_v] = per output column / row the
 * first input index, the tap count and PIL's 22-bit fixed-point bicubic taps; lut f32 [3,256] = uint8 -> normalised
 * value per channel (host-built in the reference's float64/float32 arithmetic).
 * clip_preprocess: out f32 [batch, out_channels (3, or 4 with the ones mask), size, size].
 * patchify_u8: the same pixels written straight as bf16 im2col rows (layout of mla_patchify). */
int mla_clip_preprocess(const void* frames_u8, const void* tab_h, const void* tab_v, const void* lut, void* out,
                        int32_t batch, int32_t h, int32_t w, int32_t size, int32_t taps_h, int32_t taps_v,
                        int32_t out_channels, void* stream);
int mla_patchify_u8(const void* frames_u8, const void* tab_h, const void* tab_v, const void* lut, void* out, int32_t batch,
                    int32_t h, int32_t w, int32_t size, int32_t taps_h, int32_t taps_v, int32_t patch,
                    int32_t conv_stride, int32_t k_pad, void* stream);

/* ---- multimodal sequence splice (models/vlm/prismatic.py:949-1042,:1121-1124) -----------------------------------
 * Builds, without host syncs, the row map of [BOS | fused(n_fused) | text.. | proprio,t,x.. | EOS..] per sample:
 * src_idx int32 [B,S] into a row table (text at text_base+b*lt+j, fused at fused_base+b*n_fused+j, inserted rows at
 * ins_base+b*n_ins+j), the fused attention mask uint8 [B,S], labels int64 [B,S] (optional), lti int32 [B] and the
 * flat rows of the n_x noisy-action tokens (head_rows int32 [B,n_x]).  S = n_fused + lt + n_ins.  err_flag (int32)
 * is set to 1 if a sample has no `eos_id` (the reference raises IndexError at :983). */
int mla_splice_index(const void* input_ids, const void* attn_mask, const void* labels, int32_t batch, int32_t lt,
                     int32_t n_fused, int32_t n_ins, int32_t n_x, int64_t eos_id, int32_t text_base,
                     int32_t fused_base, int32_t ins_base, void* src_idx, void* mask_out, void* labels_out,
                     void* lti_out, void* head_rows, void* err_flag, void* stream);
/* Shared-prefix splice (models/mla/model_mla.py:147-176 repeats each sample `repeats` times although only
 * the [t | x] rows differ): sample b -> one packed sequence [prefix (prefix_len[b] rows) | repeats x (t | x.. | EOS)] of
 * S' = n_fused + lt + repeats * (n_x + 2) rows (filler rows masked).  Row sources are offsets into one row table
 * (text | fused | proprio [batch] | t [batch*repeats] | x [batch*repeats, n_x]); copy e = r * batch + b as `.repeat(R, ..)`
 * orders them.  Outputs: src_idx, mask, rope_pos [batch, S'], prefix_len [batch], lti [batch*repeats], head_rows
 * [batch*repeats, n_x] (flat packed rows of the noisy-action tokens, prismatic.py:1121-1124). */
int mla_splice_index_shared(const void* input_ids, const void* attn_mask, int32_t batch, int32_t lt, int32_t n_fused,
                            int32_t n_x, int32_t repeats, int64_t eos_id, int32_t text_base, int32_t fused_base,
                            int32_t pr_base, int32_t t_base, int32_t x_base, void* src_idx, void* mask_out, void* rope_pos,
                            void* prefix_len, void* lti_out, void* head_rows, void* err_flag, void* stream);

/* ---- InfoNCE alignment losses (models/mla/fuser/contrastive.py) and patch correspondence ---------------------------
 * l2norm: F.normalize(p=2, dim=-1) (:192-193), fp32 math, bf16 out, norms f32 [rows] saved for backward.
 * infonce: CoordinateAwareContrastiveLoss.forward (:196-215) on a materialised bf16 similarity matrix sim [n,n]
 *   (sim = a . b^T from mla_gemm_bf16); rows/cols with valid[i] == 0 are excluded (the reference compacts them away);
 *   out[0] = (CE(rows) + CE(cols)) / 2, out[1] = M (#valid).  bwd overwrites sim with d(loss)/d(sim) (bf16).
 * tac_nce: TactileContrastiveLoss (:241-258), one (query, 256 keys, positive index) problem per (batch, arm);
 *   loss_rows f32 [batch*arms] (mean them), probs f32 [batch*arms, k]; bwd: dq f32 [batch*arms,d], dkeys f32 += .
 * project_points: project_3d_to_2d_672_* (:5-131): xyz f32 [n,3]; cam f32 [21] = R(9) | t(3) | K(9);
 *   patch_idx int64 [n,2] = (row, col) clamped, valid uint8 [n]. */
int mla_l2norm_fwd(const void* x, void* y, void* norms, int64_t rows, int32_t d, int64_t ldx, float eps, void* stream);
int mla_l2norm_bwd(const void* x, const void* norms, const void* dy, void* dx, int64_t rows, int32_t d, int64_t ldx,
                   void* stream);
size_t mla_infonce_workspace(int32_t n);
int mla_infonce_fwd(const void* sim, const void* valid, void* workspace, void* out, int32_t n, float temperature,
                    void* stream);
int mla_infonce_bwd(void* sim_inout, const void* valid, const void* workspace, const void* out, const void* gscale,
                    int32_t n, float temperature, void* stream);
int mla_tac_nce_fwd(const void* q, const void* keys, const void* pos, void* probs, void* loss_rows, int32_t batch,
                    int32_t arms, int32_t k, int32_t d, float temperature, void* stream);
int mla_tac_nce_bwd(const void* q, const void* keys, const void* pos, const void* probs, const void* gscale, void* dq,
                    void* dkeys, int32_t batch, int32_t arms, int32_t k, int32_t d, float temperature, void* stream);
int mla_project_points(const void* xyz, const void* cam, int64_t n, float sx, float sy, float total_stride,
                       int32_t patch_h, int32_t patch_w, float img_w, float img_h, void* patch_idx, void* valid,
                       void* stream);
/* Tactile positives (models/vlm/prismatic.py:742-749, torch.cdist + topk(1)): gripper_xyz f32 [batch,arms,3], centers f32
 * [batch,groups,3], patch_idx int64 [batch,groups,2] (from mla_project_points) -> pos_pc int64 [batch,arms] = nearest
 * centre (lowest index on ties), lin_img int64 [batch,arms] = row*patch_w + col of that centre's image patch. */
int mla_nearest_center(const void* gripper_xyz, const void* centers, const void* patch_idx, int32_t batch, int32_t arms,
                       int32_t groups, int32_t patch_w, void* pos_pc, void* lin_img, void* stream);

/* ---- point-cloud tokenizer (models/mla/pointcloud/backbone/Point_PN.py) -----------------------------------------
 * fps: furthest_point_sample (:6-21): xyz f32 [B,n,3], start int64 [B] (the reference's torch.randint draw) ->
 *   idx int32 [B,npoint], centers f32 [B,npoint,3]; lowest-index tie-break.
 * knn: knn_point (:62-73): k nearest of each query f32 [B,groups,3] among xyz; bf16_dist=1 reproduces the autocast
 *   distance arithmetic (bf16 matmul + in-place bf16 accumulation); ties by lowest index; idx int32 [B,groups,k].
 * group_pose: LGA 'scan' normalisation + concat(knn feat, centre feat) + PosE_Geo (:125-150,:228-249):
 *   feat [B,n,c] (bf16 or f32) -> rows (b,g,k) x 2c channels, written as f32 and as a bf16 copy.
 * bn_*: train-mode BatchNorm over rows of a bf16 [rows,c] matrix: sums f32 [2c] (zero it first), coef f32 [2c] =
 *   mean | invstd, running stats updated with `momentum`; bn_relu: out = bf16(relu(bn(y)));
 *   bn_res_relu: v = relu(bf16(bn(y)) + x) (Linear2Layer :219); whichever of x_f32_out, x_bf16_out and pooled
 *   (max over the k rows of each group, Pooling :161-170) is non-NULL is written. */
int mla_fps(const void* xyz, const void* start, void* idx_out, void* centers, int32_t batch, int32_t n,
            int32_t npoint, void* stream);
int mla_knn(const void* xyz, const void* query, void* knn_idx, int32_t batch, int32_t n, int32_t groups, int32_t k,
            int32_t bf16_dist, void* stream);
int mla_group_pose(const void* xyz, const void* feat, int32_t feat_is_bf16, const void* fps_idx, const void* knn_idx,
                   const void* dim_embed, void* x_f32, void* x_bf16, int32_t batch, int32_t n, int32_t groups,
                   int32_t k, int32_t c, float beta, void* stream);
int mla_bn_stats(const void* y, void* sums, int64_t rows, int32_t c, void* stream);
int mla_bn_finalize(const void* sums, void* coef, void* running_mean, void* running_var, int64_t rows, int32_t c,
                    float eps, float momentum, void* stream);
int mla_bn_relu(const void* y, const void* coef, const void* w, const void* bias, void* out, int64_t rows, int32_t c,
                void* stream);
int mla_bn_res_relu(const void* y, const void* coef, const void* w, const void* bias, const void* x, void* x_f32_out,
                    void* x_bf16_out, void* pooled, int64_t groups, int32_t k, int32_t c, void* stream);

/* ---- tokenizer backward: stage "pretrain" trains vision_tower_2d / vision_tower_3d (models/vlm/prismatic.py:427-434)
 * The GEMM gradients run on mla_gemm_bf16; these are the autograd formulas of the ops between them.
 * local_attn_bwd: backward of mla_local_attn (vision_tokenizer.py:40-45): dout [groups,c] -> dq [groups,c],
 *   dkv [groups*win, 2c] (bf16); the 9-way softmax is recomputed from q, kv.
 * layernorm_bwd: backward of mla_layernorm_fwd (nn.LayerNorm of LocalAttention.q[0] / kv[0], :19-24) on bf16 rows;
 *   statistics recomputed; dx = LN-backward (+ dres[row]) (+ dgrp[row / win] / win): the optional addends are the
 *   `reduced_features +` residual (:46) and F.avg_pool2d's broadcast (:28) into the same tensor.  dw, db f32 [h] +=.
 * bn_bwd: backward of the train-mode BatchNorm of mla_bn_stats/finalize (Point_PN.py:173-219) over bf16 rows y
 *   [rows,c] with coef = mean | invstd.  Upstream `up` (f32 if up_is_f32 else bf16) is masked first: mode 0 none;
 *   mode 1 by relu(bn(y)) > 0 (conv-BN-ReLU, :176-181,:192-196); mode 2 by xnew > 0, xnew f32 = relu(bn(y) + x) the
 *   Linear2Layer output (:219) — then the masked f32 gradient is also the residual input's and is written to dres.
 *   sums f32 [2c]: on return db = sums[0..c), dw = sums[c..2c).  dy bf16 [rows,c] = gradient of the conv output.
 * maxpool_bwd: backward of x.max(-1) over the k neighbours (:157): d_xnew f32 [groups,k,c] from d_pooled [groups,c].
 * group_pose_bwd: backward of mla_group_pose's two gathers (index_points, :41-58,:116-122): dx f32 [B*groups*k, 2c]
 *   -> dfeat f32 [B*n, c] += (zero it first).
 * diag_block_sum: out f32 [m,n] = sum_s x[s*m+i, s*n+j], x f32 [parts*m, parts*n]: folds a weight-gradient GEMM
 *   whose very long reduction (rows of neighbours) was spread over `parts` row-interleaved tiles. */
int mla_local_attn_bwd(const void* q, const void* kv, const void* dout, void* dq, void* dkv, int64_t groups, int32_t c,
                       int32_t heads, int32_t win, float scale, void* stream);
int mla_layernorm_bwd(const void* dy, const void* x, const void* w, const void* dres, const void* dgrp, int32_t win,
                      void* dx, void* dw, void* db, int64_t rows, int32_t h, float eps, void* stream);
int mla_bn_bwd(const void* up, int32_t up_is_f32, const void* y, const void* coef, const void* w, const void* bias,
               const void* xnew, int32_t mode, void* sums, void* dy, void* dres, int64_t rows, int32_t c, void* stream);
int mla_maxpool_bwd(const void* xnew, const void* dpooled, void* dxnew, int64_t groups, int32_t k, int32_t c,
                    void* stream);
int mla_group_pose_bwd(const void* dx, const void* fps_idx, const void* knn_idx, void* dfeat, int32_t batch, int32_t n,
                       int32_t groups, int32_t k, int32_t c, void* stream);
int mla_diag_block_sum(const void* x, void* out, int32_t m, int32_t n, int32_t parts, void* stream);

/* ---- inference denoise loop (MLA.predict_action_diff, models/mla/model_mla.py:592-775) -----------------------------
 * The decoder prefix (everything in front of [t | x_0..x_T]) runs once on the training-path kernels; its post-RoPE K/V
 * stay in the per-layer q|k|v buffers and each DDIM step runs only the suffix rows.
 * gemv_bf16: skinny nn.Linear for m <= 16 rows (HBM-bound weight streaming): out[m,n] = bf16(bf16(x w^T) + residual),
 *   x bf16 [m,k] (pitch ldx), w bf16 [n,k] (pitch ldw), residual optional; k and pitches multiples of 8.
 * decode_attn: the last len_q positions of a length-len_k sequence attend to the cached keys with the bottom-right
 *   aligned causal mask of flash-attn (modeling_llama.py:540-557): q row (b,i) at q + (b*len_q+i)*ldq + h*d, K/V row
 *   (b,j) at k|v + (b*len_k+j)*ldkv + h*d; query i sees j <= len_k - len_q + i.  head_dim in {32,64,128,256}.
 * ddim_step: x_{t-1} of GaussianDiffusion.ddim_sample with eta = 0, clip_denoised=False, epsilon-predicting model
 *   (models/diffusion/gaussian_diffusion.py:342-352,:522-571): x, out f32 [n]; eps bf16 (or f32); coef f32 [4] =
 *   sqrt(1/ac_t), sqrt(1/ac_t - 1), sqrt(ac_prev), sqrt(1 - ac_prev).  Bit-exact with the reference's fp32 op order. */
int mla_gemv_bf16(const void* x, const void* w, void* out, const void* residual, int32_t m, int32_t n, int32_t k,
                  int64_t ldx, int64_t ldw, int64_t ldo, int64_t ldr, void* stream);
/* gemv_fused: the same skinny linear with an optional prologue applied to the activations as they are loaded (m <= 4
 * with k <= 4096, or m <= 2 with k <= 12288): prologue 1 = LlamaRMSNorm (modeling_llama.py:85-90) with ln_weight bf16 [k] and eps; prologue 2 = SwiGLU
 * of LlamaMLP (:240), x = [gate | up] bf16 [m, 2k].  prologue 0 = mla_gemv_bf16. */
typedef struct mla_gemv_args {
  const void *x, *w;
  void* out;
  const void *residual, *ln_weight;
  int32_t m, n, k;
  int64_t ldx, ldw, ldo, ldr;
  int32_t prologue;
  float eps;
} mla_gemv_args;
int mla_gemv_fused(const mla_gemv_args* a, void* stream);
/* skinny_gemm: the same skinny linear (same arguments and prologues as mla_gemv_fused, m <= 32) on the tensor cores,
 * "swap-AB": the weights are the 128-row A operand of tcgen05.mma streamed by TMA, the activation rows its 16/32-column
 * B operand (written by the kernel's prologue warps straight into the swizzled operand layout), split-K over the SMs
 * with a fixed-order last-arrival sum (csrc/skinny_sm100.cu).  The weights never pass through the CUDA cores.
 * workspace: mla_skinny_gemm_workspace(n, m) bytes, ZEROED once before the first launch (the arrival counters re-arm
 * themselves); one workspace per (n, m <= 16 | m <= 32) in flight on a stream. */
int mla_skinny_gemm(const mla_gemv_args* a, void* workspace, void* stream);
size_t mla_skinny_gemm_workspace(int32_t n, int32_t m);
/* The gemv kernels are launched with programmaticStreamSerialization (the weight prefetch of the next kernel overlaps
 * the previous kernel's tail; consumers griddepcontrol.wait before touching activations).  0 turns it off (env
 * MLA_DECODE_PDL=0 does the same). */
int mla_decode_set_pdl(int32_t on);
/* decode_stack: the n suffix rows per sample (batch*n <= 2) through ALL decoder layers in ONE persistent launch
 * (csrc/decode_stack.cu) — replaces layers x [gemv_fused, decode_attn_rope, gemv_fused, gemv_fused, gemv_fused]
 * (LlamaModel.decode; modeling_llama.py:405-597 per layer on the suffix rows of MLA.predict_action_diff,
 * model_mla.py:592-775).  One CTA per SM: a producer thread streams every layer's weights through a shared-memory ring
 * without ever waiting for activations, the consumer warps walk the five phases of each layer separated by a grid
 * barrier.  Bit-identical to the per-op path.
 *   w_qkv / w_o / w_gate_up / w_down / ln1 / ln2 / kv_cache: DEVICE arrays of `layers` device pointers (bf16, contiguous:
 *   [3h,h] [h,h] [2*ffn,h] [h,ffn] [h] [h] and the head-major prefix cache [batch, 2, heads, prefix, head_dim]).
 *   x bf16 [batch*n, h]: layer-0 input on entry, last layer's output on return (final norm not applied).
 *   qkv [batch*n, 3h], ctx / x_mid [batch*n, h], gate_up [batch*n, 2*ffn]: bf16 scratch.
 *   cos_t / sin_t bf16 [n, head_dim/2]: RoPE table rows of positions prefix..prefix+n-1.
 *   workspace: mla_decode_stack_workspace(...) bytes, ZEROED once before the first launch (the grid barrier's
 *   counters re-arm themselves).  head_dim in {32, 64, 128}; h, ffn multiples of 8 and <= 12288. */
typedef struct mla_decode_stack_args {
  const void* const* w_qkv;
  const void* const* w_o;
  const void* const* w_gate_up;
  const void* const* w_down;
  const void* const* ln1;
  const void* const* ln2;
  const void* const* kv_cache;
  void *x, *qkv, *ctx, *x_mid, *gate_up;
  const void *cos_t, *sin_t;
  void* workspace;
  int32_t layers, batch, n, prefix, heads, head_dim, ffn;
  float eps, scale;
  void* trace; /* null, or int64 [SMs][layers][5][3]: globaltimer ns per CTA at phase entry / work done / barrier passed */
} mla_decode_stack_args;
int mla_decode_stack(const mla_decode_stack_args* a, void* stream);
size_t mla_decode_stack_workspace(int32_t batch, int32_t n, int32_t prefix, int32_t heads, int32_t head_dim);
/* Profiling only (results become meaningless): 1 = consumers release the ring slots without doing the math, 2 = no grid
 * barriers, 4 = no attention phase.  0 restores the product behaviour. */
int mla_decode_stack_set_debug(int32_t flags);
/* How many weight groups (32-44 KB each, per SM) the producer's L2 prefetch runs in front of its shared-memory ring
 * (default 12 = ~57 MB chip-wide; env MLA_DECODE_STACK_AHEAD); 0 = no run-ahead. */
int mla_decode_stack_set_ahead(int32_t groups);
/* Cap the shared-memory ring (KB; 0 = as large as fits, 192 KB): tuning / profiling. */
int mla_decode_stack_set_ring_kb(int32_t kb);
/* Weight rows per ring slot where K > 4096 (the down projection): 1 or 2 (default 2); tuning. */
int mla_decode_stack_set_rows_per_slot_big(int32_t rows);
/* rope_cache: the n new rows per sample of a packed q|k|v projection bf16 [batch*n, 3*heads*head_dim] at positions
 * prefix..prefix+n-1: RoPE (modeling_llama.py:184-208) on q in place and on k into cache row (b*(prefix+n) + prefix + i),
 * v copied beside it; cache bf16 [batch*(prefix+n), 2*heads*head_dim] = k | v; cos/sin bf16 [n, head_dim/2]. */
int mla_rope_cache(void* qkv, void* cache, const void* cos_t, const void* sin_t, int32_t batch, int32_t n,
                   int32_t prefix, int32_t heads, int32_t head_dim, void* stream);
int mla_decode_attn(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* o, int64_t ldo,
                    int32_t batch, int32_t heads, int32_t len_q, int32_t len_k, int32_t head_dim, float scale,
                    void* stream);
/* decode_attn_rope: the same attention reading q and the len_q NEW key/value rows un-rotated from the packed projection
 * qkv bf16 [batch*len_q, 3*heads*head_dim] (q | k | v) and applying RoPE on the fly (cos/sin bf16 [len_q, head_dim/2]
 * = table rows of positions len_k-len_q..len_k-1).  The cache holds the rotated prefix rows only, row (b, h, j),
 * j < len_k - len_q, at k|v_cache + b*kv_stride_b + h*kv_stride_h + j*kv_stride_j (elements; a head-major cache,
 * kv_stride_j = head_dim, streams contiguously).  Nothing is written back to the cache. */
int mla_decode_attn_rope(const void* qkv, int64_t ldqkv, const void* k_cache, const void* v_cache, int64_t kv_stride_b,
                         int64_t kv_stride_h, int64_t kv_stride_j, const void* cos_t, const void* sin_t, void* o,
                         int64_t ldo, int32_t batch, int32_t heads, int32_t len_q, int32_t len_k, int32_t head_dim,
                         float scale, void* workspace, void* counters, void* stream);
/* Split-K scratch of decode_attn_rope: with workspace (mla_decode_attn_workspace bytes, f32) and counters (int32
 * [batch*heads*len_q], zero before the first use; the kernel re-arms them) the keys are split over ceil(len_k/128) CTAs
 * per (sample, head, query) — one DRAM round trip each — and the last CTA to arrive merges.  NULL = one CTA per query. */
size_t mla_decode_attn_workspace(int32_t batch, int32_t heads, int32_t len_q, int32_t len_k, int32_t head_dim);
int mla_ddim_step(const void* x, const void* eps, int32_t eps_is_f32, const void* coef, void* out, int64_t n,
                  void* stream);

/* ---- optimizer step of the data-parallel trainer (training/strategies/fsdp.py:242-257,:310) ------------------
 * sumsq: out[0] += sum(x^2) (f32).  clip_coef: scale[0] = min(1, max_norm/(||g||*inv_world + 1e-6)) * inv_world,
 * scale[1] = the mean-gradient norm — clip_grad_norm_ without the host round trip (g holds rank-summed gradients).
 * adamw: torch.optim.AdamW update of an f32 tensor with the gradient pre-scaled by grad_scale[0] (device scalar, may be
 * NULL); optionally refreshes the bf16 compute copy in the same pass. */
int mla_sumsq_f32(const void* x, int64_t n, void* out, void* stream);
int mla_clip_coef(const void* sumsq, float max_norm, float inv_world, void* scale, void* stream);
int mla_adamw_f32(void* p, const void* g, void* m, void* v, void* p_bf16, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int64_t step, const void* grad_scale, void* stream);
/* n > 0: launch AdamW with n CTAs of 128 threads per SM (small register footprint) so that it can share the SMs with
 * the persistent GEMM CTAs of the next step's forward (trainer, MLA_ADAM_STREAM=1); 0 = the standalone launch shape. */
int mla_adamw_set_lean(int32_t ctas_per_sm);

/* ---- vocabulary cross-entropy (modeling_llama.py:1254-1269) ------------------------------------------------------
 * logits bf16 [rows = B*seq, vocab] (pitch ld) from the lm_head GEMM; labels int64 [B, seq] UNshifted: row (b,s) is
 * scored against labels[b,s+1]; -100 and the last position are ignored; loss[0] = mean over scored rows.
 * lse f32 [rows], acc2 f32 [2] (sum, count) are kept for backward, which overwrites logits with d(loss)/d(logits). */
int mla_ce_fwd(const void* logits, int64_t ld, const void* labels, int64_t rows, int32_t seq, int32_t vocab, void* lse,
               void* acc2, void* loss, void* stream);
int mla_ce_bwd(void* logits_inout, int64_t ld, const void* labels, int64_t rows, int32_t seq, int32_t vocab,
               const void* lse, const void* acc2, const void* gscale, void* stream);

/* ---- ActionTokenizer (vla/action_tokenizer.py:43-71) — integer results bit-exact ------------------------------------
 * digitize: ids[i] = vocab_size - np.digitize(np.clip(x[i], lo, hi), edges)  (edges: float64 [bins] = np.linspace);
 * decode: out[i] = centers[clip(vocab_size - ids[i] - 1, 0, n_centers-1)] (float64). */
int mla_action_digitize(const void* x, int32_t x_is_f64, int64_t n, const void* edges, int32_t bins, double lo,
                        double hi, int64_t vocab_size, void* ids_out, void* stream);
int mla_action_decode(const void* ids, int64_t n, const void* centers, int32_t n_centers, int64_t vocab_size,
                      void* out, void* stream);

/* ---- post-training generation heads (models/mla/generation/models.py, gen_loss.py, utils.py; prismatic.py:771-838,
 *      :1075-1113) — SURVEY.md §8 A14 ----------------------------------------------------------------------------
 * mha_fwd / mha_bwd: the attention core of nn.MultiheadAttention / nn.TransformerDecoderLayer (generation/models.py
 * :44,:103-122,:405-413): non-causal softmax(scale * Q K^T) V for any even head_dim <= 1024 (multiple of 4), Lq != Lk;
 * bf16 in/out, fp32 math.  q row (b,i) = q + (b*len_q + i)*ldq + head*head_dim (likewise k, v with len_k), so packed
 * in_proj outputs are consumed in place.  keep_mask (u8 [batch, heads, len_q, len_k], may be NULL) + keep_scale
 * = 1/(1-p) is the attention-probability dropout of the reference's train mode.  lse f32 [batch, heads, len_q].
 * Backward (delta f32 [batch, heads, len_q] is scratch) writes dq, dk, dv with their own row strides. */
typedef struct mla_mha_args {
  const void *q, *k, *v;
  int64_t ldq, ldk, ldv;
  void* o;
  int64_t ldo;
  void* lse;
  const void* keep_mask;
  float keep_scale;
  int32_t batch, heads, len_q, len_k, head_dim;
  float scale;
  const void* d_o;
  int64_t ld_do;
  void* delta;
  void *dq, *dk, *dv;
  int64_t ld_dq, ld_dk, ld_dv;
} mla_mha_args;
int mla_mha_fwd(const mla_mha_args* a, void* stream);
int mla_mha_bwd(const mla_mha_args* a, void* stream);

/* nn.LayerNorm in fp32 (CUDA autocast runs layer_norm in fp32): x, y f32 [rows, h]; w, b f32 [h]; y_bf16 optional bf16
 * copy of y; mean, rstd f32 [rows] saved for backward.  Backward: dx f32; dw, db f32 [h] are ACCUMULATED (atomics). */
int mla_ln_f32_fwd(const void* x, const void* w, const void* b, void* y, void* y_bf16, void* mean, void* rstd,
                   int64_t rows, int32_t h, float eps, void* stream);
int mla_ln_f32_bwd(const void* dy, const void* x, const void* w, const void* mean, const void* rstd, void* dx, void* dw,
                   void* db, int64_t rows, int32_t h, void* stream);
/* helpers of the heads' fp32 residual stream: bf16 -> f32 cast; out = a(f32) + b(bf16) (optionally rounded to bf16, the
 * value a bf16 + bf16 add of the reference produces); y = keep[i / per_mask] ? x * scale : 0 (dropout / DropPath masks
 * drawn by the host-side generator); mean over the sequence [B, S, C] -> [B, C] (generation/models.py:352) and its
 * backward. */
int mla_cast_bf16_f32(const void* x, void* y, int64_t n, void* stream);
int mla_add_f32_bf16(const void* a, const void* b, void* out, int64_t n, int32_t round_bf16, void* stream);
int mla_mask_scale_bf16(const void* x, const void* keep, void* y, int64_t n, int64_t per_mask, float scale, void* stream);
int mla_seq_mean_fwd(const void* x, void* y, int32_t B, int32_t S, int32_t C, void* stream);
int mla_seq_mean_bwd(const void* dy, void* dx, int32_t B, int32_t S, int32_t C, void* stream);
/* nn.BatchNorm1d in train mode (+ReLU) over rows: x, y, dx bf16 [R, C] (the reference's [B, C, L] transposed,
 * generation/models.py:332-337); w, b f32; batch statistics over the R rows, running stats updated (momentum, unbiased
 * variance).  Backward writes dw, db f32 [C]. */
int mla_bn_rows_fwd(const void* x, const void* w, const void* b, void* y, void* mean, void* rstd, void* running_mean,
                    void* running_var, int32_t R, int32_t C, float eps, float momentum, int32_t relu, void* stream);
int mla_bn_rows_bwd(const void* dy, const void* x, const void* y, const void* w, const void* mean, const void* rstd,
                    void* dx, void* dw, void* db, int32_t R, int32_t C, int32_t relu, void* stream);
/* tile_rows: learned table f32 [P*h] (queries / positional embedding; bf16 in the reference) repeated for B samples ->
 * f32 [B*P*h] holding the bf16-rounded values; backward sums over samples.  mask_tokens: MAE decoder input
 * (generation/models.py:183-187) = (roi ? mask_token : image feature) + pos, bf16 arithmetic, written as f32 [B*P, h];
 * backward writes d_feat bf16, d_pos f32 [P, h] and ACCUMULATES d_mask_token f32 [h] (caller zeroes it). */
int mla_tile_rows_fwd(const void* table, void* out, int64_t table_elems, int32_t B, void* stream);
int mla_tile_rows_bwd(const void* d_out, void* d_table, int64_t table_elems, int32_t B, void* stream);
int mla_mask_tokens_fwd(const void* feat, const void* roi, const void* mask_token, const void* pos, void* out, int32_t B,
                        int32_t P, int32_t h, void* stream);
int mla_mask_tokens_bwd(const void* d_out, const void* roi, void* d_feat, void* d_mask_token, void* d_pos, int32_t B,
                        int32_t P, int32_t h, void* stream);
/* create_roi_mask_from_indices + dilate_mask (generation/utils.py:41-70): patch_idx i64 [B, n_pts, 2] (row, col) ->
 * mask u8 [B, G, G] dilated by a ksize x ksize max-pool. */
int mla_roi_mask(const void* patch_idx, void* mask, int32_t B, int32_t n_pts, int32_t G, int32_t ksize, void* stream);

/* Image head tail + image losses in one pass per patch (ImageGenerationModule.forward tail and
 * _generate_generated_patches, generation/models.py:193-286; compute_generation_losses prismatic.py:779-816).
 * cur / nxt: f32 images, element strides per image / channel (row stride = width); sample b uses image b % n_images
 * (MLA.forward tiles the batch, model_mla.py:160-167); patch n = (b, gy, gx), n_patches = B_eff * grid * grid.
 * delta_raw bf16 [n_patches, ld_delta >= 3*patch^2], ao_raw bf16 [n_patches, ld_ao >= 3] = raw (alpha, offset_x,
 * offset_y) head outputs; roi u8 [n_patches].  Forward writes blended f32 [n_patches, 3*patch^2] (image_generation),
 * delta_all / alpha_all / offset_all (bf16), sums f32 [5] (scratch), losses f32 [4] = {image_gen_loss, roi, bg,
 * delta reward} and coef f32 [4] for backward.  Backward (grad_scale f32 [1] = upstream gradient of image_gen_loss)
 * writes d_delta_raw / d_ao_raw with the raw tensors' strides. */
typedef struct mla_gen_image_args {
  const void *cur, *nxt;
  int64_t cur_stride_b, cur_stride_c, nxt_stride_b, nxt_stride_c;
  int32_t n_images, width, patch, grid, n_patches;
  const void* delta_raw;
  int64_t ld_delta;
  const void* ao_raw;
  int64_t ld_ao;
  const void* roi;
  float delta_clip, max_shift, gen_weight;
  void *blended, *delta_all, *alpha_all, *offset_all, *sums, *losses, *coef;
  const void* grad_scale;
  void *d_delta_raw, *d_ao_raw;
} mla_gen_image_args;
int mla_gen_image_fwd(const mla_gen_image_args* a, void* stream);
int mla_gen_image_bwd(const mla_gen_image_args* a, void* stream);

/* chamfer_distance_l2 (generation/gen_loss.py:12-18): pred bf16 [B, N1, 3], gt f32 [n_gt, N2, 3] (sample b uses
 * gt[b % n_gt]); idx / dist keep the arg-min pairs for backward; loss f32 [1].  Backward accumulates into dpred_f32
 * [B, N1, 3] (caller zeroes it), scaled by grad_scale f32 [1]. */
int mla_chamfer_fwd(const void* pred, const void* gt, int32_t B, int32_t N1, int32_t N2, int32_t n_gt, void* idx1,
                    void* dist1, void* idx2, void* dist2, void* loss, void* stream);
int mla_chamfer_bwd(const void* pred, const void* gt, int32_t B, int32_t N1, int32_t N2, int32_t n_gt, const void* idx1,
                    const void* dist1, const void* idx2, const void* dist2, const void* grad_scale, void* dpred_f32,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLA_B200_H */
