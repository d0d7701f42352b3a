/*
 * vault_b200 -- C ABI of the B200 (sm_100a) kernels behind the VAuLT hot path.
 *
 * The reference (gchochla/VAuLT) has no FFI/plugin layer: its hot path is Python glue
 * (ref:vault/models/vault/model.py:151-218, 512-570) over HuggingFace ViLT/BERT modules that dispatch to ATen.  This
 * header is the boundary the new implementation introduces *below* that Python class API: plain
 * `extern "C"` functions, raw device pointers and sizes, no torch types.  Each entry point names the reference
 * computation it replaces (`ref:` = /root/reference, `HF:` = transformers==4.48.0 modelling code the reference calls).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (outputs and workspaces included); nothing is allocated,
 *     freed or retained;
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue and return (no device synchronisation);
 *   - return 0 on success, a negative VAULT_ERR_* otherwise; text via vault_last_error();
 *   - activations bf16 unless stated, statistics / residual stream / master weights / gradients fp32;
 *   - there is NO CPU fallback: a missing device or a non-sm_100 device is an error.
 */
#ifndef VAULT_B200_H_
#define VAULT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAULT_OK 0
#define VAULT_ERR_INVALID (-1) /* bad shape / alignment / argument */
#define VAULT_ERR_LAUNCH (-2)  /* cudaGetLastError() after a launch */
#define VAULT_ERR_DRIVER (-3)  /* tensor-map encode / driver entry point */
#define VAULT_ERR_ARCH (-4)    /* device is not sm_100 */

int vault_version(void);
/* copies the last error message of the calling thread; returns its length */
size_t vault_last_error(char* buf, size_t cap);
/* 0 if device `dev` is an sm_100 part, VAULT_ERR_ARCH otherwise */
int vault_check_device(int dev);

/* ------------------------------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 / TMEM, operands fed by TMA.
 * Replaces every nn.Linear on the path and its autograd:
 *   ViLT  HF:models/vilt/modeling_vilt.py:306-428 (query/key/value, attention.output.dense, intermediate, output)
 *   BERT  HF:models/bert/modeling_bert.py:143-356
 * forward   y = x W^T + b      : A = x [M,K] (a_mn=0), B = W [N,K] (b_mn=0)
 * dgrad     dx = dy W          : A = dy [M,N'] (a_mn=0), B = W stored [N',K'] read as MN-major (b_mn=1)
 * wgrad     dW = dy^T x        : A = dy stored [M',N] read MN-major (a_mn=1), B = x stored [M',K] read MN-major (b_mn=1)
 * ------------------------------------------------------------------------------------------------------------------ */
enum {
  VAULT_EPI_BIAS_BF16 = 0,       /* out(bf16) = acc + bias                                                   */
  VAULT_EPI_BIAS_GELU_BF16 = 1,  /* out2(bf16, optional) = acc + bias ; out(bf16) = gelu_erf(acc + bias)    */
  VAULT_EPI_BIAS_RESID_F32 = 2,  /* out(f32) = resid(f32) + dropout_p(acc + bias)                            */
  VAULT_EPI_PLAIN_BF16 = 3,      /* out(bf16) = acc                                                          */
  VAULT_EPI_DGELU_BF16 = 4,      /* out(bf16) = acc * gelu_erf'(aux(bf16))                                   */
  VAULT_EPI_ATOMIC_F32 = 5,      /* out(f32) += acc   (split-K partial sums; caller zero-fills)               */
  VAULT_EPI_BIAS_F32 = 6,        /* out(f32) = acc + bias (bias optional)                                    */
  VAULT_EPI_STORE_F32 = 7,       /* out(f32) = acc        (wgrad without split-K)                             */
  VAULT_EPI_ATOMIC_BIAS_DROP_F32 = 8, /* out(f32) += dropout_p(acc + bias): split-K form of RESID, out already holds the residual */
  VAULT_EPI_BIAS_GELU_GRAD_BF16 = 9,  /* training forward: out(bf16) = gelu_erf(acc + bias) ; out2(bf16, optional) = gelu_erf'(acc + bias)  */
  VAULT_EPI_MUL_AUX_BF16 = 10         /* its backward: out(bf16) = acc * aux(bf16), aux = the derivative saved by epilogue 9                */
};

typedef struct vault_gemm_args {
  int32_t M, N, K;        /* logical C[M,N] = sum_k A[m,k] * B[n,k] */
  const void* A;          /* bf16 */
  int64_t lda;            /* elements between consecutive rows of the STORED matrix */
  int32_t a_mn;           /* 0: stored [M,K] (K contiguous)   1: stored [K,M] (M contiguous) */
  const void* B;          /* bf16 */
  int64_t ldb;
  int32_t b_mn;           /* 0: stored [N,K] (K contiguous)   1: stored [K,N] (N contiguous) */
  int32_t epilogue;       /* VAULT_EPI_* */
  const float* bias;      /* [N] or NULL */
  const float* resid;     /* [M,N] fp32 (EPI_BIAS_RESID_F32) */
  int64_t ldr;
  const void* aux;        /* [M,N] bf16 (EPI_DGELU_BF16) */
  int64_t ldaux;
  void* out;
  int64_t ldo;
  void* out2;             /* optional */
  int64_t ldo2;
  float dropout_p;        /* EPI_BIAS_RESID_F32 only; 0 = off */
  uint64_t seed;
  const uint64_t* seed_dev; /* optional DEVICE word added to `seed` at run time (lets a captured CUDA graph advance its masks) */
  uint32_t site;          /* dropout site id: forward and backward of the same site regenerate the same mask */
  int32_t split_k;        /* >=1; >1 only with EPI_ATOMIC_F32 */
  int32_t block_n;        /* 0 = choose; else 64 / 128 / 256 */
  int32_t max_ctas;       /* 0 = all SMs */
  int32_t* sched;         /* optional DEVICE int32[2], zero-initialised, re-armed by the kernel: tiles are claimed dynamically (atomic
                             counter) instead of round-robin, so SMs held by a concurrent kernel / collective do not stretch the launch.
                             One pair per launch that may run concurrently with another. */
  int32_t cluster;        /* CTA-pair operand multicast: 0/-1 off, 1 pair along M (B tile shared), 2 pair along N (A tile shared) */
  float* a_colsum;        /* optional DEVICE fp32 [M], ACCUMULATED (caller zero-fills): a_colsum[m] += sum_k A[m,k].  a_mn = 1 and an fp32
                             epilogue only: in a weight-gradient launch A = dy^T, so this is the bias gradient of the same nn.Linear
                             (replaces a separate column-sum pass over dy); the sums are taken from the A tiles already in shared memory. */
} vault_gemm_args;

int vault_gemm_bf16(const vault_gemm_args* args, void* stream);
/* Up to 4 weight-gradient problems (a_mn = b_mn = 1, VAULT_EPI_ATOMIC_F32, per-problem split_k and optional a_colsum; 128x256 tiles) as ONE
 * persistent launch over the pooled tile list: the four dW of a transformer layer (QKV, attention output, MLP-1, MLP-2: same tokens as
 * the contraction) share one launch head / tail and fill the SMs in whole waves.  Fields other than M, N, K, A, lda, B, ldb, out, ldo,
 * split_k, a_colsum (and max_ctas of problem 0) are ignored.  Same results as the single launches (fp32 atomics: same values, any order). */
int vault_gemm_wgrad_grouped(const vault_gemm_args* args, int32_t n, void* stream);

/* Patch embedding, im2col-free: Conv2d(C,N,k=32,s=32) as a TF32 tcgen05 GEMM whose A tiles are 5-D TMA boxes taken straight
 * from the NCHW fp32 pixels (no patch matrix is materialised).  Replaces ViltPatchEmbeddings.forward,
 * HF:models/vilt/modeling_vilt.py:293-303.
 *   pixels [B,C,Hi,Wi] fp32, weight [N, C*32*32] fp32 (= projection.weight.view(N,-1)), bias [N] fp32 or NULL
 *   out [B*(Hi/32)*(Wi/32), N] fp32, row = b*gh*gw + i*gw + j                                                     */
int vault_patch_embed_fwd(const float* pixels, const float* weight, const float* bias, float* out, int32_t B, int32_t C,
                          int32_t Hi, int32_t Wi, int32_t P, int32_t N, void* stream);
/* Its weight gradient, im2col-free as well (backward of HF:models/vilt/modeling_vilt.py:293-303 w.r.t. projection.weight):
 *   dW[n, (c,kh,kw)] += sum over patches of dpatch[patch, n] * pixel(patch, c, kh, kw)
 * TF32 tcgen05 GEMM with the contraction over the patches: the pixel operand comes through the SAME 5-D TMA map as the forward
 * (one box = [patch rows] x [32 kw] = a 32-wide atom of an MN-major operand), dpatch fp32 [B*gh*gw, N] is read MN-major too.
 * dW fp32 [N, C*32*32] is ACCUMULATED into (red.global.add; caller zero-fills).  vault_patch_embed_wgrad_ok() says whether the
 * patch grid admits the kernel's k-blocks (whole patch rows of <= 64 patches: images up to 2048 px wide); returns 1 / 0.     */
int vault_patch_embed_wgrad_ok(int32_t C, int32_t Hi, int32_t Wi, int32_t P, int32_t N);
int vault_patch_embed_wgrad(const float* pixels, const float* dpatch_f32, float* dW, int32_t B, int32_t C, int32_t Hi,
                            int32_t Wi, int32_t P, int32_t N, void* stream);
/* fp32 patch-gradient rows for the kernel above, gathered from the gradient of the assembled sequence, + the projection's
 * bias gradient:  dpatch[b, i*gw+j] = dX[b, T+1+i*w_b+j] inside the sample's valid rectangle hw[b] = (h_b, w_b), zero rows
 * elsewhere;  dbias[n] += column sums (atomics; may be NULL).  dX fp32 [B, T+1+Pmax, H]; dpatch fp32 [B*gh*gw, H].        */
int vault_patch_grad_rows_f32(const float* dX, const int32_t* hw, float* dpatch_f32, float* dbias, int32_t B, int32_t T,
                              int32_t Pmax, int32_t gh, int32_t gw, int32_t H, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * LayerNorm (HF nn.LayerNorm call sites: 25 in the LM, 26 in ViLT).  x fp32 [rows, cols].
 *   fwd: y = (x - mean) * rstd * gamma + beta -> y_bf16 and/or y_f32 (either may be NULL); saves mean, rstd [rows]
 *   bwd: dy = (dy_f32 ? dy_f32 : 0) + (dy_bf16 ? dy_bf16 : 0);
 *        dx = LN'(dy) + (dres_f32 ? dres_f32 : 0) -> dx_f32 (required), dx_bf16 (optional);
 *        dgamma, dbeta [cols] are ACCUMULATED with atomics (caller zero-fills), either may be NULL (frozen)
 * ------------------------------------------------------------------------------------------------------------------ */
int vault_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean,
                        float* rstd, int64_t rows, int32_t cols, float eps, void* stream);
int vault_layernorm_bwd(const float* dy_f32, const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* dres_f32, float* dx_f32, void* dx_bf16, float* dgamma,
                        float* dbeta, int64_t rows, int32_t cols, void* stream);
/* Dropout-aware variants (BERT stack in training; every mask is Philox(seed + *seed_dev, site) keyed by element index).
 *   fwd: dropout on the LayerNorm OUTPUT (BertEmbeddings: dropout(LayerNorm(.)), HF:models/bert/modeling_bert.py:110-111)
 *   bwd: in_*  : that same output mask, applied to dy before differentiating the norm;
 *        out_* : mask of the dropout that sat on the GEMM output feeding this LayerNorm's input (BertSelfOutput / BertOutput,
 *                HF:models/bert/modeling_bert.py:294-297, 352-355): applied to the bf16 copy of dx only -- that copy is the dy
 *                operand of the dgrad / wgrad / bias-grad of that GEMM, while dx_f32 continues down the residual path. */
int vault_layernorm_fwd_drop(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean,
                             float* rstd, int64_t rows, int32_t cols, float eps, float dropout_p, uint64_t seed,
                             const uint64_t* seed_dev, uint32_t site, void* stream);
/*        dcolsum (optional, [cols], accumulated): column sums of the bf16 dx copy = bias gradient of the Linear whose dy it is */
int vault_layernorm_bwd_drop(const float* dy_f32, const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                             const float* gamma, const float* dres_f32, float* dx_f32, void* dx_bf16, float* dgamma,
                             float* dbeta, float* dcolsum, int64_t rows, int32_t cols, float in_p, uint32_t in_site,
                             float out_p, uint32_t out_site, uint64_t seed, const uint64_t* seed_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused masked-softmax attention over the variable-length text+image sequence.
 * Replaces ViltSelfAttention.forward HF:models/vilt/modeling_vilt.py:325-365 and BERT eager_attention_forward
 * HF:models/bert/modeling_bert.py:115-140 plus the additive mask of HF:modeling_utils.py:902-949 (never materialised).
 *   qkv  [B*S, 3*heads*64] bf16, row = [q(h0..), k(h0..), v(h0..)]      key_mask [B,S] uint8 (1 = attend)
 *   ctx  [B*S, heads*64] bf16         lse [B,heads,S] fp32 (natural-log sum-exp of scaled, masked scores)
 *   dropout on the probabilities (BERT stack in training) is Philox(seed, site) keyed by (b,h,q,k).
 * ------------------------------------------------------------------------------------------------------------------ */
int vault_attn_fwd(const void* qkv, const uint8_t* key_mask, void* ctx, float* lse, int32_t B, int32_t S, int32_t heads,
                   float dropout_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site, void* stream);
/* dqkv [B*S, 3*heads*64] bf16 is fully overwritten; delta [B,heads,S] fp32 is workspace */
int vault_attn_bwd(const void* qkv, const uint8_t* key_mask, const void* ctx, const void* dctx, const float* lse,
                   float* delta, void* dqkv, int32_t B, int32_t S, int32_t heads, float dropout_p, uint64_t seed,
                   const uint64_t* seed_dev, uint32_t site, void* stream);
/* Kernel family used by vault_attn_fwd / vault_attn_bwd: 0 = automatic (pipelined tcgen05 kernels for 193..384 keys, and for 65..384
 * keys with probability dropout; whole-row tcgen05 kernels for dropout-free calls with 65..192 keys; mma.sync kernels otherwise),
 * 1 = mma.sync kernels only,
 * 2 = whole-row tcgen05 kernels wherever their shape limits allow, 3 = pipelined tcgen05 kernels for every dropout-free call with
 * <= 384 keys, 4 = as 3 with the alternative forward that gives each softmax warpgroup its own tile (measured slower, see
 * attention_sm100.cu).  Process-wide; meant for tests and A/B measurements. */
int vault_attn_set_impl(int32_t impl);

/* ------------------------------------------------------------------------------------------------------------------
 * Embedding assembly.
 * LM:    BertEmbeddings.forward HF:models/bert/modeling_bert.py:72-112 / RobertaEmbeddings HF:models/roberta/
 *        modeling_roberta.py:79-150 (pad-offset position ids when roberta_pad >= 0):
 *        x = word[ids] + type[tt] + pos[pid]  -> x_sum fp32 [B*T, H] (input of the embedding LayerNorm)
 * ViLT text: TextEmbeddings.forward with inputs_embeds HF:models/vilt/modeling_vilt.py:240-272 (4.48.0 gate):
 *        x = inputs_embeds + type[tt] (+ pos[t] if pos != NULL)
 *        The same pair serves the LM when the caller passes text inputs_embeds instead of input_ids
 *        (ref:vault/models/vault/model.py:170-190; BertEmbeddings / RobertaEmbeddings with inputs_embeds): pos then points at the
 *        LM's position table, offset by pad+1 rows for RoBERTa (create_position_ids_from_inputs_embeds).
 * ------------------------------------------------------------------------------------------------------------------ */
/* also converts attention_mask int64 [B,T] -> key_mask uint8 [B,T] for the attention kernels (both optional) */
int vault_lm_embed_fwd(const int64_t* ids, const int64_t* tt, const float* word, const float* type, const float* pos,
                       float* x_sum, const int64_t* attention_mask, uint8_t* key_mask, int32_t B, int32_t T, int32_t H,
                       int32_t roberta_pad, void* stream);
/* scatter-add dx [B*T,H] fp32 into dword/dtype/dpos (fp32, atomics; caller zero-fills); any table grad may be NULL */
/* word_pad: nn.Embedding(padding_idx=pad_token_id) row that receives no gradient (-1: none) */
int vault_lm_embed_bwd(const int64_t* ids, const int64_t* tt, const float* dx, float* dword, float* dtype, float* dpos,
                       int32_t B, int32_t T, int32_t H, int32_t roberta_pad, int32_t word_pad, void* stream);
int vault_vilt_text_embed_fwd(const float* inputs_embeds, const int64_t* tt, const float* type, const float* pos,
                              float* x_sum, int32_t B, int32_t T, int32_t H, void* stream);
int vault_vilt_text_embed_bwd(const int64_t* tt, const float* dx, float* dtype, float* dpos, int32_t B, int32_t T,
                              int32_t H, void* stream);

/* Per-sample valid patch grid (h_b, w_b) from the pixel mask: nearest down-sampling, HF:models/vilt/modeling_vilt.py:95-98.
 * pixel_mask [B,Hi,Wi] int64 (mask_is_f32 = 0) or fp32 (1);  hw [B,2] int32 */
int vault_patch_grid(const void* pixel_mask, int32_t mask_is_f32, int32_t* hw, int32_t B, int32_t Hi, int32_t Wi, int32_t P,
                     void* stream);

/* ViLT sequence assembly (ViltEmbeddings.visual_embed :101-175 + ViltEmbeddings.forward :203-216), raster order:
 *   X[b, t]        = text_ln[b,t] + modality[0]                                   t <  T
 *   X[b, T]        = cls + pos_table[0] + modality[img_type]
 *   X[b, T+1+p]    = patch[b, i*gw+j] + bilinear_{align_corners}(pos_table[1:], (h_b,w_b))[i,j] + modality[img_type]
 *                    for p = i*w_b + j < h_b*w_b, zero rows after; key_mask [B,S] uint8 = [attention_mask, 1, valid]
 * X fp32 [B,S,H], S = T+1+Pmax;  text_ln fp32 [B*T,H];  patch fp32 [B*gh*gw, H];  pos_table [1+grid*grid, H]        */
int vault_vilt_assemble_fwd(const float* text_ln, const float* patch, const float* cls, const float* pos_table,
                            const float* modality, const int64_t* attention_mask, const int32_t* hw, float* X,
                            uint8_t* key_mask, int32_t B, int32_t T, int32_t Pmax, int32_t gh, int32_t gw, int32_t grid,
                            int32_t H, int32_t img_type, void* stream);
/* backward of the assembly: dX fp32 [B,S,H] ->
 *   dtext_ln fp32 [B*T,H] (= dX text rows), dpatch bf16 [B*gh*gw, H] (zero rows where invalid),
 *   dcls [H], dpos_table [1+grid*grid, H], dmodality [2.., H]  (fp32, atomics; caller zero-fills) */
int vault_vilt_assemble_bwd(const float* dX, const int32_t* hw, float* dtext_ln, void* dpatch_bf16, float* dcls,
                            float* dpos_table, float* dmodality, int32_t B, int32_t T, int32_t Pmax, int32_t gh, int32_t gw,
                            int32_t grid, int32_t H, int32_t img_type, void* stream);

/* im2col of NCHW fp32 pixels into bf16 patch rows [B*gh*gw, C*P*P] (k = c*P*P + kh*P + kw): the B operand of the
 * patch-projection wgrad (dW = dpatch^T * patches). */
/* ------------------------------------------------------------------------------------------------------------------
 * Image pre-processing of the ViLT processor: per-image bicubic resize in uint8 -- bit-identical to
 * Pillow's ImagingResample as called by transformers==4.48.0 image_transforms.resize(resample=BICUBIC) from
 * HF:models/vilt/image_processing_vilt.py -- then rescale 1/255 + normalise through lut256 (fp32 [3,256], one row per channel), zero padding to [Hmax, Wmax] and the
 * pixel mask.  The host computes the output sizes and the fixed-point filter taps (vault_b200/image_processing.py).
 *   src: packed uint8 HWC RGB images; descs[B] (device); coefs / bounds: int32 tap tables (Pillow precompute_coeffs +
 *   normalize_coeffs_8bpc, 22 fractional bits; bounds = {first source index, tap count} per output index); tmp: uint8 scratch for
 *   the horizontal pass; pixel_values [B,3,Hmax,Wmax] fp32 and pixel_mask [B,Hmax,Wmax] int64 are fully overwritten.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t src_off, tmp_off;             /* byte offsets of this image in src / tmp */
  int32_t h_in, w_in, h_out, w_out;
  int32_t ksize_h, ksize_v;             /* taps per output column / row (row stride of the coefficient tables) */
  int32_t coef_h, bound_h, coef_v, bound_v; /* element offsets into coefs / bounds */
} vault_image_desc;
int vault_image_preprocess(const uint8_t* src, const vault_image_desc* descs, const int32_t* coefs, const int32_t* bounds, uint8_t* tmp,
                           const float* lut256, float* pixel_values, int64_t* pixel_mask, int32_t B, int32_t Hmax, int32_t Wmax,
                           int32_t max_h_in, int32_t max_w_out, void* stream);

/* Pre-embedded image tokens (ViltEmbeddings.forward with image_embeds=..., HF:models/vilt/modeling_vilt.py:196-201; the TomViLT
 * path ref:vault/models/tomvilt/model.py:281-287): X [B,T+P,H] = [text_ln + modality[0] | image_embeds + modality[img_type]] (no CLS, no
 * position table), key_mask = [attention_mask | image_mask (uint8 [B,P], NULL = all valid)].  Backward: dtext_ln / dimage_embeds are
 * overwritten (either may be NULL), dmodality rows 0 and img_type are accumulated. */
int vault_vilt_assemble_embeds_fwd(const float* text_ln, const float* image_embeds, const float* modality, const int64_t* attention_mask,
                                   const uint8_t* image_mask, float* X, uint8_t* key_mask, int32_t B, int32_t T, int32_t P, int32_t H,
                                   int32_t img_type, void* stream);
int vault_vilt_assemble_embeds_bwd(const float* dX, float* dtext_ln, float* dimage_embeds, float* dmodality, int32_t B, int32_t T, int32_t P,
                                   int32_t H, int32_t img_type, void* stream);

int vault_patchify_bf16(const float* pixels, void* out_bf16, int32_t B, int32_t C, int32_t Hi, int32_t Wi, int32_t P,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Pooler + TMSC head + loss (ViltPooler HF:models/vilt/modeling_vilt.py:663-675; VaultForTMSC classifier
 * ref:vault/models/vault/model.py:547-550,569; CE mean ref:vault/tmsc_utils/trainer.py:228-242).  All fp32.
 * small_linear: y[r,n] = act(sum_k x[r*ldx + k] W[n,k] + b[n]); act 0 none, 1 tanh
 * ------------------------------------------------------------------------------------------------------------------ */
int vault_small_linear_fwd(const float* x, int64_t ldx, const float* W, const float* b, float* y, int32_t rows, int32_t N,
                           int32_t K, int32_t act, void* stream);
/* dy [rows,N]; if act==1, y is the saved tanh output and dy is multiplied by (1-y^2) first.
 * dx[r*lddx + k] (+)= sum_n dy W[n,k] (accumulate_dx: add into existing), dW [N,K] = dy^T x, db [N] = colsum(dy): overwritten */
int vault_small_linear_bwd(const float* dy, const float* y, const float* x, int64_t ldx, const float* W, float* dx,
                           int64_t lddx, int32_t accumulate_dx, float* dW, float* db, int32_t rows, int32_t N, int32_t K,
                           int32_t act, void* stream);
/* dropout on fp32 [n] (head dropout): y = x * keep / (1-p); linear, so the same call on dy is its backward */
int vault_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                      void* stream);
/* softmax cross-entropy, mean over rows: loss[0] = mean_r(-log softmax(logits[r])[label[r]]);
 * dlogits = (softmax - onehot) * grad_scale / rows  (dlogits may be NULL) */
int vault_ce_loss(const float* logits, const int64_t* labels, float* loss, float* dlogits, int32_t rows, int32_t n_classes,
                  float grad_scale, void* stream);
/* All three losses of the reference's trainers behind one call (mean over rows; dlogits optional):
 *   kind 0: cross-entropy, labels int64 [rows]                (ref:vault/tmsc_utils/trainer.py:228-242; MVSA pre-processed)
 *   kind 1: BCE-with-logits, n_classes = 1, labels fp32 [rows] (ref:vault/models/vault/trainer.py:42-56, Bloomberg)
 *   kind 2: 0.5 * (CE(first half, labels[:,0]) + CE(second half, labels[:,1])), labels int64 [rows,2]   (ref :114-137, MVSA raw) */
int vault_head_loss(const float* logits, const void* labels, float* loss, float* dlogits, int32_t rows, int32_t n_classes, int32_t kind,
                    float grad_scale, void* stream);

/* column sums of a bf16 [rows, cols] matrix accumulated into fp32 out[cols] (bias gradients); caller zero-fills */
int vault_colsum_bf16(const void* x, int64_t ldx, float* out, int64_t rows, int32_t cols, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * One fused AdamW step over a flat fp32 parameter range with the transformers==4.48.0 rule
 * (transformers.optimization.AdamW, used by ref:vault/tmsc_utils/trainer.py:244-254 with correct_bias=False):
 *   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= step_size * m / (sqrt(v) + eps) ; p -= lr*wd*p (if wd > 0)
 *   step_size = lr * sqrt(1-b2^t)/(1-b1^t) if correct_bias else lr.      grad_scale multiplies g first (DP averaging).
 * Also refreshes the bf16 shadow copy used by the GEMMs (shadow may be NULL).
 * sched_dev (optional, DEVICE float[2] = {step_size, lr*weight_decay}) overrides the host-computed scalars at run time so a
 * captured CUDA graph can follow the linear-warmup schedule (HF:optimization.py:101-131) without re-capture.
 * ------------------------------------------------------------------------------------------------------------------ */
/* g: fp32 gradients, or bf16 if grad_is_bf16 (the data-parallel path all-reduces a bf16 copy of each gradient range) */
int vault_adamw_step(float* p, const void* g, int32_t grad_is_bf16, float* m, float* v, void* shadow_bf16, int64_t n, double lr,
                     double beta1, double beta2, double eps, double weight_decay, int32_t correct_bias, int32_t step,
                     float grad_scale, const float* sched_dev, void* stream);
/* Data-parallel form of the step above as ONE kernel over NVSwitch multicast (replaces NCCL all-reduce + a full-range AdamW per gradient
 * range; optimizer rule ref:vault/tmsc_utils/trainer.py:244-254, data parallelism SURVEY.md 8e).  The caller owns slice [0, n) of a
 * finished gradient range (1/world of it, n % 8 == 0); every pointer is already offset to that slice:
 *   g_mc      multicast address of the gradients, fp32 or (grad_is_bf16) a bf16 copy: multimem.ld_reduce returns the SUM over all ranks,
 *             added in the switch with fp32 accumulation
 *   p_local   this rank's fp32 masters;  p_mc  multicast address of the same range: the new masters are stored to every replica, except
 *             in the 64-parameter blocks whose bit is set in local_only_bits (DEVICE bitmap over the whole flat buffer, block b = parameters
 *             [64b, 64b+64), or NULL = none; first_param = flat index of the slice's first parameter): those masters stay SHARDED -- a
 *             rank's copy is current only for the slices it owns, vault_mc_broadcast_f32 consolidates them.  Meant for the dense
 *             projection matrices, which the forward / backward only ever read through the bf16 shadow
 *   shadow_mc multicast address of the bf16 shadow: always stored to EVERY replica (the next forward reads it)
 *   m, v      this rank's AdamW moments of the slice (only the owner of a slice ever touches them)
 * The flat buffers must be symmetric memory mapped into one multicast object (torch.distributed._symmetric_memory).  Ordering across
 * ranks is the caller's: a barrier before the launch (all ranks' gradients of the range are final) and one after the last launch of
 * the step (all slices written everywhere).  `ctas` bounds the grid (the kernel is NVLink-latency bound; two CTAs fit on an SM). */
int vault_mc_adamw_step(float* p_local, float* p_mc, const uint32_t* local_only_bits, int64_t first_param, const void* g_mc,
                        int32_t grad_is_bf16, float* m, float* v, void* shadow_mc, int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay, int32_t correct_bias,
                        int32_t step, float grad_scale, const float* sched_dev, int32_t ctas, void* stream);
/* src_local fp32 [n] -> the same range of every replica through its multicast address (n % 4 == 0) */
int vault_mc_broadcast_f32(const float* src_local, float* dst_mc, int64_t n, int32_t ctas, void* stream);
/* fp32 -> bf16 cast of a flat range (shadow refresh after an external optimizer touched the masters) */
int vault_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAULT_B200_H_ */
