/* vitcap_b200 -- C ABI of the B200-native ViTCAP caption-generation hot path.
 *
 * The reference (jacobswan1/ViTCAP) is 100 % Python/PyTorch: it has no FFI of its own. Each entry point below
 * therefore names the reference *operator* (file:line under /root/reference) whose ATen call sequence it replaces;
 * a Python host (vitcap_b200/ops.py, via ctypes) binds them exactly as listed. INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a CUDA device pointer owned by the caller; kernels never allocate; nothing is retained
 *     after the call returns except cached TMA descriptors keyed by (pointer, shape)
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it, no host sync, graph-capturable
 *   - return 0 on success, negative on error (VC_ERR_*); vc_last_error() gives the message (thread local)
 *   - `bf16` flags select the storage/operand type of activations and weights: 1 = bfloat16 operands on the
 *     tcgen05 tensor cores (fast mode), 0 = fp32 operands on CUDA cores (exact mode). Accumulation, LayerNorm
 *     statistics, softmax and the residual stream are fp32 in both modes.
 *   - matrices are row-major; `ld*` are row pitches in elements
 */
#ifndef VITCAP_B200_H
#define VITCAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC_OK 0
#define VC_ERR_BAD_ARG (-1)
#define VC_ERR_UNSUPPORTED (-2)
#define VC_ERR_LAUNCH (-3)
#define VC_ERR_DRIVER (-4)

#define VC_ACT_NONE 0
#define VC_ACT_GELU 1 /* exact erf GELU: activations.py:16-23, nn.GELU in vision_transformer.py:143 */
#define VC_ACT_TANH 2 /* BertPooler, modeling_bert.py:524-527 */

const char* vc_last_error(void);
int vc_abi_version(void);
/* number of kernels launched through this library since the last vc_reset_launch_count() (bench.py gpu_launches) */
long long vc_launch_count(void);
void vc_reset_launch_count(void);
/* programmatic dependent launch between consecutive kernels of this library (griddepcontrol: launch latency and prologue of
 * kernel i+1 hide under kernel i; every kernel waits for its predecessor's completion before touching global memory).
 * mode 0 = never, 1 = launches captured into a CUDA graph only (default: the decode loop), 2 = every launch; environment
 * VITCAP_PDL presets it; takes effect for launches (and graph captures) made afterwards. */
void vc_set_pdl(int mode);
int vc_get_pdl(void);
/* VC_OK when the current device can run this library (compute capability 10.x: the kernels are sm_100a-only tcgen05 / TMEM /
 * TMA code), else VC_ERR_UNSUPPORTED with a message. Every compute entry point performs the same check and returns
 * VC_ERR_UNSUPPORTED instead of failing late inside a launch. (The reference has no counterpart: torch dispatches per device.) */
int vc_check_device(void);

/* out[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) (+ resid[M,N]);  replaces torch.nn.functional.linear behind
 *   vision_transformer.py:152-158 (Mlp fc1/fc2), :169-201 (Attention qkv/proj), :267-275 (PatchEmbed conv as GEMM),
 *   modeling_bert.py:307-313 (query/key/value), :353-357, :402-405, :415-419 (dense layers), :524-563 (pooler, heads).
 * bf16=1: A, W bfloat16, tcgen05.mma kind::f16 (cta_group::2 on CTA pairs for large M) with TMA-fed 128B-swizzled smem
 *         tiles, fp32 accumulator in TMEM
 *         (requires K % 64 == 0, 16-byte aligned pointers, pitches % 8 == 0).
 * bf16=0: A, W fp32, CUDA-core FFMA tiles (K % 16 == 0).
 * out_f32: output element type (1 = fp32, 0 = bf16 in fast mode / fp32 in exact mode is selected by the caller).
 * bias may be NULL; resid (fp32, pitch ldr) may be NULL and may alias out when out_f32 = 1.
 * bf16=1 stores tiles with TMA, which clips at 16-byte granularity: if N * sizeof(out element) is not a multiple of 16,
 * the pad columns up to the next 16-byte boundary of each row (< ldo) are overwritten.
 * bf16=1 treats W as a weight: its first tiles are fetched before the kernel waits for its predecessor in the stream, so with
 * programmatic dependent launch active (vc_set_pdl) W must not be written by the kernel launched immediately before. */
int vc_linear(int bf16, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32,
              int act, const float* resid, int ldr, int M, int N, int K, void* stream);
/* same contract, forcing the CUDA-core kernel for bf16 operands (cross-check of the tensor-core kernel in tests) */
int vc_linear_simt(int in_bf16, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo,
                   int out_f32, int act, const float* resid, int ldr, int M, int N, int K, void* stream);
/* tcgen05 kernel with an explicit N-tile width (64/128/256 = single-CTA kernel, 512 = the CTA-pair kernel with 256x256
 * tiles; 0 = heuristic: CTA pairs when the problem has at least two rounds of pair tiles) -- for tests and tuning */
int vc_linear_tc(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32, int act,
                 const float* resid, int ldr, int M, int N, int K, int tile_n, void* stream);

/* LayerNorm folded around two GEMMs of a pre-LN block (Block.forward, vision_transformer.py:233-250: x + mlp(norm2(x)) followed
 * by the next block's attn(norm1(x))): a stand-alone LayerNorm pass re-reads the 0.9 GB fp32 stream.
 * vc_linear_ln_emit = vc_linear(bf16, fp32 out, residual) that ALSO stores xb = bf16(out) [M,N] and
 *   stats[M, ceil(N/256), 2] = per-256-column partial (sum, sum of squares) of every fp32 output row.   N % 64 == 0.
 * vc_linear_ln_fold: out(bf16)[M,N] = act(LayerNorm(x)[M,K] * W[N,K]^T + b) evaluated from the raw copy A = xb and those
 *   statistics (st_tiles partials per row):  rstd * (xb * Wf^T - mean * colsum) + bias_f  with  Wf = bf16(gamma o W),
 *   colsum[n] = sum_k Wf[n,k] (fp32), bias_f = b + W beta, prepared by the caller.  act: VC_ACT_NONE / VC_ACT_GELU. */
int vc_linear_ln_emit(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo, const float* resid,
                      int ldr, void* xb, int ldxb, float* stats, int M, int N, int K, void* stream);
/* vc_linear_ln_emit for POST-LN layers (BertSelfOutput / BertOutput, modeling_bert.py:353-357, 415-419: LayerNorm(dense(h) + x)
 * where x is itself the output of the previous LayerNorm): that LayerNorm output is never materialised. resid_raw [M, N] is the
 * RAW fp32 row the previous emit produced, rstats [M, rst_tiles, 2] its partial sums, rgamma / rbeta / r_eps its LayerNorm;
 * the epilogue adds (resid_raw - mean) * rstd * rgamma + rbeta. Together with vc_linear_ln_fold on the consumer side no
 * stand-alone LayerNorm pass (2.3 GB per pass at 512 images) remains in the decoder prefill. */
int vc_linear_ln_emit_postln(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo,
                             const float* resid_raw, int ldr, const float* rstats, int rst_tiles, const float* rgamma,
                             const float* rbeta, float r_eps, void* xb, int ldxb, float* stats, int M, int N, int K, void* stream);
int vc_linear_ln_fold(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum, const float* stats,
                      int st_tiles, float ln_eps, void* out, int ldo, int act, int M, int N, int K, void* stream);

/* image fp32 [B,3,S,S] -> patch matrix [B*(S/p)^2, 3*p*p] (column order = Conv2d weight.flatten(1));
 * PatchEmbed.forward, vision_transformer.py:267-275 */
int vc_patchify(int bf16, const float* image, void* out, int B, int img_size, int patch, void* stream);
/* same from 8-bit pixels: image uint8 [B,S,S,3] (HWC, channel order BGR if bgr else RGB) with the tail of the reference's test
 * transform fused in: BGR2RGB, ToTensor (x/255, HWC->CHW), Normalize(0.5, 0.5) -- uni_pipeline.py:1233-1256, transform.py:47-50 --
 * bit-identical to the fp32 path fed with the host-transformed tensor; resize / center-crop stay on the host. */
int vc_patchify_u8(int bf16, const uint8_t* image, void* out, int B, int img_size, int patch, int bgr, void* stream);
/* Head of the same test transform on the device, for a ragged batch of decoded 8-bit images (HWC, 3 channels, order kept):
 * Resize(resize_to, BICUBIC) of the shorter edge + CenterCrop(crop) -- uni_pipeline.py:1246-1250 (crop_pct 1.0 in the shipped
 * yaml => resize_to == crop). Bit-identical to torchvision's output geometry and Pillow's 8-bit ImagingResample (the two
 * third-party libraries the reference calls there): double-precision Keys-cubic windows -> 22-bit fixed point, horizontal pass
 * into an 8-bit intermediate, vertical pass, saturating shift. Only the cropped window is computed.
 *   vc_resize_crop_plan  (HOST function, no device work): hw = host int32 [B][2] (height, width). Returns the widest window
 *                        `kmax`, the tallest image `max_rows` and tmp_off host int64 [B+1] (byte offsets of each image's
 *                        intermediate rows; tmp_off[B] = bytes of `tmp`). Fails if an image would resize below the crop.
 *   vc_resize_crop_u8    src = packed source pixels, src_off device int64 [B] byte offsets, hw device int32 [B][2],
 *                        coef = workspace int32 [B][2][kmax+2][crop], tmp = workspace of tmp_off[B] bytes (4-byte aligned),
 *                        tmp_off device int64 [B], out uint8 [B, crop, crop, 3] (feeds vc_patchify_u8). crop % 4 == 0. */
int vc_resize_crop_plan(const int* hw, int B, int resize_to, int crop, int* kmax, long long* tmp_off, int* max_rows);
int vc_resize_crop_u8(const uint8_t* src, const long long* src_off, const int* hw, int B, int resize_to, int crop, int kmax,
                      int max_rows, int* coef, uint8_t* tmp, const long long* tmp_off, uint8_t* out, void* stream);
/* x[b,0] = cls + pos[0]; x[b,1+i] = patch_out[b*P+i] + pos[1+i]; forward_features, vision_transformer.py:423-427 */
int vc_assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* x, int B, int P, int H, void* stream);

/* LayerNorm over the last dim of fp32 rows (nn.LayerNorm eps 1e-6 in ViT blocks, vision_transformer.py:218/229/352;
 * eps 1e-12 in BERT layers/heads, modeling_bert.py:219/350/412/538). out_t: operand copy (bf16 or fp32), may be NULL;
 * out_f: fp32 copy, may be NULL.
 * bf16 = 2 (VC_OPERAND_BF16X3): out_t is the split-bf16 operand [hi | lo | hi], 3H columns (ld_t >= 3H), hi = bf16(y),
 * lo = bf16(y - hi). Multiplied on the tensor cores against the weight laid out as [w_hi | w_hi | w_lo] (K' = 3K) it gives
 * y_hi w_hi + y_lo w_hi + y_hi w_lo with fp32 accumulation: ~2^-16 relative operand error instead of bf16's 2^-9, the
 * precision the decode-step Linear layers of the fast mode need for the reference's argmax decisions
 * (modeling_utils.py:815-822). The first H columns are the plain bf16 copy. */
#define VC_OPERAND_F32 0
#define VC_OPERAND_BF16 1
#define VC_OPERAND_BF16X3 2
int vc_layernorm(int bf16, const float* in, int ld_in, const float* gamma, const float* beta, float eps, void* out_t, int ld_t,
                 float* out_f, int ld_f, int rows, int H, void* stream);
/* fp32 rows [rows, K] (pitch ld_in) -> the same split-bf16 operand [rows, 3K] (pitch ld_out >= 3K); K % 8 == 0.
 * Used for the GELU output of BertIntermediate (modeling_bert.py:395-407) on its way into BertOutput.dense. */
int vc_split_bf16x3(const float* in, int ld_in, void* out, int ld_out, int rows, int K, void* stream);
/* out(fp32)[M,N] = A W^T + bias + resid for split operands A3 = [a_hi | a_lo | a_hi] [M, K3], W3 = [w_hi | w_hi | w_lo] [N, K3]
 * (K3 = 3K, K % 64 == 0): the same three products as vc_linear(A3, W3, K3) up to fp32 summation order, but the four distinct
 * tiles of a k-block are loaded once and multiplied three ways, 2/3 of the operand bytes per SM. BertOutput.dense of a decode
 * step (modeling_bert.py:409-419: K = 3072, N = 768), whose launch is bound by the L2 -> SM feed of its 96 CTAs. */
int vc_linear_x3(const void* A3, int lda, const void* W3, int ldw, const float* bias, float* out, int ldo, const float* resid,
                 int ldr, int M, int N, int K3, void* stream);
/* out[r,:] = cast(in[r*row_stride : +H]) -- e.g. hidden_states[:, 0] of BertPooler (modeling_bert.py:524) */
int vc_gather_rows(int bf16, const float* in, size_t row_stride, void* out, int ld_out, int rows, int H, void* stream);
/* ctx[b] = [tag_feats[b,0] ; cap_feats[b,0..N-1]] (modeling_bert.py:1493), fp32 copy + operand copy */
int vc_assemble_ctx(int bf16, const float* cap, const float* tag, float* ctx_f, void* ctx_t, int B, int N, int H, void* stream);
/* the same with rows_per_image >= N + 1 context rows allocated per image (the label rows below follow the N + 1 rows) */
int vc_assemble_ctx_pitched(int bf16, const float* cap, const float* tag, float* ctx_f, void* ctx_t, int B, int N, int H,
                            int rows_per_image, void* stream);
/* od/tag label rows of the context, for callers whose attention mask makes the label region visible (dataset.py:405-408):
 * row (b, i), i < K, = word[tag_i]                                  recipe_ln = 0  (modeling_bert.py:1447-1470)
 *                    = LN(word[tag_i] + pos[pos0 + i] + type0)      recipe_ln = 1  (encode_tag_to_embedding, :1381-1406, :1472-1489)
 * with tag_i = tag_idx[b, i] (int32 [B,K], the sorted top-K concepts) and the last slot forced to sep_id (:1447 / :1477).
 * Written to context rows row0 .. row0 + K - 1 of image b in ctx_f (fp32) and ctx_t (operand copy). */
int vc_label_rows(int bf16, const int* tag_idx, int K, int sep_id, int recipe_ln, int pos0, const float* word, const float* pos,
                  const float* type0, const float* gamma, const float* beta, float eps, float* ctx_f, void* ctx_t, int B,
                  int rows_per_image, int row0, int H, void* stream);

/* softmax(Q K^T * scale) V for packed qkv [B,N,3*heads*64] -> out [B,N,heads*64]; scores stay on chip.
 * Attention.forward vision_transformer.py:174-200 (mask is all-zero, modeling_bert.py:1415) and BertSelfAttention
 * modeling_bert.py:303-340 over the context rows. bf16=1: tcgen05 flash kernel; bf16=0: CUDA-core kernel. */
int vc_attention(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, void* stream);
int vc_attention_simt(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, void* stream);
/* the same under the label-region mask of the seq2seq family (dataset.py:395-408 + the image rows of
 * tagger_caption_uni_pipeline_expanding_bertemb.py:57-85): rows below n_base (tag-CLS + image tokens) see the keys below
 * n_base only; rows from n_base on (label rows) also see the first n_extra[b] label keys. n_extra: int32 [B], required. */
int vc_attention_labels(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, int n_base, const int* n_extra,
                        void* stream);
int vc_attention_labels_simt(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, int n_base,
                             const int* n_extra, void* stream);

/* single-query attention: out[b] = softmax(q[b] K_b^T * scale) V_b with K, V taken from the packed qkv [B,N,3H] and ONE query
 * row per image (q [B,H], pitch ldq). Used for the last block of the concept branch, of which only the CLS row is consumed
 * (pooler, modeling_bert.py:1424; tag token of the context, modeling_bert.py:1493; Attention.forward vision_transformer.py:174-200
 * restricted to query row 0). */
int vc_cls_attention(int bf16, const void* q, int ldq, const void* qkv, void* out, int ldo, int B, int N, int heads, float scale,
                     void* stream);

/* concept head selection: sigmoid -> topk(K) sorted desc -> count(prob >= thresh); modeling_bert.py:1429-1432 */
int vc_tag_topk(const float* logits, int ld, int rows, int V, int K, float thresh, int* out_idx, float* out_prob, int* out_len,
                void* stream);

/* decode-step text embedding, BertEmbeddings.forward modeling_bert.py:222-237 for rows [last token @ cur_len-1, MASK @ cur_len]
 * of each of R sequences; ids int32 [R,max_len]; out_f fp32 [2R,H], out_t operand copy */
int vc_embed_ln(int bf16, const int* ids, int max_len, int cur_len, int mask_id, const float* word, const float* pos,
                const float* type0, const float* gamma, const float* beta, float eps, float* out_f, void* out_t, int R, int H,
                void* stream);

/* one decoder self-attention step over the KV cache (BertSelfAttention with history, modeling_bert.py:303-340; mask
 * semantics of modeling_bert.py:1494-1501 / dataset.py:371-390 encoded structurally). ctx_qkv [B,C,3H]: prefill QKV of the
 * context rows; step_qkv [max_len, 2*B*E, 3H]: per-step QKV rows (2r = token, 2r+1 = MASK); anc int32 [max_len, B*E]
 * ancestor rows for beam search (NULL = identity); out [2*B*E, H]. E = beams*samples per image.
 * bf16=1: cp.async-pipelined mma.sync kernel, one CTA per (image, head) covering up to 8 sequences (16 query rows), so the
 * shared context K/V is read once per image; bf16=0: CUDA-core kernel on fp32 storage. */
int vc_decode_attention(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C, int heads,
                        int E, int cur_len, float scale, void* stream);
/* same contract on the CUDA-core kernel for either storage type (bf16=0 is what vc_decode_attention runs in exact mode;
 * bf16=1 cross-checks the mma.sync kernel in tests) */
int vc_decode_attention_simt(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C,
                             int heads, int E, int cur_len, float scale, void* stream);
/* the same with rows_per_image context rows allocated per image of which only the first ctx_vis[b] (int32 [B]) are visible:
 * the caption rows of image b see [tag-CLS | image tokens | its visible label rows] (C-L block of dataset.py:407-408) */
int vc_decode_attention_labels(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B,
                               int rows_per_image, const int* ctx_vis, int heads, int E, int cur_len, float scale, void* stream);
int vc_decode_attention_labels_simt(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B,
                                    int rows_per_image, const int* ctx_vis, int heads, int E, int cur_len, float scale,
                                    void* stream);
/* fast mode (bf16 K/V): vc_decode_attention_labels (ctx_vis may be NULL) that does no work for finished captions. The reference
 * keeps computing every row until ALL sequences of the batch have finished (modeling_utils.py:858-867; beam search: until every
 * image is done, :1003-1011, :1096) and discards what it computed for the finished ones; here a sequence with
 * seq_unfinished[b * E + e] == 0 (int32 [B * E], may be NULL) or an image with img_done[b] != 0 (int32 [B], may be NULL; takes
 * precedence) is skipped before its K/V rows are read, and its output rows are left untouched (stale but finite). */
int vc_decode_attention_skip(const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int rows_per_image,
                             const int* ctx_vis, int heads, int E, int cur_len, float scale, const int* seq_unfinished,
                             const int* img_done, void* stream);

/* greedy / sampled next token + log-prob + state update for `rows` sequences, modeling_utils.py:839-862.
 * logits fp32 [rows, ld]; sampling = Gumbel-max with Philox4x32-10 noise keyed by (seed; vocab idx/4, row, cur_len).
 * seed_dev: NULL, or a device pointer to the 64-bit seed, read when the kernel RUNS (it then replaces `seed`): a decode loop
 * captured once in a CUDA graph draws fresh noise on every replay after the caller rewrites that word. */
int vc_token_step(const float* logits, int ld, int rows, int V, int do_sample, float temperature, uint64_t seed,
                  const uint64_t* seed_dev, int cur_len, int max_len, int pad_id, const int* eos_ids, int n_eos, int* ids,
                  int* unfinished, float* sum_lp, int* n_steps, void* stream);
/* ---- fused decode step (fast mode): the Linear layers of a decode step on CTA-pair tiles sized for M = 2 x sequences rows.
 * They replace the same reference operators as vc_linear (modeling_bert.py:307-313 q/k/v, :353-357 BertSelfOutput.dense,
 * :402-405 BertIntermediate, :415-419 BertOutput.dense, :524-563 prediction head); what differs is the boundary between kernels.
 * A [M, K] and W [N, K] 16-bit, row pitches lda / ldw. Operand format `fmt`:
 *   VC_DEC_FMT_BF16   bf16 x bf16
 *   VC_DEC_FMT_BF16X3 split operands, A = [a_hi | a_lo | (unused)] and W = [w_hi | w_hi | w_lo] along K (K = 3 Kt,
 *                     VC_OPERAND_BF16X3): out = a_hi w_hi + a_lo w_hi + a_hi w_lo, every distinct tile loaded once
 *   VC_DEC_FMT_F16    IEEE half x IEEE half (ABI 9): 11-bit significands in ONE product at the rate and bytes of bf16; the
 *                     range (6e-8 .. 65504) holds LayerNorm / GELU outputs and weights, conversions saturate
 *   mode VC_DEC_PARTIAL    out = fp32 [splits, m_pad, N] (pitch ldo): plane s = A W^T over the s-th slice of K, NO bias
 *                          (split-K over the SMs a 2-sequence-row GEMM leaves idle; vc_finish_ln sums the planes).
 *                          splits == 1 accepts a bias and any N: out = fp32 [M, N] = A W^T + bias, the materialised
 *                          vocabulary logits of beam search / sampling (the store works in 16-byte units: columns
 *                          [N, round_up(N, 4)) of a row, which the pitch ldo % 4 == 0 guarantees to exist, may receive zeros)
 *   mode VC_DEC_BF16       out = bf16 (A W^T + bias)
 *   mode VC_DEC_GELU_BF16  out = GELU(A W^T + bias) as bf16 (as IEEE halves with VC_DEC_FMT_F16: the next GEMM's operand)
 *   mode VC_DEC_GELU_SPLIT out = bf16 [M, >= 2N]: columns [0, N) = hi, [N, 2N) = lo of GELU(A W^T + bias) (the split operand
 *                          of the next x3 GEMM, written by the epilogue instead of an fp32 round trip + vc_split_bf16x3)
 * splits must divide K / 64 (Kt / 64) and be 1 for the bf16 modes; m_pad >= M, multiple of 128. */
#define VC_DEC_PARTIAL 0
#define VC_DEC_BF16 1
#define VC_DEC_GELU_BF16 2
#define VC_DEC_GELU_SPLIT 3
#define VC_DEC_FMT_BF16 0
#define VC_DEC_FMT_BF16X3 1
#define VC_DEC_FMT_F16 2
int vc_dec_linear(int mode, int fmt, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int M,
                  int N, int K, int splits, int m_pad, void* stream);
/* vocabulary projection of greedy decoding WITHOUT materialising the logits (BertLMPredictionHead.decoder + bias,
 * modeling_bert.py:551-563, then argmax -> log_softmax -> gather, modeling_utils.py:849-853): part = float4 [M, n_part],
 * n_part = 2 * ceil(N / 208); entry (row, 2 * tile + g) = (max, first arg max as int bits, sum of exp(x - max), 0) of
 * A W^T + bias over the 32-column chunks g, g + 2, ... of the 208-column tile (208: the 30522-word vocabulary then makes 3.97
 * rounds of tiles over the 74 CTA pairs). vc_token_step_partials reduces them. */
int vc_dec_vocab_argmax(int fmt, const void* A, int lda, const void* W, int ldw, const float* bias, void* part, int n_part, int M,
                        int N, int K, void* stream);
/* y = LayerNorm(act(sum_s part[s] + bias) + resid) per row: the split-K reduction, bias / residual of BertSelfOutput /
 * BertOutput (modeling_bert.py:353-357, 415-419), their LayerNorm -- or, with gelu = 1 and resid = NULL, the GELU + LayerNorm
 * of the prediction-head transform (modeling_bert.py:524-537) -- and the operand copy for the next GEMM in one pass.
 * part: `splits` fp32 planes [rows, ld_p], `plane` elements apart; out_f (fp32, may be NULL) the normalised row;
 * out_mode 0 = no operand copy, 1 = bf16 [rows, ld_t], 2 = split pair: columns [0, H) = hi, [H, 2H) = lo (ld_t >= 2H),
 * 3 = [hi | lo | hi] (ld_t >= 3H; the K-concatenated form vc_linear reads for VC_OPERAND_BF16X3), 4 = IEEE half [rows, ld_t],
 * 5 = [bf16 | half] (ld_t >= 2H: the q|k|v projection reads the bf16 columns, a VC_DEC_FMT_F16 GEMM the half columns). */
int vc_finish_ln(const float* part, int splits, size_t plane, int ld_p, const float* bias, int gelu, const float* resid, int ld_r,
                 const float* gamma, const float* beta, float eps, float* out_f, int ld_f, void* out_t, int ld_t, int out_mode,
                 int rows, int H, void* stream);
/* vc_token_step for greedy decoding from the partials of vc_dec_vocab_argmax (first index on ties, as torch.argmax) */
int vc_token_step_partials(const void* part, int n_part, int rows, int cur_len, int max_len, int pad_id, const int* eos_ids,
                           int n_eos, int* ids, int* unfinished, float* sum_lp, int* n_steps, void* stream);
/* modeling_utils.py:869-886: force EOS, mean log-prob, int64 ids [R,max_len] */
int vc_greedy_finalize(const int* ids, const int* unfinished, const float* sum_lp, const int* n_steps, int eos0, int max_len, int R,
                       long long* out_ids, float* out_lp, void* stream);

/* Early exit of a CAPTURED decode loop without a host round trip (modeling_utils.py:865-867 `if cur_unfinished.max() == 0:
 * break`, :1071-1073 `if all(done): break`). `capture_stream` must be capturing into a CUDA graph (e.g. torch.cuda.graph):
 * vc_graph_if_any_begin appends a one-block kernel that evaluates any(flags[i] != 0) (invert = 0: per-sequence `unfinished`)
 * or any(flags[i] == 0) (invert = 1: per-image `done`) over n int32 flags on the device, then a conditional IF node on that
 * value, and starts capturing `body_stream` (another stream, not capturing) into the IF node's body: everything the caller
 * launches on body_stream until vc_graph_if_end(body_stream) is skipped by a replay in which no flag is live.
 * Bodies hold kernel launches only (no allocations, no host-pageable copies). */
int vc_graph_if_any_begin(const int* flags, int n, int invert, void* capture_stream, void* body_stream);
int vc_graph_if_end(void* body_stream);
/* a non-blocking stream of the current device that belongs to the caller alone (a framework's pooled streams may alias the
 * capturing stream or carry other work): the body stream of vc_graph_if_any_begin. *stream_out receives the cudaStream_t. */
int vc_stream_create(void** stream_out);
int vc_stream_destroy(void* stream);

/* beam search step, modeling_utils.py:988-1065: (1) per-row log-sum-exp + top-(2*beams) logits, (2) per-image candidate
 * walk, hypothesis pool (BeamHypotheses, :1138-1180), next beams, ancestor table update. */
int vc_beam_row_topk(const float* logits, int ld, int rows, int V, int K, float* cand_val, int* cand_idx, float* row_max,
                     float* row_logsum, void* stream);
int vc_beam_advance(int* ids, float* beam_scores, int* done, int* anc, double* hyp_score, int* hyp_len, int* hyp_ids,
                    int* hyp_count, double* worst, const float* cand_val, const int* cand_idx, const float* row_max,
                    const float* row_logsum, int B, int num_beams, int V, int cur_len, int max_len, int keep,
                    double length_penalty, int pad_id, const int* eos_ids, int n_eos, void* stream);
/* modeling_utils.py:1074-1100 */
int vc_beam_finalize(const double* hyp_score, const int* hyp_len, const int* hyp_ids, const int* hyp_count, int B, int keep,
                     int max_len, int pad_id, int eos0, long long* out_ids, float* out_lp, void* stream);

/* top_k_top_p_filtering, modeling_utils.py:1103-1135, in place on fp32 logits (after the 1/temperature scaling) */
int vc_filter_logits(float* logits, int ld, int rows, int V, float inv_temperature, int top_k, float top_p,
                     int min_tokens_to_keep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITCAP_B200_H */
