// extern "C" surface of libvitcap_b200.so (declared in include/vitcap_b200.h). Pure marshalling: plain pointers and
// sizes in, error code out; every function forwards to one kernel launcher.
#include "common.cuh"
#include "../../include/vitcap_b200.h"

#include <atomic>

namespace vc {
const char* last_error();
int check_device();
int gemm_bf16_tc2_ln_emit(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo, const float* resid,
                          int ldr, void* xb, int ldxb, float* stats, int M, int N, int K, cudaStream_t stream,
                          const float* rstats = nullptr, int rst_tiles = 0, const float* rgamma = nullptr, const float* rbeta = nullptr,
                          float r_eps = 0.f);
int gemm_bf16_tc2_ln_fold(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum,
                          const float* stats, int st_tiles, float ln_eps, void* out, int ldo, int act, int M, int N, int K,
                          cudaStream_t stream);
int gemm_bf16_tc_x3(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo, const float* resid,
                    int ldr, int M, int N, int K3, cudaStream_t stream);
int gemm_bf16_tc(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32, int act,
                 const float* resid, int ldr, int M, int N, int K, int force_bn, cudaStream_t stream);
int gemm_simt(int in_bf16, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32,
              int act, const float* resid, int ldr, int M, int N, int K, cudaStream_t s);
int patchify(int out_bf16, const float* img, void* out, int B, int img_size, int patch, cudaStream_t s);
int patchify_u8(int out_bf16, const uint8_t* img, void* out, int B, int img_size, int patch, int bgr, cudaStream_t s);
int resize_crop_plan(const int* hw, int B, int resize_to, int S, int* kmax, long long* tmp_off, int* max_rows);
int resize_crop_u8(const uint8_t* src, const long long* src_off, const int* hw, int B, int resize_to, int S, int kmax, int max_rows,
                   int* coef, uint8_t* tmp, const long long* tmp_off, uint8_t* out, cudaStream_t s);
int assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* x, int B, int P, int H, cudaStream_t s);
int layernorm(int out_bf16, const float* in, int ld_in, const float* gamma, const float* beta, float eps, void* out_t, int ld_t,
              float* out_f, int ld_f, int rows, int H, cudaStream_t s);
int split_bf16x3(const float* in, int ld_in, void* out, int ld_out, int rows, int K, cudaStream_t s);
int gather_rows(int out_bf16, const float* in, size_t row_stride, void* out, int ld_out, int rows, int H, cudaStream_t s);
int assemble_ctx(int out_bf16, const float* cap, const float* tag, float* ctx_f, void* ctx_t, int B, int N, int H, int Cp,
                 cudaStream_t s);
int label_rows(int out_bf16, const int* tag_idx, int K, int sep_id, int recipe_ln, int pos0, const float* word, const float* pos,
               const float* type0, const float* gamma, const float* beta, float eps, float* ctx_f, void* ctx_t, int B, int Cp,
               int row0, int H, cudaStream_t s);
int attention_simt(int is_bf16, const void* qkv, void* out, int B, int N, int heads, float scale, int n_base, const int* n_extra,
                   cudaStream_t s);
int attention_tc(const void* qkv, void* out, int B, int N, int heads, float scale, int n_base, const int* n_extra, cudaStream_t s);
int cls_attention(int is_bf16, const void* q, int ldq, const void* qkv, void* out, int ldo, int B, int N, int heads, float scale,
                  cudaStream_t s);
int tag_topk(const float* logits, int ld, int rows, int V, int K, float thresh, int* out_idx, float* out_prob, int* out_len,
             cudaStream_t s);
int embed_ln(int out_bf16, const int* ids, int max_len, int cur_len, int mask_id, const float* word, const float* pos,
             const float* type0, const float* gamma, const float* beta, float eps, float* out_f, void* out_t, int R, int H,
             cudaStream_t s);
int decode_attention(int is_bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C,
                     const int* ctx_vis, int heads, int E, int cur_len, float scale, cudaStream_t s);
int decode_attention_simt(int is_bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C,
                          const int* ctx_vis, int heads, int E, int cur_len, float scale, cudaStream_t s);
int decode_attention_skip(const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C, const int* ctx_vis,
                          int heads, int E, int cur_len, float scale, const int* seq_unfinished, const int* img_done, cudaStream_t s);
int token_step(const float* logits, int ld, int rows, int V, int do_sample, float temperature, uint64_t seed,
               const uint64_t* seed_dev, int cur_len, int max_len, int pad_id, const int* eos_ids, int n_eos, int* ids,
               int* unfinished, float* sum_lp, int* n_steps, cudaStream_t s);
int greedy_finalize(const int* ids, const int* unfinished, const float* sum_lp, const int* n_steps, int eos0, int max_len, int R,
                    long long* out_ids, float* out_lp, cudaStream_t s);
int beam_row_topk(const float* logits, int ld, int rows, int V, int K, float* cand_val, int* cand_idx, float* row_max,
                  float* row_logsum, cudaStream_t s);
int beam_advance(int* ids, float* beam_scores, int* done, int* anc, double* hyp_score, int* hyp_len, int* hyp_ids, int* hyp_count,
                 double* worst, const float* cand_val, const int* cand_idx, const float* row_max, const float* row_logsum, int B,
                 int nb, int V, int cur_len, int max_len, int keep, double length_penalty, int pad_id, const int* eos_ids,
                 int n_eos, cudaStream_t s);
int beam_finalize(const double* hyp_score, const int* hyp_len, const int* hyp_ids, const int* hyp_count, int B, int keep, int max_len,
                  int pad_id, int eos0, long long* out_ids, float* out_lp, cudaStream_t s);
int filter_logits(float* logits, int ld, int rows, int V, float inv_temperature, int top_k, float top_p, int min_tokens_to_keep,
                  cudaStream_t s);
int gemm_dec(int mode, int x3, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int M, int N,
             int K, int splits, int m_pad, cudaStream_t stream);
int gemm_dec_argmax(int x3, const void* A, int lda, const void* W, int ldw, const float* bias, void* part, int n_part, int M, int N,
                    int K, cudaStream_t stream);
int finish_ln(const float* part, int splits, size_t plane, int ld_p, const float* bias, int gelu, const float* resid, int ld_r,
              const float* gamma, const float* beta, float eps, float* out_f, int ld_f, void* out_t, int ld_t, int out_mode, int rows,
              int H, cudaStream_t s);
int token_step_partials(const void* part, int n_part, int rows, int cur_len, int max_len, int pad_id, const int* eos_ids, int n_eos,
                        int* ids, int* unfinished, float* sum_lp, int* n_steps, cudaStream_t s);
int graph_if_any_begin(const int* flags, int n, int invert, cudaStream_t cap, cudaStream_t body);
int graph_if_end(cudaStream_t body);
}  // namespace vc

static std::atomic<long long> g_launches{0};
#define VC_COUNT(n, expr) do { int rc__ = vc::check_device(); if (rc__ == 0) rc__ = (expr); if (rc__ == 0) g_launches += (n); return rc__; } while (0)
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* vc_last_error(void) { return vc::last_error(); }
int vc_abi_version(void) { return 9; }
int vc_check_device(void) { return vc::check_device(); }
long long vc_launch_count(void) { return g_launches.load(); }
void vc_reset_launch_count(void) { g_launches = 0; }
void vc_set_pdl(int mode) { vc::set_pdl_mode(mode); }
int vc_get_pdl(void) { return vc::pdl_mode(); }

int vc_graph_if_any_begin(const int* flags, int n, int invert, void* capture_stream, void* body_stream) {
  VC_COUNT(1, vc::graph_if_any_begin(flags, n, invert, ST(capture_stream), ST(body_stream)));
}
int vc_graph_if_end(void* body_stream) { return vc::graph_if_end(ST(body_stream)); }
int vc_stream_create(void** stream_out) {
  cudaStream_t s = nullptr;
  if (stream_out == nullptr) { vc::set_last_error("vc_stream_create: NULL"); return VC_ERR_BAD_ARG; }
  const cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  if (e != cudaSuccess) { vc::set_last_error("vc_stream_create: %s", cudaGetErrorString(e)); cudaGetLastError(); return VC_ERR_DRIVER; }
  *stream_out = s;
  return VC_OK;
}
int vc_stream_destroy(void* stream) {
  if (stream != nullptr && cudaStreamDestroy(ST(stream)) != cudaSuccess) { cudaGetLastError(); return VC_ERR_DRIVER; }
  return VC_OK;
}

int vc_linear(int bf16, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32, int act,
              const float* resid, int ldr, int M, int N, int K, void* stream) {
  if (bf16) VC_COUNT(1, vc::gemm_bf16_tc(A, lda, W, ldw, bias, out, ldo, out_f32, act, resid, ldr, M, N, K, 0, ST(stream)));
  VC_COUNT(1, vc::gemm_simt(0, A, lda, W, ldw, bias, out, ldo, 1, act, resid, ldr, M, N, K, ST(stream)));
}
int vc_linear_ln_emit(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo, const float* resid,
                      int ldr, void* xb, int ldxb, float* stats, int M, int N, int K, void* stream) {
  VC_COUNT(1, vc::gemm_bf16_tc2_ln_emit(A, lda, W, ldw, bias, out, ldo, resid, ldr, xb, ldxb, stats, M, N, K, ST(stream)));
}
int vc_linear_ln_emit_postln(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo,
                             const float* resid_raw, int ldr, const float* rstats, int rst_tiles, const float* rgamma,
                             const float* rbeta, float r_eps, void* xb, int ldxb, float* stats, int M, int N, int K, void* stream) {
  if (rstats == nullptr) { vc::set_last_error("vc_linear_ln_emit_postln: rstats is NULL"); return VC_ERR_BAD_ARG; }
  VC_COUNT(1, vc::gemm_bf16_tc2_ln_emit(A, lda, W, ldw, bias, out, ldo, resid_raw, ldr, xb, ldxb, stats, M, N, K, ST(stream), rstats,
                                        rst_tiles, rgamma, rbeta, r_eps));
}
int vc_linear_ln_fold(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum, const float* stats,
                      int st_tiles, float ln_eps, void* out, int ldo, int act, int M, int N, int K, void* stream) {
  VC_COUNT(1, vc::gemm_bf16_tc2_ln_fold(A, lda, Wf, ldw, bias_f, colsum, stats, st_tiles, ln_eps, out, ldo, act, M, N, K, ST(stream)));
}
int vc_linear_simt(int in_bf16, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32,
                   int act, const float* resid, int ldr, int M, int N, int K, void* stream) {
  VC_COUNT(1, vc::gemm_simt(in_bf16, A, lda, W, ldw, bias, out, ldo, out_f32, act, resid, ldr, M, N, K, ST(stream)));
}
int vc_linear_tc(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32, int act,
                 const float* resid, int ldr, int M, int N, int K, int tile_n, void* stream) {
  VC_COUNT(1, vc::gemm_bf16_tc(A, lda, W, ldw, bias, out, ldo, out_f32, act, resid, ldr, M, N, K, tile_n, ST(stream)));
}
int vc_patchify(int bf16, const float* image, void* out, int B, int img_size, int patch, void* stream) {
  VC_COUNT(1, vc::patchify(bf16, image, out, B, img_size, patch, ST(stream)));
}
int vc_patchify_u8(int bf16, const uint8_t* image, void* out, int B, int img_size, int patch, int bgr, void* stream) {
  VC_COUNT(1, vc::patchify_u8(bf16, image, out, B, img_size, patch, bgr, ST(stream)));
}
int vc_resize_crop_plan(const int* hw, int B, int resize_to, int crop, int* kmax, long long* tmp_off, int* max_rows) {
  return vc::resize_crop_plan(hw, B, resize_to, crop, kmax, tmp_off, max_rows);
}
int vc_resize_crop_u8(const uint8_t* src, const long long* src_off, const int* hw, int B, int resize_to, int crop, int kmax,
                      int max_rows, int* coef, uint8_t* tmp, const long long* tmp_off, uint8_t* out, void* stream) {
  VC_COUNT(3, vc::resize_crop_u8(src, src_off, hw, B, resize_to, crop, kmax, max_rows, coef, tmp, tmp_off, out, ST(stream)));
}
int vc_assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* x, int B, int P, int H, void* stream) {
  VC_COUNT(1, vc::assemble_tokens(patch_out, cls, pos, x, B, P, H, ST(stream)));
}
int vc_layernorm(int bf16, const float* in, int ld_in, const float* gamma, const float* beta, float eps, void* out_t, int ld_t,
                 float* out_f, int ld_f, int rows, int H, void* stream) {
  VC_COUNT(1, vc::layernorm(bf16, in, ld_in, gamma, beta, eps, out_t, ld_t, out_f, ld_f, rows, H, ST(stream)));
}
int vc_linear_x3(const void* A3, int lda, const void* W3, int ldw, const float* bias, float* out, int ldo, const float* resid,
                 int ldr, int M, int N, int K3, void* stream) {
  VC_COUNT(1, vc::gemm_bf16_tc_x3(A3, lda, W3, ldw, bias, out, ldo, resid, ldr, M, N, K3, ST(stream)));
}
int vc_split_bf16x3(const float* in, int ld_in, void* out, int ld_out, int rows, int K, void* stream) {
  VC_COUNT(1, vc::split_bf16x3(in, ld_in, out, ld_out, rows, K, ST(stream)));
}
int vc_gather_rows(int bf16, const float* in, size_t row_stride, void* out, int ld_out, int rows, int H, void* stream) {
  VC_COUNT(1, vc::gather_rows(bf16, in, row_stride, out, ld_out, rows, H, ST(stream)));
}
int vc_assemble_ctx(int bf16, const float* cap, const float* tag, float* ctx_f, void* ctx_t, int B, int N, int H, void* stream) {
  VC_COUNT(1, vc::assemble_ctx(bf16, cap, tag, ctx_f, ctx_t, B, N, H, N + 1, ST(stream)));
}
int vc_assemble_ctx_pitched(int bf16, const float* cap, const float* tag, float* ctx_f, void* ctx_t, int B, int N, int H,
                            int rows_per_image, void* stream) {
  VC_COUNT(1, vc::assemble_ctx(bf16, cap, tag, ctx_f, ctx_t, B, N, H, rows_per_image, ST(stream)));
}
int vc_label_rows(int bf16, const int* tag_idx, int K, int sep_id, int recipe_ln, int pos0, const float* word, const float* pos,
                  const float* type0, const float* gamma, const float* beta, float eps, float* ctx_f, void* ctx_t, int B,
                  int rows_per_image, int row0, int H, void* stream) {
  VC_COUNT(1, vc::label_rows(bf16, tag_idx, K, sep_id, recipe_ln, pos0, word, pos, type0, gamma, beta, eps, ctx_f, ctx_t, B,
                             rows_per_image, row0, H, ST(stream)));
}
int vc_attention(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, void* stream) {
  if (bf16) VC_COUNT(1, vc::attention_tc(qkv, out, B, N, heads, scale, 0, nullptr, ST(stream)));
  VC_COUNT(1, vc::attention_simt(0, qkv, out, B, N, heads, scale, 0, nullptr, ST(stream)));
}
int vc_attention_simt(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, void* stream) {
  VC_COUNT(1, vc::attention_simt(bf16, qkv, out, B, N, heads, scale, 0, nullptr, ST(stream)));
}
int vc_attention_labels(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, int n_base, const int* n_extra,
                        void* stream) {
  if (n_extra == nullptr) return VC_ERR_BAD_ARG;
  if (bf16) VC_COUNT(1, vc::attention_tc(qkv, out, B, N, heads, scale, n_base, n_extra, ST(stream)));
  VC_COUNT(1, vc::attention_simt(0, qkv, out, B, N, heads, scale, n_base, n_extra, ST(stream)));
}
int vc_attention_labels_simt(int bf16, const void* qkv, void* out, int B, int N, int heads, float scale, int n_base,
                             const int* n_extra, void* stream) {
  if (n_extra == nullptr) return VC_ERR_BAD_ARG;
  VC_COUNT(1, vc::attention_simt(bf16, qkv, out, B, N, heads, scale, n_base, n_extra, ST(stream)));
}
int vc_cls_attention(int bf16, const void* q, int ldq, const void* qkv, void* out, int ldo, int B, int N, int heads, float scale,
                     void* stream) {
  VC_COUNT(1, vc::cls_attention(bf16, q, ldq, qkv, out, ldo, B, N, heads, scale, ST(stream)));
}
int vc_tag_topk(const float* logits, int ld, int rows, int V, int K, float thresh, int* out_idx, float* out_prob, int* out_len,
                void* stream) {
  VC_COUNT(1, vc::tag_topk(logits, ld, rows, V, K, thresh, out_idx, out_prob, out_len, ST(stream)));
}
int vc_embed_ln(int bf16, const int* ids, int max_len, int cur_len, int mask_id, const float* word, const float* pos,
                const float* type0, const float* gamma, const float* beta, float eps, float* out_f, void* out_t, int R, int H,
                void* stream) {
  VC_COUNT(1, vc::embed_ln(bf16, ids, max_len, cur_len, mask_id, word, pos, type0, gamma, beta, eps, out_f, out_t, R, H, ST(stream)));
}
int vc_decode_attention(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C, int heads,
                        int E, int cur_len, float scale, void* stream) {
  VC_COUNT(1, vc::decode_attention(bf16, ctx_qkv, step_qkv, anc, out, B, C, nullptr, heads, E, cur_len, scale, ST(stream)));
}
int vc_decode_attention_skip(const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int rows_per_image,
                             const int* ctx_vis, int heads, int E, int cur_len, float scale, const int* seq_unfinished,
                             const int* img_done, void* stream) {
  VC_COUNT(1, vc::decode_attention_skip(ctx_qkv, step_qkv, anc, out, B, rows_per_image, ctx_vis, heads, E, cur_len, scale,
                                        seq_unfinished, img_done, ST(stream)));
}
int vc_decode_attention_labels(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B,
                               int rows_per_image, const int* ctx_vis, int heads, int E, int cur_len, float scale, void* stream) {
  VC_COUNT(1, vc::decode_attention(bf16, ctx_qkv, step_qkv, anc, out, B, rows_per_image, ctx_vis, heads, E, cur_len, scale,
                                   ST(stream)));
}
int vc_decode_attention_labels_simt(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B,
                                    int rows_per_image, const int* ctx_vis, int heads, int E, int cur_len, float scale,
                                    void* stream) {
  VC_COUNT(1, vc::decode_attention_simt(bf16, ctx_qkv, step_qkv, anc, out, B, rows_per_image, ctx_vis, heads, E, cur_len, scale,
                                        ST(stream)));
}
int vc_decode_attention_simt(int bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C,
                             int heads, int E, int cur_len, float scale, void* stream) {
  VC_COUNT(1, vc::decode_attention_simt(bf16, ctx_qkv, step_qkv, anc, out, B, C, nullptr, heads, E, cur_len, scale, ST(stream)));
}
int vc_token_step(const float* logits, int ld, int rows, int V, int do_sample, float temperature, uint64_t seed,
                  const uint64_t* seed_dev, int cur_len, int max_len, int pad_id, const int* eos_ids, int n_eos, int* ids,
                  int* unfinished, float* sum_lp, int* n_steps, void* stream) {
  VC_COUNT(1, vc::token_step(logits, ld, rows, V, do_sample, temperature, seed, seed_dev, cur_len, max_len, pad_id, eos_ids, n_eos,
                             ids, unfinished, sum_lp, n_steps, ST(stream)));
}
int vc_greedy_finalize(const int* ids, const int* unfinished, const float* sum_lp, const int* n_steps, int eos0, int max_len, int R,
                       long long* out_ids, float* out_lp, void* stream) {
  VC_COUNT(1, vc::greedy_finalize(ids, unfinished, sum_lp, n_steps, eos0, max_len, R, out_ids, out_lp, ST(stream)));
}
int vc_beam_row_topk(const float* logits, int ld, int rows, int V, int K, float* cand_val, int* cand_idx, float* row_max,
                     float* row_logsum, void* stream) {
  VC_COUNT(1, vc::beam_row_topk(logits, ld, rows, V, K, cand_val, cand_idx, row_max, row_logsum, ST(stream)));
}
int vc_beam_advance(int* ids, float* beam_scores, int* done, int* anc, double* hyp_score, int* hyp_len, int* hyp_ids, int* hyp_count,
                    double* worst, const float* cand_val, const int* cand_idx, const float* row_max, const float* row_logsum, int B,
                    int num_beams, int V, int cur_len, int max_len, int keep, double length_penalty, int pad_id,
                    const int* eos_ids, int n_eos, void* stream) {
  VC_COUNT(1, vc::beam_advance(ids, beam_scores, done, anc, hyp_score, hyp_len, hyp_ids, hyp_count, worst, cand_val, cand_idx,
                               row_max, row_logsum, B, num_beams, V, cur_len, max_len, keep, length_penalty, pad_id, eos_ids, n_eos,
                               ST(stream)));
}
int vc_beam_finalize(const double* hyp_score, const int* hyp_len, const int* hyp_ids, const int* hyp_count, int B, int keep,
                     int max_len, int pad_id, int eos0, long long* out_ids, float* out_lp, void* stream) {
  VC_COUNT(1, vc::beam_finalize(hyp_score, hyp_len, hyp_ids, hyp_count, B, keep, max_len, pad_id, eos0, out_ids, out_lp, ST(stream)));
}
int vc_dec_linear(int mode, int fmt, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int M,
                  int N, int K, int splits, int m_pad, void* stream) {
  VC_COUNT(1, vc::gemm_dec(mode, fmt, A, lda, W, ldw, bias, out, ldo, M, N, K, splits, m_pad, ST(stream)));
}
int vc_dec_vocab_argmax(int fmt, const void* A, int lda, const void* W, int ldw, const float* bias, void* part, int n_part, int M,
                        int N, int K, void* stream) {
  VC_COUNT(1, vc::gemm_dec_argmax(fmt, A, lda, W, ldw, bias, part, n_part, M, N, K, ST(stream)));
}
int vc_finish_ln(const float* part, int splits, size_t plane, int ld_p, const float* bias, int gelu, const float* resid, int ld_r,
                 const float* gamma, const float* beta, float eps, float* out_f, int ld_f, void* out_t, int ld_t, int out_mode,
                 int rows, int H, void* stream) {
  VC_COUNT(1, vc::finish_ln(part, splits, plane, ld_p, bias, gelu, resid, ld_r, gamma, beta, eps, out_f, ld_f, out_t, ld_t, out_mode,
                            rows, H, ST(stream)));
}
int vc_token_step_partials(const void* part, int n_part, int rows, int cur_len, int max_len, int pad_id, const int* eos_ids,
                           int n_eos, int* ids, int* unfinished, float* sum_lp, int* n_steps, void* stream) {
  VC_COUNT(1, vc::token_step_partials(part, n_part, rows, cur_len, max_len, pad_id, eos_ids, n_eos, ids, unfinished, sum_lp, n_steps,
                                      ST(stream)));
}
int vc_filter_logits(float* logits, int ld, int rows, int V, float inv_temperature, int top_k, float top_p, int min_tokens_to_keep,
                     void* stream) {
  VC_COUNT(1, vc::filter_logits(logits, ld, rows, V, inv_temperature, top_k, top_p, min_tokens_to_keep, ST(stream)));
}

}  // extern "C"
