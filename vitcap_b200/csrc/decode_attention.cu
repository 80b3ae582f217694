// Single-step decoder self-attention over the KV cache (HBM-bound).
//
// Every sequence contributes two query rows per step: the last generated token (row 2r) and the [MASK] slot whose
// output feeds the vocabulary head (row 2r+1)  -- modeling_bert.py:845-876 feeds [ids, MASK] each step; the cached
// formulation is SURVEY.md section 7.1. Keys visible to them:
//   (1) the 578 context rows [tag-CLS | image tokens] of the sequence's image: K/V written once by the prefill QKV GEMM
//       into ctx_qkv[B, C, 3H]; shared by all beams/samples of the image and read ONCE per block,
//   (2) the caption tokens generated so far: K/V rows of earlier steps' QKV GEMM outputs, step_qkv[step, 2R, 3H] (row 2r'),
//       reached through the ancestor table (beam search re-parents rows; no K/V is ever copied),
//   (3) this step's own rows; the real-token query must NOT see the MASK key (additive -10000 in the reference mask,
//       modeling_bert.py:1501, which underflows to exactly 0 after softmax).
// 8 lanes share one 128-byte (bf16) K or V row => every global access is a coalesced 16-byte load; a warp covers
// 4 keys per iteration, 4 warps split the keys, partial (max, sum, out) states merge by shuffles and shared memory.
#include "common.cuh"

namespace vc {

int decode_attention_mma(const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C, const int* ctx_vis,
                         int heads, int E, int cur_len, float scale, const int* seq_unfinished, const int* img_done, cudaStream_t s);
// CUDA-core variant (exact mode, fp32 storage); also instantiable for bf16 as a cross-check of the mma kernel
int decode_attention_simt(int is_bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C,
                          const int* ctx_vis, int heads, int E, int cur_len, float scale, cudaStream_t s);

template <typename T> __device__ __forceinline__ float fast_exp(float x);
template <> __device__ __forceinline__ float fast_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ float fast_exp<bf16>(float x) { return __expf(x); }

template <int NQ> struct AttnState {
  float m[NQ], l[NQ], o[NQ][8];
};

template <typename T, int EG>
__global__ void __launch_bounds__(128)
decode_attention_kernel(const T* __restrict__ ctx_qkv, const T* __restrict__ step_qkv, const int* __restrict__ anc,
                        T* __restrict__ out, int Cs, const int* __restrict__ ctx_vis, int H, int R, int E, int cur_len, float scale) {
  constexpr int NQ = 2 * EG, D = 64;
  const int h = blockIdx.x;
  const int groups = (E + EG - 1) / EG;
  const int b = blockIdx.y / groups;
  const int e0 = (blockIdx.y % groups) * EG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, part = lane & 7;
  const size_t ld = 3 * (size_t)H;
  const int step = cur_len - 1;
  const T* cur = step_qkv + (size_t)step * 2 * R * ld;
  // context rows of an image: Cs allocated, the first C visible (label-region masks hide a per-image tail)
  const int C = ctx_vis ? ctx_vis[b] : Cs;

  float q[NQ][8];
  AttnState<NQ> st;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int e = e0 + (i >> 1);
    const int r = b * E + (e < E ? e : E - 1);
    load8<T>(cur + (size_t)(2 * r + (i & 1)) * ld + h * D + part * 8, q[i]);
    st.m[i] = -1e30f; st.l[i] = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) { q[i][d] *= scale; st.o[i][d] = 0.f; }
  }

  // ---- phase 1: context keys, shared by all rows of the image ----
  const T* kbase = ctx_qkv + (size_t)b * Cs * ld + H + h * D + part * 8;
  const T* vbase = kbase + H;
  constexpr int U = 4;                                   // keys in flight per lane group
  for (int kb = warp * 4; kb < C; kb += 16 * U) {          // warp-uniform trip count (shuffles inside)
    const int k0 = kb + grp;
    float kf[U][8], vf[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + 16 * u;
      if (k < C) {
        load8<T>(kbase + (size_t)k * ld, kf[u]);
        load8<T>(vbase + (size_t)k * ld, vf[u]);
      } else {
#pragma unroll
        for (int d = 0; d < 8; ++d) { kf[u][d] = 0.f; vf[u][d] = 0.f; }
      }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float s[U];
      float mx = st.m[i];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) a = fmaf(q[i][d], kf[u][d], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        s[u] = (k0 + 16 * u < C) ? a : -1e30f;
        mx = fmaxf(mx, s[u]);
      }
      const float corr = fast_exp<T>(st.m[i] - mx);
      st.m[i] = mx;
      float ls = st.l[i] * corr;
#pragma unroll
      for (int d = 0; d < 8; ++d) st.o[i][d] *= corr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float p = (k0 + 16 * u < C) ? fast_exp<T>(s[u] - mx) : 0.f;
        ls += p;
#pragma unroll
        for (int d = 0; d < 8; ++d) st.o[i][d] = fmaf(p, vf[u][d], st.o[i][d]);
      }
      st.l[i] = ls;
    }
  }

  // ---- phase 2: each row's own caption keys (cached steps 0..step-1, then this step's real and MASK rows) ----
  const int nk = cur_len + 1;
#pragma unroll
  for (int el = 0; el < EG; ++el) {
    const int e = e0 + el;
    if (e < E) {
      const int r = b * E + e;
      for (int jb = warp * 4; jb < nk; jb += 16) {          // warp-uniform trip count
        const int j = jb + grp;
        const bool jvalid = j < nk;
        const T* kp;
        if (!jvalid) {
          kp = cur + (size_t)(2 * r) * ld;
        } else if (j < step) {
          const int src = anc ? anc[(size_t)j * R + r] : r;
          kp = step_qkv + ((size_t)j * 2 * R + 2 * src) * ld;
        } else {
          kp = cur + (size_t)(2 * r + (j - step)) * ld;
        }
        float kf[8], vf[8];
        load8<T>(kp + H + h * D + part * 8, kf);
        load8<T>(kp + 2 * H + h * D + part * 8, vf);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int i = 2 * el + w;
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < 8; ++d) a = fmaf(q[i][d], kf[d], a);
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          a += __shfl_xor_sync(0xffffffffu, a, 4);
          const bool visible = jvalid && !(w == 0 && j == step + 1);   // real-token query never sees the MASK key
          if (visible) {
            const float mx = fmaxf(st.m[i], a);
            const float corr = fast_exp<T>(st.m[i] - mx);
            const float p = fast_exp<T>(a - mx);
            st.m[i] = mx;
            st.l[i] = st.l[i] * corr + p;
#pragma unroll
            for (int d = 0; d < 8; ++d) st.o[i][d] = fmaf(p, vf[d], st.o[i][d] * corr);
          }
        }
      }
    }
  }

  // ---- merge the 4 lane groups of the warp ----
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
#pragma unroll
    for (int x = 8; x <= 16; x <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, st.m[i], x);
      const float l2 = __shfl_xor_sync(0xffffffffu, st.l[i], x);
      const float mx = fmaxf(st.m[i], m2);
      const float c1 = fast_exp<T>(st.m[i] - mx), c2 = fast_exp<T>(m2 - mx);
      st.l[i] = st.l[i] * c1 + l2 * c2;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const float o2 = __shfl_xor_sync(0xffffffffu, st.o[i][d], x);
        st.o[i][d] = st.o[i][d] * c1 + o2 * c2;
      }
      st.m[i] = mx;
    }
  }
  // ---- merge the 4 warps ----
  __shared__ float sm_m[4][NQ], sm_l[4][NQ], sm_o[4][NQ][D];
  if (grp == 0) {
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      if (part == 0) { sm_m[warp][i] = st.m[i]; sm_l[warp][i] = st.l[i]; }
#pragma unroll
      for (int d = 0; d < 8; ++d) sm_o[warp][i][part * 8 + d] = st.o[i][d];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NQ * 8; t += 128) {
    const int i = t >> 3, pt = t & 7;
    const int e = e0 + (i >> 1);
    if (e >= E) continue;
    float mx = sm_m[0][i];
#pragma unroll
    for (int w = 1; w < 4; ++w) mx = fmaxf(mx, sm_m[w][i]);
    float l = 0.f, o[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float c = fast_exp<T>(sm_m[w][i] - mx);
      l += sm_l[w][i] * c;
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] += sm_o[w][i][pt * 8 + d] * c;
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] *= inv;
    const int r = b * E + e;
    store8<T>(out + (size_t)(2 * r + (i & 1)) * H + h * D + pt * 8, o);
  }
}

template <typename T>
static int launch_da(const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C, const int* ctx_vis,
                     int heads, int E, int cur_len, float scale, cudaStream_t s) {
  const int H = heads * 64, R = B * E;
  const T* c = (const T*)ctx_qkv;
  const T* q = (const T*)step_qkv;
  T* o = (T*)out;
  if (E == 1) {
    decode_attention_kernel<T, 1><<<dim3(heads, B), 128, 0, s>>>(c, q, anc, o, C, ctx_vis, H, R, E, cur_len, scale);
  } else if (E == 2) {
    decode_attention_kernel<T, 2><<<dim3(heads, B), 128, 0, s>>>(c, q, anc, o, C, ctx_vis, H, R, E, cur_len, scale);
  } else if (E == 3) {
    decode_attention_kernel<T, 3><<<dim3(heads, B), 128, 0, s>>>(c, q, anc, o, C, ctx_vis, H, R, E, cur_len, scale);
  } else {
    const int groups = (E + 3) / 4;
    decode_attention_kernel<T, 4><<<dim3(heads, B * groups), 128, 0, s>>>(c, q, anc, o, C, ctx_vis, H, R, E, cur_len, scale);
  }
  return check_launch("decode_attention");
}

// ctx_qkv [B, C, 3H]; step_qkv [max_len, 2*B*E, 3H]; anc int32 [max_len, B*E] or NULL; out [2*B*E, H]
int decode_attention(int is_bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C,
                     const int* ctx_vis, int heads, int E, int cur_len, float scale, cudaStream_t s) {
  if (B <= 0 || C <= 0 || heads <= 0 || E <= 0 || cur_len < 1 || B * ((E + 3) / 4) > 65535) {
    set_last_error("decode_attention: bad args B=%d C=%d heads=%d E=%d cur_len=%d", B, C, heads, E, cur_len);
    return VC_ERR_BAD_ARG;
  }
  if (is_bf16) return decode_attention_mma(ctx_qkv, step_qkv, anc, out, B, C, ctx_vis, heads, E, cur_len, scale, nullptr, nullptr, s);
  return launch_da<float>(ctx_qkv, step_qkv, anc, out, B, C, ctx_vis, heads, E, cur_len, scale, s);
}

// fast mode only: the same, skipping sequences that have finished (seq_unfinished [B*E], 0 = finished) or images whose beam
// search is done (img_done [B], != 0 = done); their output rows are left as they are
int decode_attention_skip(const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C, const int* ctx_vis,
                          int heads, int E, int cur_len, float scale, const int* seq_unfinished, const int* img_done, cudaStream_t s) {
  if (B <= 0 || C <= 0 || heads <= 0 || E <= 0 || cur_len < 1) { set_last_error("decode_attention_skip: bad args"); return VC_ERR_BAD_ARG; }
  return decode_attention_mma(ctx_qkv, step_qkv, anc, out, B, C, ctx_vis, heads, E, cur_len, scale, seq_unfinished, img_done, s);
}

int decode_attention_simt(int is_bf16, const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C,
                          const int* ctx_vis, int heads, int E, int cur_len, float scale, cudaStream_t s) {
  if (B <= 0 || C <= 0 || heads <= 0 || E <= 0 || cur_len < 1 || B * ((E + 3) / 4) > 65535) {
    set_last_error("decode_attention_simt: bad args"); return VC_ERR_BAD_ARG;
  }
  if (is_bf16) return launch_da<bf16>(ctx_qkv, step_qkv, anc, out, B, C, ctx_vis, heads, E, cur_len, scale, s);
  return launch_da<float>(ctx_qkv, step_qkv, anc, out, B, C, ctx_vis, heads, E, cur_len, scale, s);
}

}  // namespace vc
