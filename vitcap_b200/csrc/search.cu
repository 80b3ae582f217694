// Token search on the vocabulary logits: exact row top-k (radix select), fused greedy / sampling step with state
// update, and the beam-search step. One 256-thread block per logits row; rows are read with 128-bit loads.
//
// Restated from modeling_utils.py:768-886 (_generate_no_beam_search), 888-1100 (_generate_beam_search),
// 1138-1180 (BeamHypotheses) and modeling_bert.py:1429-1432 (sigmoid -> topk(50) -> count >= 0.2).
#include "common.cuh"

namespace vc {

__device__ __forceinline__ uint32_t f2key(float f) {   // order-preserving float -> uint
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------------------------------
// block_topk: exact top-K (K <= 64) of row[0..V) sorted by (value desc, index asc).
// 3 histogram passes (11+11+10 bits) find the K-th largest key; one pass collects; bitonic sort of 64.
// ------------------------------------------------------------------------------------------
struct TopkSmem {
  uint32_t hist[2048];
  uint32_t keys[64];
  int idxs[64];
  uint32_t prefix, need, cnt_gt, cnt_eq;
  int last_idx, best_idx;
  uint32_t wsum[32];
};

__device__ void block_topk(const float* __restrict__ row, int V, int K, TopkSmem& sm, float* out_val, int* out_idx) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) { sm.prefix = 0; sm.need = (uint32_t)K; }
  uint32_t mask_bits = 0;
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = tid; i < 2048; i += nt) sm.hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = sm.prefix;
    const int sh = shifts[pass], wd = widths[pass];
    for (int i = tid; i < V; i += nt) {
      const uint32_t k = f2key(row[i]);
      if ((k & mask_bits) == prefix) atomicAdd(&sm.hist[(k >> sh) & ((1u << wd) - 1)], 1u);
    }
    __syncthreads();
    {
      // the bin in which the count from the top reaches `need`: every thread sums a run of bins (taken from the top), a block
      // scan over those sums finds the run that crosses, and its owner walks it (a serial walk of 2048 bins by one thread
      // was most of this kernel: 3 x ~60k cycles)
      const int nb = 1 << wd;
      const int run = (nb + nt - 1) / nt;                // bins per thread
      const int hi = nb - 1 - tid * run;                 // this thread's bins: hi, hi-1, .., hi-run+1 (those >= 0)
      uint32_t mine = 0;
      for (int j = 0; j < run; ++j)
        if (hi - j >= 0) mine += sm.hist[hi - j];
      uint32_t incl = mine;                              // inclusive scan in thread order == from the top bin downwards
      const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      if (lane == 31) sm.wsum[wid] = incl;
      __syncthreads();
      uint32_t above = incl - mine;                      // count in the bins above this thread's run
      for (int w2 = 0; w2 < wid; ++w2) above += sm.wsum[w2];
      const uint32_t need = sm.need;
      __syncthreads();                                   // every thread has read sm.need / sm.wsum before they are rewritten
      if (above < need && above + mine >= need) {        // exactly one thread: the crossing lies in its run
        uint32_t acc = above;
        int bin = hi;
        for (; bin > 0 && bin > hi - run + 1; --bin) {
          if (acc + sm.hist[bin] >= need) break;
          acc += sm.hist[bin];
        }
        sm.need = need - acc;          // how many are still needed from inside this bin
        sm.prefix = prefix | ((uint32_t)bin << sh);
      }
    }
    mask_bits |= ((1u << wd) - 1) << sh;
    __syncthreads();
  }
  const uint32_t thr = sm.prefix;       // key of the K-th largest element
  const uint32_t need_eq = sm.need;     // number of elements equal to thr that belong to the top-K
  if (tid == 0) { sm.cnt_gt = 0; sm.cnt_eq = 0; sm.last_idx = -1; }
  for (int i = tid; i < 64; i += nt) { sm.keys[i] = 0; sm.idxs[i] = 0x7fffffff; }
  __syncthreads();
  for (int i = tid; i < V; i += nt) {
    const uint32_t k = f2key(row[i]);
    if (k > thr) {
      const uint32_t p = atomicAdd(&sm.cnt_gt, 1u);
      sm.keys[p] = k; sm.idxs[p] = i;
    } else if (k == thr) {
      atomicAdd(&sm.cnt_eq, 1u);
    }
  }
  __syncthreads();
  const uint32_t ngt = sm.cnt_gt;
  if (sm.cnt_eq == need_eq) {
    // no tie ambiguity: take every element equal to the threshold
    __syncthreads();
    if (tid == 0) sm.cnt_eq = 0;
    __syncthreads();
    for (int i = tid; i < V; i += nt) {
      if (f2key(row[i]) == thr) {
        const uint32_t p = ngt + atomicAdd(&sm.cnt_eq, 1u);
        sm.keys[p] = thr; sm.idxs[p] = i;
      }
    }
    __syncthreads();
  } else {
    // ties at the threshold: lowest indices first (deterministic)
    for (uint32_t t = 0; t < need_eq; ++t) {
      if (tid == 0) sm.best_idx = 0x7fffffff;
      __syncthreads();
      const int last = sm.last_idx;
      int best = 0x7fffffff;
      for (int i = tid; i < V; i += nt)
        if (i > last && f2key(row[i]) == thr) { best = i; break; }
      if (best != 0x7fffffff) atomicMin(&sm.best_idx, best);
      __syncthreads();
      if (tid == 0) { sm.keys[ngt + t] = thr; sm.idxs[ngt + t] = sm.best_idx; sm.last_idx = sm.best_idx; }
      __syncthreads();
    }
  }
  // bitonic sort of 64 (key desc, idx asc); unused slots have key 0 / idx max and sink to the end
  for (int size = 2; size <= 64; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < 64) {
        const int j = tid ^ stride;
        if (j > tid) {
          const bool up = ((tid & size) == 0);
          const uint32_t ka = sm.keys[tid], kb = sm.keys[j];
          const int ia = sm.idxs[tid], ib = sm.idxs[j];
          const bool a_first = (ka > kb) || (ka == kb && ia < ib);   // a should precede b in the final order
          if (a_first != up) { sm.keys[tid] = kb; sm.keys[j] = ka; sm.idxs[tid] = ib; sm.idxs[j] = ia; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < K; i += nt) { out_val[i] = key2f(sm.keys[i]); out_idx[i] = sm.idxs[i]; }
  __syncthreads();
}

// online (max, sum exp) over a row; returns to all threads via smem
struct LseSmem { float m[32], s[32]; float row_max, row_logsum; };

__device__ void block_lse(const float* __restrict__ row, int V, float inv_temp, LseSmem& sm) {
  const int tid = threadIdx.x, nt = blockDim.x;
  float m = -INFINITY, s = 0.f;
  for (int i = tid; i < V; i += nt) {
    const float x = row[i] * inv_temp;
    // filtered logits are -inf (modeling_utils.py:1120/1134): they contribute nothing, and exp(-inf - -inf) must not be formed
    if (x > m) { s = (m == -INFINITY ? 0.f : s * expf(m - x)) + 1.f; m = x; }
    else if (x != -INFINITY) s += expf(x - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mx = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * expf(m - mx)) + (m2 == -INFINITY ? 0.f : s2 * expf(m2 - mx));
    m = mx;
  }
  if ((tid & 31) == 0) { sm.m[tid >> 5] = m; sm.s[tid >> 5] = s; }
  __syncthreads();
  if (tid == 0) {
    float mx = -INFINITY;
    for (int w = 0; w < nt / 32; ++w) mx = fmaxf(mx, sm.m[w]);
    float tot = 0.f;
    for (int w = 0; w < nt / 32; ++w) tot += (sm.m[w] == -INFINITY) ? 0.f : sm.s[w] * expf(sm.m[w] - mx);
    sm.row_max = mx;
    sm.row_logsum = logf(tot);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// concept (tag) head selection: sigmoid -> top-k -> count(prob >= thresh)   (modeling_bert.py:1429-1432)
// selection is done on the logits (sigmoid is monotonic; avoids false ties where sigmoid saturates)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tag_topk_kernel(const float* __restrict__ logits, int ld, int V, int K, float thresh, int* __restrict__ out_idx,
                float* __restrict__ out_prob, int* __restrict__ out_len) {
  __shared__ TopkSmem sm;
  __shared__ float vals[64];
  __shared__ int idxs[64];
  const float* row = logits + (size_t)blockIdx.x * ld;
  block_topk(row, V, K, sm, vals, idxs);
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < K; ++i) {
      const float p = 1.f / (1.f + expf(-vals[i]));
      out_prob[(size_t)blockIdx.x * K + i] = p;
      out_idx[(size_t)blockIdx.x * K + i] = idxs[i];
      n += (p >= thresh);
    }
    out_len[blockIdx.x] = n;
  }
}

int tag_topk(const float* logits, int ld, int rows, int V, int K, float thresh, int* out_idx, float* out_prob, int* out_len,
             cudaStream_t s) {
  if (K < 1 || K > 64 || K > V || rows <= 0) { set_last_error("tag_topk: need 1 <= K <= 64 (K=%d)", K); return VC_ERR_BAD_ARG; }
  tag_topk_kernel<<<rows, 256, 0, s>>>(logits, ld, V, K, thresh, out_idx, out_prob, out_len);
  return check_launch("tag_topk");
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (sampling noise is a pure function of (seed, step, row, vocab index))
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) {   // (0,1), 24 bits
  return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// ------------------------------------------------------------------------------------------
// greedy / sampling step: next token + log-prob + sequence state update, one block per sequence.
//   greedy : tok = argmax (first index on ties, as torch.argmax), lp = log_softmax(x)[tok]
//   sample : x' = x / T; tok = argmax_i(x'_i + Gumbel_i)  (== multinomial(softmax(x'))), lp = log_softmax(x')[tok]
//   update : modeling_utils.py:855-862
// ------------------------------------------------------------------------------------------
// Two vectorised sweeps of the row (the second one hits L1/L2): sweep 1 = row max of x/T and the arg-max of the selection
// key (x for greedy, x/T + Gumbel noise for sampling), sweep 2 = sum exp(x/T - max). Columns >= V are padding.
template <bool VEC>
__device__ __forceinline__ float4 row_load4(const float* __restrict__ row, int i4, int V) {
  float4 v;
  const int i = i4 * 4;
  if (VEC && i + 3 < V) return __ldg(reinterpret_cast<const float4*>(row) + i4);
  v.x = (i < V) ? row[i] : -INFINITY;
  v.y = (i + 1 < V) ? row[i + 1] : -INFINITY;
  v.z = (i + 2 < V) ? row[i + 2] : -INFINITY;
  v.w = (i + 3 < V) ? row[i + 3] : -INFINITY;
  return v;
}

template <bool VEC>
__global__ void __launch_bounds__(256)
token_step_kernel(const float* __restrict__ logits, int ld, int V, int do_sample, float inv_temp, uint64_t seed,
                  const uint64_t* __restrict__ seed_dev, int cur_len, int max_len, int pad_id, const int* __restrict__ eos_ids, int n_eos, int* __restrict__ ids,
                  int* __restrict__ unfinished, float* __restrict__ sum_lp, int* __restrict__ n_steps) {
  __shared__ float bv[8], bm[8], bs[8];
  __shared__ int bi[8];
  __shared__ float s_max;
  pdl_launch_dependents();
  pdl_wait();                   // programmatic dependent launch: global memory from here on
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* row = logits + (size_t)r * ld;
  const float it = do_sample ? inv_temp : 1.f;
  const int n4 = (V + 3) >> 2;
  float best = -INFINITY, m = -INFINITY;
  int besti = 0x7fffffff;
  if (!do_sample) {
    for (int i4 = tid; i4 < n4; i4 += 256) {
      const float4 v = row_load4<VEC>(row, i4, V);
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (xs[j] > best) { best = xs[j]; besti = i4 * 4 + j; }     // first index on ties (ascending scan per thread)
    }
    m = best;
  } else {
    if (seed_dev != nullptr) seed = __ldg(seed_dev);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int i4 = tid; i4 < n4; i4 += 256) {
      const float4 v = row_load4<VEC>(row, i4, V);
      const float xs[4] = {v.x, v.y, v.z, v.w};
      const uint4 rnd = philox4x32_10(make_uint4((uint32_t)i4, (uint32_t)r, (uint32_t)cur_len, 0u), key);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i4 * 4 + j < V) {
          const float xt = xs[j] * inv_temp;
          m = fmaxf(m, xt);
          const float x = xt + -logf(-logf(u32_to_unit(rr[j])));
          if (x > best) { best = x; besti = i4 * 4 + j; }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, besti, o);
    if (v2 > best || (v2 == best && i2 < besti)) { best = v2; besti = i2; }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  if ((tid & 31) == 0) { bv[tid >> 5] = best; bi[tid >> 5] = besti; bm[tid >> 5] = m; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      if (bv[w] > best || (bv[w] == best && bi[w] < besti)) { best = bv[w]; besti = bi[w]; }
      m = fmaxf(m, bm[w]);
    }
    bi[0] = besti;
    s_max = m;
  }
  __syncthreads();
  const float row_max = s_max;
  // filtered logits are -inf (modeling_utils.py:1120/1134): exp(-inf - max) == 0, they contribute nothing
  float s = 0.f;
  for (int i4 = tid; i4 < n4; i4 += 256) {
    const float4 v = row_load4<VEC>(row, i4, V);
    s += (expf(v.x * it - row_max) + expf(v.y * it - row_max)) + (expf(v.z * it - row_max) + expf(v.w * it - row_max));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((tid & 31) == 0) bs[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += bs[w];
    besti = bi[0];
    const float x = row[besti] * it;
    const float lp = (x - row_max) - logf(tot);
    const int unf = unfinished[r];
    const int tok = unf ? besti : pad_id;
    ids[(size_t)r * max_len + cur_len] = tok;
    if (unf) sum_lp[r] += lp;                  // (not lp * unf: a finished row's logits may be stale)
    n_steps[r] += unf;
    int still = unf;
    for (int e = 0; e < n_eos; ++e) still *= (tok != eos_ids[e]);
    unfinished[r] = still;
  }
}

int token_step(const float* logits, int ld, int rows, int V, int do_sample, float temperature, uint64_t seed,
               const uint64_t* seed_dev, int cur_len, int max_len, int pad_id, const int* eos_ids, int n_eos, int* ids,
               int* unfinished, float* sum_lp, int* n_steps, cudaStream_t s) {
  if (rows <= 0 || V <= 0 || cur_len < 1 || cur_len >= max_len || temperature <= 0.f) {
    set_last_error("token_step: bad args"); return VC_ERR_BAD_ARG;
  }
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec)
    launch_pdl(token_step_kernel<true>, dim3(rows), dim3(256), 0, s, logits, ld, V, do_sample, 1.f / temperature, seed, seed_dev,
               cur_len, max_len, pad_id, eos_ids, n_eos, ids, unfinished, sum_lp, n_steps);
  else
    launch_pdl(token_step_kernel<false>, dim3(rows), dim3(256), 0, s, logits, ld, V, do_sample, 1.f / temperature, seed, seed_dev,
               cur_len, max_len, pad_id, eos_ids, n_eos, ids, unfinished, sum_lp, n_steps);
  return check_launch("token_step");
}

// final: force EOS at the last position of unfinished rows, mean log-prob, widen ids to int64
// (modeling_utils.py:869-886)
__global__ void greedy_finalize_kernel(const int* __restrict__ ids, const int* __restrict__ unfinished,
                                       const float* __restrict__ sum_lp, const int* __restrict__ n_steps, int eos0, int max_len,
                                       int R, long long* __restrict__ out_ids, float* __restrict__ out_lp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  for (int t = 0; t < max_len; ++t) {
    int v = ids[(size_t)r * max_len + t];
    if (t == max_len - 1 && unfinished[r]) v = eos0;
    out_ids[(size_t)r * max_len + t] = v;
  }
  out_lp[r] = sum_lp[r] / (float)n_steps[r];
}

int greedy_finalize(const int* ids, const int* unfinished, const float* sum_lp, const int* n_steps, int eos0, int max_len, int R,
                    long long* out_ids, float* out_lp, cudaStream_t s) {
  greedy_finalize_kernel<<<(R + 127) / 128, 128, 0, s>>>(ids, unfinished, sum_lp, n_steps, eos0, max_len, R, out_ids, out_lp);
  return check_launch("greedy_finalize");
}

// ------------------------------------------------------------------------------------------
// beam search, part 1: per (image, beam) row: log-sum-exp and the top-(2*beams) logits
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
beam_row_topk_kernel(const float* __restrict__ logits, int ld, int V, int K, float* __restrict__ cand_val,
                     int* __restrict__ cand_idx, float* __restrict__ row_max, float* __restrict__ row_logsum) {
  __shared__ TopkSmem sm;
  __shared__ LseSmem lse;
  __shared__ float vals[64];
  __shared__ int idxs[64];
  pdl_launch_dependents();
  pdl_wait();                   // programmatic dependent launch: global memory from here on
  const float* row = logits + (size_t)blockIdx.x * ld;
  block_lse(row, V, 1.f, lse);
  block_topk(row, V, K, sm, vals, idxs);
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    cand_val[(size_t)blockIdx.x * K + i] = vals[i];
    cand_idx[(size_t)blockIdx.x * K + i] = idxs[i];
  }
  if (threadIdx.x == 0) { row_max[blockIdx.x] = lse.row_max; row_logsum[blockIdx.x] = lse.row_logsum; }
}

int beam_row_topk(const float* logits, int ld, int rows, int V, int K, float* cand_val, int* cand_idx, float* row_max,
                  float* row_logsum, cudaStream_t s) {
  if (K < 1 || K > 64 || K > V) { set_last_error("beam_row_topk: need K <= 64"); return VC_ERR_BAD_ARG; }
  launch_pdl(beam_row_topk_kernel, dim3(rows), dim3(256), 0, s, logits, ld, V, K, cand_val, cand_idx, row_max, row_logsum);
  return check_launch("beam_row_topk");
}

// ------------------------------------------------------------------------------------------
// beam search, part 2: one thread per image walks the merged candidates exactly like the Python loop
// (modeling_utils.py:1003-1052) and maintains the n-best pool (BeamHypotheses, 1138-1180; scores in double
// because the reference does this arithmetic on Python floats).
// ------------------------------------------------------------------------------------------
#define VC_MAX_BEAMS 8
#define VC_MAX_LEN 64

struct BeamState {
  int* ids;            // [R, max_len]
  float* beam_scores;  // [R]
  int* done;           // [B]
  int* anc;            // [max_len, R]  ancestor row of every cached caption position
  double* hyp_score;   // [B, keep]
  int* hyp_len;        // [B, keep]
  int* hyp_ids;        // [B, keep, max_len]
  int* hyp_count;      // [B]
  double* worst;       // [B]
};

__global__ void beam_advance_kernel(BeamState st, const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                    const float* __restrict__ row_max, const float* __restrict__ row_logsum, int B, int nb, int V,
                                    int cur_len, int max_len, int keep, double length_penalty, int pad_id,
                                    const int* __restrict__ eos_ids, int n_eos) {
  pdl_launch_dependents();
  pdl_wait();                   // programmatic dependent launch: global memory from here on
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int R = B * nb, K = 2 * nb, step = cur_len - 1;
  // merged candidate list: score = log_softmax(x) + beam_score (fp32, as the reference), flat index = beam*V + word
  float cs[VC_MAX_BEAMS * 2 * VC_MAX_BEAMS];
  int ci[VC_MAX_BEAMS * 2 * VC_MAX_BEAMS];
  int n = 0;
  for (int j = 0; j < nb; ++j) {
    const int r = b * nb + j;
    for (int k = 0; k < K; ++k) {
      const float lsm = (cand_val[(size_t)r * K + k] - row_max[r]) - row_logsum[r];
      cs[n] = lsm + st.beam_scores[r];
      ci[n] = j * V + cand_idx[(size_t)r * K + k];
      ++n;
    }
  }
  // partial selection sort: top-K by (score desc, flat index asc)
  for (int a = 0; a < K; ++a) {
    int best = a;
    for (int c = a + 1; c < n; ++c)
      if (cs[c] > cs[best] || (cs[c] == cs[best] && ci[c] < ci[best])) best = c;
    const float ts = cs[a]; cs[a] = cs[best]; cs[best] = ts;
    const int ti = ci[a]; ci[a] = ci[best]; ci[best] = ti;
  }
  int old_ids[VC_MAX_BEAMS][VC_MAX_LEN];
  int old_anc[VC_MAX_BEAMS][VC_MAX_LEN];
  for (int j = 0; j < nb; ++j) {
    for (int t = 0; t < cur_len; ++t) old_ids[j][t] = st.ids[(size_t)(b * nb + j) * max_len + t];
    for (int t = 0; t < step; ++t) old_anc[j][t] = st.anc[(size_t)t * R + b * nb + j];
  }
  const int hmax = max_len - 1;                     // BeamHypotheses.max_length
  int cnt = st.hyp_count[b];
  double worst = st.worst[b];
  int done = st.done[b];
  if (!done) {                                      // is_done(best_sum_logprobs) evaluated before the walk
    if (cnt >= keep) done = worst >= (double)cs[0] / pow((double)hmax, length_penalty);
  }
  float nscore[VC_MAX_BEAMS];
  int nword[VC_MAX_BEAMS], nparent[VC_MAX_BEAMS];
  int filled = 0;
  if (!done) {
    for (int a = 0; a < K && filled < nb; ++a) {
      const int beam = ci[a] / V, word = ci[a] % V;
      bool is_eos = false;
      for (int e = 0; e < n_eos; ++e) is_eos |= (word == eos_ids[e]);
      if (is_eos || cur_len + 1 == max_len) {
        // BeamHypotheses.add(ids[beam, :cur_len], score)
        const double sc = (double)cs[a] / pow((double)cur_len, length_penalty);
        if (cnt < keep || sc > worst) {
          const size_t hb = (size_t)b * keep;
          int slot;
          const bool was_full = (cnt == keep);
          if (was_full) {
            // python: append, then delete sorted([(score, idx)])[0]; worst_score = the next lowest score.
            int lo = -1;                       // -1 == the entry being appended (highest idx)
            double los = sc;
            for (int h2 = 0; h2 < cnt; ++h2) {
              const double hs = st.hyp_score[hb + h2];
              if ((lo == -1) ? (hs <= los) : (hs < los)) { lo = h2; los = hs; }
            }
            if (lo >= 0) {
              for (int h2 = lo; h2 + 1 < cnt; ++h2) {
                st.hyp_score[hb + h2] = st.hyp_score[hb + h2 + 1];
                st.hyp_len[hb + h2] = st.hyp_len[hb + h2 + 1];
                for (int t = 0; t < max_len; ++t) st.hyp_ids[(hb + h2) * max_len + t] = st.hyp_ids[(hb + h2 + 1) * max_len + t];
              }
              slot = cnt - 1;
            } else {
              slot = -1;                       // the new entry itself was the lowest: dropped again
            }
          } else {
            slot = cnt++;
          }
          if (slot >= 0) {
            st.hyp_score[hb + slot] = sc;
            st.hyp_len[hb + slot] = cur_len;
            for (int t = 0; t < max_len; ++t) st.hyp_ids[(hb + slot) * max_len + t] = (t < cur_len) ? old_ids[beam][t] : pad_id;
          }
          if (was_full) {                      // worst_score = sorted_scores[1][0]: lowest score left in the pool
            double w = st.hyp_score[hb];
            for (int h2 = 1; h2 < cnt; ++h2) w = fmin(w, st.hyp_score[hb + h2]);
            worst = w;
          } else {
            worst = fmin(sc, worst);
          }
        }
      } else {
        nscore[filled] = cs[a]; nword[filled] = word; nparent[filled] = beam;
        ++filled;
      }
    }
  }
  st.hyp_count[b] = cnt;
  st.worst[b] = worst;
  st.done[b] = done;
  for (int j = 0; j < nb; ++j) {
    const int r = b * nb + j;
    int parent = 0, word = pad_id;
    float sc = 0.f;
    if (!done && filled == nb) { parent = nparent[j]; word = nword[j]; sc = nscore[j]; }   // else: (0, PAD, 0) padding rows
    for (int t = 0; t < cur_len; ++t) st.ids[(size_t)r * max_len + t] = old_ids[parent][t];
    st.ids[(size_t)r * max_len + cur_len] = word;
    st.beam_scores[r] = sc;
    for (int t = 0; t < step; ++t) st.anc[(size_t)t * R + r] = old_anc[parent][t];
    st.anc[(size_t)step * R + r] = b * nb + parent;
  }
}

int beam_advance(int* ids, float* beam_scores, int* done, int* anc, double* hyp_score, int* hyp_len, int* hyp_ids, int* hyp_count,
                 double* worst, const float* cand_val, const int* cand_idx, const float* row_max, const float* row_logsum, int B,
                 int nb, int V, int cur_len, int max_len, int keep, double length_penalty, int pad_id, const int* eos_ids,
                 int n_eos, cudaStream_t s) {
  if (nb < 1 || nb > VC_MAX_BEAMS || max_len > VC_MAX_LEN || keep < 1) {
    set_last_error("beam_advance: num_beams <= %d, max_length <= %d required", VC_MAX_BEAMS, VC_MAX_LEN);
    return VC_ERR_BAD_ARG;
  }
  BeamState st = {ids, beam_scores, done, anc, hyp_score, hyp_len, hyp_ids, hyp_count, worst};
  launch_pdl(beam_advance_kernel, dim3((B + 31) / 32), dim3(32), 0, s, st, cand_val, cand_idx, row_max, row_logsum, B, nb, V, cur_len,
             max_len, keep, length_penalty, pad_id, eos_ids, n_eos);
  return check_launch("beam_advance");
}

// final selection (modeling_utils.py:1074-1100): best `keep` hypotheses by score, EOS appended, PAD filled, -1e5 for empty slots
__global__ void beam_finalize_kernel(const double* __restrict__ hyp_score, const int* __restrict__ hyp_len,
                                     const int* __restrict__ hyp_ids, const int* __restrict__ hyp_count, int B, int keep,
                                     int max_len, int pad_id, int eos0, long long* __restrict__ out_ids, float* __restrict__ out_lp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int cnt = hyp_count[b];
  bool used[64];
  for (int i = 0; i < cnt; ++i) used[i] = false;
  for (int j = 0; j < keep; ++j) {
    long long* o = out_ids + ((size_t)b * keep + j) * max_len;
    for (int t = 0; t < max_len; ++t) o[t] = pad_id;
    out_lp[(size_t)b * keep + j] = -1e5f;
    if (j < cnt) {
      int best = -1;
      float bs = 0.f;
      for (int i = 0; i < cnt; ++i) {
        const float sc = (float)hyp_score[(size_t)b * keep + i];   // torch.tensor(python floats) -> fp32 before topk
        if (!used[i] && (best < 0 || sc > bs)) { best = i; bs = sc; }
      }
      used[best] = true;
      const int len = hyp_len[(size_t)b * keep + best];
      for (int t = 0; t < len; ++t) o[t] = hyp_ids[((size_t)b * keep + best) * max_len + t];
      o[len] = eos0;
      out_lp[(size_t)b * keep + j] = (float)hyp_score[(size_t)b * keep + best];
    }
  }
}

int beam_finalize(const double* hyp_score, const int* hyp_len, const int* hyp_ids, const int* hyp_count, int B, int keep, int max_len,
                  int pad_id, int eos0, long long* out_ids, float* out_lp, cudaStream_t s) {
  if (keep > 64) { set_last_error("beam_finalize: num_keep_best <= 64"); return VC_ERR_BAD_ARG; }
  beam_finalize_kernel<<<(B + 31) / 32, 32, 0, s>>>(hyp_score, hyp_len, hyp_ids, hyp_count, B, keep, max_len, pad_id, eos0, out_ids, out_lp);
  return check_launch("beam_finalize");
}

}  // namespace vc

// ------------------------------------------------------------------------------------------
// top_k_top_p_filtering (modeling_utils.py:1103-1135), in place, one block per row.
//   scale : x *= inv_temperature                      (modeling_utils.py:841-842)
//   top-k : x < (k-th largest x)  -> -inf               (ties with the k-th value are kept, as `logits < kth`)
//   top-p : a token is removed iff the probability mass of all tokens ranked before it exceeds top_p
//           (cumsum in descending order, shifted right by one so the first token crossing top_p is kept).
//           Equivalent threshold form: keep key >= t*, t* = smallest key with mass{key_j > t*} <= top_p, found by
//           bisection over the 32-bit ordered key space with a deterministic block reduction per probe.
// ------------------------------------------------------------------------------------------
namespace vc {

__device__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(256)
filter_logits_kernel(float* __restrict__ logits, int ld, int V, float inv_temp, int top_k, float top_p) {
  __shared__ TopkSmem sm;
  __shared__ LseSmem lse;
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();                   // programmatic dependent launch: global memory from here on
  float* row = logits + (size_t)blockIdx.x * ld;
  const int tid = threadIdx.x;
  if (inv_temp != 1.f) {
    for (int i = tid; i < V; i += 256) row[i] *= inv_temp;
    __syncthreads();
  }
  if (top_k > 0 && top_k < V) {
    // radix select of the k-th largest key (same 3 passes as block_topk)
    if (tid == 0) { sm.prefix = 0; sm.need = (uint32_t)top_k; }
    uint32_t mask_bits = 0;
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
      for (int i = tid; i < 2048; i += 256) sm.hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = sm.prefix;
      const int sh = shifts[pass], wd = widths[pass];
      for (int i = tid; i < V; i += 256) {
        const uint32_t k = f2key(row[i]);
        if ((k & mask_bits) == prefix) atomicAdd(&sm.hist[(k >> sh) & ((1u << wd) - 1)], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t need = sm.need, acc = 0;
        int bin = (1 << wd) - 1;
        for (; bin > 0; --bin) {
          if (acc + sm.hist[bin] >= need) break;
          acc += sm.hist[bin];
        }
        sm.need = need - acc;
        sm.prefix = prefix | ((uint32_t)bin << sh);
      }
      mask_bits |= ((1u << wd) - 1) << sh;
      __syncthreads();
    }
    const uint32_t thr = sm.prefix;
    for (int i = tid; i < V; i += 256)
      if (f2key(row[i]) < thr) row[i] = -INFINITY;
    __syncthreads();
  }
  if (top_p < 1.0f) {
    block_lse(row, V, 1.f, lse);
    const float mx = lse.row_max, ls = lse.row_logsum;
    uint64_t lo = 0, hi = 0xffffffffull;
    while (lo < hi) {
      const uint32_t mid = (uint32_t)((lo + hi) >> 1);
      float part = 0.f;
      for (int i = tid; i < V; i += 256) {
        const float x = row[i];
        if (f2key(x) > mid) part += expf((x - mx) - ls);
      }
      const float mass = block_sum_256(part, red);
      if (mass <= top_p) hi = mid; else lo = (uint64_t)mid + 1;
    }
    const uint32_t tstar = (uint32_t)lo;
    __syncthreads();
    for (int i = tid; i < V; i += 256)
      if (f2key(row[i]) < tstar) row[i] = -INFINITY;
  }
}

int filter_logits(float* logits, int ld, int rows, int V, float inv_temperature, int top_k, float top_p, int min_tokens_to_keep,
                  cudaStream_t s) {
  if (min_tokens_to_keep != 1) { set_last_error("filter_logits: min_tokens_to_keep must be 1 (beam sampling unsupported)"); return VC_ERR_UNSUPPORTED; }
  if (rows <= 0 || V <= 0 || top_k < 0 || !(top_p > 0.f)) { set_last_error("filter_logits: bad args"); return VC_ERR_BAD_ARG; }
  launch_pdl(filter_logits_kernel, dim3(rows), dim3(256), 0, s, logits, ld, V, inv_temperature, top_k, top_p);
  return check_launch("filter_logits");
}

}  // namespace vc
