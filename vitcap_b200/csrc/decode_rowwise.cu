// Row-wise kernels of the fused decode step (the partners of gemm_dec.cu).
//
// finish_ln:   y = LayerNorm( act( sum_s partial[s] + bias ) + resid ) per row, written as the fp32 row (residual stream) and as
//              the operand of the next GEMM (bf16, or the split pair [hi | lo] of the three-product GEMMs).
//              Replaces, in ONE pass over rows that are L2 resident: the split-K reduction, the bias / residual epilogue of
//              BertSelfOutput.dense / BertOutput.dense (modeling_bert.py:353-357, 415-419), their LayerNorm, the GELU of the
//              prediction-head transform (modeling_bert.py:524-537) and the operand split.
// token_step_partials: greedy token selection from the vocabulary GEMM's per-tile (max, arg max, sum of exponentials) partials:
//              argmax -> log_softmax -> gather of modeling_utils.py:849-853 without the logits ever being stored, then the
//              sequence-state update of token_step_kernel (search.cu).
#include "pair.cuh"

#include <climits>

namespace vc {

// TWO warps per row (interleaved float4 columns), two rows per 128-thread block; H <= 1024, H % 128 == 0. S = number of partial
// planes as a compile-time constant (0 = run-time loop).
// The kernel is pure latency (the rows are L2 resident, 1024 rows x 3 KB x a few planes): what counts is how many dependent
// memory round trips a row makes. ncu of the one-warp-per-row form (session 4 of round 2): 75-80 % of the stall samples on the
// first use of four successive load groups (first half of the planes, second half, bias / residual, gamma / beta after the
// reductions). Now: bias, gamma and beta -- parameters, not products of the previous kernel -- are requested BEFORE the
// programmatic-dependency wait, and every plane and the residual of a thread's columns right after it, all in flight together
// (half the columns per thread keeps that within the register file): one round trip.
template <bool GELU, int S>
__global__ void __launch_bounds__(128)
finish_ln_kernel(const float* __restrict__ part, int splits, size_t plane, int ld_p, const float* __restrict__ bias,
                 const float* __restrict__ resid, int ld_r, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 float* __restrict__ out_f, int ld_f, bf16* __restrict__ out_t, int ld_t, int out_mode, int rows, int H) {
  __shared__ float red[2][2][2];                      // [row in block][pass][warp of the row]
  pdl_launch_dependents();
  const int rb = threadIdx.x >> 6;                   // row inside the block
  const int w = (threadIdx.x >> 5) & 1;              // which of the row's two warps
  const int lane = threadIdx.x & 31;
  const int row_raw = blockIdx.x * 2 + rb;
  const bool active = row_raw < rows;
  const int row = active ? row_raw : rows - 1;       // an odd tail row: compute on a valid row, store nothing (barriers below)
  const int nv = H / 128;                            // float4 per lane over the whole row; this thread owns ii = 2 i + w
  float4 g[4], be[4], bi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = 2 * i + w;
    if (ii < nv) {
      const int c = (ii * 32 + lane) * 4;
      g[i] = __ldg(reinterpret_cast<const float4*>(gamma + c));
      be[i] = __ldg(reinterpret_cast<const float4*>(beta + c));
      bi[i] = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  pdl_wait();
  const int ns = S > 0 ? S : splits;
  const float* prow = part + (size_t)row * ld_p;
  float4 acc[4][S > 0 ? S : 1];
  float4 rs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = 2 * i + w;
    if (ii < nv) {
      const int c = (ii * 32 + lane) * 4;
      if (S > 0) {
#pragma unroll
        for (int sp = 0; sp < S; ++sp) acc[i][sp] = *reinterpret_cast<const float4*>(prow + sp * plane + c);
      } else {
        acc[i][0] = *reinterpret_cast<const float4*>(prow + c);
      }
      rs[i] = resid != nullptr ? *reinterpret_cast<const float4*>(resid + (size_t)row * ld_r + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = 2 * i + w;
    if (ii < nv) {
      const int c = (ii * 32 + lane) * 4;
      float4 a = acc[i][0];
      if (S > 0) {
#pragma unroll
        for (int sp = 1; sp < S; ++sp) { a.x += acc[i][sp].x; a.y += acc[i][sp].y; a.z += acc[i][sp].z; a.w += acc[i][sp].w; }
      } else {
        for (int sp = 1; sp < ns; ++sp) {    // fixed summation order: plane 0, 1, ... (as the unrolled form)
          const float4 b = *reinterpret_cast<const float4*>(prow + sp * plane + c);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
      }
      a.x += bi[i].x; a.y += bi[i].y; a.z += bi[i].z; a.w += bi[i].w;
      if (GELU) {
        gelu_erf_tanh_x2(a.x, a.y);
        gelu_erf_tanh_x2(a.z, a.w);
      }
      a.x += rs[i].x; a.y += rs[i].y; a.z += rs[i].z; a.w += rs[i].w;
      v[i] = a;
      s += (a.x + a.y) + (a.z + a.w);
    }
  }
  s = warp_sum(s);
  if (lane == 0) red[rb][0][w] = s;
  __syncthreads();
  const float mean = (red[rb][0][0] + red[rb][0][1]) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (2 * i + w < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  q = warp_sum(q);
  if (lane == 0) red[rb][1][w] = q;
  __syncthreads();
  const float rstd = rsqrtf((red[rb][1][0] + red[rb][1][1]) / (float)H + eps);
  if (!active) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = 2 * i + w;
    if (ii < nv) {
      const int c = (ii * 32 + lane) * 4;
      const float4 o = make_float4((v[i].x - mean) * rstd * g[i].x + be[i].x, (v[i].y - mean) * rstd * g[i].y + be[i].y,
                                   (v[i].z - mean) * rstd * g[i].z + be[i].z, (v[i].w - mean) * rstd * g[i].w + be[i].w);
      if (out_f != nullptr) *reinterpret_cast<float4*>(out_f + (size_t)row * ld_f + c) = o;
      if (out_mode == 1) {
        *reinterpret_cast<uint2*>(out_t + (size_t)row * ld_t + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      } else if (out_mode == 2 || out_mode == 3) {
        uint2 hi, lo;
        split_bf16x2(o.x, o.y, hi.x, lo.x);
        split_bf16x2(o.z, o.w, hi.y, lo.y);
        bf16* op = out_t + (size_t)row * ld_t + c;
        *reinterpret_cast<uint2*>(op) = hi;
        *reinterpret_cast<uint2*>(op + H) = lo;
        if (out_mode == 3) *reinterpret_cast<uint2*>(op + 2 * H) = hi;     // the K-concatenated GEMM form reads [hi | lo | hi]
      }
      if (out_mode == 4 || out_mode == 5) {             // IEEE-half operand copy (the one-product half GEMMs of gemm_dec.cu)
        bf16* op = out_t + (size_t)row * ld_t + c;
        if (out_mode == 5) {                             // [bf16 | half]: the q|k|v projection reads the first H columns
          *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          op += H;
        }
        *reinterpret_cast<uint2*>(op) = make_uint2(pack_f16x2(o.x, o.y), pack_f16x2(o.z, o.w));
      }
    }
  }
}

// part: [splits] planes of [>= rows, ld_p] fp32, `plane` elements apart; out_mode 0 = no operand copy, 1 = bf16 [rows, ld_t],
// 2 = split pair: columns [0, H) = hi, [H, 2H) = lo (ld_t >= 2H), 3 = [hi | lo | hi] (ld_t >= 3H), 4 = IEEE half [rows, ld_t],
// 5 = [bf16 | half] (ld_t >= 2H)
int finish_ln(const float* part, int splits, size_t plane, int ld_p, const float* bias, int gelu, const float* resid, int ld_r,
              const float* gamma, const float* beta, float eps, float* out_f, int ld_f, void* out_t, int ld_t, int out_mode, int rows,
              int H, cudaStream_t s) {
  if (H % 128 || H > 1024 || rows <= 0 || splits < 1 || (ld_p % 4) || (plane % 4) || (resid && (ld_r % 4)) || (out_f && (ld_f % 4)) ||
      out_mode < 0 || out_mode > 5 || part == nullptr ||
      (out_mode && (out_t == nullptr || (ld_t % 4) || ld_t < (out_mode == 4 ? 1 : out_mode == 5 ? 2 : out_mode) * H)) ||
      gamma == nullptr || beta == nullptr) {
    set_last_error("finish_ln: need H %% 128 == 0, H <= 1024, pitches %% 4 == 0, ld_t >= H (2H for the split pair) (H=%d)", H);
    return VC_ERR_BAD_ARG;
  }
  const dim3 grid((rows + 1) / 2), block(128);
  bf16* ot = static_cast<bf16*>(out_t);
#define VC_FIN(G, S) launch_pdl(finish_ln_kernel<G, S>, grid, block, 0, s, part, splits, plane, ld_p, bias, resid, ld_r, gamma, beta, \
                                eps, out_f, ld_f, ot, ld_t, out_mode, rows, H)
  if (gelu) {
    if (splits == 3) VC_FIN(true, 3); else if (splits == 6) VC_FIN(true, 6); else VC_FIN(true, 0);
  } else {
    if (splits == 3) VC_FIN(false, 3); else if (splits == 6) VC_FIN(false, 6); else VC_FIN(false, 0);
  }
#undef VC_FIN
  return check_launch("finish_ln");
}

// one warp per sequence row
__global__ void __launch_bounds__(256)
token_step_partials_kernel(const float4* __restrict__ part, int n_part, int rows, int cur_len, int max_len, int pad_id,
                           const int* __restrict__ eos_ids, int n_eos, int* __restrict__ ids, int* __restrict__ unfinished,
                           float* __restrict__ sum_lp, int* __restrict__ n_steps) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float4* p = part + (size_t)r * n_part;
  float mx = -INFINITY;
  int best = INT_MAX;
  for (int i = lane; i < n_part; i += 32) {
    const float4 e = p[i];
    const int ei = __float_as_int(e.y);
    if (e.x > mx || (e.x == mx && ei < best)) { mx = e.x; best = ei; }       // first index on ties, as torch.argmax
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, mx, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, best, o);
    if (m2 > mx || (m2 == mx && i2 < best)) { mx = m2; best = i2; }
  }
  float s = 0.f;
  for (int i = lane; i < n_part; i += 32) {
    const float4 e = p[i];
    if (e.z > 0.f) s += e.z * expf(e.x - mx);          // a partial with no valid column carries sum 0 (max -inf)
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float lp = -logf(s);                          // log_softmax at the arg max: (max - max) - log sum exp(x - max)
    const int unf = unfinished[r];
    const int tok = unf ? best : pad_id;
    ids[(size_t)r * max_len + cur_len] = tok;
    if (unf) sum_lp[r] += lp;
    n_steps[r] += unf;
    int still = unf;
    for (int e = 0; e < n_eos; ++e) still *= (tok != eos_ids[e]);
    unfinished[r] = still;
  }
}

int token_step_partials(const void* part, int n_part, int rows, int cur_len, int max_len, int pad_id, const int* eos_ids, int n_eos,
                        int* ids, int* unfinished, float* sum_lp, int* n_steps, cudaStream_t s) {
  if (rows <= 0 || n_part < 1 || cur_len < 1 || cur_len >= max_len || part == nullptr || (reinterpret_cast<uintptr_t>(part) & 15)) {
    set_last_error("token_step_partials: bad args"); return VC_ERR_BAD_ARG;
  }
  launch_pdl(token_step_partials_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, static_cast<const float4*>(part), n_part, rows, cur_len,
             max_len, pad_id, eos_ids, n_eos, ids, unfinished, sum_lp, n_steps);
  return check_launch("token_step_partials");
}

}  // namespace vc
