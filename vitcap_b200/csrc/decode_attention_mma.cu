// Single-step decoder self-attention over the KV cache, bf16 storage (fast mode). HBM-bound: the kernel's job is to keep
// ~40 KB of K/V reads in flight per SM at all times and to spend as few issue slots per key as possible.
//
// Semantics are those of decode_attention.cu (BertSelfAttention with history, modeling_bert.py:303-340; mask of
// modeling_bert.py:1494-1501 / dataset.py:371-390 encoded structurally): every sequence has two query rows per step
// (row 2r = last generated token, row 2r+1 = [MASK]); keys = the 578 context rows of the sequence's image (shared by all
// beams / samples of the image) + the sequence's own caption rows (earlier steps via the beam ancestor table, then this
// step's token and MASK rows; the token query never sees the MASK key).
//
// One CTA = one (image, head) and up to 8 sequences of that image = up to 16 query rows = ONE mma.sync M-tile, so the
// shared context K/V is read once for all beams / samples and the per-key instruction cost does not grow with them:
//   * the key axis is the union [context keys | caption keys of sequence 0 | sequence 1 | ...]; a small shared table
//     tells which sequence a caption key belongs to (row r may see key k iff seq(k) == r / 2 and not (r even, k = MASK))
//   * 4 warps take 16-key blocks round-robin; each warp owns a private 3-stage cp.async ring (16 keys x (K 128 B + V 128 B)
//     per stage, 16-byte chunks XOR-swizzled for conflict-free ldmatrix), so loads of later blocks are always in flight
//     while a block is being multiplied and no CTA-wide barrier sits in the main loop
//   * S = Q K^T and O += P V on mma.sync.m16n8k16 (bf16 operands, fp32 accumulators); the legacy warp-level MMA is the
//     right tool here: M = 16 rows cannot feed tcgen05 (M >= 64) and the op is two orders of magnitude below the tensor
//     roofline anyway -- what matters is that one 16-key block costs ~90 issue slots instead of ~800 FFMA-path slots
//   * online softmax in the exp2 domain with lazy rescaling (the running reference only moves when the block max exceeds
//     it by more than 2^8), per-warp partial (m, l, O) merged through shared memory at the end.
#include "common.cuh"

namespace vc {

namespace {
constexpr int KB = 16;                      // keys per block
constexpr int NST = 3;                      // cp.async stages per warp
constexpr int WARPS = 4;
constexpr int STAGE_BYTES = 2 * KB * 128;   // K tile 2 KB + V tile 2 KB
constexpr int SMEM_BYTES = WARPS * NST * STAGE_BYTES;   // 48 KB
constexpr int MAX_SEQ = 8;                  // sequences per CTA (16 query rows)
constexpr int MAX_CAP = MAX_SEQ * 64;       // caption keys per CTA (max_len <= 64)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef VC_STORE_F16
#define VC_MMA_16816 "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32"
#else
#define VC_MMA_16816 "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32"
#endif
  asm volatile(VC_MMA_16816 " {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace

// UPPER: query rows 8..15 are in use (more than 4 sequences in this CTA)
template <bool UPPER>
__global__ void __launch_bounds__(128, 4)
decode_attention_mma_kernel(const bf16* __restrict__ ctx_qkv, const bf16* __restrict__ step_qkv, const int* __restrict__ anc,
                            bf16* __restrict__ out, int Cs, const int* __restrict__ ctx_vis, int H, int R, int E, int cur_len,
                            float scale_log2, const int* __restrict__ seq_unfinished, const int* __restrict__ img_done) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint8_t cap_info[MAX_CAP];          // (sequence-in-CTA << 1) | is_mask_key, per caption key
  __shared__ float sm_m[WARPS][16], sm_l[WARPS][16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;         // mma fragment coordinates
  const int lane7 = lane & 7, mi = lane >> 3;    // ldmatrix address coordinates
  const int h = blockIdx.x;
  const int groups = (E + MAX_SEQ - 1) / MAX_SEQ;
  const int b = blockIdx.y / groups;
  const int e0 = (blockIdx.y % groups) * MAX_SEQ;
  const int EC = min(MAX_SEQ, E - e0);           // sequences handled by this CTA
  // context rows of an image: Cs allocated, the first C visible (label-region masks hide a per-image tail, dataset.py:405-408)
  pdl_launch_dependents();
  pdl_wait();                                    // programmatic dependent launch: global memory from here on
  // Finished sequences (modeling_utils.py:858-867: their tokens are PAD from now on and their scores frozen; beam search: the
  // image is `done`, modeling_utils.py:1003-1011) need no attention: a CTA none of whose sequences is live returns before it
  // touches the K/V cache -- the sweep that dominates a decode step then shrinks with the number of live captions. Their
  // output rows keep the (finite) values of an earlier step; everything computed from them is discarded by the token step.
  if (img_done != nullptr) {
    if (img_done[b] != 0) return;
  } else if (seq_unfinished != nullptr) {
    bool live = false;
    for (int e = 0; e < EC; ++e) live |= (seq_unfinished[b * E + e0 + e] != 0);
    if (!live) return;
  }
  const int C = ctx_vis ? ctx_vis[b] : Cs;
  const size_t ld = 3 * (size_t)H;
  const int step = cur_len - 1;
  const int nk = cur_len + 1;                    // caption keys per sequence: steps 0..step-1, this token, this MASK
  const int NK = C + EC * nk;                    // virtual key axis
  const int nblocks = (NK + KB - 1) / KB;
  const bf16* cur = step_qkv + (size_t)step * 2 * R * ld;
  const bf16* ctx_k = ctx_qkv + (size_t)b * Cs * ld + H + h * 64;    // K of context key 0 (V is H elements further)

  for (int i = threadIdx.x; i < EC * nk; i += 128) {
    const int e = i / nk, j = i - e * nk;
    cap_info[i] = (uint8_t)((e << 1) | (j == nk - 1));
  }

  // ---- Q fragments (A operand of S = Q K^T): rows g and g + 8, 4 k-steps of 16 dims ----
  uint32_t qf[4][4];
  {
    const bf16* q_lo = nullptr;
    const bf16* q_hi = nullptr;
    if (g < 2 * EC) q_lo = cur + (size_t)(2 * (b * E + e0 + (g >> 1)) + (g & 1)) * ld + h * 64;
    if (UPPER && g + 8 < 2 * EC) q_hi = cur + (size_t)(2 * (b * E + e0 + ((g + 8) >> 1)) + (g & 1)) * ld + h * 64;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      qf[kk][0] = q_lo ? *reinterpret_cast<const uint32_t*>(q_lo + 16 * kk + 2 * t) : 0u;
      qf[kk][2] = q_lo ? *reinterpret_cast<const uint32_t*>(q_lo + 16 * kk + 8 + 2 * t) : 0u;
      qf[kk][1] = q_hi ? *reinterpret_cast<const uint32_t*>(q_hi + 16 * kk + 2 * t) : 0u;
      qf[kk][3] = q_hi ? *reinterpret_cast<const uint32_t*>(q_hi + 16 * kk + 8 + 2 * t) : 0u;
    }
  }

  // ---- per-warp cp.async ring ----
  const uint32_t ring = smem_u32(smem) + warp * (NST * STAGE_BYTES);
  const int n_it = (nblocks - warp + WARPS - 1) / WARPS;            // blocks warp, warp + 4, ... (may be <= 0)
  auto issue = [&](int it) {
    const int vk0 = (warp + it * WARPS) * KB;
    const uint32_t st = ring + (it % NST) * STAGE_BYTES;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = mi + 4 * i;                // key row inside the block; this lane moves 16-byte chunk `lane7` of it
      const int vk = vk0 + row;
      const bf16* src = ctx_k;
      uint32_t bytes = 16;
      if (vk < C) {
        src = ctx_k + (size_t)vk * ld;
      } else if (vk < NK) {
        const int c = vk - C;
        const int e = c / nk, j = c - e * nk;
        const int r = b * E + e0 + e;
        if (j < step) {
          const int srow = anc ? anc[(size_t)j * R + r] : r;
          src = step_qkv + ((size_t)j * 2 * R + 2 * srow) * ld + H + h * 64;
        } else {
          src = cur + (size_t)(2 * r + (j - step)) * ld + H + h * 64;
        }
      } else {
        bytes = 0;                               // past the end: zero-fill (V must be finite, P is 0 there)
      }
      const uint32_t dst = st + row * 128 + ((lane7 ^ (row & 7)) << 4);
      cp_async16(dst, src + lane7 * 8, bytes);
      cp_async16(dst + KB * 128, src + H + lane7 * 8, bytes);
    }
  };
#pragma unroll
  for (int it = 0; it < NST - 1; ++it) {
    if (it < n_it) issue(it);
    cp_async_commit();
  }
  __syncthreads();                               // cap_info visible

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  for (int it = 0; it < n_it; ++it) {
    cp_async_wait<NST - 2>();
    __syncwarp();
    if (it + NST - 1 < n_it) issue(it + NST - 1);
    cp_async_commit();

    const int vk0 = (warp + it * WARPS) * KB;
    const uint32_t stK = ring + (it % NST) * STAGE_BYTES;
    const uint32_t stV = stK + KB * 128;
    // ---- S = Q K^T for 16 keys: two n-tiles of 8 keys ----
    float s[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t kf[4];
        ldsm_x4(kf, stK + (8 * j + lane7) * 128 + (((4 * q + mi) ^ lane7) << 4));
        mma16816(s[j], qf[2 * q], kf[0], kf[1]);
        mma16816(s[j], qf[2 * q + 1], kf[2], kf[3]);
      }
    }
    // ---- scale, mask, block max ----
    float x[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) x[j][c] = s[j][c] * scale_log2;
    const bool plain = (vk0 + KB <= C);          // warp-uniform: a block of context keys only, everything visible
    if (!plain) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int vk = vk0 + 8 * j + 2 * t + c;
          bool vis_lo = true, vis_hi = true;
          if (vk >= NK) {
            vis_lo = vis_hi = false;
          } else if (vk >= C) {
            const int info = cap_info[vk - C];
            const int seq = info >> 1;
            const bool ismask = info & 1;
            vis_lo = (seq == (g >> 1)) && !(ismask && !(g & 1));
            vis_hi = (seq == ((g + 8) >> 1)) && !(ismask && !(g & 1));
          }
          if (!vis_lo) x[j][c] = -INFINITY;
          if (!vis_hi) x[j][2 + c] = -INFINITY;
        }
    }
    float mx_lo = fmaxf(fmaxf(x[0][0], x[0][1]), fmaxf(x[1][0], x[1][1]));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    float mx_hi = -INFINITY;
    if (UPPER) {
      mx_hi = fmaxf(fmaxf(x[0][2], x[0][3]), fmaxf(x[1][2], x[1][3]));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    }
    const bool need = (mx_lo > m_lo + 8.f) || (UPPER && mx_hi > m_hi + 8.f);
    if (__any_sync(0xffffffffu, need)) {
      // integer references (as in attention_tc.cu): the bf16 rounding of P then depends on the scores only, not on which warp
      // met which key block first, and every rescale / merge factor is an exact power of two
      const float n_lo = (mx_lo > m_lo + 8.f) ? ceilf(mx_lo) : m_lo;
      const float n_hi = (UPPER && mx_hi > m_hi + 8.f) ? ceilf(mx_hi) : m_hi;
      const float c_lo = (n_lo == m_lo) ? 1.f : ex2f(m_lo - n_lo);      // m = -inf -> 0 (nothing accumulated yet)
      const float c_hi = (n_hi == m_hi) ? 1.f : ex2f(m_hi - n_hi);
      m_lo = n_lo; m_hi = n_hi;
      l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= c_lo; o[i][1] *= c_lo; o[i][2] *= c_hi; o[i][3] *= c_hi; }
    }
    // ---- P = exp2(x - m) as bf16 A fragments ----
    float p[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      p[j][0] = (x[j][0] == -INFINITY) ? 0.f : ex2f(x[j][0] - m_lo);
      p[j][1] = (x[j][1] == -INFINITY) ? 0.f : ex2f(x[j][1] - m_lo);
      if (UPPER) {
        p[j][2] = (x[j][2] == -INFINITY) ? 0.f : ex2f(x[j][2] - m_hi);
        p[j][3] = (x[j][3] == -INFINITY) ? 0.f : ex2f(x[j][3] - m_hi);
      } else {
        p[j][2] = p[j][3] = 0.f;
      }
    }
    l_lo += (p[0][0] + p[0][1]) + (p[1][0] + p[1][1]);
    if (UPPER) l_hi += (p[0][2] + p[0][3]) + (p[1][2] + p[1][3]);
    uint32_t pa[4];
    pa[0] = pack_bf16x2(p[0][0], p[0][1]);
    pa[1] = pack_bf16x2(p[0][2], p[0][3]);
    pa[2] = pack_bf16x2(p[1][0], p[1][1]);
    pa[3] = pack_bf16x2(p[1][2], p[1][3]);
    // ---- O += P V : 8 n-tiles of 8 dims ----
#pragma unroll
    for (int xx = 0; xx < 4; ++xx) {
      uint32_t vf[4];
      ldsm_x4_t(vf, stV + ((mi & 1) * 8 + lane7) * 128 + (((2 * xx + (mi >> 1)) ^ lane7) << 4));
      mma16816(o[2 * xx], pa, vf[0], vf[1]);
      mma16816(o[2 * xx + 1], pa, vf[2], vf[3]);
    }
  }
  cp_async_wait<0>();

  // ---- merge the 4 warps: partial (m, l, O) per row through shared memory (ring storage is reused) ----
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  if (UPPER) {
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  }
  __syncthreads();                               // every warp is done with its ring
  float* sm_o = reinterpret_cast<float*>(smem);  // [WARPS][16][64]
  if (t == 0) {
    sm_m[warp][g] = m_lo; sm_l[warp][g] = l_lo;
    sm_m[warp][g + 8] = m_hi; sm_l[warp][g + 8] = l_hi;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    *reinterpret_cast<float2*>(sm_o + ((warp * 16 + g) * 64 + 8 * i + 2 * t)) = make_float2(o[i][0], o[i][1]);
    if (UPPER) *reinterpret_cast<float2*>(sm_o + ((warp * 16 + g + 8) * 64 + 8 * i + 2 * t)) = make_float2(o[i][2], o[i][3]);
  }
  __syncthreads();
  {
    const int row = threadIdx.x >> 3, pt = threadIdx.x & 7;     // 16 rows x 8 chunks of 8 dims
    if (row < 2 * EC) {
      float M = sm_m[0][row];
#pragma unroll
      for (int w = 1; w < WARPS; ++w) M = fmaxf(M, sm_m[w][row]);
      float L = 0.f, acc[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const float mw = sm_m[w][row];
        const float c = (mw == -INFINITY) ? 0.f : ex2f(mw - M);
        L += sm_l[w][row] * c;
        const float4 a0 = *reinterpret_cast<const float4*>(sm_o + ((w * 16 + row) * 64 + pt * 8));
        const float4 a1 = *reinterpret_cast<const float4*>(sm_o + ((w * 16 + row) * 64 + pt * 8 + 4));
        acc[0] += a0.x * c; acc[1] += a0.y * c; acc[2] += a0.z * c; acc[3] += a0.w * c;
        acc[4] += a1.x * c; acc[5] += a1.y * c; acc[6] += a1.z * c; acc[7] += a1.w * c;
      }
      const float inv = 1.f / L;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] *= inv;
      const int r = b * E + e0 + (row >> 1);
      store8<bf16>(out + (size_t)(2 * r + (row & 1)) * H + h * 64 + pt * 8, acc);
    }
  }
}

// ctx_qkv [B, C, 3H]; step_qkv [max_len, 2*B*E, 3H]; anc int32 [max_len, B*E] or NULL; out [2*B*E, H]; all bf16
int decode_attention_mma(const void* ctx_qkv, const void* step_qkv, const int* anc, void* out, int B, int C, const int* ctx_vis,
                         int heads, int E, int cur_len, float scale, const int* seq_unfinished, const int* img_done, cudaStream_t s) {
  const int groups = (E + MAX_SEQ - 1) / MAX_SEQ;
  if (B <= 0 || C <= 0 || heads <= 0 || E <= 0 || cur_len < 1 || cur_len + 1 > 64 || (size_t)B * groups > 65535) {
    set_last_error("decode_attention: bad args B=%d C=%d heads=%d E=%d cur_len=%d", B, C, heads, E, cur_len);
    return VC_ERR_BAD_ARG;
  }
  const int H = heads * 64, R = B * E;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(decode_attention_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(decode_attention_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("decode_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VC_ERR_LAUNCH; }
    configured = true;
  }
  const float scale_log2 = scale * 1.4426950408889634f;
  const dim3 grid(heads, B * groups);
  if (E > 4)
    launch_pdl(decode_attention_mma_kernel<true>, grid, dim3(128), SMEM_BYTES, s, (const bf16*)ctx_qkv, (const bf16*)step_qkv, anc,
               (bf16*)out, C, ctx_vis, H, R, E, cur_len, scale_log2, seq_unfinished, img_done);
  else
    launch_pdl(decode_attention_mma_kernel<false>, grid, dim3(128), SMEM_BYTES, s, (const bf16*)ctx_qkv, (const bf16*)step_qkv, anc,
               (bf16*)out, C, ctx_vis, H, R, E, cur_len, scale_log2, seq_unfinished, img_done);
  return check_launch("decode_attention_mma");
}

}  // namespace vc
