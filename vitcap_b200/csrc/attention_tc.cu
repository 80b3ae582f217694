// Flash-style self-attention on tcgen05 tensor cores (bf16 operands, fp32 softmax statistics).
//   qkv [B, N, 3H] bf16 (q | k | v, H = heads*64)  ->  out [B, N, H] bf16 = softmax(Q K^T * scale) V
// Replaces Attention.forward (vision_transformer.py:174-200; the additive mask is all zeros, modeling_bert.py:1415)
// and BertSelfAttention over the 578 context rows (modeling_bert.py:303-340). The N x N scores never leave the SM.
//
// One CTA = one (image, head, 128-query tile). TMEM: S (128 x 128 fp32, columns 0..127) and O (128 x 64 fp32, columns
// 128..191). Per 128-key chunk:  S = Q K_j^T (tcgen05.mma, both operands K-major smem)  ->  4 softmax warps, one query
// row per thread: tcgen05.ld S, online max/sum, rescale O in TMEM, P = exp2(.) as bf16 into a 128B-swizzled smem tile
// ->  O += P V_j (A = P K-major smem, B = V MN-major smem straight from the row-major qkv buffer).
// K/V chunks are double-buffered by TMA (3-D tensor map over [B, N, 3H]: rows past N are zero-filled by the hardware).
// Two CTAs are co-resident per SM (112 KB smem, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
#include "common.cuh"

namespace vc {

namespace {
constexpr int QT = 128;           // queries per CTA
constexpr int KT = 128;           // keys per chunk
constexpr int D = 64;             // head dim
constexpr int TILE_BYTES = 128 * 64 * 2;   // 16 KB: [128 rows][64 bf16], 128B swizzle
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = SMEM_Q + TILE_BYTES;              // 2 stages
constexpr int SMEM_V = SMEM_K + 2 * TILE_BYTES;          // 2 stages
constexpr int SMEM_P = SMEM_V + 2 * TILE_BYTES;          // 2 sub-tiles of 64 keys
constexpr int SMEM_BAR = SMEM_P + 2 * TILE_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR + 128;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_O = 128;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace

__global__ void __launch_bounds__(160, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap, bf16* __restrict__ out, int N, int H, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* bar_q = bars;            // Q tile landed
  uint64_t* k_full = bars + 1;       // [2]
  uint64_t* v_full = bars + 3;       // [2]
  uint64_t* s_full = bars + 5;       // S = QK^T complete (tcgen05.commit)
  uint64_t* p_full = bars + 6;       // P written + O rescaled (128 arrivals)
  uint64_t* pv_done = bars + 7;      // O += PV complete (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int nchunks = (N + KT - 1) / KT;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();    // swizzled tiles need 1024-byte alignment
    tma_prefetch_desc(&tmap);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== control warp: TMA loads + MMA issue (one elected thread) =====================
    if (lane == 0) {
      const int cq = h * D, ck = H + h * D, cv = 2 * H + h * D;
      mbar_arrive_expect_tx(bar_q, TILE_BYTES);
      tma_load_3d(smem + SMEM_Q, &tmap, bar_q, cq, q0, b);
      for (int j = 0; j < 2 && j < nchunks; ++j) {
        mbar_arrive_expect_tx(&k_full[j], TILE_BYTES);
        tma_load_3d(smem + SMEM_K + j * TILE_BYTES, &tmap, &k_full[j], ck, j * KT, b);
        mbar_arrive_expect_tx(&v_full[j], TILE_BYTES);
        tma_load_3d(smem + SMEM_V + j * TILE_BYTES, &tmap, &v_full[j], cv, j * KT, b);
      }
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KT, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);    // P (K-major) x V (MN-major)
      const uint32_t sq = smem_u32(smem + SMEM_Q);
      const uint32_t sp = smem_u32(smem + SMEM_P);
      const uint64_t qdesc = make_smem_desc_sw128(sq, 16, 1024);

      mbar_wait(bar_q, 0);
      // S_0
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(smem + SMEM_K), 16, 1024);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base + COL_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      for (int j = 0; j < nchunks; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        // O += P_j V_j
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        {
          const uint32_t sv = smem_u32(smem + SMEM_V + st * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) {
            const uint64_t pdesc = make_smem_desc_sw128(sp + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16, 1024);
            const uint64_t vdesc = make_smem_desc_sw128(sv + k * 2048, 16, 1024);   // 16 keys = 2 groups of 8 rows
            umma_f16(tmem_base + COL_O, pdesc, vdesc, idesc_o, (j | k) != 0);
          }
          umma_commit(pv_done);
        }
        // S_{j+1} right behind it (S is free: every softmax thread finished reading S_j before arriving on p_full)
        if (j + 1 < nchunks) {
          const int st1 = (j + 1) & 1;
          mbar_wait(&k_full[st1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t kdesc = make_smem_desc_sw128(smem_u32(smem + SMEM_K + st1 * TILE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base + COL_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
        // refill this K/V stage with chunk j+2 once PV_j has consumed it
        if (j + 2 < nchunks) {
          mbar_wait(pv_done, j & 1);
          mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
          tma_load_3d(smem + SMEM_K + st * TILE_BYTES, &tmap, &k_full[st], ck, (j + 2) * KT, b);
          mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
          tma_load_3d(smem + SMEM_V + st * TILE_BYTES, &tmap, &v_full[st], cv, (j + 2) * KT, b);
        }
      }
    }
  } else {
    // ===================== softmax warps: thread t owns query row t =====================
    const int t = threadIdx.x;                         // 0..127 == TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = smem + SMEM_P + t * 128;
    for (int j = 0; j < nchunks; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_base + lane_base + COL_S + c * 32, r[c]);
      tmem_ld_wait();
      const int kbase = j * KT;
      const bool ragged = (kbase + KT > N);
      float mx = m;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(r[c][i]) * scale_log2;
          if (ragged && kbase + c * 32 + i >= N) s = -INFINITY;
          r[c][i] = __float_as_uint(s);
          mx = fmaxf(mx, s);
        }
      const float corr = ex2(m - mx);                  // 0 for the first chunk (m = -inf)
      m = mx;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { p[e] = ex2(__uint_as_float(r[c][i + e]) - mx); sum += p[e]; }
          const int key = c * 32 + i;                  // 8 consecutive keys -> one 16-byte chunk
          const int sub = key >> 6, chunk = (key & 63) >> 3;
          uint4 pk = make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
          *reinterpret_cast<uint4*>(prow + sub * TILE_BYTES + ((chunk ^ (t & 7)) << 4)) = pk;
        }
      }
      l = l * corr + sum;
      if (j > 0) {
        // rescale the running output: PV_{j-1} must have landed in TMEM first
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        uint32_t o[2][32];
        tmem_ld_32x32(tmem_base + lane_base + COL_O, o[0]);
        tmem_ld_32x32(tmem_base + lane_base + COL_O + 32, o[1]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c][i] = __float_as_uint(__uint_as_float(o[c][i]) * corr);
        tmem_st_32x32(tmem_base + lane_base + COL_O, o[0]);
        tmem_st_32x32(tmem_base + lane_base + COL_O + 32, o[1]);
        tmem_st_wait();
      }
      fence_proxy_async_smem();                        // P (generic-proxy stores) -> visible to the MMA (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16 -> out[b, q0 + t, h*64 ..]
    mbar_wait(pv_done, (nchunks - 1) & 1);
    tc_fence_after();
    uint32_t o[2][32];
    tmem_ld_32x32(tmem_base + lane_base + COL_O, o[0]);
    tmem_ld_32x32(tmem_base + lane_base + COL_O + 32, o[1]);
    tmem_ld_wait();
    const int q = q0 + t;
    if (q < N) {
      const float inv = 1.f / l;
      bf16* op = out + ((size_t)b * N + q) * H + h * D;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[c][i + e]) * inv;
          store8<bf16>(op + c * 32 + i, f);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TMEM_COLS>(tmem_base);
}

int attention_tc(const void* qkv, void* out, int B, int N, int heads, float scale, cudaStream_t s) {
  if (B <= 0 || N <= 0 || heads <= 0 || B > 65535) { set_last_error("attention_tc: bad args"); return VC_ERR_BAD_ARG; }
  const int H = heads * D;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_last_error("attention_tc: pointers must be 16-byte aligned"); return VC_ERR_BAD_ARG;
  }
  CUtensorMap tm;
  int rc = get_tmap_3d_bf16(&tm, qkv, (uint64_t)B, (uint64_t)N, (uint64_t)3 * H, (uint64_t)3 * H, (uint64_t)N * 3 * H, 128, 64);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) { set_last_error("attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VC_ERR_LAUNCH; }
    configured = true;
  }
  dim3 grid((N + QT - 1) / QT, heads, B);
  const float scale_log2 = scale * 1.4426950408889634f;
  attention_tc_kernel<<<grid, 160, SMEM_TOTAL, s>>>(tm, reinterpret_cast<bf16*>(out), N, H, scale_log2);
  return check_launch("attention_tc");
}

}  // namespace vc
