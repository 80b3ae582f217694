// Self-attention for sequences of N = PEEL + 96 m tokens (ViT-B/16-384: 577 = 1 + 6 x 96; decoder context: 578 = 2 + 6 x 96):
// the variant of attention_tc.cu without padding work. Same contract: qkv [B, N, 3H] bf16 -> out [B, N, H] bf16 =
// softmax(Q K^T * scale) V (Attention.forward, vision_transformer.py:174-200; BertSelfAttention over the context rows,
// modeling_bert.py:303-340).
//
// What the 64-key kernel spends on the one token that does not fit (round-1 profile, N = 577: 10 pipeline steps per query tile,
// the tenth for ONE key; 7 % of the softmax warps' samples in that ragged step, another 30 % in per-step synchronisation):
//   * keys run in chunks of 96 = 576 / 6: six pipeline steps per 128-query tile instead of ten. TMEM: S0 | S1 (96 fp32 columns
//     each) | O (64) = 256 columns, still two CTAs per SM; K / V rings of three 12 KB stages.
//   * the PEEL leading keys (the CLS token; the tag-CLS row as well in the decoder context) never enter the tensor-core
//     pipeline: their scores are a 64-term dot product per query row on the CUDA cores (Q row read from the swizzled tile the
//     TMA delivered), their P V contribution a rank-PEEL update of the output row in the tile epilogue. They take part in
//     the row maximum and the row sum like every other key; P is rounded to bf16 for the V product exactly as the tensor-core
//     keys' P (oracle/port.py QuantPortModel.attend).
// Roles as in attention_tc.cu: warp 0 = TMA loader, warp 1 = MMA issuer, warps 2-5 = softmax (one query row per thread), P
// written back into the first 48 columns of its own S buffer and consumed from TMEM by the O += P V MMA.
#include "common.cuh"

#include <cstdlib>
#include <type_traits>

namespace vc {

namespace {
constexpr int QT = 128;                    // queries per tile
constexpr int KT = 96;                     // keys per chunk
constexpr int D = 64;                      // head dim
constexpr int Q_BYTES = QT * D * 2;        // 16 KB  [128 rows][64 bf16], 128B swizzle
constexpr int KV_BYTES = KT * D * 2;       // 12 KB  [96 keys][64 bf16]
constexpr int NKV = 3;                     // K/V ring depth
constexpr int NS = 2;                      // S buffers
constexpr int MAX_PEEL = 2;
constexpr int SMEM_Q = 0;                  // 2 buffers
constexpr int SMEM_K = SMEM_Q + 2 * Q_BYTES;
constexpr int SMEM_V = SMEM_K + NKV * KV_BYTES;
constexpr int SMEM_X = SMEM_V + NKV * KV_BYTES;        // peeled keys: kx [MAX_PEEL][64] f32, vx [MAX_PEEL][64] f32
constexpr int SMEM_BAR = SMEM_X + 2 * MAX_PEEL * D * 4;
constexpr int SMEM_TOTAL = SMEM_BAR + 256;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_O = NS * KT;  // S0 | S1 | O; P_g aliases columns [0, 48) of S_g
constexpr int HALF = KT / 2;               // the softmax warps work on 48 columns at a time (register budget)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 48 consecutive fp32 columns of this warp's lanes
__device__ __forceinline__ void tmem_ld_48(uint32_t taddr, uint32_t (&r)[48]) {
  tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
  tmem_ld_32x16(taddr + 32, *reinterpret_cast<uint32_t(*)[16]>(&r[32]));
  tmem_ld_wait();
}

// rescale the running output row by `corr` (rare: only when the exponent reference moved)
__device__ __forceinline__ void rescale_o(uint32_t taddr_o, float corr) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t o[32];
    tmem_ld_32x32(taddr_o + c * 32, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
    tmem_st_32x32(taddr_o + c * 32, o);
  }
  tmem_st_wait();
}
}  // namespace

template <int PEEL>
__global__ void __launch_bounds__(192, 2)
attention_tc96_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                      const bf16* __restrict__ qkv, bf16* __restrict__ out, int N, int H, float scale_log2) {
  static_assert(PEEL >= 0 && PEEL <= MAX_PEEL, "peeled keys");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* q_full = bars;           // [2] Q tile landed
  uint64_t* q_free = bars + 2;       // [2] every S MMA of the tile that used this Q buffer retired (tcgen05.commit)
  uint64_t* s_full = bars + 4;       // [NS] S_g complete (tcgen05.commit)
  uint64_t* p_full = bars + 6;       // [NS] P_g written to TMEM (128 arrivals)
  uint64_t* pv_done = bars + 8;      // [NS] O += P_g V complete (tcgen05.commit)
  uint64_t* o_free = bars + 10;      // [1] epilogue of the tile finished reading O (128 arrivals)
  uint64_t* k_full = bars + 11;      // [NKV]
  uint64_t* v_full = k_full + NKV;
  uint64_t* k_free = v_full + NKV;   // S MMAs that read this K stage retired (tcgen05.commit)
  uint64_t* v_free = k_free + NKV;   // PV MMAs that read this V stage retired (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_free + NKV);
  float* kx = reinterpret_cast<float*>(smem + SMEM_X);               // [MAX_PEEL][64]
  float* vx = kx + MAX_PEEL * D;                                     // [MAX_PEEL][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nch = (N - PEEL) / KT;                   // key chunks per tile (the launcher guarantees N = PEEL + 96 nch, nch >= 3)
  const int nq = (N + QT - 1) / QT;                  // query tiles
  const int total = nq * nch;                        // flattened pipeline steps

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();    // swizzled tiles need 1024-byte alignment
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int i = 0; i < NKV; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&k_free[i], 1); mbar_init(&v_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_free[i], 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&pv_done[i], 1); }
    mbar_init(&o_free[0], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();                        // programmatic dependent launch: global memory from here on

  if (warp == 0) {
    // ===================== TMA loader (one elected thread) =====================
    if (lane == 0) {
      const int cq = h * D, ck = H + h * D, cv = 2 * H + h * D;
      auto load_q = [&](int tile) {
        mbar_arrive_expect_tx(&q_full[tile & 1], Q_BYTES);
        tma_load_3d(smem + SMEM_Q + (tile & 1) * Q_BYTES, &tmap_q, &q_full[tile & 1], cq, tile * QT, b);
      };
      auto load_k = [&](int st, int j) {               // chunk j of the keys = rows [PEEL + 96 j, PEEL + 96 j + 96)
        mbar_arrive_expect_tx(&k_full[st], KV_BYTES);
        tma_load_3d(smem + SMEM_K + st * KV_BYTES, &tmap_kv, &k_full[st], ck, PEEL + j * KT, b);
      };
      auto load_v = [&](int st, int j) {
        mbar_arrive_expect_tx(&v_full[st], KV_BYTES);
        tma_load_3d(smem + SMEM_V + st * KV_BYTES, &tmap_kv, &v_full[st], cv, PEEL + j * KT, b);
      };
      load_q(0);
      for (int g = 0; g < NKV && g < total; ++g) { load_k(g, g % nch); load_v(g, g % nch); }
      if (nq > 1) load_q(1);
      // classic full/empty rings: every *_free barrier is re-armed only by this thread's own refill
      int tile = 0, j = 0;                            // step g
      int jn = NKV % nch;                             // chunk index of step g + NKV
      int st = 0;                                     // g % NKV
      uint32_t ph = 0;                                // (g / NKV) & 1
      for (int g = 0; g < total; ++g) {
        if (g + NKV < total) {
          mbar_wait(&k_free[st], ph);                 // S_g retired -> K stage free
          load_k(st, jn);
        }
        if (j == nch - 1 && tile + 2 < nq) {
          mbar_wait(&q_free[tile & 1], (tile >> 1) & 1);   // last S of this tile retired -> Q buffer free
          load_q(tile + 2);
        }
        if (g + NKV < total) {
          mbar_wait(&v_free[st], ph);                 // PV_g retired -> V stage free
          load_v(st, jn);
        }
        if (++j == nch) { j = 0; ++tile; }
        if (++jn == nch) jn = 0;
        if (++st == NKV) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    if (lane == 0) {
      const uint64_t qd0 = make_smem_desc_sw128(smem_u32(smem + SMEM_Q), 16, 1024);
      const uint64_t kd0 = make_smem_desc_sw128(smem_u32(smem + SMEM_K), 16, 1024);
      const uint64_t vd0 = make_smem_desc_sw128(smem_u32(smem + SMEM_V), 16, 1024);
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KT, 0, 0);             // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);              // P (TMEM) x V (MN-major)
      // S_g = Q_tile K_j^T into S buffer sbuf; st / ph: ring stage and phase of step g
      auto issue_s = [&](int sbuf, int tile, int j, int st, uint32_t ph) {
        if (j == 0) mbar_wait(&q_full[tile & 1], (tile >> 1) & 1);
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        const uint64_t qd = qd0 + (uint64_t)((tile & 1) * (Q_BYTES >> 4));
        const uint64_t kd = kd0 + (uint64_t)(st * (KV_BYTES >> 4));
        const uint32_t d = tmem_base + COL_S + sbuf * KT;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(d, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[sbuf]);
        umma_commit(&k_free[st]);
        if (j == nch - 1) umma_commit(&q_free[tile & 1]);
      };
      int t0 = 0, j0 = 0, st0 = 0;                      // step g: tile, chunk, ring stage
      uint32_t ph0 = 0;
      int t2 = 0, j2 = 0, st2 = 0;                      // step g + NS
      uint32_t ph2 = 0;
      for (int i = 0; i < NS && i < total; ++i) {
        issue_s(i, t2, j2, st2, ph2);
        if (++j2 == nch) { j2 = 0; ++t2; }
        if (++st2 == NKV) { st2 = 0; ph2 ^= 1; }
      }
      int pb = 0;                                       // g % NS
      uint32_t pph = 0;                                 // (g / NS) & 1
      for (int g = 0; g < total; ++g) {
        // everything the next MMAs need besides P_g is waited for FIRST, so that the issue follows the softmax warps' arrival
        mbar_wait(&v_full[st0], ph0);
        if (j0 == 0 && t0 >= 1) mbar_wait(&o_free[0], (t0 - 1) & 1);   // epilogue of the previous tile has read O
        if (g + NS < total) {
          if (j2 == 0) mbar_wait(&q_full[t2 & 1], (t2 >> 1) & 1);
          mbar_wait(&k_full[st2], ph2);
        }
        mbar_wait(&p_full[pb], pph);                   // P_g sits in TMEM (first 48 columns of S buffer pb)
        tc_fence_after();
        {
          // O_tile += P_g V_j : A from TMEM (16 keys = 8 columns of bf16 pairs per step), B = 16 keys = 2048 B of V
          const uint32_t pa = tmem_base + COL_S + pb * KT;
          const uint64_t vd = vd0 + (uint64_t)(st0 * (KV_BYTES >> 4));
          const uint32_t d = tmem_base + COL_O;
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) umma_f16_ts(d, pa + 8 * k, vd + 128 * k, idesc_o, (j0 | k) != 0);
          umma_commit(&pv_done[pb]);
          umma_commit(&v_free[st0]);
        }
        // S_{g+NS} reuses the buffer P_g occupied: issued after PV_g, the tensor pipe keeps the order
        if (g + NS < total) issue_s(pb, t2, j2, st2, ph2);
        if (++j0 == nch) { j0 = 0; ++t0; }
        if (++st0 == NKV) { st0 = 0; ph0 ^= 1; }
        if (++j2 == nch) { j2 = 0; ++t2; }
        if (++st2 == NKV) { st2 = 0; ph2 ^= 1; }
        if (++pb == NS) { pb = 0; pph ^= 1; }
      }
    }
  } else {
    // ===================== softmax warps 2..5: thread t owns query row t of the current tile =====================
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int t = quad * 32 + lane;                    // 0..127 == TMEM lane == row inside the tile
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
    const uint32_t taddr_o = tmem_base + lane_base + COL_O;
    if (PEEL > 0) {
      // the peeled keys' K and V rows of this (image, head) as fp32: threads 0..63 convert K, 64..127 V
      const int d = t & 63;
      const int col = (t < 64 ? H : 2 * H) + h * D + d;
#pragma unroll
      for (int x = 0; x < PEEL; ++x)
        (t < 64 ? kx : vx)[x * D + d] = to_f32<bf16>(qkv[((size_t)b * N + x) * (3 * H) + col]);
      named_bar_sync(1, 128);
    }
    int sb = 0, sb_prev = NS - 1;                      // g % NS, (g - 1) % NS
    uint32_t sph = 0, sph_prev = 1;                    // (g / NS) & 1 and the same for g - 1
    // deferred epilogue of the previous tile (runs after the first chunk of the next tile, off the critical path)
    bool pend = false;
    float pend_l = 0.f;
    float pend_px[MAX_PEEL] = {0.f, 0.f};
    int pend_tile = 0, pend_sb = 0;
    uint32_t pend_ph = 0;

    // per-tile epilogue of one query row: (O + sum_x px vx) / l -> bf16 -> out
    auto store_o_row = [&](float l, const float (&px)[MAX_PEEL], int q) {
      uint32_t o[2][32];
      tmem_ld_32x32(taddr_o, o[0]);
      tmem_ld_32x32(taddr_o + 32, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_free[0]);                         // the O buffer may be overwritten by the next tile
      if (q < N) {
        bf16* op = out + ((size_t)b * N + q) * H + h * D;
        const float inv = 1.f / l;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[c][i + e]);
#pragma unroll
            for (int x = 0; x < PEEL; ++x) {
              const float4 v0 = *reinterpret_cast<const float4*>(vx + x * D + c * 32 + i);
              const float4 v1 = *reinterpret_cast<const float4*>(vx + x * D + c * 32 + i + 4);
              f[0] = fmaf(px[x], v0.x, f[0]); f[1] = fmaf(px[x], v0.y, f[1]); f[2] = fmaf(px[x], v0.z, f[2]); f[3] = fmaf(px[x], v0.w, f[3]);
              f[4] = fmaf(px[x], v1.x, f[4]); f[5] = fmaf(px[x], v1.y, f[5]); f[6] = fmaf(px[x], v1.z, f[6]); f[7] = fmaf(px[x], v1.w, f[7]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] *= inv;
            store8<bf16>(op + c * 32 + i, f);
          }
      }
    };

    for (int tile = 0; tile < nq; ++tile) {
      // exponent reference (scaled log2 units); lags the true max by < 8 + 1. Always an INTEGER: P = exp2(x - m_ref) is then a
      // power-of-two multiple of exp2(x - ceil(row max)) and its bf16 rounding is a function of the scores alone
      // (oracle/port.py QuantPortModel.attend states it that way; rescale factors are exact powers of two)
      float m_ref = -INFINITY;
      float l = 0.f;
      float sx[MAX_PEEL] = {0.f, 0.f};                 // raw scores of the peeled keys
      float px[MAX_PEEL] = {0.f, 0.f};                 // their P, rounded to bf16 (kept in step with m_ref)
      const bool idle = (tile * QT + quad * 32 >= N);  // every row of this warp lies past the last query
      if (PEEL > 0 && !idle) {
        mbar_wait(&q_full[tile & 1], (tile >> 1) & 1);
        const uint8_t* qrow = smem + SMEM_Q + (tile & 1) * Q_BYTES + t * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float qf[8];
          load8<bf16>(reinterpret_cast<const bf16*>(qrow + ((c ^ (t & 7)) << 4)), qf);
#pragma unroll
          for (int x = 0; x < PEEL; ++x) {
            const float4 k0 = *reinterpret_cast<const float4*>(kx + x * D + c * 8);
            const float4 k1 = *reinterpret_cast<const float4*>(kx + x * D + c * 8 + 4);
            float a = sx[x];
            a = fmaf(qf[0], k0.x, a); a = fmaf(qf[1], k0.y, a); a = fmaf(qf[2], k0.z, a); a = fmaf(qf[3], k0.w, a);
            a = fmaf(qf[4], k1.x, a); a = fmaf(qf[5], k1.y, a); a = fmaf(qf[6], k1.z, a); a = fmaf(qf[7], k1.w, a);
            sx[x] = a;
          }
        }
        // the peeled keys give every row an exponent reference BEFORE its first chunk arrives: chunk 0 then runs speculatively
        // like every other chunk (one pass over S instead of a maximum pass followed by an exponential pass)
        float mxp = sx[0];
#pragma unroll
        for (int x = 1; x < PEEL; ++x) mxp = fmaxf(mxp, sx[x]);
        m_ref = ceilf(mxp * scale_log2);
#pragma unroll
        for (int x = 0; x < PEEL; ++x) {
          const float p = ex2(fmaf(sx[x], scale_log2, -m_ref));
          l += p;
          px[x] = to_f32<bf16>(from_f32<bf16>(p));
        }
      }
      for (int j = 0; j < nch; ++j) {
        const uint32_t taddr_s = tmem_base + lane_base + COL_S + sb * KT;
        mbar_wait(&s_full[sb], sph);
        tc_fence_after();
        if (idle) {
          // rows 608..639 of the last tile at N = 577: their S rows are exact zeros of the zero-filled Q rows, which read as
          // P = 0 where the PV MMA looks; nothing to compute, keep in step
          tc_fence_before();
          mbar_arrive(&p_full[sb]);
        } else {
          uint32_t pk[KT / 2];                         // P as packed bf16 pairs
          float csum = 0.f;
          float mx0 = -INFINITY, mx1 = -INFINITY;
          // P = exp2(s * scale - mref) of 48 columns into pk[24 half ..], their sum into csum; WITH_MAX folds the maximum of
          // the raw scores into (mx0, mx1) in the same instruction stream
          auto exp_half = [&](int half, float mref, auto with_max) {
            uint32_t r[HALF];
            tmem_ld_48(taddr_s + half * HALF, r);
            const uint64_t neg2 = pack_f32x2(-mref, -mref);
            uint64_t sum_a = 0ull, sum_b = 0ull;       // two packed partial sums (bit pattern of +0.0f pairs)
#pragma unroll
            for (int i = 0; i < HALF; i += 4) {
              const float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
              const float s2 = __uint_as_float(r[i + 2]), s3 = __uint_as_float(r[i + 3]);
              if (decltype(with_max)::value) {
                mx0 = fmaxf(mx0, fmaxf(s0, s1));
                mx1 = fmaxf(mx1, fmaxf(s2, s3));
              }
              float x0, x1, x2, x3;
              unpack_f32x2(fma_f32x2(pack_f32x2(s0, s1), scale2, neg2), x0, x1);
              unpack_f32x2(fma_f32x2(pack_f32x2(s2, s3), scale2, neg2), x2, x3);
              const float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
              sum_a = add_f32x2(sum_a, pack_f32x2(p0, p1));
              sum_b = add_f32x2(sum_b, pack_f32x2(p2, p3));
              pk[half * (HALF / 2) + (i >> 1)] = pack_bf16x2(p0, p1);
              pk[half * (HALF / 2) + (i >> 1) + 1] = pack_bf16x2(p2, p3);
            }
            float a0, a1, a2, a3;
            unpack_f32x2(sum_a, a0, a1);
            unpack_f32x2(sum_b, a2, a3);
            csum += (a0 + a1) + (a2 + a3);
          };
          bool redo;
          if (PEEL == 0 && j == 0) {
            // first chunk of a tile without peeled keys: no reference yet, the maximum has to come first
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t r[HALF];
              tmem_ld_48(taddr_s + half * HALF, r);
#pragma unroll
              for (int i = 0; i < HALF; i += 4) {
                mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
                mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
              }
            }
            m_ref = ceilf(fmaxf(mx0, mx1) * scale_log2);
            redo = true;
          } else {
            // speculate that the reference holds (it moves only when the chunk max exceeds it by 2^8): the exponentials run
            // against the old reference while the max chain proceeds beside them, off the MUFU critical path
            exp_half(0, m_ref, std::true_type{});
            exp_half(1, m_ref, std::true_type{});
            const float mxs = fmaxf(mx0, mx1) * scale_log2;
            const bool need = mxs > m_ref + 8.0f;
            redo = __any_sync(0xffffffffu, need);
            if (redo) {
              const float m_new = need ? ceilf(mxs) : m_ref;
              const float corr = ex2(m_ref - m_new);   // exactly 1 for lanes that keep their reference, else a power of two
              if (j > 0) {                             // (at j == 0 the accumulator still holds the PREVIOUS tile's rows)
                mbar_wait(&pv_done[sb_prev], sph_prev);  // every issued PV has landed in TMEM
                tc_fence_after();
                rescale_o(taddr_o, corr);
              }
              l *= corr;
#pragma unroll
              for (int x = 0; x < PEEL; ++x) px[x] *= corr;
              m_ref = m_new;
              csum = 0.f;
            }
          }
          if (redo) {
            exp_half(0, m_ref, std::false_type{});
            exp_half(1, m_ref, std::false_type{});
          }
          l += csum;
          // P_g overwrites the first half of its own S buffer
          tmem_st_32x32(taddr_s, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
          tmem_st_32x16(taddr_s + 32, *reinterpret_cast<uint32_t(*)[16]>(&pk[32]));
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[sb]);
        }
        if (pend && j == 0) {
          // epilogue of the previous tile, after this tile's first chunk has been handed to the tensor core
          mbar_wait(&pv_done[pend_sb], pend_ph);
          tc_fence_after();
          store_o_row(pend_l, pend_px, pend_tile * QT + t);
          pend = false;
        }
        sb_prev = sb; sph_prev = sph;
        if (++sb == NS) { sb = 0; sph ^= 1; }
      }
      pend = true; pend_l = l; pend_tile = tile; pend_sb = sb_prev; pend_ph = sph_prev;
#pragma unroll
      for (int x = 0; x < MAX_PEEL; ++x) pend_px[x] = px[x];
    }
    if (pend) {
      mbar_wait(&pv_done[pend_sb], pend_ph);
      tc_fence_after();
      store_o_row(pend_l, pend_px, pend_tile * QT + t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int PEEL>
static int launch_attention96(dim3 grid, const CUtensorMap& tq, const CUtensorMap& tkv, const bf16* qkv, bf16* out, int N, int H,
                              float scale_log2, cudaStream_t s) {
  auto kern = attention_tc96_kernel<PEEL>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) { set_last_error("attention_tc96: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VC_ERR_LAUNCH; }
    configured = true;
  }
  launch_pdl(kern, grid, dim3(192), SMEM_TOTAL, s, tq, tkv, qkv, out, N, H, scale_log2);
  return check_launch("attention_tc96");
}

// number of leading keys the 96-key kernel would peel for this N, or -1 when the shape is not its own (attention_tc takes it)
int attention_tc96_peel(int N) {
  const char* knob = getenv("VITCAP_ATTN96");       // tuning / A-B knob, read per call (VITCAP_ATTN96=0: the 64-key kernel)
  if (knob != nullptr && knob[0] == '0') return -1;
  for (int p = 0; p <= MAX_PEEL; ++p)
    if (N - p >= 3 * KT && (N - p) % KT == 0) return p;
  return -1;
}

int attention_tc96(const void* qkv, void* out, int B, int N, int heads, float scale, cudaStream_t s) {
  const int peel = attention_tc96_peel(N);
  if (B <= 0 || heads <= 0 || B > 65535 || peel < 0) { set_last_error("attention_tc96: bad args (N=%d)", N); return VC_ERR_BAD_ARG; }
  const int H = heads * D;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_last_error("attention_tc96: pointers must be 16-byte aligned"); return VC_ERR_BAD_ARG;
  }
  CUtensorMap tq, tkv;
  int rc = get_tmap_3d_bf16(&tq, qkv, (uint64_t)B, (uint64_t)N, (uint64_t)3 * H, (uint64_t)3 * H, (uint64_t)N * 3 * H, QT, 64);
  if (rc) return rc;
  rc = get_tmap_3d_bf16(&tkv, qkv, (uint64_t)B, (uint64_t)N, (uint64_t)3 * H, (uint64_t)3 * H, (uint64_t)N * 3 * H, KT, 64);
  if (rc) return rc;
  dim3 grid(heads, B);
  const float scale_log2 = scale * 1.4426950408889634f;
  const bf16* in = reinterpret_cast<const bf16*>(qkv);
  bf16* o = reinterpret_cast<bf16*>(out);
  switch (peel) {
    case 0: return launch_attention96<0>(grid, tq, tkv, in, o, N, H, scale_log2, s);
    case 1: return launch_attention96<1>(grid, tq, tkv, in, o, N, H, scale_log2, s);
    default: return launch_attention96<2>(grid, tq, tkv, in, o, N, H, scale_log2, s);
  }
}

}  // namespace vc
