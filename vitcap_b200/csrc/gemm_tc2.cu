// bf16 GEMM on CTA pairs (tcgen05 cta_group::2):  out[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ resid)
//
// The large GEMMs of the caption path (M = images x 577 rows) run here; gemm_tc.cu keeps the small-M decode shapes.
// Compared with the single-CTA kernel:
//   * a cluster of two CTAs (one per SM of a TPC) owns a 256 x 256 output tile; the leader's single elected thread issues
//     tcgen05.mma.cta_group::2 (M = 256), each CTA holds its 128 accumulator rows in its own TMEM and stages only its own
//     128 rows of A and HALF of the W tile, so the operand traffic into shared memory per SM drops from 48 KB to 32 KB per
//     k-block and the freed space deepens the TMA pipeline to 5-6 stages (round-1 profile: 3-4 stages of 48 KB left the
//     tensor pipe at 67-78 % on the single-CTA kernel)
//   * every CTA's TMA loads complete on the LEADER's full barrier (peer bit of the barrier address cleared); the leader's
//     tcgen05.commit multicasts the stage-free / accumulator-ready arrivals to both CTAs; the peer's epilogue warps release
//     the accumulator stage by a remote mbarrier arrive on the leader
//   * epilogue warps come in one or two groups of four (one warp per TMEM lane quadrant and group). Two groups split the
//     32/64-column chunks of a tile between them, which keeps two warps per scheduler busy on the activation math
//     (the erf-GELU epilogue was issue/latency bound with a single warp per scheduler)
//   * GELU is x/2 * (1 + tanh(z * (c0 + c1 z^2 + c2 z^4))), z = x / sqrt(2): a minimax fit of erf(z) by one MUFU.TANH
//     (|gelu error| < 3e-5 + tanh.approx error, far below the bf16 output resolution), 10 issue slots instead of 17.
//   * LayerNorm folded around the GEMM (LN = 1 / 2). Under the board's power cap every GB of HBM traffic costs ~0.11 ms
//     whether or not it overlaps tensor work (tools/power_probe.py), and a stand-alone LayerNorm pass is 1.36 GB. So the
//     residual GEMM that PRODUCES a stream row (LN = 1, "emit") also stores a bf16 copy of the row and its partial
//     (sum, sum of squares) per 256-column tile, and the GEMM that CONSUMES the normalised row (LN = 2, "fold") multiplies the
//     RAW bf16 row by gamma-scaled weights and applies the normalisation in its epilogue:
//         LN(x) W^T + b  =  rstd * (x (gamma o W)^T  -  mean * colsum(gamma o W))  +  (b + W beta)
//     (bf16 rounding is relative, so rounding x before instead of after the affine map gives the same error bound as long as
//     |mean| is not much larger than the row's standard deviation; the statistics are fp32 over the fp32 row).
#include "pair.cuh"

#include <cstdlib>

namespace vc {

namespace {

enum { ACT2_NONE = 0, ACT2_GELU = 1, ACT2_TANH = 2 };
enum { LN_NONE = 0, LN_EMIT = 1, LN_FOLD = 2, LN_EMIT_TMA = 3 };   // 3 = emit with the bf16 copy staged in shared memory (TMA store)

template <bool OUT_F32, bool RESID, int G, int LN = 0> struct Gemm2Cfg {
  static constexpr int BM = 128;                        // rows per CTA (256 per pair)
  static constexpr int BN = 256;                        // columns per pair tile; each CTA stages 128 rows of W
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;           // 16 KB
  static constexpr int B_BYTES = (BN / 2) * BK * 2;     // 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 16384;               // staging tile: 128 rows x 128 B
  static constexpr int NBUF_G = (G == 2) ? 2 : (RESID ? 4 : 2);     // staging tiles per epilogue group
  static constexpr int NBUF = NBUF_G * G;
  // LN = 1: the emitted bf16 row copy goes to global memory straight from registers (staging it would cost the fifth pipeline
  // stage of the K = 3072 GEMM); LN = 3: it is staged in two 128 x 64 tiles and stored by TMA (the HBM-bound K = 768 GEMM, where
  // the row-per-thread stores hurt and four stages are plenty)
  static constexpr int XB_BUFS = (LN == 3) ? 2 : 0;
  static constexpr int BUDGET = 227 * 1024 - 1024 - 512 - (NBUF + XB_BUFS) * EPI_BYTES;
  static constexpr int STAGES_MAX = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_MAX > 8 ? 8 : STAGES_MAX;
  static constexpr int TMEM_COLS = 512;                 // two accumulator stages of 256 columns
  static constexpr int CW = OUT_F32 ? 32 : 64;          // output columns per staging tile
  static constexpr int NCH = BN / CW;
  static constexpr int THREADS = 64 + 128 * G;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (NBUF + XB_BUFS) * EPI_BYTES + 1024 + 512;
  static_assert(!(RESID && G == 2), "the residual epilogue runs with one group (4 staging tiles)");
  static_assert((LN != 1 && LN != 3) || (OUT_F32 && RESID && G == 1), "emit: fp32 residual epilogue");
  static_assert(LN != 2 || (!OUT_F32 && !RESID), "fold: bf16 output");
  static_assert(STAGES >= 4, "pipeline too shallow");
};

template <int ACT> __device__ __forceinline__ float apply_act(float v) {
  if (ACT == ACT2_GELU) return gelu_erf_tanh(v);
  if (ACT == ACT2_TANH) return tanhf(v);
  return v;
}

}  // namespace

// LN = 1: xb = bf16 [M, N] copy of the output row (pitch ldxb), stats = float [M, n_tiles, 2] partial (sum, sum of squares)
// LN = 2: stats = the producer's partials [M, st_tiles, 2] of the K-wide input row, colsum = float [N] row sums of W (bf16
//         values, fp32 sum), inv_k = 1 / K, ln_eps; bias already holds b + W beta
// RLN (post-LN layers, BertSelfOutput / BertOutput, modeling_bert.py:353-357, 415-419): the residual is itself a LayerNorm output
//         that is never materialised: the residual tile holds the RAW fp32 row of the producer, rstats its partial sums
//         [M, rst_tiles, 2] (over r_inv_n = 1 / row width), and the epilogue adds (raw - mean) * rstd * rgamma + rbeta
struct LnArgs {
  bf16* xb;
  int ldxb;
  float* stats;
  const float* colsum;
  float ln_eps;
  float inv_k;
  int st_tiles;
  const float* rstats;
  const float* rgamma;
  const float* rbeta;
  float r_eps;
  float r_inv_n;
  int rst_tiles;
};

template <int ACT, bool OUT_F32, bool RESID, int G, int LN, bool RLN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * G, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                const __grid_constant__ CUtensorMap tmap_xb, const float* __restrict__ bias, LnArgs ln, int M, int N, int K) {
  using C = Gemm2Cfg<OUT_F32, RESID, G, LN>;
  constexpr bool EMIT = (LN == LN_EMIT || LN == LN_EMIT_TMA);
  static_assert(!RLN || (RESID && OUT_F32), "a normalised residual needs the fp32 residual epilogue");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + C::STAGES * C::STAGE_BYTES;
  uint8_t* xb_stage = epi + C::NBUF * C::EPI_BYTES;     // [XB_BUFS] bf16 staging tiles (LN = 3)
  uint64_t* bars = reinterpret_cast<uint64_t*>(xb_stage + C::XB_BUFS * C::EPI_BYTES);
  uint64_t* full_bar = bars;                            // [STAGES] used in the leader only (both CTAs' TMA bytes land here)
  uint64_t* empty_bar = bars + C::STAGES;               // [STAGES] per CTA, armed by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * C::STAGES;           // [2] per CTA, multicast commit
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;      // [2] leader only: 2 CTAs x 4G warps arrive
  uint64_t* res_full = bars + 2 * C::STAGES + 4;        // [NBUF] residual tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4 + C::NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_tiles = (M + 2 * C::BM - 1) / (2 * C::BM);
  const int n_tiles = (N + C::BN - 1) / C::BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = K / C::BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (RESID) tma_prefetch_desc(&tmap_res);
    if (LN == LN_EMIT_TMA) tma_prefetch_desc(&tmap_xb);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * 4 * G);
    }
    for (int i = 0; i < C::NBUF; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();                                      // (the cluster barrier below already orders the allocator's write of
                                                        // tmem_slot; this CTA barrier is what compute-sanitizer racecheck models)
  cluster_sync_all();                                   // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();                                           // programmatic dependent launch: global memory from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A rows, own half of the W tile) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / n_tiles) * (2 * C::BM) + (int)rank * C::BM;
        const int n0 = (tile % n_tiles) * C::BN + (int)rank * (C::BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * C::BK, m0);
          tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kb * C::BK, n0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * C::BM, C::BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * C::BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty_bar[stage]);          // frees this stage in both CTAs
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full[as]);               // accumulator complete -> both epilogues
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: G groups of 4 warps; thread <-> accumulator row of this CTA =====================
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
    const int grp = (warp - 2) >> 2;                    // epilogue group 0 / 1
    const int t = quad * 32 + lane;                     // row inside this CTA's half of the pair tile
    const bool leader = (((warp - 2) & 3) == 0 && lane == 0);
    const uint32_t sw = (uint32_t)(t & 7);
    uint8_t* gepi = epi + grp * C::NBUF_G * C::EPI_BYTES;
    uint64_t* gres = res_full + grp * C::NBUF_G;
    const int bar_id = 1 + grp;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t g = 0;                                     // this group's running chunk counter (buffer = g % NBUF_G)
    uint32_t pf_g = 0;                                  // residual prefetch cursor (leader of the group, RESID => G == 1)
    int pf_tile = cluster_id, pf_c = 0;
    auto prefetch_resid = [&]() {
      if (pf_tile >= num_tiles) return;
      const int pm0 = (pf_tile / n_tiles) * (2 * C::BM) + (int)rank * C::BM;
      const int pn0 = (pf_tile % n_tiles) * C::BN + pf_c * C::CW;
      const int b = pf_g % C::NBUF_G;
      mbar_arrive_expect_tx(&gres[b], C::EPI_BYTES);
      tma_load_2d(gepi + b * C::EPI_BYTES, &tmap_res, &gres[b], pn0, pm0);
      ++pf_g;
      ++pf_c;
      if (pf_c == C::NCH || (pf_tile % n_tiles) * C::BN + pf_c * C::CW >= N) { pf_c = 0; pf_tile += num_clusters; }
    };
    if (RESID && leader) {
      for (int i = 0; i < C::NBUF_G - 2; ++i) prefetch_resid();
    }
    // LN = 2: the partial (sum, sum of squares) of the NEXT tile's row are requested one tile ahead and only SUMMED when that
    // tile starts: the loads stay in flight behind the current tile's epilogue (round-2 profile: summing them at once put
    // 5 % of the epilogue warps' samples on the first add). Up to four partials in registers (K <= 1024), the rest summed at once.
    float2 nx_p[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float nx_s = 0.f, nx_q = 0.f;
    auto load_row_stats = [&](int row) {
#pragma unroll
      for (int i = 0; i < 4; ++i) nx_p[i] = make_float2(0.f, 0.f);
      nx_s = 0.f;
      nx_q = 0.f;
      if (row < M) {
        const float2* sp = reinterpret_cast<const float2*>(ln.stats) + (size_t)row * ln.st_tiles;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < ln.st_tiles) nx_p[i] = __ldg(sp + i);
        for (int i = 4; i < ln.st_tiles; ++i) { const float2 pq = __ldg(sp + i); nx_s += pq.x; nx_q += pq.y; }
      }
    };
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / n_tiles) * (2 * C::BM) + (int)rank * C::BM;
      const int n0 = (tile % n_tiles) * C::BN;
      // LN = 2: mean / rstd of this thread's input row from the producer's partial sums. The sums of the NEXT tile's row are
      // requested here and consumed one tile later, so their latency never sits in front of the accumulator read
      float ln_rstd = 0.f, ln_nm = 0.f;
      if (LN == LN_FOLD) {
        if (tile == cluster_id) load_row_stats(m0 + t);
        // (the same left-to-right order as a running sum over the partials)
        const float row_s = (((nx_p[0].x + nx_p[1].x) + nx_p[2].x) + nx_p[3].x) + nx_s;
        const float row_q = (((nx_p[0].y + nx_p[1].y) + nx_p[2].y) + nx_p[3].y) + nx_q;
        const float mean = row_s * ln.inv_k;
        const float var = fmaxf(fmaf(row_q, ln.inv_k, -mean * mean), 0.f);
        ln_rstd = (m0 + t < M) ? rsqrtf(var + ln.ln_eps) : 0.f;
        ln_nm = -mean * ln_rstd;
        const int nt = tile + num_clusters;
        if (nt < num_tiles) load_row_stats((nt / n_tiles) * (2 * C::BM) + (int)rank * C::BM + t);
      }
      float st_s = 0.f, st_q = 0.f;                      // LN = 1: this row's partial statistics over the tile's columns
      float r_rstd = 0.f, r_nm = 0.f;                    // RLN: the residual row's 1 / std and -mean / std
      if (RLN && m0 + t < M) {
        const float2* sp = reinterpret_cast<const float2*>(ln.rstats) + (size_t)(m0 + t) * ln.rst_tiles;
        float rs = 0.f, rq = 0.f;
        for (int i = 0; i < ln.rst_tiles; ++i) { const float2 pq = __ldg(sp + i); rs += pq.x; rq += pq.y; }
        const float mean = rs * ln.r_inv_n;
        r_rstd = rsqrtf(fmaxf(fmaf(rq, ln.r_inv_n, -mean * mean), 0.f) + ln.r_eps);
        r_nm = -mean * r_rstd;
      }
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = grp; c < C::NCH; c += G) {
        const int col0 = n0 + c * C::CW;
        if (col0 >= N) break;
        const int b = g % C::NBUF_G;
        uint8_t* ebuf = gepi + b * C::EPI_BYTES;
        uint8_t* myrow = ebuf + t * 128;
        // (1) the staging tile written NBUF_G chunks ago must have been read by its TMA store
        if (leader) {
          bulk_wait_read<1>();
          if (RESID) prefetch_resid();
        }
        named_bar_sync(bar_id, 128);
        const bool full = (col0 + C::CW <= N);
        // (2) accumulator row -> registers -> bias / activation / residual -> swizzled staging row
#pragma unroll
        for (int hh = 0; hh < C::CW / 32; ++hh) {
          const int cb = col0 + hh * 32;
          // LN = 2: the chunk's bias' and column sums are requested BEFORE the accumulator read (the TMEM load's "memory"
          // clobber would otherwise pin them behind its wait: 6 % of the epilogue warps' samples sat on the first use)
          float4 fb[LN == LN_FOLD ? 8 : 1], fc[LN == LN_FOLD ? 8 : 1];
          if (LN == LN_FOLD) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              fb[j] = fc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (full || cb + 4 * j < N) {             // N % 4 == 0 is checked by the launcher
                fb[j] = __ldg(reinterpret_cast<const float4*>(bias + cb + 4 * j));
                fc[j] = __ldg(reinterpret_cast<const float4*>(ln.colsum + cb + 4 * j));
              }
            }
          }
          // RLN: likewise the previous LayerNorm's gamma / beta of the chunk (re-applied to the raw residual tile below)
          float4 rg[RLN ? 8 : 1], rb[RLN ? 8 : 1];
          if (RLN) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              rg[j] = rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (full || cb + 4 * j < N) {
                rg[j] = __ldg(reinterpret_cast<const float4*>(ln.rgamma + cb + 4 * j));
                rb[j] = __ldg(reinterpret_cast<const float4*>(ln.rbeta + cb + 4 * j));
              }
            }
          }
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * C::BN + c * C::CW + hh * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (LN == LN_FOLD) {
            // v = rstd * acc + (nm * colsum + bias'),  nm = -mean * rstd   (columns past N: clipped by the TMA store)
            const uint64_t r2 = pack_f32x2(ln_rstd, ln_rstd), n2 = pack_f32x2(ln_nm, ln_nm);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              unpack_f32x2(fma_f32x2(r2, pack_f32x2(v[4 * j], v[4 * j + 1]), fma_f32x2(n2, pack_f32x2(fc[j].x, fc[j].y), pack_f32x2(fb[j].x, fb[j].y))),
                           v[4 * j], v[4 * j + 1]);
              unpack_f32x2(fma_f32x2(r2, pack_f32x2(v[4 * j + 2], v[4 * j + 3]), fma_f32x2(n2, pack_f32x2(fc[j].z, fc[j].w), pack_f32x2(fb[j].z, fb[j].w))),
                           v[4 * j + 2], v[4 * j + 3]);
            }
          } else if (bias != nullptr) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + cb + j));
                unpack_f32x2(add_f32x2(pack_f32x2(v[j], v[j + 1]), pack_f32x2(bb.x, bb.y)), v[j], v[j + 1]);
                unpack_f32x2(add_f32x2(pack_f32x2(v[j + 2], v[j + 3]), pack_f32x2(bb.z, bb.w)), v[j + 2], v[j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (cb + j < N) ? __ldg(bias + cb + j) : 0.f;
            }
          }
          if (ACT == ACT2_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_erf_tanh_x2(v[j], v[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act<ACT>(v[j]);
          }
          if (OUT_F32) {
            if (RESID) {
              mbar_wait(&gres[b], (g / C::NBUF_G) & 1);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 rr = *reinterpret_cast<const float4*>(myrow + ((j ^ sw) << 4));
                if (RLN) {
                  const float4 gg = rg[j], bb = rb[j];
                  v[4 * j] += fmaf(fmaf(rr.x, r_rstd, r_nm), gg.x, bb.x);
                  v[4 * j + 1] += fmaf(fmaf(rr.y, r_rstd, r_nm), gg.y, bb.y);
                  v[4 * j + 2] += fmaf(fmaf(rr.z, r_rstd, r_nm), gg.z, bb.z);
                  v[4 * j + 3] += fmaf(fmaf(rr.w, r_rstd, r_nm), gg.w, bb.w);
                } else {
                  v[4 * j] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(myrow + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (EMIT) {
              // statistics of the fp32 row (columns >= N hold zeros: TMA zero-fills the operand tiles and the residual tile,
              // and bias is not added there) and the bf16 copy: two 32-column chunks fill one 64-column staging tile
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float xv = (full || cb + j < N) ? v[j] : 0.f;
                st_s += xv;
                st_q = fmaf(xv, xv, st_q);
              }
              if (LN == LN_EMIT_TMA) {
                uint8_t* xrow = xb_stage + ((g >> 1) & 1) * C::EPI_BYTES + t * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(xrow + ((((c & 1) * 4 + j) ^ sw) << 4)) =
                      make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
              } else if (m0 + t < M) {
                // 64 contiguous bytes of this thread's row; the K = 3072 tile leaves the epilogue ample time for the 4 stores
                uint4* xp = reinterpret_cast<uint4*>(ln.xb + (size_t)(m0 + t) * ln.ldxb + cb);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  xp[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 pk = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                          pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
              *reinterpret_cast<uint4*>(myrow + (((hh * 4 + j) ^ sw) << 4)) = pk;
            }
          }
        }
        // (3) staging tile -> global (rows >= M and columns >= N are clipped by the tensor map)
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (leader) {
          tma_store_2d(&tmap_out, ebuf, col0, m0);
          // LN = 3: the bf16 copy rides in the bulk group of its second chunk (N % 64 == 0: chunks pair up), so the staging-tile
          // release logic of step (1) covers it: the tile is rewritten four chunks later
          if (LN == LN_EMIT_TMA && (c & 1)) tma_store_2d(&tmap_xb, xb_stage + ((g >> 1) & 1) * C::EPI_BYTES, col0 - C::CW, m0);
          bulk_commit();
        }
        ++g;
      }
      if (EMIT && m0 + t < M) {
        const int nt = tile % n_tiles;
        reinterpret_cast<float2*>(ln.stats)[(size_t)(m0 + t) * n_tiles + nt] = make_float2(st_s, st_q);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (leader) bulk_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();                                   // the peer's smem / TMEM stay alive until the leader's MMAs retired
  if (warp == 1) tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
template <int ACT, bool OUT_F32, bool RESID, int G, int LN = 0, bool RLN = false>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr, const float* bias,
                   int M, int N, int K, cudaStream_t stream, LnArgs ln = LnArgs(), const CUtensorMap* txb = nullptr) {
  using C = Gemm2Cfg<OUT_F32, RESID, G, LN>;
  auto kern = gemm_tc2_kernel<ACT, OUT_F32, RESID, G, LN, RLN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("gemm_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VC_ERR_LAUNCH; }
    configured = true;
  }
  const int tiles = ((M + 2 * C::BM - 1) / (2 * C::BM)) * ((N + C::BN - 1) / C::BN);
  int clusters = sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  launch_pdl(kern, dim3(2 * clusters), dim3(C::THREADS), C::SMEM_BYTES, stream, ta, tb, to, tr, txb ? *txb : to, bias, ln, M, N, K);
  return check_launch("gemm_tc2");
}

template <int ACT>
static int launch2_act(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr, const float* bias,
                       int out_f32, bool resid, int M, int N, int K, cudaStream_t s) {
  if (out_f32) {
    if (resid) return launch2<ACT, true, true, 1>(ta, tb, to, tr, bias, M, N, K, s);
    return launch2<ACT, true, false, 2>(ta, tb, to, tr, bias, M, N, K, s);
  }
  return launch2<ACT, false, false, 2>(ta, tb, to, tr, bias, M, N, K, s);
}

// same contract as gemm_bf16_tc (gemm_tc.cu)
int gemm_bf16_tc2(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32, int act,
                  const float* resid, int ldr, int M, int N, int K, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 64) != 0) { set_last_error("gemm_tc2: need K %% 64 == 0 (K=%d)", K); return VC_ERR_BAD_ARG; }
  if (resid && !out_f32) { set_last_error("gemm_tc2: a residual needs fp32 output"); return VC_ERR_BAD_ARG; }
  CUtensorMap ta, tb, to, tr;
  int rc = get_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 128, 64);
  if (rc) return rc;
  if (out_f32) rc = get_tmap_2d_f32(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 128, 32);
  else rc = get_tmap_2d_bf16(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 128, 64);
  if (rc) return rc;
  tr = to;
  if (resid) {
    rc = get_tmap_2d_f32(&tr, resid, (uint64_t)M, (uint64_t)N, (uint64_t)ldr, 128, 32);
    if (rc) return rc;
  }
  switch (act) {
    case ACT2_NONE: return launch2_act<ACT2_NONE>(ta, tb, to, tr, bias, out_f32, resid != nullptr, M, N, K, stream);
    case ACT2_GELU: return launch2_act<ACT2_GELU>(ta, tb, to, tr, bias, out_f32, resid != nullptr, M, N, K, stream);
    case ACT2_TANH: return launch2_act<ACT2_TANH>(ta, tb, to, tr, bias, out_f32, resid != nullptr, M, N, K, stream);
  }
  set_last_error("gemm_tc2: unknown activation %d", act);
  return VC_ERR_BAD_ARG;
}

// out (fp32) = A W^T + bias + resid, plus xb = bf16(out) and stats[M, ceil(N/256), 2] = per-256-column partial (sum, sum of
// squares) of every output row: the producer half of the folded LayerNorm
int gemm_bf16_tc2_ln_emit(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo, const float* resid,
                          int ldr, void* xb, int ldxb, float* stats, int M, int N, int K, cudaStream_t stream, const float* rstats,
                          int rst_tiles, const float* rgamma, const float* rbeta, float r_eps) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 64) != 0 || (N % 64) != 0 || resid == nullptr || xb == nullptr || stats == nullptr) {
    set_last_error("gemm_tc2_ln_emit: need K %% 64 == 0, N %% 64 == 0, a residual, xb and stats (N=%d K=%d)", N, K);
    return VC_ERR_BAD_ARG;
  }
  if ((lda % 8) || (ldw % 8) || (ldo % 4) || (ldr % 4) || (ldxb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(resid) & 15) ||
      (reinterpret_cast<uintptr_t>(xb) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15) || (reinterpret_cast<uintptr_t>(stats) & 7)) {
    set_last_error("gemm_tc2_ln_emit: pointers must be 16-byte aligned and row pitches multiples of 16 bytes");
    return VC_ERR_BAD_ARG;
  }
  CUtensorMap ta, tb, to, tr;
  int rc = get_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_f32(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 128, 32);
  if (rc) return rc;
  rc = get_tmap_2d_f32(&tr, resid, (uint64_t)M, (uint64_t)N, (uint64_t)ldr, 128, 32);
  if (rc) return rc;
  LnArgs ln = LnArgs();
  ln.xb = static_cast<bf16*>(xb);
  ln.ldxb = ldxb;
  ln.stats = stats;
  const bool rln = (rstats != nullptr);
  if (rln) {
    if (rst_tiles < 1 || rgamma == nullptr || rbeta == nullptr || (reinterpret_cast<uintptr_t>(rstats) & 7) ||
        (reinterpret_cast<uintptr_t>(rgamma) & 15) || (reinterpret_cast<uintptr_t>(rbeta) & 15)) {
      set_last_error("gemm_tc2_ln_emit: a normalised residual needs rstats (8-byte aligned), rst_tiles >= 1, rgamma and rbeta");
      return VC_ERR_BAD_ARG;
    }
    ln.rstats = rstats;
    ln.rst_tiles = rst_tiles;
    ln.rgamma = rgamma;
    ln.rbeta = rbeta;
    ln.r_eps = r_eps;
    ln.r_inv_n = 1.0f / (float)N;                        // the residual row is as wide as the output row
  }
  static const int force_tma = (getenv("VITCAP_EMIT_TMA") != nullptr) ? atoi(getenv("VITCAP_EMIT_TMA")) : -1;   // tuning knob
  if (force_tma == 1 || (force_tma != 0 && K <= 768)) {                       // HBM-bound shape: stage the copy in shared memory
    CUtensorMap tx;
    rc = get_tmap_2d_bf16(&tx, xb, (uint64_t)M, (uint64_t)N, (uint64_t)ldxb, 128, 64);
    if (rc) return rc;
    if (rln) return launch2<ACT2_NONE, true, true, 1, LN_EMIT_TMA, true>(ta, tb, to, tr, bias, M, N, K, stream, ln, &tx);
    return launch2<ACT2_NONE, true, true, 1, LN_EMIT_TMA>(ta, tb, to, tr, bias, M, N, K, stream, ln, &tx);
  }
  if (rln) return launch2<ACT2_NONE, true, true, 1, LN_EMIT, true>(ta, tb, to, tr, bias, M, N, K, stream, ln);
  return launch2<ACT2_NONE, true, true, 1, LN_EMIT>(ta, tb, to, tr, bias, M, N, K, stream, ln);
}

// out (bf16) = act(LN(x) W^T + b) computed from the RAW bf16 row copy A = xb and the producer's statistics:
// Wf = bf16(gamma o W), colsum[n] = sum_k float(Wf[n, k]), bias_f = b + W beta (all prepared by the host)
int gemm_bf16_tc2_ln_fold(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum,
                          const float* stats, int st_tiles, float ln_eps, void* out, int ldo, int act, int M, int N, int K,
                          cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 64) != 0 || (N % 4) != 0 || bias_f == nullptr || colsum == nullptr || stats == nullptr ||
      st_tiles < 1) {
    set_last_error("gemm_tc2_ln_fold: need K %% 64 == 0, N %% 4 == 0, bias, colsum and stats (N=%d K=%d)", N, K);
    return VC_ERR_BAD_ARG;
  }
  if ((lda % 8) || (ldw % 8) || (ldo % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Wf) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(bias_f) & 15) ||
      (reinterpret_cast<uintptr_t>(colsum) & 15) || (reinterpret_cast<uintptr_t>(stats) & 7)) {
    set_last_error("gemm_tc2_ln_fold: pointers must be 16-byte aligned and row pitches multiples of 16 bytes");
    return VC_ERR_BAD_ARG;
  }
  CUtensorMap ta, tb, to;
  int rc = get_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&tb, Wf, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 128, 64);
  if (rc) return rc;
  LnArgs ln = LnArgs();
  ln.stats = const_cast<float*>(stats);
  ln.colsum = colsum;
  ln.ln_eps = ln_eps;
  ln.inv_k = 1.0f / (float)K;
  ln.st_tiles = st_tiles;
  switch (act) {
    case ACT2_NONE: return launch2<ACT2_NONE, false, false, 2, LN_FOLD>(ta, tb, to, to, bias_f, M, N, K, stream, ln);
    case ACT2_GELU: return launch2<ACT2_GELU, false, false, 2, LN_FOLD>(ta, tb, to, to, bias_f, M, N, K, stream, ln);
  }
  set_last_error("gemm_tc2_ln_fold: unsupported activation %d", act);
  return VC_ERR_BAD_ARG;
}

}  // namespace vc
