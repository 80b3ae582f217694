// The GEMMs of a DECODE STEP (2 rows per sequence: M = 2 x sequences, at most a few thousand rows) on CTA pairs.
//
// What the general kernels (gemm_tc.cu / gemm_tc2.cu) leave on the table at these shapes (round-1 profile, M = 1024):
//   * fc2 (N = 768, K = 3072, split-bf16 operands): 96 single-CTA 128 x 64 tiles, each pulling 2.3 MB through ONE SM's L2 port:
//     33 us for 11 us of tensor work, 52 SMs idle  ->  256 x 256 pair tiles (3 x the flops per operand byte) and SPLIT-K over the
//     idle SMs; the fp32 partial planes are summed by the row-wise finish kernel (rowwise.cu) that applies bias, residual and
//     LayerNorm anyway
//   * fc1 / vocabulary head on split operands ran the K-concatenated form [hi | lo | hi] x [w_hi | w_hi | w_lo], loading a_hi and
//     w_hi twice  ->  a stage holds the four distinct tiles of a k-block and the three products are issued from them (X3)
//   * the GELU output of fc1 made an fp32 round trip to a split kernel  ->  the epilogue writes the split pair (hi, lo) itself
//   * the vocabulary logits (62 MB fp32 per step) were written, then read twice by the token kernel  ->  greedy decoding takes
//     per-tile (max, arg max, sum of exponentials) partials from the epilogue and never materialises the logits
//     (argmax -> log_softmax -> gather of modeling_utils.py:849-853 restated as an online reduction)
//
// Structure (as gemm_tc2.cu): cluster of two CTAs per 256 x 256 tile, cta_group::2 MMAs issued by the leader's elected thread,
// every CTA stages its own 128 A rows and 128 W rows by TMA; warp 0 = producer, warp 1 = MMA, warps 2-9 = two epilogue groups.
// Shared memory holds ONLY the operand pipeline (192 KB); the epilogue stages its output tiles in the pipeline's memory once
// the tile's MMAs have retired (at these shapes a cluster normally owns one tile, so there is nothing to overlap; with more
// tiles per cluster the producer waits for the epilogue -- EPI_ARGMAX, which stages nothing, keeps the mainloop running ahead
// through the second accumulator stage).
#include "pair.cuh"

#include <climits>

namespace vc {

namespace {

enum { EPI_PARTIAL = 0, EPI_BF16 = 1, EPI_GELU_BF16 = 2, EPI_GELU_SPLIT = 3, EPI_ARGMAX = 4 };

// BN: columns per pair tile (each CTA stages BN / 2 rows of W). 256 by default; 128 for the q|k|v projection (twice the tiles: the
// 2-sequence-row GEMM otherwise leaves half of the SMs idle); 208 for the vocabulary head (30522 columns = 147 tiles x 2 m-tiles =
// 294 tiles = 3.97 rounds of 74 clusters, where 256-wide tiles need 4 rounds for 3.24 rounds of work)
template <bool X3, int BN_> struct DecCfg {
  static constexpr int BM = 128;                        // rows per CTA (256 per pair)
  static constexpr int BN = BN_;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;           // 16 KB
  static constexpr int B_BYTES = (BN / 2) * BK * 2;     // <= 16 KB, a multiple of 1 KB (BN % 16 == 0)
  static constexpr int HALF_BYTES = A_BYTES + B_BYTES;  // (a_hi, w_hi); X3 adds (a_lo, w_lo) behind it
  static constexpr int STAGE_BYTES = X3 ? 2 * HALF_BYTES : HALF_BYTES;
  static constexpr int STAGES_FIT = (192 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;            // <= 192 KB
  static constexpr int TILE_BYTES = 16384;              // epilogue staging tile: 128 rows x 128 B (aliases the pipeline)
  static constexpr int G = 2;                           // epilogue groups of four warps
  static constexpr int THREADS = 64 + 128 * G;
  static constexpr int ACC_COLS = 256;                  // TMEM columns per accumulator stage
  static constexpr int TMEM_COLS = 512;                 // two accumulator stages
  static constexpr int BIAS_PAD = 256;                  // floats per staged bias slice (EPI_ARGMAX: [G][2 accumulator stages][256])
  static constexpr int BIAS_BYTES = G * 2 * BIAS_PAD * 4;
  static constexpr int SMEM_BYTES = PIPE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + BIAS_BYTES;
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "UMMA N of a CTA pair: multiple of 16, at most 256");
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert((BN / 32) * TILE_BYTES <= PIPE_BYTES, "the staging tiles of one output tile must fit in the pipeline memory");
};

struct DecArgs {
  const float* bias;    // [N] or NULL (EPI_PARTIAL: added only to a single plane, splits == 1 -- the fp32 logits form)
  float4* part;         // EPI_ARGMAX: [M, n_part] (max, arg max as int bits, sum of exp(x - max), unused)
  int n_part;
  int M, N, K;          // K = operand pitch along K (3 Kt with X3)
  int splits;           // split-K factor (EPI_PARTIAL); tiles are (m, n, split) triples
  int m_pad;            // rows per partial plane (multiple of 128)
  int f16;              // operands are IEEE halves (one product); the GELU epilogue then writes halves as well
};

}  // namespace

template <bool X3, int EPI, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DecCfg<X3, BN>::THREADS, 1)
gemm_dec_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_out, DecArgs ar) {
  using C = DecCfg<X3, BN>;
  static_assert(EPI == EPI_ARGMAX || BN % 64 == 0, "staged epilogues work in 32 / 64-column chunks");
  constexpr bool STAGED = (EPI != EPI_ARGMAX);          // the epilogue writes output tiles through the pipeline memory
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::PIPE_BYTES);
  uint64_t* full_bar = bars;                            // [STAGES] used in the leader only (both CTAs' TMA bytes land here)
  uint64_t* empty_bar = bars + C::STAGES;               // [STAGES] per CTA, armed by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * C::STAGES;           // [2] per CTA, multicast commit
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;      // [2] leader only: 2 CTAs x 4G warps arrive
  uint64_t* epi_free = bars + 2 * C::STAGES + 4;        // [1] per CTA: the staging tiles have been read by their TMA stores
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 5);
  float* sbias = reinterpret_cast<float*>(smem + C::PIPE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_tiles = (ar.M + 2 * C::BM - 1) / (2 * C::BM);
  const int n_tiles = (ar.N + C::BN - 1) / C::BN;
  const int num_tiles = m_tiles * n_tiles * ar.splits;
  const int Kt = X3 ? ar.K / 3 : ar.K;                  // X3: columns [0,Kt) = a_hi / w_hi, [Kt,2Kt) = a_lo, [2Kt,3Kt) = w_lo
  const int nkb = (Kt / C::BK) / ar.splits;             // k-blocks per tile
  // tile -> (m, n, split): consecutive clusters take the m-tiles of one n-tile, so that a W tile is in flight to them together
  auto tile_m = [&](int tile) { return (tile / ar.splits) % m_tiles; };
  auto tile_n = [&](int tile) { return (tile / ar.splits) / m_tiles; };
  auto tile_s = [&](int tile) { return tile % ar.splits; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (STAGED) tma_prefetch_desc(&tmap_out);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * 4 * C::G);
    }
    mbar_init(&epi_free[0], C::G);
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

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A rows, own half of the W tile) =====================
    if (lane == 0) {
      // W is a weight matrix, never written inside the decode loop: the W tiles of this cluster's first output tile are
      // requested BEFORE the dependency wait and travel while the predecessor kernel is still running
      int pre = 0;
      if (cluster_id < num_tiles) {
        const int n0 = tile_n(cluster_id) * C::BN + (int)rank * (C::BN / 2);
        const int kb0 = tile_s(cluster_id) * nkb;
        pre = nkb < C::STAGES ? nkb : C::STAGES;
        for (int i = 0; i < pre; ++i) {
          uint8_t* sb = smem + i * C::STAGE_BYTES + C::A_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[i], 2 * C::STAGE_BYTES);
          tma_load_2d_pair(sb, &tmap_b, &full_bar[i], (kb0 + i) * C::BK, n0);
          if (X3) tma_load_2d_pair(sb + C::HALF_BYTES, &tmap_b, &full_bar[i], 2 * Kt + (kb0 + i) * C::BK, n0);
        }
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0, ephase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = tile_m(tile) * (2 * C::BM) + (int)rank * C::BM;
        const int n0 = tile_n(tile) * C::BN + (int)rank * (C::BN / 2);
        const int kb0 = tile_s(tile) * nkb;
        if (STAGED && tile != cluster_id) {             // the previous tile's epilogue still owns the pipeline memory
          mbar_wait(&epi_free[0], ephase);
          ephase ^= 1;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const int kc = (kb0 + kb) * C::BK;
          if (pre > 0) {                                // barrier armed and the W half already in flight (fresh stage)
            --pre;
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kc, n0);
            if (X3) tma_load_2d_pair(sb + C::HALF_BYTES, &tmap_b, &full_bar[stage], 2 * Kt + kc, n0);
          }
          tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kc, m0);
          if (X3) tma_load_2d_pair(sa + C::HALF_BYTES, &tmap_a, &full_bar[stage], Kt + kc, m0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      pdl_wait();
      const uint32_t idesc = ar.f16 ? make_idesc_f16(2 * C::BM, C::BN, 0, 0) : make_idesc_bf16(2 * C::BM, C::BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * C::ACC_COLS;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          if (X3) {
            const uint64_t adesc_lo = make_smem_desc_sw128(sa + C::HALF_BYTES, 16, 1024);
            const uint64_t bdesc_lo = make_smem_desc_sw128(sb + C::HALF_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < C::BK / 16; ++k) umma_f16_pair(d_tmem, adesc_lo + 2 * k, bdesc + 2 * k, idesc, true);   // a_lo w_hi
#pragma unroll
            for (int k = 0; k < C::BK / 16; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc_lo + 2 * k, idesc, true);   // a_hi w_lo
          }
          umma_commit_pair(&empty_bar[stage]);          // frees this stage in both CTAs
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full[as]);               // accumulator complete -> both epilogues
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps; thread <-> accumulator row of this CTA =====================
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
    const int grp = (warp - 2) >> 2;                    // epilogue group 0 / 1: alternate column chunks
    const int t = quad * 32 + lane;                     // row inside this CTA's half of the pair tile
    const bool leader = (((warp - 2) & 3) == 0 && lane == 0);
    const uint32_t sw = (uint32_t)(t & 7);
    const int bar_id = 1 + grp;
    int as = 0;
    uint32_t aphase = 0;
    pdl_wait();                                         // global memory (bias, output) from here on
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = tile_m(tile) * (2 * C::BM) + (int)rank * C::BM;
      const int n0 = tile_n(tile) * C::BN;
      const int row = m0 + t;
      const float* sb = sbias + (grp * 2 + as) * C::BIAS_PAD;
      if (EPI == EPI_ARGMAX) {
        // the tile's bias slice goes to shared memory WHILE its MMAs run: every thread of a warp needs the same 32 values per
        // chunk, and as 32 scalar global loads per chunk they were what the epilogue waited for (ncu, session 4 of round 2:
        // 37 % of the kernel's stall samples on the FADD that consumes them, 21 000 cycles per tile against 5 000 of tensor
        // work). One buffer per (group, accumulator stage): the group's barrier of the NEXT tile separates reuse from reads.
        float* sw = sbias + (grp * 2 + as) * C::BIAS_PAD;
#pragma unroll
        for (int i = t; i < C::BIAS_PAD; i += 128)
          sw[i] = (ar.bias != nullptr && i < C::BN && n0 + i < ar.N) ? __ldg(ar.bias + n0 + i) : 0.f;
        named_bar_sync(bar_id, 128);
      }
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + as * C::ACC_COLS;

      if (EPI == EPI_ARGMAX) {
        // online (max, first arg max, sum of exp) over this thread's row and this group's chunks of 32 columns
        float mx = -INFINITY, sum = 0.f;
        int best = INT_MAX;
#pragma unroll 1
        for (int c = grp; c < (C::BN + 31) / 32; c += C::G) {
          const int col0 = n0 + c * 32;
          if (col0 >= ar.N) break;
          const bool half = (C::BN % 32 != 0) && (c == C::BN / 32);   // the 16-column remainder of a 208-wide tile
          uint32_t r[32];
          if (half) {
            uint32_t r16[16];
            tmem_ld_32x16(tacc + c * 32, r16);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) { r[j] = r16[j]; r[16 + j] = 0u; }
          } else {
            tmem_ld_32x32(tacc + c * 32, r);
            tmem_ld_wait();
          }
          const int lim = min(ar.N, half ? col0 + 16 : col0 + 32);    // columns of this chunk that belong to this tile and to N
          float v[32];
          float cm = -INFINITY;
          int ci = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const bool in = (col0 + j < lim);
            v[j] = in ? __uint_as_float(r[j]) + sb[c * 32 + j] : -INFINITY;
            if (v[j] > cm) { cm = v[j]; ci = j; }       // ascending scan: the first index on ties
          }
          if (cm > mx) {
            sum *= exp2f((mx - cm) * 1.4426950408889634f);            // mx = -inf: sum is 0 and stays 0
            mx = cm;
            best = col0 + ci;
          }
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += exp2f((v[j] - mx) * 1.4426950408889634f);   // exp2(-inf) = 0 outside
          sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
        if (row < ar.M)
          ar.part[(size_t)row * ar.n_part + tile_n(tile) * C::G + grp] = make_float4(mx, __int_as_float(best), sum, 0.f);
      } else {
        constexpr int CW = (EPI == EPI_PARTIAL) ? 32 : 64;            // output columns per 128-byte staging row
        constexpr int NCH = C::BN / CW;
        const int orow = (EPI == EPI_PARTIAL) ? tile_s(tile) * ar.m_pad + m0 : m0;
#pragma unroll 1
        for (int c = grp; c < NCH; c += C::G) {
          const int col0 = n0 + c * CW;
          if (col0 >= ar.N) break;
          // staging tiles: one per chunk (two for the split pair); nothing is reused within a tile, so no waits in here
          uint8_t* tile_hi = smem + (EPI == EPI_GELU_SPLIT ? 2 * c : c) * C::TILE_BYTES;
          uint8_t* row_hi = tile_hi + t * 128;
          uint8_t* row_lo = row_hi + C::TILE_BYTES;
#pragma unroll
          for (int hh = 0; hh < CW / 32; ++hh) {
            uint32_t r[32];
            tmem_ld_32x32(tacc + c * CW + hh * 32, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (EPI == EPI_PARTIAL) {
              if (ar.bias != nullptr) {                 // single plane with bias = the fp32 logits (N need not be a multiple of 4)
                const int cb = col0 + hh * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (cb + j < ar.N) v[j] += __ldg(ar.bias + cb + j);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(row_hi + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              const int cb = col0 + hh * 32;
              if (ar.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  if (cb + j < ar.N) {                  // N % 4 == 0 (launcher)
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(ar.bias + cb + j));
                    v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
                  }
                }
              }
              if (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_SPLIT) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) gelu_erf_tanh_x2(v[j], v[j + 1]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 hi, lo;
                if (EPI == EPI_GELU_SPLIT) {
                  split_bf16x2(v[8 * j], v[8 * j + 1], hi.x, lo.x);
                  split_bf16x2(v[8 * j + 2], v[8 * j + 3], hi.y, lo.y);
                  split_bf16x2(v[8 * j + 4], v[8 * j + 5], hi.z, lo.z);
                  split_bf16x2(v[8 * j + 6], v[8 * j + 7], hi.w, lo.w);
                  *reinterpret_cast<uint4*>(row_lo + (((hh * 4 + j) ^ sw) << 4)) = lo;
                } else if (EPI == EPI_GELU_BF16 && ar.f16) {
                  hi = make_uint4(pack_f16x2(v[8 * j], v[8 * j + 1]), pack_f16x2(v[8 * j + 2], v[8 * j + 3]),
                                  pack_f16x2(v[8 * j + 4], v[8 * j + 5]), pack_f16x2(v[8 * j + 6], v[8 * j + 7]));
                } else {
                  hi = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                  pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                }
                *reinterpret_cast<uint4*>(row_hi + (((hh * 4 + j) ^ sw) << 4)) = hi;
              }
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          // rows >= M and columns >= N are clipped by the tensor map; a half tile that lies entirely past M stores nothing
          // (in a partial-plane buffer its rows would belong to the NEXT plane when m_pad is an odd multiple of 128)
          if (leader && m0 < ar.M) {
            tma_store_2d(&tmap_out, tile_hi, col0, orow);
            if (EPI == EPI_GELU_SPLIT) tma_store_2d(&tmap_out, tile_hi + C::TILE_BYTES, ar.N + col0, orow);   // lo half
            bulk_commit();
          }
        }
        if (leader) {
          bulk_wait_read<0>();                          // the stores have read their tiles: the pipeline memory is free again
          mbar_arrive(&epi_free[0]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();                                   // the peer's smem / TMEM stay alive until the leader's MMAs retired
  if (warp == 1) tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
template <bool X3, int EPI, int BN>
static int launch_dec(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const DecArgs& ar, cudaStream_t stream) {
  using C = DecCfg<X3, BN>;
  auto kern = gemm_dec_kernel<X3, EPI, BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("gemm_dec: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VC_ERR_LAUNCH; }
    configured = true;
  }
  const int tiles = ((ar.M + 2 * C::BM - 1) / (2 * C::BM)) * ((ar.N + BN - 1) / BN) * ar.splits;
  int clusters = sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  launch_pdl(kern, dim3(2 * clusters), dim3(C::THREADS), C::SMEM_BYTES, stream, ta, tb, to, ar);
  return check_launch("gemm_dec");
}

constexpr int VOCAB_BN = 208;

// operand format `fmt`: 0 = bf16, 1 = split bf16 (three products, K = 3 Kt), 2 = IEEE half (one product)
static int dec_check(const char* what, const void* A, int lda, const void* W, int ldw, int M, int N, int K, int fmt) {
  if (fmt < 0 || fmt > 2) { set_last_error("%s: operand format %d (0 bf16, 1 split bf16, 2 half)", what, fmt); return VC_ERR_BAD_ARG; }
  const int unit = fmt == 1 ? 192 : 64;
  if (M <= 0 || N <= 0 || K <= 0 || (K % unit) != 0) {
    set_last_error("%s: need K %% %d == 0 (M=%d N=%d K=%d)", what, unit, M, N, K);
    return VC_ERR_BAD_ARG;
  }
  if ((lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) {
    set_last_error("%s: pointers must be 16-byte aligned and row pitches multiples of 16 bytes", what);
    return VC_ERR_BAD_ARG;
  }
  return VC_OK;
}

// mode 0: out = fp32 partial planes [splits, m_pad, N] of A W^T, plane s = the s-th slice of K; no bias -- except that a single
//         plane (splits == 1) takes one: out = fp32 A W^T + bias with any N (the materialised vocabulary logits)
// mode 1: out = bf16 (A W^T + bias);  mode 2: out = GELU(A W^T + bias) as bf16 (as halves with fmt 2: the next GEMM's operand)
// mode 3: out = bf16 [M, >= 2N]: columns [0, N) = hi, [N, 2N) = lo of the split of GELU(A W^T + bias)
int gemm_dec(int mode, int fmt, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int M, int N,
             int K, int splits, int m_pad, cudaStream_t stream) {
  int rc = dec_check("gemm_dec", A, lda, W, ldw, M, N, K, fmt);
  if (rc) return rc;
  const int x3 = fmt == 1;
  const int num_kb = (x3 ? K / 3 : K) / 64;
  if (mode < 0 || mode > 3 || out == nullptr || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15)) {
    set_last_error("gemm_dec: bad mode / output"); return VC_ERR_BAD_ARG;
  }
  if (mode == EPI_PARTIAL) {
    if (splits < 1 || num_kb % splits || m_pad < M || (m_pad % 128) || (ldo % 4) || ((N % 4) && splits > 1) || (bias && splits > 1)) {
      set_last_error("gemm_dec: partial planes need splits | k-blocks (%d), m_pad %% 128 == 0 >= M, N %% 4 == 0, a bias only with "
                     "splits == 1", num_kb);
      return VC_ERR_BAD_ARG;
    }
  } else {
    if (splits != 1 || (N % 64) || (ldo % 8) || (mode == EPI_GELU_SPLIT && ldo < 2 * N)) {
      set_last_error("gemm_dec: bf16 outputs need splits == 1, N %% 64 == 0 (and ldo >= 2N for the split pair)");
      return VC_ERR_BAD_ARG;
    }
  }
  // 128-wide tiles when the 256-wide ones would fill less than half of the clusters (q|k|v at 2 x 512 rows: 36 -> 72 tiles).
  // The tile width never changes the summation order of an output element, so results do not depend on this choice.
  const int tiles256 = ((M + 255) / 256) * ((N + 255) / 256) * splits;
  const bool narrow = (mode == EPI_BF16 || mode == EPI_GELU_BF16) && !x3 && 2 * tiles256 <= sm_count() / 2 && (N % 128) == 0;
  CUtensorMap ta, tb, to;
  rc = get_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, narrow ? 64 : 128, 64);
  if (rc) return rc;
  // (a single plane is clipped at M rows: the logits buffer has no padding rows)
  if (mode == EPI_PARTIAL) rc = get_tmap_2d_f32(&to, out, splits == 1 ? (uint64_t)M : (uint64_t)splits * m_pad, (uint64_t)N, (uint64_t)ldo, 128, 32);
  else rc = get_tmap_2d_bf16(&to, out, (uint64_t)M, (uint64_t)(mode == EPI_GELU_SPLIT ? 2 * N : N), (uint64_t)ldo, 128, 64);
  if (rc) return rc;
  DecArgs ar = {mode == EPI_PARTIAL && splits > 1 ? nullptr : bias, nullptr, 0, M, N, K, splits, m_pad, fmt == 2};
  if (x3) {
    switch (mode) {
      case EPI_PARTIAL: return launch_dec<true, EPI_PARTIAL, 256>(ta, tb, to, ar, stream);
      case EPI_BF16: return launch_dec<true, EPI_BF16, 256>(ta, tb, to, ar, stream);
      case EPI_GELU_BF16: return launch_dec<true, EPI_GELU_BF16, 256>(ta, tb, to, ar, stream);
      default: return launch_dec<true, EPI_GELU_SPLIT, 256>(ta, tb, to, ar, stream);
    }
  }
  switch (mode) {
    case EPI_PARTIAL: return launch_dec<false, EPI_PARTIAL, 256>(ta, tb, to, ar, stream);
    case EPI_BF16: return narrow ? launch_dec<false, EPI_BF16, 128>(ta, tb, to, ar, stream) : launch_dec<false, EPI_BF16, 256>(ta, tb, to, ar, stream);
    case EPI_GELU_BF16: return narrow ? launch_dec<false, EPI_GELU_BF16, 128>(ta, tb, to, ar, stream)
                                      : launch_dec<false, EPI_GELU_BF16, 256>(ta, tb, to, ar, stream);
    default: return launch_dec<false, EPI_GELU_SPLIT, 256>(ta, tb, to, ar, stream);
  }
}

// part [M, n_part] float4, n_part = 2 * ceil(N / 208): per (row, 208-column tile, epilogue group) the maximum of
// A W^T + bias over the group's columns, its first arg max, and the sum of exp(x - max)
int gemm_dec_argmax(int fmt, const void* A, int lda, const void* W, int ldw, const float* bias, void* part, int n_part, int M, int N,
                    int K, cudaStream_t stream) {
  int rc = dec_check("gemm_dec_argmax", A, lda, W, ldw, M, N, K, fmt);
  if (rc) return rc;
  const int x3 = fmt == 1;
  if (part == nullptr || (reinterpret_cast<uintptr_t>(part) & 15) || n_part != 2 * ((N + VOCAB_BN - 1) / VOCAB_BN)) {
    set_last_error("gemm_dec_argmax: part must be 16-byte aligned with n_part = 2 * ceil(N / 208)"); return VC_ERR_BAD_ARG;
  }
  CUtensorMap ta, tb;
  rc = get_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, VOCAB_BN / 2, 64);
  if (rc) return rc;
  DecArgs ar = {bias, static_cast<float4*>(part), n_part, M, N, K, 1, 0, fmt == 2};
  if (x3) return launch_dec<true, EPI_ARGMAX, VOCAB_BN>(ta, tb, ta, ar, stream);
  return launch_dec<false, EPI_ARGMAX, VOCAB_BN>(ta, tb, ta, ar, stream);
}

}  // namespace vc
