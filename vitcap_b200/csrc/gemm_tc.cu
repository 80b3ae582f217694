// bf16 GEMM on the 5th-gen tensor cores:  out[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ resid)
//
//   * A (activations) and W (nn.Linear weight, [out_features, in_features]) are both K-major, so both operands are
//     staged by TMA into 128B-swizzled K-major shared-memory tiles and fed to tcgen05.mma.kind::f16 through
//     shared-memory descriptors; the fp32 accumulator lives in TMEM (two stages: epilogue(i) overlaps mainloop(i+1)).
//   * persistent, warp-specialised CTA (one per SM): warp0 = TMA producer, warp1 = MMA issuer (one elected thread),
//     warps 2-5 = epilogue.
//   * epilogue: tcgen05.ld (one accumulator row per thread) -> bias / activation / fp32 residual -> 128B-swizzled smem
//     staging tile -> TMA store. The residual tile is itself prefetched by TMA into the staging buffer two chunks
//     ahead, so every global access of the kernel is a full-line bulk transfer (round-1 finding: row-per-thread
//     LDG/STG in the epilogue kept L1 at 60-76 % and capped proj+residual at 19 % tensor-pipe activity).
//   * replaces the cuBLAS calls behind nn.Linear in vision_transformer.py:152-210 (qkv/proj/fc1/fc2) and
//     modeling_bert.py:303-419, 524-563 (q/k/v/dense/intermediate/output/pooler/heads).
#include "common.cuh"

namespace vc {

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_TANH = 2 };

// X3: split-bf16 operands A = [a_hi | a_lo | a_hi], W = [w_hi | w_hi | w_lo] (K = 3 Kt). Instead of walking the concatenated K, a
// stage holds the four distinct tiles (a_hi, w_hi, a_lo, w_lo) of one Kt block and the three products are issued from them:
// 2/3 of the bytes each SM pulls from L2 (the long-K decode GEMMs are bound by exactly that).
template <int BN, bool OUT_F32, bool RESID, bool X3 = false> struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;                         // 64 bf16 = one 128-byte swizzle row
  static constexpr int A_BYTES = BM * BK * 2;           // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int HALF_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = (X3 ? 2 : 1) * HALF_BYTES;
  // epilogue staging tiles (128 rows x 128 B); the residual tile of chunk g + NBUF - 2 is prefetched by TMA while chunk g is
  // processed (8 tiles with a 128-wide N tile were measured slower for the K = 768 residual GEMM: 0.51 vs 0.43 ms)
  static constexpr int NBUF = RESID ? 4 : 2;
  static constexpr int EPI_BYTES = 16384;
  static constexpr int BUDGET = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - NBUF * EPI_BYTES;
  static constexpr int STAGES_MAX = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_MAX > 8 ? 8 : STAGES_MAX;
  static constexpr int TMEM_COLS = 2 * BN;              // two accumulator stages
  static constexpr int CW = OUT_F32 ? 32 : 64;          // output columns per staging tile (128 B per row)
  static constexpr int NCH = BN / CW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NBUF * EPI_BYTES + 1024 + 512;
};

template <int BN, int ACT, bool OUT_F32, bool RESID, bool X3 = false>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
               const float* __restrict__ bias, int M, int N, int K) {
  using C = GemmCfg<BN, OUT_F32, RESID, X3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + C::NBUF * C::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
  uint64_t* res_full = bars + 2 * C::STAGES + 4;      // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4 + C::NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + C::BM - 1) / C::BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int Kt = X3 ? K / 3 : K;            // X3: columns [0,Kt) = hi, [Kt,2Kt) = a_lo / w_hi, [2Kt,3Kt) = a_hi / w_lo
  const int num_kb = Kt / C::BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (RESID) tma_prefetch_desc(&tmap_res);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);   // one arrive per epilogue warp
    }
    for (int i = 0; i < C::NBUF; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();      // the next kernel's prologue may start; it blocks in its own pdl_wait() until this grid is done

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // W is a weight matrix: never written by the preceding kernel, so its tiles of this CTA's first output tile are
      // requested BEFORE the dependency wait (the first STAGES k-blocks into their pipeline stages, the rest into L2) and
      // travel while the predecessor is still running; the activations follow after the wait.
      int pre = 0;
      if ((int)blockIdx.x < num_tiles) {
        const int n0 = ((int)blockIdx.x % n_tiles) * BN;
        pre = num_kb < C::STAGES ? num_kb : C::STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_arrive_expect_tx(&full_bar[kb], C::STAGE_BYTES);
          tma_load_2d(smem + kb * C::STAGE_BYTES + C::A_BYTES, &tmap_b, &full_bar[kb], kb * C::BK, n0);
          if (X3) tma_load_2d(smem + kb * C::STAGE_BYTES + C::HALF_BYTES + C::A_BYTES, &tmap_b, &full_bar[kb], 2 * Kt + kb * C::BK, n0);
        }
        for (int kb = pre; kb < num_kb; ++kb) {
          tma_prefetch_l2_2d(&tmap_b, kb * C::BK, n0);
          if (X3) tma_prefetch_l2_2d(&tmap_b, 2 * Kt + kb * C::BK, n0);
        }
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * C::BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if (pre > 0) {                      // stage armed and its W half already in flight (fresh barriers: nothing to wait for)
            --pre;
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * C::BK, m0);
            if (X3) tma_load_2d(sa + C::HALF_BYTES, &tmap_a, &full_bar[stage], Kt + kb * C::BK, m0);
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * C::BK, m0);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * C::BK, n0);
            if (X3) {
              tma_load_2d(sa + C::HALF_BYTES, &tmap_a, &full_bar[stage], Kt + kb * C::BK, m0);
              tma_load_2d(sb + C::HALF_BYTES, &tmap_b, &full_bar[stage], 2 * Kt + kb * C::BK, n0);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      pdl_wait();
      constexpr uint32_t idesc = make_idesc_bf16(C::BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < C::BK / 16; ++k) {
            // advance 16 bf16 (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if (X3) {
            const uint64_t adesc_lo = make_smem_desc_sw128(sa + C::HALF_BYTES, 16, 1024);
            const uint64_t bdesc_lo = make_smem_desc_sw128(sb + C::HALF_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < C::BK / 16; ++k) umma_f16(d_tmem, adesc_lo + 2 * k, bdesc + 2 * k, idesc, true);     // a_lo w_hi
#pragma unroll
            for (int k = 0; k < C::BK / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc_lo + 2 * k, idesc, true);     // a_hi w_lo
          }
          umma_commit(&empty_bar[stage]);     // frees the smem stage once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);          // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5, 128 threads; thread <-> accumulator row) =====================
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int t = quad * 32 + lane;           // row inside the tile
    const bool leader = (warp == 2 && lane == 0);
    const uint32_t sw = (uint32_t)(t & 7);
    int as = 0;
    uint32_t aphase = 0;
    uint32_t g = 0;                           // running staging-chunk counter (buffer = g % NBUF)
    // residual prefetch cursor (leader only): chunk index pf_g, its tile and chunk-in-tile
    uint32_t pf_g = 0;
    int pf_tile = blockIdx.x, pf_c = 0;
    pdl_wait();                               // programmatic dependent launch: global memory (residual, bias, output) from here on
    auto prefetch_resid = [&]() {
      if (pf_tile >= num_tiles) return;
      const int pm0 = (pf_tile / n_tiles) * C::BM;
      const int pn0 = (pf_tile % n_tiles) * BN + pf_c * C::CW;
      const int b = pf_g % C::NBUF;
      mbar_arrive_expect_tx(&res_full[b], C::EPI_BYTES);
      tma_load_2d(epi + b * C::EPI_BYTES, &tmap_res, &res_full[b], pn0, pm0);
      ++pf_g;
      ++pf_c;
      if (pf_c == C::NCH || (pf_tile % n_tiles) * BN + pf_c * C::CW >= N) { pf_c = 0; pf_tile += gridDim.x; }
    };
    if (RESID && leader) {
      for (int i = 0; i < C::NBUF - 2; ++i) prefetch_resid();
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * C::BM;
      const int n0 = (tile % n_tiles) * BN;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < C::NCH; ++c) {
        const int col0 = n0 + c * C::CW;
        if (col0 >= N) break;
        const int b = g % C::NBUF;
        uint8_t* ebuf = epi + b * C::EPI_BYTES;
        uint8_t* myrow = ebuf + t * 128;
        // (1) the staging tile written two chunks ago must have been read by its TMA store
        if (leader) {
          bulk_wait_read<1>();
          if (RESID) prefetch_resid();          // chunk g + NBUF - 2 goes into the tile that store just released
        }
        named_bar_sync(1, 128);
        // (2) accumulator row -> registers
        const bool full = (col0 + C::CW <= N);
        if (OUT_F32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (bias != nullptr) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (col0 + j < N) ? __ldg(bias + col0 + j) : 0.f;
            }
          }
          if (ACT == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_tanh(v[j]);
          } else if (ACT == ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
          }
          if (RESID) {
            mbar_wait(&res_full[b], (g / C::NBUF) & 1);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 rr = *reinterpret_cast<const float4*>(myrow + ((j ^ sw) << 4));
              v[4 * j] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(myrow + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + c * 64 + hh * 32, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            const int cb = col0 + hh * 32;
            if (bias != nullptr) {
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 bb = __ldg(reinterpret_cast<const float4*>(bias + cb + j));
                  v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += (cb + j < N) ? __ldg(bias + cb + j) : 0.f;
              }
            }
            if (ACT == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf_tanh(v[j]);
            } else if (ACT == ACT_TANH) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                    pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
              *reinterpret_cast<uint4*>(myrow + (((hh * 4 + j) ^ sw) << 4)) = pk;
            }
          }
        }
        // (3) staging tile -> global (rows >= M and columns >= N are clipped by the tensor map)
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (leader) {
          tma_store_2d(&tmap_out, ebuf, col0, m0);
          bulk_commit();
        }
        ++g;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (leader) bulk_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
template <int BN, int ACT, bool OUT_F32, bool RESID>
static int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                      const float* bias, int M, int N, int K, cudaStream_t stream) {
  using C = GemmCfg<BN, OUT_F32, RESID>;
  static_assert(C::STAGES >= 3, "not enough shared memory for a 3-stage pipeline");
  auto kern = gemm_tc_kernel<BN, ACT, OUT_F32, RESID>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VC_ERR_LAUNCH; }
    configured = true;
  }
  const int tiles = ((M + C::BM - 1) / C::BM) * ((N + BN - 1) / BN);
  int grid = tiles < sm_count() ? tiles : sm_count();
  launch_pdl(kern, dim3(grid), dim3(192), C::SMEM_BYTES, stream, ta, tb, to, tr, bias, M, N, K);
  return check_launch("gemm_tc");
}

// split-bf16 operands, the four distinct tiles loaded once per k-block (fc2 of a decode step: fp32 output + residual)
int gemm_bf16_tc_x3(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo, const float* resid,
                    int ldr, int M, int N, int K3, cudaStream_t stream) {
  constexpr int BN = 64;
  using C = GemmCfg<BN, true, true, true>;
  static_assert(C::STAGES >= 3, "not enough shared memory for a 3-stage pipeline");
  if (M <= 0 || N <= 0 || K3 <= 0 || (K3 % 192) != 0 || resid == nullptr) {
    set_last_error("gemm_tc_x3: need K3 %% 192 == 0 and a residual (K3=%d)", K3);
    return VC_ERR_BAD_ARG;
  }
  if ((lda % 8) || (ldw % 8) || (ldo % 4) || (ldr % 4) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(resid) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15)) {
    set_last_error("gemm_tc_x3: pointers must be 16-byte aligned and row pitches multiples of 16 bytes");
    return VC_ERR_BAD_ARG;
  }
  CUtensorMap ta, tb, to, tr;
  int rc = get_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K3, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K3, (uint64_t)ldw, (uint32_t)BN, 64);
  if (rc) return rc;
  rc = get_tmap_2d_f32(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 128, 32);
  if (rc) return rc;
  rc = get_tmap_2d_f32(&tr, resid, (uint64_t)M, (uint64_t)N, (uint64_t)ldr, 128, 32);
  if (rc) return rc;
  auto kern = gemm_tc_kernel<BN, ACT_NONE, true, true, true>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("gemm_tc_x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VC_ERR_LAUNCH; }
    configured = true;
  }
  const int tiles = ((M + C::BM - 1) / C::BM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  launch_pdl(kern, dim3(grid), dim3(192), C::SMEM_BYTES, stream, ta, tb, to, tr, bias, M, N, K3);
  return check_launch("gemm_tc_x3");
}

template <int BN, int ACT>
static int launch_act(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                      const float* bias, int out_f32, bool resid, int M, int N, int K, cudaStream_t s) {
  if (out_f32) {
    if (resid) return launch_one<BN, ACT, true, true>(ta, tb, to, tr, bias, M, N, K, s);
    return launch_one<BN, ACT, true, false>(ta, tb, to, tr, bias, M, N, K, s);
  }
  return launch_one<BN, ACT, false, false>(ta, tb, to, tr, bias, M, N, K, s);
}

template <int BN>
static int launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                     const float* bias, int out_f32, int act, bool resid, int M, int N, int K, cudaStream_t s) {
  switch (act) {
    case ACT_NONE: return launch_act<BN, ACT_NONE>(ta, tb, to, tr, bias, out_f32, resid, M, N, K, s);
    case ACT_GELU: return launch_act<BN, ACT_GELU>(ta, tb, to, tr, bias, out_f32, resid, M, N, K, s);
    case ACT_TANH: return launch_act<BN, ACT_TANH>(ta, tb, to, tr, bias, out_f32, resid, M, N, K, s);
  }
  set_last_error("gemm_tc: unknown activation %d", act);
  return VC_ERR_BAD_ARG;
}

int gemm_bf16_tc2(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32, int act,
                  const float* resid, int ldr, int M, int N, int K, cudaStream_t stream);

int gemm_bf16_tc(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int out_f32,
                 int act, const float* resid, int ldr, int M, int N, int K, int force_bn, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 64) != 0) { set_last_error("gemm_tc: need K %% 64 == 0 (K=%d)", K); return VC_ERR_BAD_ARG; }
  if (resid && !out_f32) { set_last_error("gemm_tc: a residual needs fp32 output"); return VC_ERR_BAD_ARG; }
  if ((lda % 8) || (ldw % 8) || (ldo % (out_f32 ? 4 : 8)) || (resid && (ldr % 4)) ||
      (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(resid) & 15) ||
      (reinterpret_cast<uintptr_t>(bias) & 15)) {
    set_last_error("gemm_tc: pointers must be 16-byte aligned and row pitches multiples of 16 bytes");
    return VC_ERR_BAD_ARG;
  }
  // large problems run on CTA pairs (gemm_tc2.cu): 256 x 256 tiles, at least two rounds of tiles per cluster
  {
    const long pair_tiles = (long)((M + 255) / 256) * ((N + 255) / 256);
    // (the K <= 768 residual GEMM is bound by its fp32 residual traffic, where the single-CTA kernel measured 3 % faster)
    const bool hbm_bound_resid = resid != nullptr && K <= 768;
    if (force_bn == 512 || (force_bn == 0 && N >= 256 && pair_tiles >= 2 * (sm_count() / 2) && !hbm_bound_resid))
      return gemm_bf16_tc2(A, lda, W, ldw, bias, out, ldo, out_f32, act, resid, ldr, M, N, K, stream);
  }
  // tile width: the widest N tile that still gives every SM work
  int bn = force_bn;
  if (bn == 0) {
    const int mt = (M + 127) / 128;
    bn = 256;
    while (bn > 64 && mt * ((N + bn - 1) / bn) < sm_count()) bn >>= 1;
  }
  CUtensorMap ta, tb, to, tr;
  int rc = get_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = get_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)bn, 64);
  if (rc) return rc;
  if (out_f32) rc = get_tmap_2d_f32(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 128, 32);
  else rc = get_tmap_2d_bf16(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 128, 64);
  if (rc) return rc;
  tr = to;
  if (resid) {
    rc = get_tmap_2d_f32(&tr, resid, (uint64_t)M, (uint64_t)N, (uint64_t)ldr, 128, 32);
    if (rc) return rc;
  }
  switch (bn) {
    case 256: return launch_bn<256>(ta, tb, to, tr, bias, out_f32, act, resid != nullptr, M, N, K, stream);
    case 128: return launch_bn<128>(ta, tb, to, tr, bias, out_f32, act, resid != nullptr, M, N, K, stream);
    case 64: return launch_bn<64>(ta, tb, to, tr, bias, out_f32, act, resid != nullptr, M, N, K, stream);
  }
  set_last_error("gemm_tc: unsupported tile width %d", bn);
  return VC_ERR_BAD_ARG;
}

}  // namespace vc
