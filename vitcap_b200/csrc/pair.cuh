// CTA-pair (tcgen05 cta_group::2) helpers shared by gemm_tc2.cu and gemm_dec.cu: a cluster of two CTAs (the two SMs of a TPC)
// owns one 256-row accumulator tile; the leader's elected thread issues the MMAs, every CTA stages its own operand rows and
// its TMA bytes are credited to the leader's barrier.
#pragma once
#include "common.cuh"

namespace vc {

constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address -> rank 0 of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into OWN shared memory whose completion bytes are credited to the barrier at the same offset in CTA rank 0
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
template <int kCols> __device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols> __device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in CTA rank 0 of the pair (local arrive when executed by rank 0).
// Default semantics (release at CTA scope, the form CUTLASS's ClusterBarrier::arrive uses): the barrier only tells the MMA thread
// that this warp's tcgen05.ld reads of the accumulator stage have completed (tcgen05.wait::ld + tcgen05.fence::before_thread_sync
// precede it); no generic-proxy data is published through it. Spelled .release.cluster, every epilogue warp paid
// MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR per tile (ncu: 25 % of the kernel's stall samples); this form is one SYNCS.ARRIVE.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n"
      ::"r"(smem_u32(bar)) : "memory");
}

// hi = bf16(x), lo = bf16(x - hi): the split operand of the three-product GEMMs (x ~ hi + lo to 2^-17)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  float ha, hb;
  unpack_bf16x2(hi, ha, hb);
  lo = pack_bf16x2(a - ha, b - hb);
}

}  // namespace vc
