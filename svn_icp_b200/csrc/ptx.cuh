// ptx.cuh -- inline-PTX helpers shared by the kernels: mbarrier + 1-D bulk TMA, system-scope flags for the peer exchange.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace svn {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// the same wait for a thread that has nothing else to do (TMA producer warps): back off between polls so that the spin does
// not take issue slots from the consumer warps of the same SM sub-partition (measured in k_gn: 7 % of all issued instructions)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT_B:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_B;\n"
      "nanosleep.u32 256;\n"
      "bra LAB_WAIT_B;\n"
      "DONE_B:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared (TMA, no tensor map); completion is signalled on the mbarrier as transferred bytes
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Whole CTA: wait until every other rank has published sequence number `seq` (or later) for `kind` in THIS rank's flag block.
// One thread per source rank polls; the trailing __syncthreads orders every thread's later loads after the acquire.
// A wait that exceeds the timeout sets ctrl->error and falls through (the scan's results are then reported as an error by the
// host) -- a lost peer must never hang the GPU.
__device__ __forceinline__ void peer_wait(const PeerTable &pt, int kind, unsigned seq, Ctrl *c) {
  if (pt.n_ranks > 1 && (int)threadIdx.x < pt.n_ranks && (int)threadIdx.x != pt.rank) {
    const unsigned *f = pt.flag[pt.rank] + kind * MAX_RANKS + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    unsigned v;
    while ((int)((v = ld_acquire_sys(f)) - seq) < 0) {
      if (global_timer_ns() - t0 > pt.timeout_ns) { atomicExch(&c->error, 1); break; }
      __nanosleep(64);
    }
    // a peer is never more than one iteration ahead: a flag beyond seq + 1 means the ranks disagree on the sequence numbers
    // (e.g. they restarted with a scan) and the waits protect nothing -- caught by the debug build
    SVN_CHECK(c, (int)(v - seq) <= 1 || (int)(v - seq) < 0, 40);
  }
  __syncthreads();
}

}  // namespace svn
