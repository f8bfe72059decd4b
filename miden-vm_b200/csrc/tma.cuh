// Bulk asynchronous copies global -> shared (the TMA engine's 1-D path, `cp.async.bulk`, SASS UBLKCP) completed on an
// mbarrier: one thread arms the barrier with the byte count and issues the copy, every thread of the block waits on
// the barrier's phase.  Used for the tables and contiguous chunks the NTT passes stage in shared memory (ntt2.cuh): the
// copy engine moves them while the threads load the strided / padded part of the tile with ordinary loads.
// Source and destination must be 16-byte aligned and the size a multiple of 16.
// Device only; the host build of ntt2.cuh (tests/cpp) and the CPU kernel emulator take the plain-loop path.
#pragma once
#include <cstdint>
#if defined(__CUDA_ARCH__)
namespace tma {
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // make the init visible to the async proxy
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(saddr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t phase) {
    asm volatile("{\n\t.reg .pred p;\n\tTMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra TMA_DONE;\n\tbra TMA_WAIT;\n\tTMA_DONE:\n\t}"
                 ::"r"(saddr(bar)), "r"(phase) : "memory");
}
}  // namespace tma
#endif
