// pbh_tma.cuh — hand-written PTX wrappers for the Tensor Memory Accelerator (cp.async.bulk.tensor) and mbarrier,
// sm_90+/sm_100a.  Used to move [planes x 256-item] byte tiles between HBM and shared memory: one elected thread issues
// the bulk copies, completion is signalled on an mbarrier (loads) or tracked by bulk async-groups (stores).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace pbh {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the barrier initialisation (generic proxy) visible to the async proxy that will complete transactions on it
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// global -> shared tile load; (c0, c1) = (item index, plane index) of the tile's origin
__device__ __forceinline__ void load_2d(void* dst_smem, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// shared -> global tile store (out-of-range items of the last tile are clipped by the hardware)
__device__ __forceinline__ void store_2d(const CUtensorMap* map, const void* src_smem, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src_smem)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until every committed store group has finished READING shared memory (the buffer may be overwritten)
__device__ __forceinline__ void store_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tma
}  // namespace pbh
