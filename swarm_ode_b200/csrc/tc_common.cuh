// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarriers, TMA / bulk copies, UMMA descriptors and issue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnode {
namespace tc {

// Barrier waits are bounded in TIME (%globaltimer), not in polls: a slow clock, a profiler replay or time slicing must
// not turn a healthy wait into a spurious timeout.  A wait that does expire flags the status word (surfaced by
// gnode_tc_status and by the deferred checks of the Python layer) and the kernel winds down.
constexpr uint64_t WAIT_LIMIT_NS = 4000000000ull;   // 4 s
constexpr uint32_t FAST_POLLS = 2048;               // polls before the clock is consulted at all
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One lane of a converged warp (elect.sync).  The tcgen05 / bulk-copy issue sites use this instead of `lane == 0`: behind
// a lane test the compiler treats the operands as divergent and wraps EVERY tcgen05.mma in an ELECT + 6 x R2UR + branch
// waterfall (~180 cycles per instruction, measured: scripts/dev/probe_umma_rate.cu); behind elect.sync with warp-uniform
// operands the descriptors live in uniform registers and the instruction issues directly.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: returns false (and flags the status word) on timeout
__device__ __forceinline__ bool mbar_wait(uint32_t addr, uint32_t parity, int* status, int code) {
  for (uint32_t i = 0; i < FAST_POLLS; ++i)
    if (mbar_try_wait(addr, parity)) return true;
  const uint64_t t0 = globaltimer_ns();
  do {
    for (uint32_t i = 0; i < 256; ++i)
      if (mbar_try_wait(addr, parity)) return true;
  } while (globaltimer_ns() - t0 < WAIT_LIMIT_NS);
  if (status) atomicExch(status, code);
  return false;
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// UMMA shared-memory descriptor, K-major, no swizzle: 8-row x 16-byte core matrices, rows contiguous
// (SBO = 128 B between 8-row groups), LBO bytes between the two K-chunks of one K=8 step.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc(int bn, int bm = 128) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(bm >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
      : "memory");
}


}  // namespace tc
}  // namespace gnode
