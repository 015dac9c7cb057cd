// smem_tma.cuh — shared-memory / bulk-copy (TMA) helpers of the one-block-per-store kernels (pir_batch.cu, pir_search.cu).
#pragma once
#include "pir_device.cuh"

namespace lpc {

// ---- PTX helpers: mbarrier + bulk async copies (TMA, 1-D) --------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
    "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while(!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Issue a large global->shared copy as <= 32 KB bulk pieces (sizes are multiples of 16 by construction).
__device__ __forceinline__ void bulk_g2s_chunked(char* dst, const char* src, unsigned bytes, unsigned long long* bar) {
  for(unsigned o = 0; o < bytes; o += 32768u) bulk_g2s(dst + o, src + o, min(32768u, bytes - o), bar);
}

// Shared-state-space accessors with 32-bit addresses: the ring slot is selected at run time, so through generic
// pointers the compiler falls back to generic LD / ATOM and 64-bit address arithmetic.
__device__ __forceinline__ int2 lds_itv(unsigned addr) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_s32(unsigned addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_u8(unsigned addr) {
  int v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void reds_max(unsigned addr, int v) { asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_min(unsigned addr, int v) { asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Join into the shared-memory store (only called when something tightened or an operand was empty).
__device__ __forceinline__ int commit_smem(unsigned addr, int2 old, const Itv& nw) {
  int f = 0;
  if(nw.lb > old.x) { reds_max(addr, nw.lb); f = 1; }
  if(nw.ub < old.y) { reds_min(addr + 4, nw.ub); f = 1; }
  if(f && nw.lb > nw.ub) f |= 2;
  return f;
}


} // namespace lpc
