// Bulk asynchronous copies (the TMA unit's non-tensor form, PTX `cp.async.bulk`; SASS
// UBLKCP) with mbarrier completion, for staging a CTA's tile of chain state between global
// and shared memory: in the component-major layout a row of the tile (kBlock consecutive
// chains of one component) is one contiguous, 16-byte-aligned run, so ONE elected thread
// moves the whole [2 Q][kBlock] tile with 2 Q copy instructions instead of every thread
// issuing 2 Q loads / stores.
#pragma once
#include <cstdint>

#include "mlb_common.cuh"

namespace mlb {

MLB_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

MLB_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
MLB_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
MLB_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, completion counted in bytes on `bar`; bytes % 16 == 0, both 16 B aligned
MLB_DEV void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_addr(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}
// shared -> global (bulk async-group of the issuing thread)
MLB_DEV void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_addr(src_smem)), "r"(bytes)
               : "memory");
}
MLB_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's groups have finished READING shared memory
MLB_DEV void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory made visible to the async proxy (before a s2g copy)
MLB_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// a [rows][kB] tile of a component-major array can go through the copy unit
MLB_DEV bool bulk_tile_ok(const void* base, int64_t es, int64_t cs) {
  return cs == 1 && (es & 1) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
}

}  // namespace mlb
