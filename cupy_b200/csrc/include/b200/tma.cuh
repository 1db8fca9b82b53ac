// b200/tma.cuh -- device-side wrappers of the sm_100a bulk-asynchronous copy
// engine (TMA) and its mbarrier completion mechanism.  Used by the pipelined scan
// (scan_tma.cuh) and by the transposing elementwise tiler: tiles travel
// HBM -> shared memory -> HBM without occupying LSU slots or registers, so the
// bytes in flight per SM are set by the stage count, not by occupancy.
#pragma once
#include "base.cuh"

namespace b200 {

B200_DEVICE uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

B200_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make freshly initialised barriers visible to the async proxy
B200_DEVICE void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
B200_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
B200_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
B200_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
B200_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// the same primitives on raw shared-window addresses (tilers that carve dynamic shared memory)
B200_DEVICE void mbar_init_a(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
B200_DEVICE void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
B200_DEVICE void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
B200_DEVICE void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
B200_DEVICE uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// generic-proxy writes to shared memory -> visible to the async proxy (before a TMA store)
B200_DEVICE void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 2-D tiled load: box at (c0 = inner coordinate, c1 = outer coordinate) -> smem, completes on `bar`
B200_DEVICE void tma_load_2d(void* smem_dst, const void* tmap, int32_t c0, int32_t c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
B200_DEVICE void tma_load_3d(void* smem_dst, const void* tmap, int32_t c0, int32_t c1, int32_t c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
B200_DEVICE void tma_load_5d(uint32_t smem_dst, const void* tmap, int32_t c0, int32_t c1, int32_t c2, int32_t c3,
                            int32_t c4, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
// 2-D tiled store smem -> global (bulk-group completion)
B200_DEVICE void tma_store_2d(const void* tmap, int32_t c0, int32_t c1, const void* smem_src) {
    asm volatile(
        "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
        ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}
B200_DEVICE void tma_store_3d(const void* tmap, int32_t c0, int32_t c1, int32_t c2, const void* smem_src) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
        ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(smem_src)) : "memory");
}
B200_DEVICE void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N committed store groups still READ their shared-memory source
template <int N>
B200_DEVICE void tma_wait_group_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
B200_DEVICE void tma_wait_group() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
B200_DEVICE void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// Byte offset of 16-byte chunk `c` of 128-byte row `r` inside a tile written by TMA
// with CU_TENSOR_MAP_SWIZZLE_128B (tile base 1024-byte aligned): the chunk index is
// XORed with the low three bits of the row index, which is what makes "one row per
// thread" shared-memory access conflict-free.
B200_DEVICE uint32_t swz128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

}  // namespace b200
