// 2-D box transfers between global memory and a dense shared-memory tile.
//
// Device: TMA (cp.async.bulk.tensor.2d) with an mbarrier for loads and bulk-group completion for stores;
// out-of-range box elements are zero-filled on load and dropped on store by the hardware.  Constraints measured
// on B200 (tools/probe/): the box origin's inner coordinate must be a multiple of 16 bytes for loads and stores
// (else "illegal instruction"); loads accept negative origins, stores do not (overshoot on the high side is
// fine for both).  Shared-memory tiles are 128-byte aligned, base/pitch of the tensor 16-byte aligned.  When a tensor
// cannot be described by a tensor map (base not 16-byte aligned, pitch not a multiple of 16 bytes) the
// same transfer is done with ordinary loads/stores by `box_load_generic` / `box_store_generic`, which have
// identical semantics; the host emulation (tests/host_emu) uses those too.
#pragma once
#include "pysp_common.cuh"

namespace pysp {

// box [box_h][box_w] at (y, x) of `v` -> dense smem tile; OOB -> 0.  Work is spread over the CTA.
PYSP_HD void box_load_generic(void* dst, const View2D& v, int x, int y, int box_w, int box_h) {
    const int n = box_w * box_h;
    PYSP_ITEMS(i, n) {
        int r = i / box_w, c = i - r * box_w;
        int gy = y + r, gx = x + c;
        bool in = gy >= 0 && gy < v.rows && gx >= 0 && gx < v.cols;
        const char* src = (const char*)v.base + (long long)gy * v.pitch + (long long)gx * v.elem;
        if (v.elem == 2) ((uint16_t*)dst)[i] = in ? pysp_ldg((const uint16_t*)src) : (uint16_t)0;
        else ((uint32_t*)dst)[i] = in ? pysp_ldg((const uint32_t*)src) : 0u;
    }
}

// dense smem tile [box_h][box_w] of 4-byte elements -> box at (y, x) of `v`, clipped to the tensor
PYSP_HD void box_store_generic(const void* src, const View2D& v, int x, int y, int box_w, int box_h) {
    const int n = box_w * box_h;
    PYSP_ITEMS(i, n) {
        int r = i / box_w, c = i - r * box_w;
        int gy = y + r, gx = x + c;
        if (gy >= 0 && gy < v.rows && gx >= 0 && gx < v.cols)
            *(uint32_t*)((char*)v.base + (long long)gy * v.pitch + (long long)gx * 4) = ((const uint32_t*)src)[i];
    }
}

#ifndef PYSP_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses to smem before this point are ordered before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* src, const CUtensorMap* map, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(x), "r"(y), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores have finished READING their shared-memory source
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
#endif

}  // namespace pysp
