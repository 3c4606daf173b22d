// vis_fused_common.cuh — PTX helpers and small device utilities shared by the fused kernels
// (vis_fused_ws.cu: general warp-specialised persistent CTA; vis_fused_sched*.cu: statically scheduled kernels).
#pragma once
#include <climits>

#include "vis_internal.h"

namespace visf {

__host__ __device__ inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// ---- PTX helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
#ifndef VIS_SPIN_NS
#define VIS_SPIN_NS 32          // back-off between polls of a phase that is not complete yet
#endif
#ifndef VIS_WAIT_HINT_NS
#define VIS_WAIT_HINT_NS 0      // > 0: pass a suspend-time hint to try_wait in the slow path (the warp sleeps in hardware)
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// slow path of a wait, kept out of line so the hot loops stay small (the kernels are instruction-cache sensitive):
// back off between polls, and trap instead of hanging the GPU if the phase never completes (lost bulk copy)
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
#if VIS_WAIT_HINT_NS > 0
    for (unsigned spins = 0; !mbar_try_wait_hint(bar, parity, VIS_WAIT_HINT_NS); ++spins)
        if (spins > (1u << 24)) __trap();
#else
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 26)) __trap();
        __nanosleep(VIS_SPIN_NS);
    }
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void stg128(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// clip8 of Pillow: arithmetic >> 22, clamp to 0..255 (one VIMNMX.RELU)
__device__ __forceinline__ int clip8i(int acc) { return __vimin_s32_relu(acc >> VIS_PRECISION_BITS, 255); }

template <int KT>
struct Rec {                        // one coefficient record held in registers (warp-uniform values)
    int k[KT];
    int last;
};

template <int KT, int STRIDE>
__device__ __forceinline__ void load_rec(Rec<KT>& r, const int* p) {      // p: shared or global, 16-byte aligned
    int tmp[STRIDE];
#pragma unroll
    for (int q = 0; q < STRIDE / 4; ++q) {
        const int4 v = *reinterpret_cast<const int4*>(p + 4 * q);
        tmp[4 * q] = v.x; tmp[4 * q + 1] = v.y; tmp[4 * q + 2] = v.z; tmp[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int t = 0; t < KT; ++t) r.k[t] = tmp[t];
    r.last = tmp[STRIDE - 1];
}

// same, from a shared-memory address (explicit ld.shared: never a generic load)
template <int KT, int STRIDE>
__device__ __forceinline__ void load_rec_s(Rec<KT>& r, uint32_t addr) {
    int tmp[STRIDE];
#pragma unroll
    for (int q = 0; q < STRIDE / 4; ++q) {
        const uint4 v = lds128(addr + 16 * q);
        tmp[4 * q] = (int)v.x; tmp[4 * q + 1] = (int)v.y; tmp[4 * q + 2] = (int)v.z; tmp[4 * q + 3] = (int)v.w;
    }
#pragma unroll
    for (int t = 0; t < KT; ++t) r.k[t] = tmp[t];
    r.last = tmp[STRIDE - 1];
}

}  // namespace visf
