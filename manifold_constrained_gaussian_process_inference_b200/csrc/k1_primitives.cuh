// Shared device primitives of the K1 kernels (banded_kernel.cuh: windowed, warp-specialised; flow_kernel.cuh: dataflow):
// quad reductions, mbarrier / TMA bulk-copy wrappers, named barriers and the compile-time dimension dispatch.
#pragma once
#include <type_traits>
#include "magi_common.cuh"

namespace magi {

__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// ---- TMA bulk copy + mbarrier helpers (fragment ring) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
#pragma unroll 1
    for (unsigned it = 0; it < (1u << 22); ++it) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// waits for two barriers at once (the two try_wait round trips overlap)
__device__ __forceinline__ void mbar_wait2(unsigned long long* bar_a, unsigned parity_a, unsigned long long* bar_b, unsigned parity_b) {
    unsigned da = 0, db = 0;
#pragma unroll 1
    for (unsigned it = 0; it < (1u << 22); ++it) {
        asm volatile("{\n.reg .pred p;\n.reg .pred r;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\nmbarrier.try_wait.parity.shared::cta.b64 r, [%4], %5;\n"
                     "selp.u32 %0, 1, 0, p;\nselp.u32 %1, 1, 0, r;\n}\n"
                     : "=r"(da), "=r"(db) : "r"(smem_u32(bar_a)), "r"(parity_a), "r"(smem_u32(bar_b)), "r"(parity_b) : "memory");
        if (da & db) return;
    }
    __trap();
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void named_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}

// calls f(std::integral_constant<int, d>) for the runtime (warp-uniform) dimension d: the model functors then see a
// compile-time dimension and compile to straight-line code
template <int D, class F> __device__ __forceinline__ void dispatch_dim(int d, F& f) {
    if constexpr (D >= 1) { if (d == 0) { f(std::integral_constant<int, 0>{}); return; } }
    if constexpr (D >= 2) { if (d == 1) { f(std::integral_constant<int, 1>{}); return; } }
    if constexpr (D >= 3) { if (d == 2) { f(std::integral_constant<int, 2>{}); return; } }
    if constexpr (D >= 4) { if (d == 3) { f(std::integral_constant<int, 3>{}); return; } }
    if constexpr (D >= 5) { if (d == 4) { f(std::integral_constant<int, 4>{}); return; } }
}

}  // namespace magi
