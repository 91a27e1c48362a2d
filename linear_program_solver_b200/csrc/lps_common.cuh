// Shared device helpers for the sm_100a simplex pivot kernels: mbarrier + TMA bulk copies,
// streaming load/store flavours, warp/block arg-min reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

namespace lps {

constexpr int kWarp = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) -----------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LPS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LPS_DONE;\n"
        "bra LPS_WAIT;\n"
        "LPS_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- 128-bit global accesses, default and streaming (evict-first) flavours -------------------
template <bool STREAM>
__device__ __forceinline__ double2 ld_f64x2(const double* p) {
    double2 v;
    if (STREAM)
        asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else
        asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
template <bool STREAM>
__device__ __forceinline__ void st_f64x2(double* p, double2 v) {
    if (STREAM)
        asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    else
        asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
template <bool STREAM>
__device__ __forceinline__ uint4 ld_u32x4(const uint32_t* p) {
    uint4 v;
    if (STREAM)
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <bool STREAM>
__device__ __forceinline__ void st_u32x4(uint32_t* p, uint4 v) {
    if (STREAM)
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- arg-min with the reference's tie-break (lowest index among equal keys) -------------------
// idx == INT_MAX marks "no candidate".
struct ArgMinF64 {
    double key;
    int idx;
};
__device__ __forceinline__ ArgMinF64 argmin_pick(ArgMinF64 a, ArgMinF64 b) {
    if (b.idx == INT_MAX) return a;
    if (a.idx == INT_MAX) return b;
    if (b.key < a.key || (b.key == a.key && b.idx < a.idx)) return b;
    return a;
}
__device__ __forceinline__ ArgMinF64 warp_argmin(ArgMinF64 v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ArgMinF64 o;
        o.key = __shfl_xor_sync(0xffffffffu, v.key, off);
        o.idx = __shfl_xor_sync(0xffffffffu, v.idx, off);
        v = argmin_pick(v, o);
    }
    return v;
}
// Block-wide arg-min; result valid in every thread.  `scratch` holds >= 32 entries.
__device__ __forceinline__ ArgMinF64 block_argmin(ArgMinF64 v, ArgMinF64* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    v = warp_argmin(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    ArgMinF64 r = (lane < nwarps) ? scratch[lane] : ArgMinF64{0.0, INT_MAX};
    r = warp_argmin(r);
    return r;
}
__device__ __forceinline__ int warp_min_int(int v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ int block_min_int(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    v = warp_min_int(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int r = (lane < nwarps) ? scratch[lane] : INT_MAX;
    return warp_min_int(r);
}

}  // namespace lps
