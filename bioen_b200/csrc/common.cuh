// common.cuh -- device-side building blocks shared by all bioen_b200 kernels (sm_100a only).
//
//  * error handling that turns CUDA failures into C++ exceptions (the C-ABI layer converts them to codes)
//  * mbarrier / TMA (cp.async.bulk[.tensor]) PTX wrappers used by the yTilde streaming kernels
//  * fixed-order warp / block reductions and the "last block finishes the sum" grid reduction that makes
//    every scalar this library produces run-to-run bit-reproducible (the GPU counterpart of the
//    reference's `_fast_openmp_flag == 0` mode, bioen/optimize/ext/c_bioen_common.c:46-55)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <cstdio>
#include <stdexcept>
#include <string>

namespace bioen {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "bioen_b200: CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e),
                 file, line, what);
        throw CudaError(buf);
    }
}
#define CUDA_CHECK(x) ::bioen::cuda_check((x), #x, __FILE__, __LINE__)

constexpr int kWarp = 32;

// NVTX range (header-only NVTX3: a no-op unless a profiler has injected itself).  Ranges mark the two halves of an
// evaluation, minimiser iterations, uploads and exchanges, so that an nsys / ncu --nvtx timeline reads in the terms of
// DESIGN.md.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// ------------------------------------------------------------------------------------------------
// mbarrier + TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make mbarrier.init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// L2 eviction policies for the streamed matrix (evict_first) vs. data that should stay resident
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// 2-D tiled TMA load global -> shared; completion is signalled on `bar` (complete_tx::bytes).
// c0 = coordinate in the contiguous (column) dimension, c1 = row.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned src/dst, size multiple of 16)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// same with an L2 eviction policy (streamed data: evict_first)
__device__ __forceinline__ void bulk_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// fixed-order reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of K values per thread.  Result valid in thread 0.  `red` needs K*32 doubles of smem.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();  // protect `red` against a previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[k * 32 + wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = lane < nw ? red[k * 32 + lane] : 0.0;
            v[k] = warp_sum(x);
        }
    }
}

// Grid-wide sum of K values: every block deposits its K partial sums, the block that arrives last adds all
// partials in a fixed order and calls fin(sum[K]) from its thread 0.  `partials` holds gridDim.x*K doubles,
// `ticket` is a zero-initialised counter that is reset for the next launch.  Deterministic because the
// final summation order depends only on the launch geometry.
template <int K, class Fin>
__device__ __forceinline__ void grid_sum(double (&v)[K], double* partials, unsigned int* ticket, double* red,
                                         Fin fin);

// Variant that hands the finish to the caller: returns true (to ALL threads of the block, block-uniformly) in the
// block that arrived last; there thread 0 holds the K totals in v[] and the ticket is already reset.  Used by the
// kernels that follow their grid reduction with an exchange between the ranks (all threads of the last block take
// part in it, see comm.cuh).
template <int K>
__device__ __forceinline__ bool grid_sum_last(double (&v)[K], double* partials, unsigned int* ticket, double* red) {
    __shared__ bool is_last;
    block_sum<K>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) partials[(size_t)blockIdx.x * K + k] = v[k];
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __ldcg(&partials[(size_t)b * K + k]);
    }
    block_sum<K>(acc, red);
    if (threadIdx.x == 0) {
        *ticket = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = acc[k];
    }
    return true;
}

template <int K, class Fin>
__device__ __forceinline__ void grid_sum(double (&v)[K], double* partials, unsigned int* ticket, double* red,
                                         Fin fin) {
    if (grid_sum_last<K>(v, partials, ticket, red) && threadIdx.x == 0) fin(v);
}

// Same, for K sums plus one maximum (v[K] is the max-reduced value); finish handed to the caller (see grid_sum_last).
template <int K>
__device__ __forceinline__ bool grid_sum_max_last(double (&v)[K + 1], double* partials, unsigned int* ticket,
                                                  double* red) {
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    auto block_reduce = [&](double (&x)[K + 1]) {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = warp_sum(x[k]);
        x[K] = warp_max(x[K]);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k <= K; ++k) red[k * 32 + wid] = x[k];
        }
        __syncthreads();
        if (wid == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = warp_sum(lane < nw ? red[k * 32 + lane] : 0.0);
            x[K] = warp_max(lane < nw ? red[K * 32 + lane] : -1.7976931348623157e308);
        }
    };
    block_reduce(v);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k <= K; ++k) partials[(size_t)blockIdx.x * (K + 1) + k] = v[k];
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double acc[K + 1];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    acc[K] = -1.7976931348623157e308;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __ldcg(&partials[(size_t)b * (K + 1) + k]);
        acc[K] = fmax(acc[K], __ldcg(&partials[(size_t)b * (K + 1) + K]));
    }
    block_reduce(acc);
    if (threadIdx.x == 0) {
        *ticket = 0;
#pragma unroll
        for (int k = 0; k <= K; ++k) v[k] = acc[k];
    }
    return true;
}

template <int K, class Fin>
__device__ __forceinline__ void grid_sum_max(double (&v)[K + 1], double* partials, unsigned int* ticket,
                                             double* red, Fin fin) {
    if (grid_sum_max_last<K>(v, partials, ticket, red) && threadIdx.x == 0) fin(v);
}

}  // namespace bioen
