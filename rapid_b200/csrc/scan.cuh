// Exclusive prefix sum of int32 on the device: three small kernels (block-local scan + block totals, scan of the
// totals, add-back).  Used by the batch preprocessing (slot assignment, segment heads) and by the vote tally (the
// exact arrival index at which a proposal reaches the quorum).
#pragma once

#include "common.cuh"

namespace rapid {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                         // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 elements per block

__device__ __forceinline__ int32_t scan_block_exclusive(int32_t v, int32_t* warp_sums, int32_t* block_total) {
    // exclusive scan of one value per thread across the block (SCAN_THREADS threads)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int32_t s = lane < (SCAN_THREADS >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t x = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += x;
        }
        if (lane < (SCAN_THREADS >> 5)) warp_sums[lane] = s;      // inclusive scan of the warp totals
    }
    __syncthreads();
    const int32_t warp_off = wid ? warp_sums[wid - 1] : 0;
    if (block_total) *block_total = warp_sums[(SCAN_THREADS >> 5) - 1];
    return warp_off + inc - v;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int32_t* src, int32_t* data, int64_t n, int32_t* __restrict__ sums) {
    __shared__ int32_t warp_sums[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS], tsum = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) { v[q] = base + q < n ? src[base + q] : 0; tsum += v[q]; }
    int32_t total;
    int32_t run = scan_block_exclusive(tsum, warp_sums, &total);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) { if (base + q < n) data[base + q] = run; run += v[q]; }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile totals (any count; SCAN_THREADS at a time), grand total to *total
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(int32_t* __restrict__ sums, int32_t n_tiles, int32_t* __restrict__ total) {
    __shared__ int32_t warp_sums[SCAN_THREADS / 32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int32_t b = 0; b < n_tiles; b += SCAN_THREADS) {
        const int32_t i = b + threadIdx.x;
        const int32_t v = i < n_tiles ? sums[i] : 0;
        int32_t tot;
        const int32_t ex = scan_block_exclusive(v, warp_sums, &tot);
        const int32_t c = carry;
        if (i < n_tiles) sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int32_t* __restrict__ data, int64_t n, const int32_t* __restrict__ sums) {
    const int32_t off = sums[blockIdx.x];
    if (off == 0) return;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) if (base + q < n) data[base + q] += off;
}

// data[n] -> exclusive prefix sums in place; *total_dev (optional, device) = sum.  `sums` must hold ceil(n / 2048) ints.
static inline int32_t exclusive_scan_i32(int32_t* data, int64_t n, DevBuf<int32_t>& sums, int32_t* total_dev, cudaStream_t s, int* launches,
                                         const int32_t* src = nullptr) {
    if (n <= 0) {
        if (total_dev) RAPID_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int32_t), s));
        return RAPID_OK;
    }
    const int32_t tiles = (int32_t)ceil_div<int64_t>(n, SCAN_TILE);
    RAPID_CHECK(sums.reserve((size_t)tiles));
    k_scan_tiles<<<tiles, SCAN_THREADS, 0, s>>>(src ? src : data, data, n, sums.p);
    k_scan_sums<<<1, SCAN_THREADS, 0, s>>>(sums.p, tiles, total_dev);
    if (tiles > 1) k_scan_add<<<tiles, SCAN_THREADS, 0, s>>>(data, n, sums.p);
    RAPID_KERNEL_CHECK();
    if (launches) *launches += tiles > 1 ? 3 : 2;
    return RAPID_OK;
}

// dst[n] = exclusive prefix sums of src[n] (src untouched)
static inline int32_t exclusive_scan_i32_to(const int32_t* src, int32_t* dst, int64_t n, DevBuf<int32_t>& sums, cudaStream_t s, int* launches = nullptr) {
    if (n <= 0) return RAPID_OK;
    return exclusive_scan_i32(dst, n, sums, nullptr, s, launches, src);
}

}  // namespace rapid
