// Stable LSD radix sort of (uint64 or uint32 key, int32 value) pairs, hand-written for sm_100a — the sort behind the K rings of a
// view (MembershipView's TreeSet order, view.cu), the NodeId order of the configuration id, the arrival-order grouping of the
// classic-Paxos tallies and the failure detectors' notification order.  One read of the keys counts all digits of all passes
// (global digit offsets); then every 8-bit pass is three launches over tiles of 4096 pairs:
//     k_rs_count     digit counts per tile
//     k_rs_tilescan  per digit, exclusive prefix of the counts over the tiles (+ the digit's global offset)
//     k_rs_scatter   ranks the tile's items per digit (warp-private counters + __match_any groups: stable, no atomics), sorts the
//                    tile by digit in shared memory and writes every digit's run with coalesced stores
// No spinning on other blocks (a decoupled look-back version stalled for ~150 us per pass when all tiles start together).
// Traffic per pass: 20 B read + 12 B written per pair — HBM/L2-bound like everything else on this path; no tensor cores.
#pragma once

#include "common.cuh"

namespace rapid {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                            // per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;          // 4096 pairs per tile
constexpr int RS_BINS = 256;

// ---- counts of every digit of every pass in one read of the keys ---------------------------------------------------------
template <typename KeyT>
static __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const KeyT* __restrict__ keys, int64_t n, int shift0, int passes,
                                                               uint32_t* __restrict__ hist /* [passes][256] */) {
    __shared__ uint32_t s_h[8 * RS_BINS];
    for (int i = threadIdx.x; i < passes * RS_BINS; i += RS_THREADS) s_h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
        const KeyT k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&s_h[p * RS_BINS + (int)((k >> (shift0 + 8 * p)) & 0xFF)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_BINS; i += RS_THREADS)
        if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

// exclusive prefix of each pass's 256 counts (one warp per pass)
static __global__ void __launch_bounds__(RS_THREADS) k_rs_scan(uint32_t* __restrict__ hist, int passes) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w >= passes) return;
    uint32_t* h = hist + w * RS_BINS;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = h[lane * 8 + q]; sum += v[q]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    uint32_t run = inc - sum;
#pragma unroll
    for (int q = 0; q < 8; ++q) { h[lane * 8 + q] = run; run += v[q]; }
}

template <typename KeyT>
struct RsSmem {
    uint32_t wcnt[RS_WARPS][RS_BINS];                   // per-warp digit counters -> exclusive offsets of the warp inside the tile's digit run
    uint32_t dstart[RS_BINS];                           // start of the digit's run inside the (digit-sorted) tile
    uint32_t gbase[RS_BINS];                            // where the tile's run of the digit starts in the output
    KeyT key[RS_TILE];
    int32_t val[RS_TILE];
};

// digit counts of every tile: counts[tile][256]
template <typename KeyT>
static __global__ void __launch_bounds__(RS_THREADS) k_rs_count(const KeyT* __restrict__ kin, int64_t n, int shift, uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_c[RS_BINS];
    s_c[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_c[(int)((kin[i] >> shift) & 0xFF)], 1u);
    }
    __syncthreads();
    counts[(size_t)blockIdx.x * RS_BINS + threadIdx.x] = s_c[threadIdx.x];
}

// block d: exclusive prefix over the tiles of counts[.][d], plus the digit's global offset -> tbase[tile][d]
static __global__ void __launch_bounds__(RS_THREADS) k_rs_tilescan(const uint32_t* __restrict__ counts, int64_t tiles,
                                                                   const uint32_t* __restrict__ gpref, uint32_t* __restrict__ tbase) {
    __shared__ uint32_t s_w[RS_WARPS];
    __shared__ uint32_t s_carry;
    const int d = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) s_carry = gpref[d];
    __syncthreads();
    for (int64_t b0 = 0; b0 < tiles; b0 += RS_THREADS) {
        const int64_t i = b0 + t;
        const uint32_t v = i < tiles ? counts[(size_t)i * RS_BINS + d] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int ww = 0; ww < w; ++ww) woff += s_w[ww];
        const uint32_t carry = s_carry;
        if (i < tiles) tbase[(size_t)i * RS_BINS + d] = carry + woff + inc - v;
        __syncthreads();
        if (t == RS_THREADS - 1) s_carry = carry + woff + inc;
        __syncthreads();
    }
}

template <typename KeyT>
static __global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const KeyT* __restrict__ kin, const int32_t* __restrict__ vin,
                                                                  KeyT* __restrict__ kout, int32_t* __restrict__ vout, int64_t n, int shift,
                                                                  const uint32_t* __restrict__ tbase /* [tiles][256] */) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem<KeyT>& sm = *reinterpret_cast<RsSmem<KeyT>*>(rs_raw);
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&sm.wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = blockIdx.x;
    const int64_t base = (int64_t)tile * RS_TILE + (int64_t)w * (32 * RS_ITEMS);   // this warp's contiguous chunk of the tile
    KeyT key[RS_ITEMS];
    int32_t val[RS_ITEMS];
    uint32_t off[RS_ITEMS];                             // rank of the item among the warp's items of the same digit
    // ---- rank (stable: chunks in warp order, rounds in order, lanes in order) ---------------------------------------------
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? kin[i] : (KeyT)0;
        val[r] = valid ? vin[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < n;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        off[r] = 0;
        if (valid) {
            const int d = (int)((key[r] >> shift) & 0xFF);
            const unsigned peers = __match_any_sync(act, d);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) { old = sm.wcnt[w][d]; sm.wcnt[w][d] = old + __popc(peers); }
            old = __shfl_sync(peers, old, leader);
            off[r] = old + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- per digit (thread d): offsets of the warps inside the tile's run, start of the run in the sorted tile ---------------
    {
        const int d = t;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const uint32_t c = sm.wcnt[ww][d]; sm.wcnt[ww][d] = run; run += c; }
        const uint32_t cnt = run;
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        __shared__ uint32_t s_ws[RS_WARPS];
        if (lane == 31) s_ws[w] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int ww = 0; ww < w; ++ww) woff += s_ws[ww];
        sm.dstart[d] = woff + inc - cnt;
        sm.gbase[d] = tbase[(size_t)tile * RS_BINS + d];
    }
    __syncthreads();
    // ---- the tile sorted by digit in shared memory ------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < n) {
            const int d = (int)((key[r] >> shift) & 0xFF);
            const uint32_t p = sm.dstart[d] + sm.wcnt[w][d] + off[r];
            sm.key[p] = key[r];
            sm.val[p] = val[r];
        }
    }
    __syncthreads();
    // ---- coalesced copy-out: consecutive threads write consecutive positions of a digit's run -----------------------------------
    const int64_t tile_left = n - (int64_t)tile * RS_TILE;
    const int64_t tile_n = tile_left < RS_TILE ? tile_left : RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int p = r * RS_THREADS + t;
        if (p < tile_n) {
            const KeyT k = sm.key[p];
            const int d = (int)((k >> shift) & 0xFF);
            const size_t g = (size_t)sm.gbase[d] + (size_t)(p - (int)sm.dstart[d]);
            kout[g] = k;
            vout[g] = sm.val[p];
        }
    }
}

struct RadixScratch {
    DevBuf<uint32_t> hist;                              // [8][256] global digit offsets of every pass
    DevBuf<uint32_t> counts, tbase;                     // [tiles][256] per-tile digit counts / output offsets of the pass in flight
    DevBuf<uint64_t> ktmp, ktmp2;                       // ping-pong buffers (also hold uint32 keys)
    DevBuf<int32_t> vtmp, vtmp2, vtmp3;
};

// Sorts n (key, value) pairs by bits [bit_begin, bit_end) of the key, stable.  The result is in (keys_out, vals_out).
// may_clobber_input: (keys_in, vals_in) may serve as a ping-pong buffer (saves a scratch pair); otherwise they are only read.
// vals_in == nullptr sorts keys only (vals_out is then ignored).  n < 2^30.
template <typename KeyT>
static inline int32_t radix_sort_pairs(RadixScratch& sc, KeyT* keys_in, int32_t* vals_in, KeyT* keys_out, int32_t* vals_out,
                                       int64_t n, int bit_begin, int bit_end, cudaStream_t s, bool may_clobber_input = true,
                                       int* launches = nullptr) {
    if (n <= 0) return RAPID_OK;
    if (n >= (1LL << 30)) { set_error("radix_sort_pairs: too many pairs"); return RAPID_EINVAL; }
    const int passes = (bit_end - bit_begin + 7) / 8;
    const bool keys_only = vals_in == nullptr;
    RAPID_CHECK(sc.vtmp.reserve((size_t)n));
    if (keys_only) { vals_in = sc.vtmp.p; vals_out = sc.vtmp.p; }     // the values ride along but are never looked at
    if (passes <= 0) {
        RAPID_CUDA(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * sizeof(KeyT), cudaMemcpyDeviceToDevice, s));
        if (!keys_only) RAPID_CUDA(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        return RAPID_OK;
    }
    if (passes > 8) { set_error("radix_sort_pairs: more than 8 digit passes"); return RAPID_EINVAL; }
    const int64_t tiles = ceil_div<int64_t>(n, RS_TILE);
    RAPID_CHECK(sc.hist.reserve(8 * RS_BINS + 8));
    RAPID_CHECK(sc.counts.reserve((size_t)tiles * RS_BINS)); RAPID_CHECK(sc.tbase.reserve((size_t)tiles * RS_BINS));
    RAPID_CUDA(cudaMemsetAsync(sc.hist.p, 0, (8 * RS_BINS + 8) * sizeof(uint32_t), s));
    RAPID_CUDA(cudaFuncSetAttribute(k_rs_scatter<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem<KeyT>)));
    const unsigned hgrid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(n, RS_THREADS * 8), 148 * 8);
    k_rs_hist<KeyT><<<hgrid, RS_THREADS, 0, s>>>(keys_in, n, bit_begin, passes, sc.hist.p);
    k_rs_scan<<<1, RS_THREADS, 0, s>>>(sc.hist.p, passes);
    RAPID_KERNEL_CHECK();
    // ping-pong so that the LAST pass lands in (keys_out, vals_out): intermediate passes alternate between X and Y, where
    // X / Y are the caller's input pair (if it may be clobbered) and scratch pairs
    KeyT* kx = keys_out; int32_t* vx = vals_out; KeyT* ky = keys_out; int32_t* vy = vals_out;
    if (passes > 1) {
        RAPID_CHECK(sc.ktmp.reserve((size_t)n)); RAPID_CHECK(sc.vtmp2.reserve((size_t)n));
        ky = reinterpret_cast<KeyT*>(sc.ktmp.p); vy = sc.vtmp2.p;
        if (may_clobber_input) { kx = keys_in; vx = vals_in; }
        else {
            RAPID_CHECK(sc.ktmp2.reserve((size_t)n)); RAPID_CHECK(sc.vtmp3.reserve((size_t)n));
            kx = reinterpret_cast<KeyT*>(sc.ktmp2.p); vx = sc.vtmp3.p;
        }
    }
    const KeyT* ksrc = keys_in; const int32_t* vsrc = vals_in;
    for (int p = 0; p < passes; ++p) {
        // pass p writes Y when p is even, X when odd; the last pass writes the output
        KeyT* kdst = p == passes - 1 ? keys_out : ((p & 1) ? kx : ky);
        int32_t* vdst = p == passes - 1 ? vals_out : ((p & 1) ? vx : vy);
        k_rs_count<KeyT><<<(unsigned)tiles, RS_THREADS, 0, s>>>(ksrc, n, bit_begin + 8 * p, sc.counts.p);
        k_rs_tilescan<<<RS_BINS, RS_THREADS, 0, s>>>(sc.counts.p, tiles, sc.hist.p + p * RS_BINS, sc.tbase.p);
        k_rs_scatter<KeyT><<<(unsigned)tiles, RS_THREADS, sizeof(RsSmem<KeyT>), s>>>(ksrc, vsrc, kdst, vdst, n, bit_begin + 8 * p, sc.tbase.p);
        RAPID_KERNEL_CHECK();
        ksrc = kdst; vsrc = vdst;
    }
    if (launches) *launches += 2 + 3 * passes;
    return RAPID_OK;
}

}  // namespace rapid
