// Stable LSD radix sort of (uint64 key, int32 value) pairs, hand-written for sm_100a — the sort behind the K rings of a view
// (MembershipView's TreeSet order, view.cu), the NodeId order of the configuration id and the arrival-order grouping of the
// classic-Paxos tallies.  "Onesweep" structure: ONE pass over the keys counts all digits of all passes, then every 8-bit pass is a
// single kernel in which a tile of 4096 pairs
//     ranks its items per digit (warp-private counters + __match_any groups: stable, no atomics),
//     learns how many items of each digit precede it in earlier tiles by decoupled look-back over a status word per (tile, digit),
//     sorts the tile by digit in shared memory and writes every digit's run with coalesced stores.
// Tiles take their index from an atomic ticket, so a tile only ever waits for tiles that are already running.
// Traffic per pass: 12 B read + 12 B written per pair — HBM-bound like everything else on this path; no tensor cores.
#pragma once

#include "common.cuh"

namespace rapid {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                            // per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;          // 4096 pairs per tile
constexpr int RS_BINS = 256;
constexpr uint32_t RS_FLAG_LOCAL = 1u << 30, RS_FLAG_INCL = 2u << 30, RS_VALUE = (1u << 30) - 1u;

// ---- counts of every digit of every pass in one read of the keys ---------------------------------------------------------
static __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, int64_t n, int shift0, int passes,
                                                               uint32_t* __restrict__ hist /* [passes][256] */) {
    __shared__ uint32_t s_h[8 * RS_BINS];
    for (int i = threadIdx.x; i < passes * RS_BINS; i += RS_THREADS) s_h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_THREADS) {
        const uint64_t k = keys[i];
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

struct RsSmem {
    uint32_t wcnt[RS_WARPS][RS_BINS];                   // per-warp digit counters -> exclusive offsets of the warp inside the tile's digit run
    uint32_t dstart[RS_BINS];                           // start of the digit's run inside the (digit-sorted) tile
    uint32_t gbase[RS_BINS];                            // where the tile's run of the digit starts in the output
    uint64_t key[RS_TILE];
    int32_t val[RS_TILE];
    uint32_t tile;
};

static __global__ void __launch_bounds__(RS_THREADS) k_rs_pass(const uint64_t* __restrict__ kin, const int32_t* __restrict__ vin,
                                                               uint64_t* __restrict__ kout, int32_t* __restrict__ vout, int64_t n, int shift,
                                                               const uint32_t* __restrict__ gpref /* [256] exclusive global digit offsets */,
                                                               uint32_t* __restrict__ status /* [tiles][256], zero */,
                                                               uint32_t* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_raw);
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    if (t == 0) sm.tile = atomicAdd(ticket, 1u);
    for (int i = t; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&sm.wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    const int64_t base = (int64_t)tile * RS_TILE + (int64_t)w * (32 * RS_ITEMS);   // this warp's contiguous chunk of the tile
    uint64_t key[RS_ITEMS];
    int32_t val[RS_ITEMS];
    uint32_t off[RS_ITEMS];                             // rank of the item among the warp's items of the same digit
    // ---- rank (stable: chunks in warp order, rounds in order, lanes in order) ---------------------------------------------
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? kin[i] : 0ull;
        val[r] = valid ? vin[i] : 0;
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
    // ---- per digit (thread d): offsets of the warps inside the tile's run, the tile's count, look-back -----------------------
    {
        const int d = t;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { const uint32_t c = sm.wcnt[ww][d]; sm.wcnt[ww][d] = run; run += c; }
        const uint32_t cnt = run;
        // exclusive scan of the digit counts over the block -> start of each digit's run in the sorted tile
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        __shared__ uint32_t s_ws[RS_WARPS];
        if (lane == 31) s_ws[w] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int ww = 0; ww < w; ++ww) woff += s_ws[ww];
        sm.dstart[d] = woff + inc - cnt;
        // decoupled look-back: how many items of digit d sit in earlier tiles
        volatile uint32_t* st = status;
        st[(size_t)tile * RS_BINS + d] = RS_FLAG_LOCAL | cnt;
        uint32_t excl = 0;
        for (int64_t t2 = (int64_t)tile - 1; t2 >= 0; --t2) {
            uint32_t v;
            do { v = st[(size_t)t2 * RS_BINS + d]; } while ((v & (RS_FLAG_LOCAL | RS_FLAG_INCL)) == 0);
            excl += v & RS_VALUE;
            if (v & RS_FLAG_INCL) break;
        }
        __threadfence();
        st[(size_t)tile * RS_BINS + d] = RS_FLAG_INCL | (excl + cnt);
        sm.gbase[d] = gpref[d] + excl;
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
            const uint64_t k = sm.key[p];
            const int d = (int)((k >> shift) & 0xFF);
            const size_t g = (size_t)sm.gbase[d] + (size_t)(p - (int)sm.dstart[d]);
            kout[g] = k;
            vout[g] = sm.val[p];
        }
    }
}

struct RadixScratch {
    DevBuf<uint32_t> hist;                              // [8][256] + ticket words
    DevBuf<uint32_t> status;                            // [tiles][256]
    DevBuf<uint64_t> ktmp;
    DevBuf<int32_t> vtmp;
};

// Sorts n (key, value) pairs by bits [bit_begin, bit_end) of the key, stable.  The result is in (keys_out, vals_out); keys_in /
// vals_in are overwritten when an odd number of passes needs the ping-pong.  n < 2^30.
static inline int32_t radix_sort_pairs(RadixScratch& sc, uint64_t* keys_in, int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out,
                                       int64_t n, int bit_begin, int bit_end, cudaStream_t s, int* launches = nullptr) {
    if (n <= 0) return RAPID_OK;
    if (n >= (1LL << 30)) { set_error("radix_sort_pairs: too many pairs"); return RAPID_EINVAL; }
    const int passes = (bit_end - bit_begin + 7) / 8;
    if (passes <= 0) {
        RAPID_CUDA(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        RAPID_CUDA(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        return RAPID_OK;
    }
    const int64_t tiles = ceil_div<int64_t>(n, RS_TILE);
    RAPID_CHECK(sc.hist.reserve(8 * RS_BINS + 8));
    RAPID_CHECK(sc.status.reserve((size_t)tiles * RS_BINS));
    RAPID_CUDA(cudaMemsetAsync(sc.hist.p, 0, (8 * RS_BINS + 8) * sizeof(uint32_t), s));
    RAPID_CUDA(cudaFuncSetAttribute(k_rs_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
    const unsigned hgrid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(n, RS_THREADS * 8), 148 * 8);
    k_rs_hist<<<hgrid, RS_THREADS, 0, s>>>(keys_in, n, bit_begin, passes, sc.hist.p);
    k_rs_scan<<<1, RS_THREADS, 0, s>>>(sc.hist.p, passes);
    RAPID_KERNEL_CHECK();
    // ping-pong so that the LAST pass lands in (keys_out, vals_out): odd pass counts bounce between `in` and `out`, even ones
    // between `in` and a scratch pair
    uint64_t* ky = keys_out; int32_t* vy = vals_out;
    if (!(passes & 1)) {
        RAPID_CHECK(sc.ktmp.reserve((size_t)n)); RAPID_CHECK(sc.vtmp.reserve((size_t)n));
        ky = sc.ktmp.p; vy = sc.vtmp.p;
    }
    const uint64_t* ksrc = keys_in; const int32_t* vsrc = vals_in;
    for (int p = 0; p < passes; ++p) {
        uint64_t* kdst = p == passes - 1 ? keys_out : ((p & 1) ? keys_in : ky);
        int32_t* vdst = p == passes - 1 ? vals_out : ((p & 1) ? vals_in : vy);
        RAPID_CUDA(cudaMemsetAsync(sc.status.p, 0, (size_t)tiles * RS_BINS * sizeof(uint32_t), s));
        k_rs_pass<<<(unsigned)tiles, RS_THREADS, sizeof(RsSmem), s>>>(ksrc, vsrc, kdst, vdst, n, bit_begin + 8 * p, sc.hist.p + p * RS_BINS,
                                                                      sc.status.p, sc.hist.p + 8 * RS_BINS + p);
        RAPID_KERNEL_CHECK();
        ksrc = kdst; vsrc = vdst;
    }
    if (launches) *launches += 2 + passes;
    return RAPID_OK;
}

}  // namespace rapid
