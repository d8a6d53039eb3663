// The subject-bucketed cut-detection kernels (the fast path of rapid_cd_apply_batch on SERVICE handles).
//
// Reference semantics: MultiNodeCutDetector.java:84-128 applied per cell in arrival order, then :137-164, driven
// by MembershipService.java:300-354.  The Java walks the batch once per process and probes a hash map per cell.
// Here the batch is regrouped BY SUBJECT once (counting sort by subject slot inside cd_prepare.cu), and every receiver's
// 16-bit ring mask for a subject is read ONCE, updated in a register and written ONCE:
//
//   traffic = 4 bytes x (#subjects in the batch) x (#receivers)          (SURVEY.md §8d "4·S·R")
//
// instead of 4 bytes x #cells x #receivers for a per-cell sweep.  What makes that legal is that the sequential
// rule "emit when updatesInProgress returns to 0" only depends on, per subject, the two moments at which its
// report count crosses L and H (t_L, t_H): a proposal is emitted at t_H(s) iff no other subject s' has
// t_L(s') <= t_H(s) < t_H(s') (SURVEY.md §7 "key reformulation").  Each (subject, receiver) visit yields
// (t_L, t_H); per receiver we keep a handful of order-independent reductions of them and classify:
//     all pre-proposals resolved  -> everything at >= H is emitted              (EMIT_ALL)
//     an unresolved one starts before the first H-crossing -> nothing emitted   (NOEMIT)
//     otherwise                                                                  (MIXED: exact interval analysis
//                                                                                 for that receiver only)
// followed by the implicit-invalidation pass over the (few) subjects left in the unstable band.
//
// "Moments" are cell indices for uniform delivery and the receiver's own permutation keys for PERMUTED delivery,
// so a per-receiver order needs no per-receiver sort.
//
// Kernels: k_apply_uniform (SWAR: 8 receivers per thread, 128-bit loads/stores, write-only fresh-subject path),
// k_apply_generic (one receiver per thread: bitmap / permuted delivery), k_finalize1, k_mixed_pass / k_mixed_update /
// k_mixed_commit / k_mixed_mark (interval analysis), k_flip, k_inval_pairs, k_finalize2.
#include <algorithm>
#include <cstdlib>

#include "cd_internal.cuh"

namespace rapid {

constexpr int TILE_R = 1024;          // receivers per tile (uniform kernel: 128 threads x 8 receivers)
constexpr int UNI_THREADS = 128;
constexpr int GEN_THREADS = 256;
constexpr int STAGE = 32;             // batch subjects staged in shared memory at a time
constexpr int MAXK = RAPID_MAX_K;
constexpr int SMALL_SEG = 4;          // generic visit: subjects with at most this many cells in the batch stay in registers
constexpr uint32_t T32_NONE = 0xFFFFFFFFu;
constexpr uint64_t T64_NONE = ~0ULL;
constexpr uint32_t RF_K3 = 16u;       // receiver enters the invalidation pass of the batch in flight
constexpr uint32_t RF_ACTIVE = 32u;   // receiver processed the batch in flight
constexpr uint32_t RF_MIXED_EMIT = 64u;   // announced a proposal found by the interval analysis: "emitted" == t_H <= estar[r]

// per-receiver state of the interval analysis (MIXED receivers)
constexpr uint32_t MX_ON = 1u, MX_NEG = 2u, MX_HAS_E = 4u, MX_DONE = 8u;

// partial-accumulator flags
constexpr uint32_t PF_SEEN = 1u;      // a valid DOWN cell was delivered
constexpr uint32_t PF_NEGINF = 2u;    // a subject already in the unstable band stayed there (its t_L is "before the batch")

struct ChunkAcc {                     // what the FRESH subjects of a chunk contribute to EVERY active receiver
    uint32_t nLH, tpUn, fl, minTH, minTLun, pad_;
    uint64_t h1, h2;
};

struct Partials {                     // [n_chunks][Rpad] structure of arrays
    uint4* cnt;                       // x = nL | nH << 16, y = touched_pre | nUn << 16, z = flags, w = 0
    uint64_t* minTH;
    uint64_t* minTLun;
    uint64_t* h1;
    uint64_t* h2;
    int32_t* flag;                    // [n_chunks][n_tiles] 1 = the per-receiver partials of this (chunk, tile) were written
    ChunkAcc* chunk;                  // [n_chunks] used for (chunk, tile) pairs whose flag is 0
    int n_tiles;
};

struct Bucketed {
    DevBuf<int32_t> sidx;                     // cell indices grouped by subject (arrival order inside a subject)
    DevBuf<int32_t> seg_cnt, seg_pos;         // [slot] scratch of the prepare kernel (seg_cnt all zero between batches)
    DevBuf<SubjDesc> desc;
    DevBuf<SubjWalk> walk;
    DevBuf<uint8_t> s_ring, s_status;         // per sorted cell
    DevBuf<int32_t> p_flag;
    DevBuf<ChunkAcc> p_chunk;
    DevBuf<uint4> p_cnt;
    DevBuf<uint64_t> p_minTH, p_minTLun, p_h1, p_h2;
    DevBuf<uint32_t> mx_fl;                   // [Rpad] MX_* flags
    DevBuf<uint64_t> mx_a, mx_cand, mx_emax;  // [Rpad] start of the never-closing component / its next candidate / e* candidate
    DevBuf<uint64_t> mx_p1, mx_p2;            // [Rpad] fingerprint of `proposal` before the batch
    DevBuf<int32_t> mx_pc;
    DevBuf<unsigned long long> mx_e1, mx_e2;  // [Rpad] fingerprint of the batch subjects emitted explicitly
    DevBuf<int32_t> mx_ec;
    DevBuf<uint64_t> estar;                   // [Rpad] last explicit emission moment of RF_MIXED_EMIT receivers
    DevBuf<int32_t> mx_changed;               // [1]
    DevBuf<int32_t> batch_index;              // [slot] -> index of the subject in the batch in flight
    DevBuf<int32_t> k3_res;
    DevBuf<unsigned long long> k3_h1, k3_h2;
    DevBuf<int2> pre_pairs;
    DevBuf<int32_t> pre_count;
    DevBuf<int32_t> in_list;                  // [slot][n_tiles]
    size_t in_list_slots = 0;
    int n_tiles = 0;
    size_t part_cap = 0;
    int slots_uniform = 0, slots_generic = 0;   // resident blocks of the apply kernels on this device
};

// ------------------------------------------------------------------------------------------------------------------
// the (subject, receiver) visit, uniform delivery: old 16-bit state -> crossings and moments
// ------------------------------------------------------------------------------------------------------------------
struct Visit {
    int c0, c1;
    bool crossL, crossH;
    uint32_t tL, tH;
};

__device__ __forceinline__ Visit visit_uniform(uint32_t ur, const SubjDesc& d, const SubjWalk& w, int L, int H) {
    Visit v;
    v.tL = 0; v.tH = 0;
    if (ur == 0) {                                      // fresh subject: the descriptor already knows the answer
        v.c0 = 0; v.c1 = d.nr;
        v.tL = d.tLf; v.tH = d.tHf;
    } else {
        int c = __popc(ur);
        v.c0 = c;
        const int nr = d.nr;
        for (int q = 0; q < nr; ++q) {
            const int k = w.ring[q];
            const bool isnew = !((ur >> k) & 1u);
            c += isnew;
            if (isnew && c == L) v.tL = w.time[q];
            if (isnew && c == H) v.tH = w.time[q];
        }
        v.c1 = c;
    }
    v.crossL = v.c0 < L && v.c1 >= L;
    v.crossH = v.c0 < H && v.c1 >= H;
    return v;
}

struct Acc {
    uint32_t nL = 0, nH = 0, tp = 0, nUn = 0, flags = 0;
    uint32_t minTH = T32_NONE, minTLun = T32_NONE;
    uint64_t h1 = 0, h2 = 0;
};

__device__ __forceinline__ bool accumulate(Acc& a, const Visit& v, const SubjDesc& d, int L, int H) {
    if (v.c0 >= L && v.c0 < H) a.tp++;
    if (v.crossL) a.nL++;
    if (v.crossH) { a.nH++; a.minTH = min(a.minTH, v.tH); a.h1 += d.mix1; a.h2 += d.mix2; }
    if (v.c1 >= L && v.c1 < H) {                        // still in the unstable band after the batch
        if (v.crossL) { a.nUn++; a.minTLun = min(a.minTLun, v.tL); }
        else a.flags |= PF_NEGINF;
        return true;
    }
    return false;
}

struct ApplyArgs {
    uint16_t* masks;
    const uint8_t* cur;
    size_t Rpad;
    int K, H, L;
    int64_t R, rbegin;
    const uint32_t* rflags;
    DeliveryDev dl;
    int Sb, chunk;
    const SubjDesc* desc;
    const SubjWalk* walk;
    const int32_t* slot_subject;
    const int32_t* sidx;          // sorted cell indices
    const uint8_t* s_ring;
    const uint8_t* s_status;
    Partials part;
    int n_tiles;
    int32_t* in_list;
    int2* pre_pairs;
    int32_t* pre_count;
    int32_t pre_cap;
    int32_t S_before;             // slots >= S_before were assigned by this batch: their state is known-zero (never read)
};

__device__ __forceinline__ void note_unresolved(const ApplyArgs& a, int tile, int32_t slot) {
    int32_t* f = a.in_list + (size_t)slot * a.n_tiles + tile;
    if (atomicExch(f, 1) == 0) {
        const int32_t at = atomicAdd(a.pre_count, 1);
        if (at < a.pre_cap) a.pre_pairs[at] = make_int2(tile, slot);
    }
}

// ---- uniform delivery: every active receiver gets every valid cell in array order -------------------------------------
// One thread owns 8 consecutive receivers (one 128-bit load + one 128-bit store per subject).  The common case is
// that all of a thread's ACTIVE receivers hold the same state for the subject (they saw the same history): the
// visit is computed once, merged into the new word with a SWAR mask, and accumulated in registers ("com").  Only
// when active neighbours disagree (partitions) do we fall back to a per-receiver visit whose accumulators live in
// the thread's own slice of the global partial arrays.
__device__ __forceinline__ void part_store(const Partials& p, size_t at, uint32_t nLH, uint32_t tpUn, uint32_t fl,
                                           uint32_t mTH, uint32_t mTL, uint64_t h1, uint64_t h2) {
    p.cnt[at] = make_uint4(nLH, tpUn, fl, 0u);
    p.minTH[at] = mTH == T32_NONE ? T64_NONE : (uint64_t)mTH;
    p.minTLun[at] = mTL == T32_NONE ? T64_NONE : (uint64_t)mTL;
    p.h1[at] = h1;
    p.h2[at] = h2;
}

// Subjects whose slot was assigned by this very batch ("fresh") have known-zero state for EVERY receiver: nothing is
// read, the new word is the batch's ring mask under the thread's activity mask, and their contribution to the
// per-receiver accumulators is the same for every active receiver — it is reduced once per stage by warp 0 and added
// at the end.  Only carried subjects (reports from earlier batches) take the load / compare / visit path.
struct StageAcc {
    uint32_t nLH, tpUn, fl, minTH, minTLun;
    uint64_t h1, h2;
};

__global__ void __launch_bounds__(UNI_THREADS) k_apply_uniform(const ApplyArgs a) {
    __shared__ SubjDesc sd[STAGE];
    __shared__ SubjWalk sw[STAGE];
    __shared__ const uint16_t* s_src[STAGE];
    __shared__ uint16_t* s_dst[STAGE];
    __shared__ uint32_t s_nw[STAGE];          // (batch ring mask) replicated in both half-words
    __shared__ int s_unres[STAGE];
    __shared__ StageAcc s_facc;               // fresh-subject accumulators of this block's chunk (same for every receiver)

    const int tile = blockIdx.x, chunk = blockIdx.y, t = threadIdx.x;
    const int s0 = chunk * a.chunk, s1 = min(a.Sb, s0 + a.chunk);
    const int64_t r0 = (int64_t)tile * TILE_R + (int64_t)t * 8;
    const size_t pbase = (size_t)chunk * a.Rpad + (size_t)r0;
    const uint32_t RM = (1u << a.K) - 1u;
    const int L = a.L, H = a.H;

    uint32_t act = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t r = r0 + j;
        if (r < a.R) {
            const bool on = !(a.rflags[r] & RF_ANNOUNCED) && !((a.dl.flags & RAPID_DELIVERY_BLOCKED) && a.dl.blocked[r]);
            act |= (on ? 1u : 0u) << j;
        }
    }
    // SWAR masks: 0xFFFF per active half-word
    uint32_t am[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) am[q] = (((act >> (2 * q)) & 1u) ? 0x0000FFFFu : 0u) | (((act >> (2 * q + 1)) & 1u) ? 0xFFFF0000u : 0u);
    if (t == 0) { s_facc.nLH = 0; s_facc.tpUn = 0; s_facc.fl = 0; s_facc.minTH = T32_NONE; s_facc.minTLun = T32_NONE; s_facc.h1 = 0; s_facc.h2 = 0; }
    const int block_active = __syncthreads_or(act != 0);
    Acc com;                                              // carried subjects: shared by all ACTIVE receivers of this thread
    bool had_exc = false;                                 // the thread's global partial slots hold per-receiver extras
    bool carried = false;                                 // this thread visited a carried subject (com is not just zeros)

    for (int base = s0; base < s1; base += STAGE) {
        const int n = min(STAGE, s1 - base);
        __syncthreads();
        if (t < 32) {                                     // warp 0 stages the descriptors (STAGE == 32)
            uint32_t nLH = 0, tpUn = 0, mTH = T32_NONE, mTL = T32_NONE;
            uint64_t h1 = 0, h2 = 0;
            if (t < n) {
                const SubjDesc d = a.desc[base + t];
                sd[t] = d;
                const uint8_t c = a.cur[d.slot];
                const bool fresh = d.slot >= a.S_before;
                s_src[t] = fresh ? nullptr : a.masks + ((size_t)d.slot * 2 + c) * a.Rpad;
                s_dst[t] = a.masks + ((size_t)d.slot * 2 + (c ^ 1)) * a.Rpad;
                s_nw[t] = (uint32_t)d.bmask * 0x10001u;
                int un = 0;
                if (fresh) {
                    const int nr = d.nr;
                    if (nr >= L) nLH += 1u;
                    if (nr >= H) { nLH += 1u << 16; mTH = d.tHf; h1 = d.mix1; h2 = d.mix2; }
                    else if (nr >= L) { tpUn += 1u << 16; mTL = d.tLf; un = block_active; }
                } else {
                    sw[t] = a.walk[base + t];
                }
                s_unres[t] = un;
            }
            // warp reduction of the fresh subjects' contribution
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                nLH += __shfl_down_sync(0xffffffffu, nLH, o);
                tpUn += __shfl_down_sync(0xffffffffu, tpUn, o);
                mTH = min(mTH, __shfl_down_sync(0xffffffffu, mTH, o));
                mTL = min(mTL, __shfl_down_sync(0xffffffffu, mTL, o));
                h1 += __shfl_down_sync(0xffffffffu, h1, o);
                h2 += __shfl_down_sync(0xffffffffu, h2, o);
            }
            if (t == 0) {
                s_facc.nLH += nLH; s_facc.tpUn += tpUn; s_facc.minTH = min(s_facc.minTH, mTH);
                s_facc.minTLun = min(s_facc.minTLun, mTL); s_facc.h1 += h1; s_facc.h2 += h2;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const uint16_t* src = s_src[i];
            uint16_t* dst = s_dst[i];
            const uint32_t nwb = s_nw[i];
            if (src == nullptr) {                          // fresh subject: write-only
                *reinterpret_cast<uint4*>(dst + r0) = make_uint4(nwb & am[0], nwb & am[1], nwb & am[2], nwb & am[3]);
                continue;
            }
            uint4 w = *reinterpret_cast<const uint4*>(src + r0);
            bool unres = false;
            if (act) {
                const SubjDesc& d = sd[i];
                // all active half-words equal  <=>  AND over them == OR over them
                const uint32_t andw = (w.x | ~am[0]) & (w.y | ~am[1]) & (w.z | ~am[2]) & (w.w | ~am[3]);
                const uint32_t orw = (w.x & am[0]) | (w.y & am[1]) | (w.z & am[2]) | (w.w & am[3]);
                const uint32_t andv = andw & (andw >> 16) & 0xFFFFu, st = (orw | (orw >> 16)) & 0xFFFFu;
                carried = true;
                if (andv == st) {
                    const Visit v = visit_uniform(st & RM, d, sw[i], L, H);
                    unres = accumulate(com, v, d, L, H);
                    const uint32_t nw = st * 0x10001u | nwb;
                    w.x = (w.x & ~am[0]) | (nw & am[0]);
                    w.y = (w.y & ~am[1]) | (nw & am[1]);
                    w.z = (w.z & ~am[2]) | (nw & am[2]);
                    w.w = (w.w & ~am[3]) | (nw & am[3]);
                } else {
                    if (!had_exc) {
                        had_exc = true;
#pragma unroll
                        for (int j = 0; j < 8; ++j) part_store(a.part, pbase + j, 0u, 0u, 0u, T32_NONE, T32_NONE, 0ull, 0ull);
                    }
                    uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (!((act >> j) & 1u)) continue;
                        const uint32_t sj = (words[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
                        const Visit v = visit_uniform(sj & RM, d, sw[i], L, H);
                        Acc ex;
                        unres |= accumulate(ex, v, d, L, H);
                        const size_t at = pbase + j;
                        uint4 c = a.part.cnt[at];
                        c.x += ex.nL | (ex.nH << 16); c.y += ex.tp | (ex.nUn << 16); c.z |= ex.flags;
                        a.part.cnt[at] = c;
                        if (ex.minTH != T32_NONE) { const uint64_t o = a.part.minTH[at]; if ((uint64_t)ex.minTH < o) a.part.minTH[at] = ex.minTH; }
                        if (ex.minTLun != T32_NONE) { const uint64_t o = a.part.minTLun[at]; if ((uint64_t)ex.minTLun < o) a.part.minTLun[at] = ex.minTLun; }
                        if (ex.nH) { a.part.h1[at] += ex.h1; a.part.h2[at] += ex.h2; }
                        words[j >> 1] |= (uint32_t)d.bmask << ((j & 1) * 16);
                    }
                    w = make_uint4(words[0], words[1], words[2], words[3]);
                }
            }
            *reinterpret_cast<uint4*>(dst + r0) = w;       // the non-current row becomes the new state
            if (__any_sync(0xffffffffu, unres) && (t & 31) == 0) s_unres[i] = 1;
        }
        __syncthreads();
        if (t < n && s_unres[t]) note_unresolved(a, tile, sd[t].slot);
    }
    // Fresh subjects contribute the same to every active receiver: that goes to ONE record per chunk.  Per-receiver
    // partials are only written by blocks in which some thread met a carried subject; k_finalize1 adds the two.
    const int need = __syncthreads_or((carried || had_exc) ? 1 : 0);
    if (t == 0) {
        a.part.flag[(size_t)chunk * a.part.n_tiles + tile] = need;
        if (tile == 0) {
            const StageAcc f = s_facc;
            ChunkAcc c;
            c.nLH = f.nLH; c.tpUn = f.tpUn; c.fl = f.fl; c.minTH = f.minTH; c.minTLun = f.minTLun; c.pad_ = 0; c.h1 = f.h1; c.h2 = f.h2;
            a.part.chunk[chunk] = c;
        }
    }
    if (!need) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const size_t at = pbase + j;
        const bool on = (act >> j) & 1u;
        if (!had_exc) {
            if (on) part_store(a.part, at, com.nL | (com.nH << 16), com.tp | (com.nUn << 16), com.flags, com.minTH, com.minTLun, com.h1, com.h2);
            else part_store(a.part, at, 0u, 0u, 0u, T32_NONE, T32_NONE, 0ull, 0ull);
        } else if (on) {
            uint4 c = a.part.cnt[at];
            c.x += com.nL | (com.nH << 16); c.y += com.tp | (com.nUn << 16); c.z |= com.flags;
            a.part.cnt[at] = c;
            if (com.minTH != T32_NONE) { const uint64_t o = a.part.minTH[at]; if ((uint64_t)com.minTH < o) a.part.minTH[at] = com.minTH; }
            if (com.minTLun != T32_NONE) { const uint64_t o = a.part.minTLun[at]; if ((uint64_t)com.minTLun < o) a.part.minTLun[at] = com.minTLun; }
            a.part.h1[at] += com.h1; a.part.h2[at] += com.h2;
        }
    }
}

// ---- generic delivery (per-receiver subset and/or per-receiver order): one receiver per thread -------------------
struct GVisit {
    int c0, c1;
    bool crossL, crossH, seen_down;
    uint64_t tL, tH;
    uint32_t have;        // rings delivered to this receiver in this batch
};

// moments of a subject's cells as seen by one receiver; `rs` is the receiver's permutation seed
__device__ __forceinline__ GVisit visit_generic(uint32_t ur, const SubjDesc& d, const int32_t* __restrict__ sidx,
                                                const uint8_t* __restrict__ s_ring, const uint8_t* __restrict__ s_status,
                                                const DeliveryDev& dl, int64_t r, uint64_t rs, int L, int H) {
    GVisit v;
    v.have = 0; v.seen_down = false; v.tL = 0; v.tH = 0;
    const bool has_bitmap = dl.flags & RAPID_DELIVERY_BITMAP, permuted = dl.flags & RAPID_DELIVERY_PERMUTED;
    if (d.seg_len <= SMALL_SEG) {
        // Few cells (the common case: ~1-3 reports of a subject per batch): work on the cells themselves instead of a
        // per-ring table — first occurrence of each not-yet-reported ring, then its rank by moment, all in registers.
        uint64_t tm[SMALL_SEG];
        int rk[SMALL_SEG];
        bool ok[SMALL_SEG];
#pragma unroll
        for (int j = 0; j < SMALL_SEG; ++j) {
            ok[j] = false; tm[j] = 0; rk[j] = 0;
            if (j < (int)d.seg_len) {
                const int32_t ci = sidx[d.seg_begin + j];
                if (!has_bitmap || ((dl.bitmap[(size_t)ci * dl.words + (r >> 5)] >> (r & 31)) & 1u)) {
                    ok[j] = true;
                    rk[j] = s_ring[d.seg_begin + j];
                    tm[j] = permuted ? splitmix64(rs ^ (uint64_t)ci) : (uint64_t)ci + 1ull;
                    if (s_status[d.seg_begin + j] == RAPID_EDGE_DOWN) v.seen_down = true;
                    v.have |= 1u << rk[j];
                }
            }
        }
        const uint32_t fresh = v.have & ~ur;
        v.c0 = __popc(ur);
        v.c1 = v.c0 + __popc(fresh);
        v.crossL = v.c0 < L && v.c1 >= L;
        v.crossH = v.c0 < H && v.c1 >= H;
        if (v.crossL || v.crossH) {
            const int wantL = L - v.c0 - 1, wantH = H - v.c0 - 1;
            bool first[SMALL_SEG];
#pragma unroll
            for (int j = 0; j < SMALL_SEG; ++j) {             // first report of a ring this receiver had not counted yet
                first[j] = ok[j] && !((ur >> rk[j]) & 1u);
#pragma unroll
                for (int i = 0; i < SMALL_SEG; ++i)
                    if (i != j && ok[i] && rk[i] == rk[j] && tm[i] < tm[j]) first[j] = false;
            }
#pragma unroll
            for (int j = 0; j < SMALL_SEG; ++j) {
                if (!first[j]) continue;
                int rank = 0;
#pragma unroll
                for (int i = 0; i < SMALL_SEG; ++i) rank += (first[i] && tm[i] < tm[j]) ? 1 : 0;
                if (rank == wantL) v.tL = tm[j];
                if (rank == wantH) v.tH = tm[j];
            }
        }
        return v;
    }
    uint64_t tmin[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) tmin[k] = 0;
    const uint32_t e = d.seg_begin + d.seg_len;
    for (uint32_t j = d.seg_begin; j < e; ++j) {
        const int32_t ci = sidx[j];
        if (has_bitmap && !((dl.bitmap[(size_t)ci * dl.words + (r >> 5)] >> (r & 31)) & 1u)) continue;
        if (s_status[j] == RAPID_EDGE_DOWN) v.seen_down = true;
        const int k = s_ring[j];
        const uint64_t tm = permuted ? splitmix64(rs ^ (uint64_t)ci) : (uint64_t)ci + 1ull;
        const bool had = (v.have >> k) & 1u;
#pragma unroll
        for (int kk = 0; kk < MAXK; ++kk)
            if (kk == k && (!had || tm < tmin[kk])) tmin[kk] = tm;
        v.have |= 1u << k;
    }
    const uint32_t fresh = v.have & ~ur;
    v.c0 = __popc(ur);
    v.c1 = v.c0 + __popc(fresh);
    v.crossL = v.c0 < L && v.c1 >= L;
    v.crossH = v.c0 < H && v.c1 >= H;
    if (v.crossL || v.crossH) {
        const int wantL = L - v.c0 - 1, wantH = H - v.c0 - 1;       // rank (0-based) among the new rings' first moments
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (!((fresh >> k) & 1u)) continue;
            int rank = 0;
#pragma unroll
            for (int q = 0; q < MAXK; ++q) rank += (((fresh >> q) & 1u) && tmin[q] < tmin[k]) ? 1 : 0;
            if (rank == wantL) v.tL = tmin[k];
            if (rank == wantH) v.tH = tmin[k];
        }
    }
    return v;
}

__global__ void __launch_bounds__(GEN_THREADS, 4) k_apply_generic(const ApplyArgs a) {
    __shared__ SubjDesc sd[STAGE];
    __shared__ const uint16_t* s_src[STAGE];
    __shared__ uint16_t* s_dst[STAGE];
    __shared__ int s_unres[STAGE];
    const int t = threadIdx.x, chunk = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.x * GEN_THREADS + t;
    const int tile = (int)(((int64_t)blockIdx.x * GEN_THREADS) / TILE_R);
    const int s0 = chunk * a.chunk, s1 = min(a.Sb, s0 + a.chunk);
    const uint32_t RM = (1u << a.K) - 1u;
    const int L = a.L, H = a.H;
    const bool in_range = r < a.R;
    const bool active = in_range && !(a.rflags[r] & RF_ANNOUNCED) && !((a.dl.flags & RAPID_DELIVERY_BLOCKED) && a.dl.blocked[r]);
    const uint64_t rs = splitmix64(a.dl.perm_seed + (uint64_t)(a.rbegin + r));
    uint32_t nL = 0, nH = 0, tp = 0, nUn = 0, fl = 0;
    uint64_t minTH = T64_NONE, minTLun = T64_NONE, h1 = 0, h2 = 0;
    bool haveTH = false, haveTL = false;
    for (int base = s0; base < s1; base += STAGE) {
        const int n = min(STAGE, s1 - base);
        __syncthreads();
        if (t < n) {
            const SubjDesc d = a.desc[base + t];
            sd[t] = d;
            const uint8_t c = a.cur[d.slot];
            s_src[t] = d.slot >= a.S_before ? nullptr : a.masks + ((size_t)d.slot * 2 + c) * a.Rpad;
            s_dst[t] = a.masks + ((size_t)d.slot * 2 + (c ^ 1)) * a.Rpad;
            s_unres[t] = 0;
        }
        __syncthreads();
        // the visit is a dependent chain (state word -> cells -> moments): keep 8 state loads in flight per thread
        uint32_t pre[8];
        for (int i = 0; i < n; ++i) {
            if ((i & 7) == 0) {
#pragma unroll
                for (int u = 0; u < 8; ++u) pre[u] = (i + u < n && s_src[i + u]) ? s_src[i + u][r] : 0u;
            }
            const SubjDesc& d = sd[i];
            bool unres = false;
            if (r < (int64_t)a.Rpad) {
                uint32_t st = pre[i & 7];
                if (active) {
                    const GVisit v = visit_generic(st & RM, d, a.sidx, a.s_ring, a.s_status, a.dl, r, rs, L, H);
                    if (v.seen_down) fl |= PF_SEEN;
                    if (v.c0 >= L && v.c0 < H) tp++;
                    if (v.crossL) nL++;
                    if (v.crossH) {
                        nH++;
                        if (!haveTH || v.tH < minTH) { minTH = v.tH; haveTH = true; }
                        h1 += d.mix1; h2 += d.mix2;
                    }
                    if (v.c1 >= L && v.c1 < H) {
                        unres = true;
                        if (v.crossL) { nUn++; if (!haveTL || v.tL < minTLun) { minTLun = v.tL; haveTL = true; } }
                        else fl |= PF_NEGINF;
                    }
                    st |= v.have;
                }
                s_dst[i][r] = (uint16_t)st;
            }
            if (__any_sync(0xffffffffu, unres) && (t & 31) == 0) s_unres[i] = 1;
        }
        __syncthreads();
        if (t < n && s_unres[t]) note_unresolved(a, tile, sd[t].slot);   // a 256-receiver block lies inside one tile
    }
    if (t == 0) {
        if ((blockIdx.x * GEN_THREADS) % TILE_R == 0) a.part.flag[(size_t)chunk * a.part.n_tiles + tile] = 1;
        if (blockIdx.x == 0) {
            ChunkAcc c;
            c.nLH = 0; c.tpUn = 0; c.fl = 0; c.minTH = T32_NONE; c.minTLun = T32_NONE; c.pad_ = 0; c.h1 = 0; c.h2 = 0;
            a.part.chunk[chunk] = c;
        }
    }
    if (r < (int64_t)a.Rpad) {
        const size_t p = (size_t)chunk * a.Rpad + (size_t)r;
        a.part.cnt[p] = make_uint4(nL | (nH << 16), tp | (nUn << 16), fl, 0u);
        a.part.minTH[p] = minTH;
        a.part.minTLun[p] = minTLun;
        a.part.h1[p] = h1;
        a.part.h2[p] = h2;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// finalize 1: combine the chunk partials of every receiver, classify, keep the scalars
// ------------------------------------------------------------------------------------------------------------------
struct FinArgs {
    int64_t R;
    size_t Rpad;
    int n_chunks;
    Partials part;
    DeliveryDev dl;
    int any_down_uniform;        // uniform delivery: the batch holds a valid DOWN cell
    int uniform;
    int32_t* n_pre;
    uint32_t* rflags;
    uint64_t* pend_h1;
    uint64_t* pend_h2;
    int32_t* pend_cnt;
    uint64_t* out_h1;
    uint64_t* out_h2;
    int32_t* out_len;
    uint32_t* mx_fl;
    uint64_t* mx_a;
    uint64_t* mx_cand;
    uint64_t* mx_emax;
    uint64_t* mx_p1;
    uint64_t* mx_p2;
    int32_t* mx_pc;
    BatchCounts* bc;
};

__global__ void k_finalize1(const FinArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    a.mx_fl[r] = 0;
    uint32_t flags = a.rflags[r] & ~(RF_ANN_NOW | RF_K3 | RF_ACTIVE);
    a.out_h1[r] = 0; a.out_h2[r] = 0; a.out_len[r] = 0;
    const bool active = !(flags & RF_ANNOUNCED) && !((a.dl.flags & RAPID_DELIVERY_BLOCKED) && a.dl.blocked[r]);
    if (!active) { a.rflags[r] = flags; return; }
    flags |= RF_ACTIVE;
    uint32_t nL = 0, nH = 0, tp = 0, nUn = 0, fl = 0;
    uint64_t minTH = T64_NONE, minTLun = T64_NONE, h1 = 0, h2 = 0;
    bool haveTH = false, haveTL = false;
    const int tile = (int)(r / TILE_R);
    for (int c = 0; c < a.n_chunks; ++c) {
        {   // fresh subjects of the chunk: the same for every active receiver
            const ChunkAcc k = a.part.chunk[c];
            const uint32_t cH = k.nLH >> 16, cUn = k.tpUn >> 16;
            nL += k.nLH & 0xFFFFu; nH += cH; nUn += cUn; fl |= k.fl;
            if (cH && (!haveTH || (uint64_t)k.minTH < minTH)) { minTH = k.minTH; haveTH = true; }
            if (cUn && (!haveTL || (uint64_t)k.minTLun < minTLun)) { minTLun = k.minTLun; haveTL = true; }
            h1 += k.h1; h2 += k.h2;
        }
        if (!a.part.flag[(size_t)c * a.part.n_tiles + tile]) continue;
        const size_t p = (size_t)c * a.Rpad + (size_t)r;
        const uint4 q = a.part.cnt[p];
        const uint32_t cH = q.x >> 16, cUn = q.y >> 16;
        nL += q.x & 0xFFFFu; nH += cH; tp += q.y & 0xFFFFu; nUn += cUn; fl |= q.z;
        if (cH) { const uint64_t v = a.part.minTH[p]; if (!haveTH || v < minTH) { minTH = v; haveTH = true; } }
        if (cUn) { const uint64_t v = a.part.minTLun[p]; if (!haveTL || v < minTLun) { minTLun = v; haveTL = true; } }
        h1 += a.part.h1[p]; h2 += a.part.h2[p];
    }
    if (a.uniform ? a.any_down_uniform : (fl & PF_SEEN)) flags |= RF_SEEN_DOWN;
    const int32_t npre_old = a.n_pre[r];
    const int32_t npre_new = npre_old + (int32_t)nL - (int32_t)nH;
    uint64_t ph1 = a.pend_h1[r] + h1, ph2 = a.pend_h2[r] + h2;
    int32_t pc = a.pend_cnt[r] + (int32_t)nH;
    const int32_t untouched_pre = npre_old - (int32_t)tp;
    if (nH > 0 && npre_new == 0) {
        // EMIT_ALL: the last H-crossing of the batch leaves updatesInProgress == 0, so every subject at >= H has
        // left in some proposal of this batch (MultiNodeCutDetector.java:110-121); the union is what is announced.
        a.out_h1[r] = ph1; a.out_h2[r] = ph2; a.out_len[r] = pc;
        ph1 = 0; ph2 = 0; pc = 0;
        flags |= RF_ANNOUNCED | RF_ANN_NOW | RF_RULE_GE_H;
    } else if (nH > 0 && !(untouched_pre > 0 || (fl & PF_NEGINF) || (haveTL && minTLun < minTH))) {
        // MIXED: some proposals may have been emitted before the unresolved subjects entered the band
        // (haveTL holds here: npre_new > 0 and every unresolved subject entered the band inside the batch)
        atomicAdd(&a.bc->n_mixed, 1);
        a.mx_fl[r] = MX_ON;
        a.mx_a[r] = minTLun; a.mx_cand[r] = T64_NONE; a.mx_emax[r] = 0;
        a.mx_p1[r] = a.pend_h1[r]; a.mx_p2[r] = a.pend_h2[r]; a.mx_pc[r] = a.pend_cnt[r];
    }
    if (npre_new > 0 && (flags & RF_SEEN_DOWN)) flags |= RF_K3;
    a.n_pre[r] = npre_new;
    a.pend_h1[r] = ph1; a.pend_h2[r] = ph2; a.pend_cnt[r] = pc;
    a.rflags[r] = flags;
}

// ------------------------------------------------------------------------------------------------------------------
// MIXED receivers: exact interval analysis as data-parallel passes over (batch subjects x receivers).
//
// For one receiver the batch's subjects give intervals [t_L, t_H) (t_L = "before the batch" if the subject started
// inside the band, t_H = never if it does not reach H).  A proposal is emitted at t_H(s) iff no other interval covers
// it.  Let a = start of the connected component of intervals that contains the never-closing ones: everything closing
// after `a` is covered, and e* = the latest closing moment before `a` is the last explicit emission; what left is
// {t_H <= e*} plus whatever was pending before the batch.  `a` is found as a fixpoint: a <- min{t_L : t_H > a},
// starting from the earliest never-closing start (k_finalize1).  One FIX pass recomputes every flagged receiver's
// intervals from the PRE-batch rows (nothing is stored per (subject, receiver)); uniform delivery typically makes
// every receiver MIXED in the same way, so the passes are shaped like the apply kernels, not like a rare fallback.
// ------------------------------------------------------------------------------------------------------------------
struct MixArgs {
    ApplyArgs ap;
    uint32_t* mx_fl;
    uint64_t* mx_a;
    uint64_t* mx_cand;
    uint64_t* mx_emax;
    uint64_t* estar;
    unsigned long long* mx_e1;
    unsigned long long* mx_e2;
    int32_t* mx_ec;
    int uniform;
};

struct IVisit { int c0; bool crossL, crossH; uint64_t tL, tH; };

__device__ __forceinline__ IVisit interval_visit(const ApplyArgs& a, int uniform, uint32_t ur, const SubjDesc& d, const SubjWalk* w,
                                                 int64_t r, uint64_t rs) {
    IVisit o;
    if (uniform) {
        const Visit v = visit_uniform(ur, d, *w, a.L, a.H);
        o.c0 = v.c0; o.crossL = v.crossL; o.crossH = v.crossH; o.tL = v.tL; o.tH = v.tH;
    } else {
        const GVisit v = visit_generic(ur, d, a.sidx, a.s_ring, a.s_status, a.dl, r, rs, a.L, a.H);
        o.c0 = v.c0; o.crossL = v.crossL; o.crossH = v.crossH; o.tL = v.tL; o.tH = v.tH;
    }
    return o;
}

template <int MODE>      // 0: FIX pass (next candidate for `a`, e* candidate)   1: SUM pass (fingerprint of {t_H <= e*})
__global__ void __launch_bounds__(GEN_THREADS) k_mixed_pass(const MixArgs m) {
    __shared__ SubjDesc sd[STAGE];
    __shared__ SubjWalk sw[STAGE];
    __shared__ const uint16_t* s_old[STAGE];
    const ApplyArgs& a = m.ap;
    const int t = threadIdx.x, chunk = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.x * GEN_THREADS + t;
    const uint32_t RM = (1u << a.K) - 1u;
    const int L = a.L, H = a.H;
    const uint32_t fl = r < a.R ? m.mx_fl[r] : 0u;
    const bool on = MODE == 0 ? ((fl & MX_ON) && !(fl & MX_DONE)) : ((fl & MX_DONE) && (fl & MX_HAS_E));
    if (!__syncthreads_or(on ? 1 : 0)) return;
    const uint64_t ref = on ? (MODE == 0 ? m.mx_a[r] : m.estar[r]) : 0ull;
    const uint64_t rs = splitmix64(a.dl.perm_seed + (uint64_t)(a.rbegin + r));
    const int s0 = chunk * a.chunk, s1 = min(a.Sb, s0 + a.chunk);
    uint64_t cand = T64_NONE, emax = 0, h1 = 0, h2 = 0;
    bool neg = false, has_e = false;
    int cnt = 0;
    for (int base = s0; base < s1; base += STAGE) {
        const int n = min(STAGE, s1 - base);
        __syncthreads();
        if (t < n) {
            const SubjDesc d = a.desc[base + t];
            sd[t] = d;
            const bool fresh = d.slot >= a.S_before;
            s_old[t] = fresh ? nullptr : a.masks + ((size_t)d.slot * 2 + a.cur[d.slot]) * a.Rpad;   // pre-batch row (not flipped yet)
            if (!fresh && m.uniform) sw[t] = a.walk[base + t];
        }
        __syncthreads();
        if (!on) continue;
        for (int i = 0; i < n; ++i) {
            const SubjDesc& d = sd[i];
            const uint32_t st = s_old[i] ? s_old[i][r] : 0u;
            const IVisit v = interval_visit(a, m.uniform, st & RM, d, &sw[i], r, rs);
            if (MODE == 0) {
                if (!v.crossH) continue;                                  // never closes, or never in the band
                const bool starts_in = v.c0 >= L && v.c0 < H;
                if (v.tH > ref) { if (starts_in) neg = true; else if (v.tL < cand) cand = v.tL; }
                else if (!has_e || v.tH > emax) { emax = v.tH; has_e = true; }
            } else {
                if (v.crossH && v.tH <= ref) { h1 += d.mix1; h2 += d.mix2; ++cnt; }
            }
        }
    }
    if (!on) return;
    if (MODE == 0) {
        if (neg) atomicOr(&m.mx_fl[r], MX_NEG);
        if (cand != T64_NONE) atomicMin((unsigned long long*)&m.mx_cand[r], (unsigned long long)cand);
        if (has_e) { atomicMax((unsigned long long*)&m.mx_emax[r], (unsigned long long)emax); atomicOr(&m.mx_fl[r], MX_HAS_E); }
    } else if (cnt) {
        atomicAdd(&m.mx_e1[r], (unsigned long long)h1); atomicAdd(&m.mx_e2[r], (unsigned long long)h2); atomicAdd(&m.mx_ec[r], cnt);
    }
}

__global__ void k_mixed_update(int64_t R, uint32_t* __restrict__ mx_fl, uint64_t* __restrict__ mx_a, uint64_t* __restrict__ mx_cand,
                               uint64_t* __restrict__ mx_emax, uint64_t* __restrict__ estar, int32_t* __restrict__ changed) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const uint32_t fl = mx_fl[r];
    if (!(fl & MX_ON) || (fl & MX_DONE)) return;
    if (fl & MX_NEG) { mx_fl[r] = 0; return; }              // covered since before the batch: nothing was emitted
    const uint64_t c = mx_cand[r];
    if (c < mx_a[r]) {                                       // the component grows leftwards: another pass
        mx_a[r] = c; mx_cand[r] = T64_NONE; mx_emax[r] = 0; mx_fl[r] = fl & ~MX_HAS_E;
        atomicAdd(changed, 1);
    } else if (fl & MX_HAS_E) {
        estar[r] = mx_emax[r];
        mx_fl[r] = fl | MX_DONE;
    } else {
        mx_fl[r] = 0;                                        // no closing moment before the component: nothing emitted
    }
}

struct MixCommitArgs {
    int64_t R;
    const uint32_t* mx_fl;
    const uint64_t* mx_p1;
    const uint64_t* mx_p2;
    const int32_t* mx_pc;
    unsigned long long* mx_e1;
    unsigned long long* mx_e2;
    int32_t* mx_ec;
    uint32_t* rflags;
    uint64_t* pend_h1;
    uint64_t* pend_h2;
    int32_t* pend_cnt;
    uint64_t* out_h1;
    uint64_t* out_h2;
    int32_t* out_len;
};

__global__ void k_mixed_commit(const MixCommitArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const uint32_t fl = a.mx_fl[r];
    if (!((fl & MX_DONE) && (fl & MX_HAS_E))) return;
    // what left explicitly = everything pending before the batch + the batch subjects that closed by e*
    const uint64_t o1 = a.mx_p1[r] + a.mx_e1[r], o2 = a.mx_p2[r] + a.mx_e2[r];
    const int32_t oc = a.mx_pc[r] + a.mx_ec[r];
    a.out_h1[r] = o1; a.out_h2[r] = o2; a.out_len[r] = oc;
    a.pend_h1[r] -= o1; a.pend_h2[r] -= o2; a.pend_cnt[r] -= oc;
    a.mx_e1[r] = 0; a.mx_e2[r] = 0; a.mx_ec[r] = 0;
    a.rflags[r] = (a.rflags[r] | RF_ANNOUNCED | RF_ANN_NOW | RF_MIXED_EMIT) & ~RF_RULE_GE_H;
}

// Did subject slot `s` leave in an explicit proposal of the batch in flight, for an RF_MIXED_EMIT receiver?  Rows have been
// flipped: the pre-batch row is the non-current one.
struct EmitCtx {
    ApplyArgs ap;
    const int32_t* touch;
    const int32_t* batch_index;
    int32_t serial;
    const uint64_t* estar;
    int uniform;
};

__device__ __forceinline__ bool emitted_in_batch(const EmitCtx& e, int32_t s, int64_t r, uint32_t w_new, uint64_t rs) {
    const ApplyArgs& a = e.ap;
    const uint32_t RM = (1u << a.K) - 1u;
    // untouched by the batch: it left (with the first explicit proposal) iff it was pending, i.e. at >= H and NOT raised there by
    // this batch's invalidation pass (bit 14, cleared by k_inval_unmark once the batch is done)
    if (e.touch[s] != e.serial) return __popc(w_new & RM) >= a.H && !(w_new & CD_BIT_CALL);
    const int b = e.batch_index[s];
    const SubjDesc d = a.desc[b];
    const uint32_t old = s >= a.S_before ? 0u : (a.masks + ((size_t)s * 2 + (a.cur[s] ^ 1)) * a.Rpad)[r];
    if (__popc(old & RM) >= a.H) return true;                              // pending before the batch
    const SubjWalk* w = e.uniform ? &a.walk[b] : nullptr;
    SubjWalk wl;
    if (e.uniform) { wl = *w; }
    const IVisit v = interval_visit(a, e.uniform, old & RM, d, &wl, r, rs);
    return v.crossH && v.tH <= e.estar[r];
}

// RF_MIXED_EMIT receivers whose invalidation pass did not emit announce only the explicit part: give it bit 15 so that
// rapid_cd_get_proposal can list it later (the pre-batch rows are gone by then).
__global__ void __launch_bounds__(GEN_THREADS) k_mixed_mark(const EmitCtx e, int32_t S, int slots_per_block, const uint32_t* __restrict__ rflags) {
    const ApplyArgs& a = e.ap;
    const int64_t r = (int64_t)blockIdx.x * GEN_THREADS + threadIdx.x;
    if (r >= a.R) return;
    const uint32_t f = rflags[r];
    if (!(f & RF_MIXED_EMIT) || !(f & RF_ANN_NOW) || (f & RF_RULE_GE_H)) return;
    const uint64_t rs = splitmix64(a.dl.perm_seed + (uint64_t)(a.rbegin + r));
    const int32_t s0 = blockIdx.y * slots_per_block, s1 = min(S, s0 + slots_per_block);
    for (int32_t s = s0; s < s1; ++s) {
        uint16_t* p = a.masks + ((size_t)s * 2 + a.cur[s]) * a.Rpad + r;
        const uint32_t w = *p;
        if (!(w & CD_BIT_EMIT) && emitted_in_batch(e, s, r, w, rs)) *p = (uint16_t)(w | CD_BIT_EMIT);
    }
}

__global__ void k_flip(int Sb, const SubjDesc* __restrict__ desc, uint8_t* __restrict__ cur) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < Sb) cur[desc[b].slot] ^= 1;
}

// ------------------------------------------------------------------------------------------------------------------
// invalidateFailingEdges (MultiNodeCutDetector.java:137-164) over the (tile, subject) pairs known to hold an
// unstable subject: implicit reports from observers that are themselves in proposal U preProposal.
// ------------------------------------------------------------------------------------------------------------------
struct InvArgs {
    uint16_t* masks;
    const uint8_t* cur;
    size_t Rpad;
    int K, H, L;
    int64_t R;
    const uint32_t* rflags;
    const int2* pre_pairs;
    const int32_t* pre_count;
    int32_t pre_cap;
    const int32_t* slot_subject;
    const int32_t* slot_of;
    const int32_t* obs;
    int32_t* k3_res;
    unsigned long long* k3_h1;
    unsigned long long* k3_h2;
    int mixed;                 // some receiver announced through the interval analysis in this batch
    EmitCtx ec;                // (valid when mixed)
};

__global__ void __launch_bounds__(256) k_inval_pairs(const InvArgs a) {
    __shared__ int32_t so[MAXK];
    const uint32_t RM = (1u << a.K) - 1u;
    const int n = min(*a.pre_count, a.pre_cap);
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int2 pr = a.pre_pairs[e];
        const int32_t slot = pr.y, subject = a.slot_subject[slot];
        __syncthreads();
        if (threadIdx.x < a.K) {
            const int32_t o = a.obs[(size_t)subject * a.K + threadIdx.x];
            so[threadIdx.x] = o < 0 ? -1 : a.slot_of[o];
        }
        __syncthreads();
        uint16_t* row = a.masks + ((size_t)slot * 2 + a.cur[slot]) * a.Rpad;
        for (int q = 0; q < TILE_R / 256; ++q) {
            const int64_t r = (int64_t)pr.x * TILE_R + q * 256 + threadIdx.x;
            if (r >= a.R) continue;
            const uint32_t rf = a.rflags[r];
            if (!(rf & RF_K3)) continue;
            const bool mx = a.mixed && (rf & RF_MIXED_EMIT);
            const uint64_t rs = mx ? splitmix64(a.ec.ap.dl.perm_seed + (uint64_t)(a.ec.ap.rbegin + r)) : 0ull;
            const uint32_t w = row[r];
            const int c = __popc(w & RM);
            if (c < a.L || c >= a.H) continue;                      // not in this receiver's preProposal
            uint32_t implicit = 0;
            for (int k = 0; k < a.K; ++k) {
                if ((w >> k) & 1u) continue;
                const int32_t s2 = so[k];
                if (s2 < 0) continue;
                const uint32_t wo = (a.masks + ((size_t)s2 * 2 + a.cur[s2]) * a.Rpad)[r];
                if ((wo & CD_BIT_EMIT) || __popc(wo & RM) < a.L) continue;            // observer not in proposal U preProposal
                // a receiver that already announced explicit proposals in this batch: those subjects left `proposal`.
                // (bit 14 = raised to >= H by this very pass, i.e. it was in the band at entry, not pending)
                if (mx && emitted_in_batch(a.ec, s2, r, wo, rs)) continue;
                implicit |= 1u << k;
            }
            if (!implicit) continue;
            uint32_t nw = w | implicit;
            const bool raised = __popc(nw & RM) >= a.H;
            if (raised && a.mixed) nw |= CD_BIT_CALL;                // transient marker, cleared by k_inval_unmark
            row[r] = (uint16_t)nw;
            if (raised) {                                            // moved preProposal -> proposal
                atomicAdd(&a.k3_res[r], 1);
                atomicAdd(&a.k3_h1[r], (unsigned long long)fp_mix1(subject));
                atomicAdd(&a.k3_h2[r], (unsigned long long)fp_mix2(subject));
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_inval_unmark(const InvArgs a) {
    const int n = min(*a.pre_count, a.pre_cap);
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int2 pr = a.pre_pairs[e];
        uint16_t* row = a.masks + ((size_t)pr.y * 2 + a.cur[pr.y]) * a.Rpad;
        for (int q = 0; q < TILE_R / 256; ++q) {
            const int64_t r = (int64_t)pr.x * TILE_R + q * 256 + threadIdx.x;
            if (r >= a.R) continue;
            const uint32_t w = row[r];
            if (w & CD_BIT_CALL) row[r] = (uint16_t)(w & ~CD_BIT_CALL);
        }
    }
}

struct Fin2Args {
    int64_t R;
    int32_t* n_pre;
    uint32_t* rflags;
    uint64_t* pend_h1;
    uint64_t* pend_h2;
    int32_t* pend_cnt;
    uint64_t* out_h1;
    uint64_t* out_h2;
    int32_t* out_len;
    uint8_t* out_ann;
    int32_t* k3_res;
    unsigned long long* k3_h1;
    unsigned long long* k3_h2;
    BatchCounts* bc;
};

__global__ void k_finalize2(const Fin2Args a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    uint32_t flags = a.rflags[r];
    if (flags & RF_K3) {
        const int32_t res = a.k3_res[r];
        if (res > 0) {
            const int32_t npre = a.n_pre[r] - res;
            uint64_t ph1 = a.pend_h1[r] + a.k3_h1[r], ph2 = a.pend_h2[r] + a.k3_h2[r];
            int32_t pc = a.pend_cnt[r] + res;
            if (npre == 0) {
                // the last unstable subject resolved inside invalidateFailingEdges: proposal (all of it) is emitted
                a.out_h1[r] += ph1; a.out_h2[r] += ph2; a.out_len[r] += pc;
                ph1 = 0; ph2 = 0; pc = 0;
                flags |= RF_ANNOUNCED | RF_ANN_NOW | RF_RULE_GE_H;
            }
            a.n_pre[r] = npre;
            a.pend_h1[r] = ph1; a.pend_h2[r] = ph2; a.pend_cnt[r] = pc;
            a.k3_res[r] = 0; a.k3_h1[r] = 0; a.k3_h2[r] = 0;
        }
    }
    flags &= ~RF_K3;
    if ((flags & RF_MIXED_EMIT) && (flags & RF_ANN_NOW) && !(flags & RF_RULE_GE_H)) atomicAdd(&a.bc->n_inval, 1);   // needs bit-15 marks
    a.rflags[r] = flags;
    a.out_ann[r] = (flags & RF_ANNOUNCED) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------------------------
static Bucketed* state(CD* cd) {
    if (!cd->bucketed_state) cd->bucketed_state = new Bucketed();
    return static_cast<Bucketed*>(cd->bucketed_state);
}

void bucketed_destroy(CD* cd) {
    if (cd->bucketed_state) { delete static_cast<Bucketed*>(cd->bucketed_state); cd->bucketed_state = nullptr; }
}

int32_t bucketed_prep_buffers(CD* cd, int64_t A, PrepOut* po) {
    Bucketed* b = state(cd);
    const size_t a = (size_t)std::max<int64_t>(A, 1);
    RAPID_CHECK(b->desc.reserve(a)); RAPID_CHECK(b->walk.reserve(a));
    RAPID_CHECK(b->sidx.reserve(a)); RAPID_CHECK(b->s_ring.reserve(a)); RAPID_CHECK(b->s_status.reserve(a));
    const size_t slots = (size_t)std::min<int64_t>((int64_t)cd->S + A, std::max<int64_t>(cd->ntot_cap, 1));
    RAPID_CHECK(b->batch_index.reserve(slots));
    RAPID_CHECK(b->seg_pos.reserve(slots));
    if (slots > b->seg_cnt.cap) {
        RAPID_CHECK(b->seg_cnt.reserve(slots));
        RAPID_CUDA(cudaMemsetAsync(b->seg_cnt.p, 0, b->seg_cnt.cap * sizeof(int32_t), cd->stream));
    }
    po->desc = b->desc.p; po->walk = b->walk.p; po->sidx = b->sidx.p; po->s_ring = b->s_ring.p; po->s_status = b->s_status.p;
    po->batch_index = b->batch_index.p; po->seg_cnt = b->seg_cnt.p; po->seg_pos = b->seg_pos.p;
    return RAPID_OK;
}

int32_t bucketed_pair_count(const CD* cd) {
    if (!cd->bucketed_state) return 0;
    const Bucketed* b = static_cast<const Bucketed*>(cd->bucketed_state);
    int32_t n = 0;
    if (b->pre_count.p) cudaMemcpy(&n, b->pre_count.p, sizeof(n), cudaMemcpyDeviceToHost);
    return n;
}

__global__ void k_clear_pairs(const int2* __restrict__ pairs, int32_t* __restrict__ count, int32_t cap, int32_t* __restrict__ in_list,
                              int n_tiles) {
    const int n = min(*count, cap);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int2 pr = pairs[e];
        in_list[(size_t)pr.y * n_tiles + pr.x] = 0;
    }
}
__global__ void k_zero_i32(int32_t* p) { *p = 0; }

int32_t bucketed_clear(CD* cd) {
    if (!cd->bucketed) return RAPID_OK;
    Bucketed* b = state(cd);
    b->n_tiles = (int)(cd->Rpad / TILE_R);
    cudaStream_t s = cd->stream;
    if (!b->pre_count.p) {
        RAPID_CHECK(b->pre_count.reserve(1));
        RAPID_CUDA(cudaMemsetAsync(b->pre_count.p, 0, sizeof(int32_t), s));
    }
    if (b->in_list.p && b->in_list_slots) {
        // undo only the (tile, subject) pairs that were noted: O(#pairs), not O(slots x tiles)
        const int32_t cap = (int32_t)std::min<size_t>(b->in_list_slots * (size_t)b->n_tiles, 0x7fffffff);
        k_clear_pairs<<<64, 256, 0, s>>>(b->pre_pairs.p, b->pre_count.p, cap, b->in_list.p, b->n_tiles);
        k_zero_i32<<<1, 1, 0, s>>>(b->pre_count.p);
        RAPID_KERNEL_CHECK();
    }
    if (!b->k3_res.p) {      // zero once: k_finalize2 leaves them zero after every batch
        RAPID_CHECK(b->k3_res.reserve(cd->Rpad));
        RAPID_CHECK(b->k3_h1.reserve(cd->Rpad));
        RAPID_CHECK(b->k3_h2.reserve(cd->Rpad));
        RAPID_CUDA(cudaMemsetAsync(b->k3_res.p, 0, cd->Rpad * sizeof(int32_t), s));
        RAPID_CUDA(cudaMemsetAsync(b->k3_h1.p, 0, cd->Rpad * sizeof(unsigned long long), s));
        RAPID_CUDA(cudaMemsetAsync(b->k3_h2.p, 0, cd->Rpad * sizeof(unsigned long long), s));
    }
    return RAPID_OK;
}

static int32_t ensure_pre_capacity(CD* cd, Bucketed* b) {
    const size_t slots = cd->S_cap;
    if (slots <= b->in_list_slots) return RAPID_OK;
    const size_t n_old = b->in_list_slots * (size_t)b->n_tiles, n_new = slots * (size_t)b->n_tiles;
    DevBuf<int32_t> nl;
    RAPID_CHECK(nl.reserve(n_new));
    RAPID_CUDA(cudaMemsetAsync(nl.p, 0, n_new * sizeof(int32_t), cd->stream));
    if (n_old) RAPID_CUDA(cudaMemcpyAsync(nl.p, b->in_list.p, n_old * sizeof(int32_t), cudaMemcpyDeviceToDevice, cd->stream));
    DevBuf<int2> np;
    RAPID_CHECK(np.reserve(n_new));
    if (n_old) RAPID_CUDA(cudaMemcpyAsync(np.p, b->pre_pairs.p, n_old * sizeof(int2), cudaMemcpyDeviceToDevice, cd->stream));
    RAPID_CUDA(cudaStreamSynchronize(cd->stream));
    std::swap(b->in_list.p, nl.p); std::swap(b->in_list.cap, nl.cap);
    std::swap(b->pre_pairs.p, np.p); std::swap(b->pre_pairs.cap, np.cap);
    b->in_list_slots = slots;
    return RAPID_OK;
}

int32_t bucketed_apply(CD* cd, int64_t A, const DeliveryDev& dl, const BatchCounts& bc) {
    Bucketed* b = state(cd);
    cudaStream_t s = cd->stream;
    const int TB = 256;
    const int Sb = bc.n_batch_subj;
    const int32_t n_valid = bc.n_valid;
    const bool uniform = !(dl.flags & (RAPID_DELIVERY_BITMAP | RAPID_DELIVERY_PERMUTED));
    b->n_tiles = (int)(cd->Rpad / TILE_R);
    RAPID_CHECK(ensure_pre_capacity(cd, b));

    int n_chunks = 1, chunk = std::max(Sb, 1);
    if (Sb > 0) {
        // ---- grid: tiles x subject chunks, a few waves of 148 SMs -----------------------------------------------------
        const int rblocks = uniform ? b->n_tiles : (int)(cd->Rpad / GEN_THREADS);
        // Pick the number of subject chunks so that (tiles x chunks) blocks fill whole waves of resident blocks:
        // a bandwidth-bound grid whose last wave is mostly empty pays almost a full wave for it.  More chunks also
        // mean more per-receiver partials (48 B each), so cap them at ~8 % of the mask traffic.
        if (b->slots_uniform == 0) {
            int dev = 0, sms = 148, per_u = 8, per_g = 4;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_u, k_apply_uniform, UNI_THREADS, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_g, k_apply_generic, GEN_THREADS, 0);
            b->slots_uniform = sms * std::max(per_u, 1);
            b->slots_generic = sms * std::max(per_g, 1);
        }
        const int slots = uniform ? b->slots_uniform : b->slots_generic;
        const int cmax = std::max(1, std::min(Sb, std::max(Sb / 300, ceil_div(slots, rblocks))));
        double best = -1.0;
        for (int c = 1; c <= cmax; ++c) {
            const int ch = ceil_div(Sb, c), cc = ceil_div(Sb, ch);
            // cost model fitted to measurements on B200 (977 tiles: 1 chunk 0.92, 2 chunks 0.94 of peak; 489 tiles x 3
            // chunks = 1.24 waves: 0.79)
            const double w = (double)rblocks * cc / slots;
            const double f = w - std::floor(w);
            double eff;
            if (w <= 1.0) eff = 0.85 + 0.15 * w;                                   // one partial wave: a little less occupancy
            else eff = w / (std::floor(w) + (f > 0 ? std::max(f, 0.6) : 0.0));     // tail wave: needs ~60 % of the slots to saturate HBM
            eff -= 0.001 * cc;                                                     // per-chunk prologue (fresh subjects cost no partials)
            if (eff > best) { best = eff; n_chunks = cc; chunk = ch; }
        }
        if (const char* ov = getenv("RAPID_B200_CHUNKS")) {          // tuning aid: force the number of subject chunks
            const int c = std::max(1, std::min(Sb, atoi(ov)));
            chunk = ceil_div(Sb, c);
            n_chunks = ceil_div(Sb, chunk);
        }
    }
    const size_t pn = (size_t)n_chunks * cd->Rpad;
    RAPID_CHECK(b->p_cnt.reserve(pn)); RAPID_CHECK(b->p_minTH.reserve(pn)); RAPID_CHECK(b->p_minTLun.reserve(pn));
    RAPID_CHECK(b->p_h1.reserve(pn)); RAPID_CHECK(b->p_h2.reserve(pn));
    RAPID_CHECK(b->mx_fl.reserve(cd->Rpad)); RAPID_CHECK(b->mx_a.reserve(cd->Rpad)); RAPID_CHECK(b->mx_cand.reserve(cd->Rpad));
    RAPID_CHECK(b->mx_emax.reserve(cd->Rpad)); RAPID_CHECK(b->mx_p1.reserve(cd->Rpad)); RAPID_CHECK(b->mx_p2.reserve(cd->Rpad));
    RAPID_CHECK(b->mx_pc.reserve(cd->Rpad)); RAPID_CHECK(b->estar.reserve(cd->Rpad)); RAPID_CHECK(b->mx_changed.reserve(1));
    if (!b->mx_e1.p) {
        RAPID_CHECK(b->mx_e1.reserve(cd->Rpad)); RAPID_CHECK(b->mx_e2.reserve(cd->Rpad)); RAPID_CHECK(b->mx_ec.reserve(cd->Rpad));
        RAPID_CUDA(cudaMemsetAsync(b->mx_e1.p, 0, cd->Rpad * sizeof(unsigned long long), s));
        RAPID_CUDA(cudaMemsetAsync(b->mx_e2.p, 0, cd->Rpad * sizeof(unsigned long long), s));
        RAPID_CUDA(cudaMemsetAsync(b->mx_ec.p, 0, cd->Rpad * sizeof(int32_t), s));
    }
    RAPID_CHECK(b->p_flag.reserve((size_t)n_chunks * std::max(b->n_tiles, 1)));
    RAPID_CHECK(b->p_chunk.reserve((size_t)n_chunks));
    Partials part{b->p_cnt.p, b->p_minTH.p, b->p_minTLun.p, b->p_h1.p, b->p_h2.p, b->p_flag.p, b->p_chunk.p, b->n_tiles};

    ApplyArgs ap;
    ap.masks = cd->masks.p; ap.cur = cd->cur.p; ap.Rpad = cd->Rpad;
    ap.K = cd->K; ap.H = cd->H; ap.L = cd->L; ap.R = cd->R; ap.rbegin = cd->rbegin;
    ap.rflags = cd->rflags.p; ap.dl = dl; ap.Sb = Sb; ap.chunk = chunk;
    ap.desc = b->desc.p; ap.walk = b->walk.p; ap.slot_subject = cd->slot_subject.p;
    ap.sidx = b->sidx.p; ap.s_ring = b->s_ring.p; ap.s_status = b->s_status.p;
    ap.part = part; ap.n_tiles = b->n_tiles; ap.in_list = b->in_list.p; ap.pre_pairs = b->pre_pairs.p;
    ap.S_before = cd->S_before;
    ap.pre_count = b->pre_count.p; ap.pre_cap = (int32_t)std::min<size_t>(b->in_list_slots * (size_t)b->n_tiles, 0x7fffffff);

    RAPID_CUDA(cudaEventRecord(cd->evk0, s));
    if (Sb > 0) {
        if (uniform) {
            dim3 grid((unsigned)b->n_tiles, (unsigned)n_chunks);
            k_apply_uniform<<<grid, UNI_THREADS, 0, s>>>(ap);
            cd->last_path = 2;
        } else {
            dim3 grid((unsigned)(cd->Rpad / GEN_THREADS), (unsigned)n_chunks);
            k_apply_generic<<<grid, GEN_THREADS, 0, s>>>(ap);
            cd->last_path = 3;
        }
        RAPID_KERNEL_CHECK();
        cd->last_launches += 1;
    } else {
        RAPID_CUDA(cudaMemsetAsync(b->p_flag.p, 0, (size_t)n_chunks * std::max(b->n_tiles, 1) * sizeof(int32_t), s));
        RAPID_CUDA(cudaMemsetAsync(b->p_chunk.p, 0, (size_t)n_chunks * sizeof(ChunkAcc), s));
        cd->last_path = uniform ? 2 : 3;
    }
    RAPID_CUDA(cudaEventRecord(cd->evk1, s));

    FinArgs fa;
    fa.R = cd->R; fa.Rpad = cd->Rpad; fa.n_chunks = n_chunks; fa.part = part; fa.dl = dl;
    fa.any_down_uniform = bc.any_down; fa.uniform = uniform ? 1 : 0;
    fa.n_pre = cd->n_pre.p; fa.rflags = cd->rflags.p; fa.pend_h1 = cd->pend_h1.p; fa.pend_h2 = cd->pend_h2.p;
    fa.pend_cnt = cd->pend_cnt.p; fa.out_h1 = cd->out_h1.p; fa.out_h2 = cd->out_h2.p; fa.out_len = cd->out_len.p;
    fa.mx_fl = b->mx_fl.p; fa.mx_a = b->mx_a.p; fa.mx_cand = b->mx_cand.p; fa.mx_emax = b->mx_emax.p;
    fa.mx_p1 = b->mx_p1.p; fa.mx_p2 = b->mx_p2.p; fa.mx_pc = b->mx_pc.p; fa.bc = cd->counts.p;
    const unsigned gr = (unsigned)ceil_div<int64_t>(cd->R, TB);
    k_finalize1<<<gr, TB, 0, s>>>(fa);
    RAPID_KERNEL_CHECK();
    cd->last_launches += 1;

    // ---- receivers needing the exact interval analysis (one small readback tells whether there are any) ----------------
    RAPID_CUDA(cudaMemcpyAsync(cd->h_counts.p, cd->counts.p, sizeof(BatchCounts), cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    const int n_mixed = cd->h_counts.p->n_mixed;
    EmitCtx ectx;
    ectx.ap = ap; ectx.touch = cd->touch.p; ectx.batch_index = b->batch_index.p; ectx.serial = cd->batch_serial;
    ectx.estar = b->estar.p; ectx.uniform = uniform ? 1 : 0;
    if (n_mixed > 0 && Sb > 0) {
        MixArgs ma;
        ma.ap = ap; ma.mx_fl = b->mx_fl.p; ma.mx_a = b->mx_a.p; ma.mx_cand = b->mx_cand.p; ma.mx_emax = b->mx_emax.p;
        ma.estar = b->estar.p; ma.mx_e1 = b->mx_e1.p; ma.mx_e2 = b->mx_e2.p; ma.mx_ec = b->mx_ec.p; ma.uniform = uniform ? 1 : 0;
        const int rblocks = (int)(cd->Rpad / GEN_THREADS);
        int mchunks = std::max(1, std::min(ceil_div(Sb, STAGE), ceil_div(2 * std::max(b->slots_generic, 148), rblocks)));
        ma.ap.chunk = ceil_div(Sb, mchunks);
        mchunks = ceil_div(Sb, ma.ap.chunk);
        dim3 mgrid((unsigned)rblocks, (unsigned)mchunks);
        for (int iter = 0; iter <= Sb + 1; ++iter) {
            RAPID_CUDA(cudaMemsetAsync(b->mx_changed.p, 0, sizeof(int32_t), s));
            k_mixed_pass<0><<<mgrid, GEN_THREADS, 0, s>>>(ma);
            k_mixed_update<<<gr, TB, 0, s>>>(cd->R, b->mx_fl.p, b->mx_a.p, b->mx_cand.p, b->mx_emax.p, b->estar.p, b->mx_changed.p);
            RAPID_KERNEL_CHECK();
            cd->last_launches += 2;
            int32_t changed = 0;
            RAPID_CUDA(cudaMemcpyAsync(&changed, b->mx_changed.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            RAPID_CUDA(cudaStreamSynchronize(s));
            if (changed == 0) break;
        }
        k_mixed_pass<1><<<mgrid, GEN_THREADS, 0, s>>>(ma);
        MixCommitArgs mc;
        mc.R = cd->R; mc.mx_fl = b->mx_fl.p; mc.mx_p1 = b->mx_p1.p; mc.mx_p2 = b->mx_p2.p; mc.mx_pc = b->mx_pc.p;
        mc.mx_e1 = b->mx_e1.p; mc.mx_e2 = b->mx_e2.p; mc.mx_ec = b->mx_ec.p; mc.rflags = cd->rflags.p;
        mc.pend_h1 = cd->pend_h1.p; mc.pend_h2 = cd->pend_h2.p; mc.pend_cnt = cd->pend_cnt.p;
        mc.out_h1 = cd->out_h1.p; mc.out_h2 = cd->out_h2.p; mc.out_len = cd->out_len.p;
        k_mixed_commit<<<gr, TB, 0, s>>>(mc);
        RAPID_KERNEL_CHECK();
        cd->last_launches += 2;
    }
    if (Sb > 0) {
        k_flip<<<(unsigned)ceil_div(Sb, TB), TB, 0, s>>>(Sb, b->desc.p, cd->cur.p);
        RAPID_KERNEL_CHECK();
        cd->last_launches += 1;
        ectx.ap.cur = cd->cur.p;
    }
    InvArgs ia;
    ia.masks = cd->masks.p; ia.cur = cd->cur.p; ia.Rpad = cd->Rpad; ia.K = cd->K; ia.H = cd->H; ia.L = cd->L; ia.R = cd->R;
    ia.rflags = cd->rflags.p; ia.pre_pairs = b->pre_pairs.p; ia.pre_count = b->pre_count.p; ia.pre_cap = ap.pre_cap;
    ia.slot_subject = cd->slot_subject.p; ia.slot_of = cd->slot_of.p; ia.obs = cd->view->obs.p;
    ia.k3_res = b->k3_res.p; ia.k3_h1 = b->k3_h1.p; ia.k3_h2 = b->k3_h2.p;
    ia.mixed = n_mixed > 0 ? 1 : 0; ia.ec = ectx;
    k_inval_pairs<<<148 * 16, 256, 0, s>>>(ia);
    Fin2Args f2;
    f2.R = cd->R; f2.n_pre = cd->n_pre.p; f2.rflags = cd->rflags.p; f2.pend_h1 = cd->pend_h1.p; f2.pend_h2 = cd->pend_h2.p;
    f2.pend_cnt = cd->pend_cnt.p; f2.out_h1 = cd->out_h1.p; f2.out_h2 = cd->out_h2.p; f2.out_len = cd->out_len.p;
    f2.out_ann = cd->out_ann.p; f2.k3_res = b->k3_res.p; f2.k3_h1 = b->k3_h1.p; f2.k3_h2 = b->k3_h2.p; f2.bc = cd->counts.p;
    k_finalize2<<<gr, TB, 0, s>>>(f2);
    RAPID_KERNEL_CHECK();
    cd->last_launches += 2;
    if (n_mixed > 0) {
        RAPID_CUDA(cudaMemcpyAsync(cd->h_counts.p, cd->counts.p, sizeof(BatchCounts), cudaMemcpyDeviceToHost, s));
        RAPID_CUDA(cudaStreamSynchronize(s));
        if (cd->h_counts.p->n_inval > 0 && cd->S > 0) {
            // receivers that announce only the explicit part: persist it as bit 15 while the pre-batch rows still exist
            const int spb = 64;
            dim3 kgrid((unsigned)(cd->Rpad / GEN_THREADS), (unsigned)ceil_div(cd->S, spb));
            k_mixed_mark<<<kgrid, GEN_THREADS, 0, s>>>(ectx, cd->S, spb, cd->rflags.p);
            RAPID_KERNEL_CHECK();
            cd->last_launches += 1;
        }
        k_inval_unmark<<<148 * 4, 256, 0, s>>>(ia);          // only now: the marks above still needed bit 14
        RAPID_KERNEL_CHECK();
        cd->last_launches += 1;
    }
    return RAPID_OK;
}

}  // namespace rapid
