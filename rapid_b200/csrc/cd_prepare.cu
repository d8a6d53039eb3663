// Batch preparation in ONE cooperative kernel launch (grid-wide barriers instead of ~20 small launches):
//
//   P1  filter every alert cell (MembershipService.java:644-675) and find, per not-yet-known subject, its first cell
//   P2  count the new subjects per block                                                   ── grid.sync between phases
//   P3  give them slots in first-appearance order (prefix over blocks + in-block scan)
//   P4  cell -> slot, cells per slot, distinct subjects of the batch
//   P5  prefix over the slots: index of each touched subject in the batch and the start of its segment
//   P6  publish them; reset the per-slot counters
//   P7  scatter the cell indices into their subject's segment
//   P8  per subject: sort its (few) indices back into arrival order, build the descriptor the apply kernels consume
//       (ring mask, first-occurrence ring sequence with moments, the fresh-subject answers, fingerprint mixes)
//
// Phases 5-8 replace a device radix sort: a batch has ~10 cells per subject, so a counting sort by slot plus a tiny
// per-segment sort is all the regrouping needs.  Every block owns a contiguous range of cells (and, in P5/P6, of slots) so
// the prefix sums are exact and deterministic.
#include <cooperative_groups.h>

#include <algorithm>
#include <climits>

#include "cd_internal.cuh"

namespace cg = cooperative_groups;

namespace rapid {

constexpr int PREP_THREADS = 256;

struct PrepArgs {
    int64_t A;
    const int32_t* dst;
    const uint8_t* ring;
    const uint8_t* status;
    const int64_t* cell_cfg;
    int64_t cfg;
    int raw, K, L, H;
    int64_t n_members, n_total;
    int32_t S_old, serial;   // S_old < 0: read it from bc->n_slots (bucketed handles keep it on the device)
    int32_t S_cap;           // slots the handle can hold (bucketed); a batch that needs more is rolled back (bc->overflow)
    WorkList wl;             // invalidation work list (bucketed): refreshed here whenever a subject gets a slot
    const int32_t* obs;      // view: [id][K]
    int32_t* slot_of;
    int32_t* first_idx;
    int32_t* slot_subject;
    int32_t* touch;
    int32_t* cell_slot;
    BatchCounts* bc;
    int32_t* blk_a;          // [grid] scratch
    int32_t* blk_b;          // [grid] scratch
    int regroup;
    PrepOut po;
};

// exclusive scan of one int per thread across the block; returns the thread's offset, *total = block sum
__device__ __forceinline__ int32_t block_excl_scan(int32_t v, int32_t* warp_sums, int32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    __syncthreads();                                   // warp_sums may still be read from a previous call
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int32_t s = lane < (PREP_THREADS >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t x = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += x;
        }
        if (lane < (PREP_THREADS >> 5)) warp_sums[lane] = s;
    }
    __syncthreads();
    *total = warp_sums[(PREP_THREADS >> 5) - 1];
    return (wid ? warp_sums[wid - 1] : 0) + inc - v;
}

__device__ __forceinline__ int32_t block_sum_before(const int32_t* arr, int upto, int32_t* warp_sums) {
    // sum of arr[0 .. upto) computed by the whole block
    int32_t s = 0;
    for (int i = threadIdx.x; i < upto; i += blockDim.x) s += arr[i];
    int32_t total;
    block_excl_scan(s, warp_sums, &total);
    return total;
}

__device__ __forceinline__ bool cell_is_new(const PrepArgs& a, int64_t i) {
    if (a.cell_slot[i] != -2) return false;
    const int32_t d = a.dst[i];
    return a.slot_of[d] < 0 && a.first_idx[d] == (int32_t)i;
}

__global__ void __launch_bounds__(PREP_THREADS) k_prepare(PrepArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int32_t warp_sums[PREP_THREADS / 32];
    const int t = threadIdx.x, G = gridDim.x, bid = blockIdx.x;
    if (a.S_old < 0) a.S_old = a.bc->n_slots;          // read by every block before anyone can change it (P3 is two barriers away)
    if (bid == 0 && t == 0) a.bc->S_before = a.S_old;
    // contiguous range of cells owned by this block (multiple of the block size)
    const int64_t per = ((a.A + G - 1) / G + PREP_THREADS - 1) / PREP_THREADS * PREP_THREADS;
    const int64_t c0 = min(a.A, (int64_t)bid * per), c1 = min(a.A, c0 + per);

    // ---- P1: validity + first occurrence of subjects without a slot -------------------------------------------------------
    for (int64_t i = c0 + t; i < c1; i += PREP_THREADS) {
        const int32_t d = a.dst[i];
        int32_t v = -2;
        if (a.ring[i] >= a.K) { atomicMax(&a.bc->bad_ring, (int32_t)i); v = -1; }
        if (d < 0 || d >= a.n_total) { atomicMax(&a.bc->bad_dst, (int32_t)i); v = -1; }
        if (v == -2 && !a.raw) {
            const bool present = d < a.n_members;                         // isHostPresent
            const int st = a.status[i];
            if (a.cell_cfg && a.cell_cfg[i] != a.cfg) v = -1;              // :653
            else if (st == RAPID_EDGE_UP && present) v = -1;               // :660-665
            else if (st == RAPID_EDGE_DOWN && !present) v = -1;            // :666-671
            else if (st != RAPID_EDGE_UP && st != RAPID_EDGE_DOWN) v = -1;
        }
        a.cell_slot[i] = v;
        if (v == -2) {
            if (a.status[i] == RAPID_EDGE_DOWN) a.bc->any_down = 1;
            if (a.slot_of[d] < 0) atomicMin(&a.first_idx[d], (int32_t)i);
        }
    }
    grid.sync();
    // ---- P2: new subjects per block ----------------------------------------------------------------------------------------
    {
        int32_t cnt = 0;
        for (int64_t i = c0 + t; i < c1; i += PREP_THREADS) cnt += cell_is_new(a, i) ? 1 : 0;
        int32_t total;
        block_excl_scan(cnt, warp_sums, &total);
        if (t == 0) a.blk_a[bid] = total;
    }
    grid.sync();
    // ---- P3: slots in first-appearance order -------------------------------------------------------------------------------
    {
        int32_t base = a.S_old + block_sum_before(a.blk_a, bid, warp_sums);
        for (int64_t i0 = c0; i0 < c1; i0 += PREP_THREADS) {
            const int64_t i = i0 + t;
            const bool isnew = i < c1 && cell_is_new(a, i);
            int32_t total;
            const int32_t off = block_excl_scan(isnew ? 1 : 0, warp_sums, &total);
            if (isnew) {
                const int32_t d = a.dst[i], slot = base + off;
                a.slot_subject[slot] = d;
                a.slot_of[d] = slot;
            }
            base += total;
        }
        if (bid == G - 1 && t == 0) a.bc->need_slots = base;                 // last block ends at S_old + all new subjects
    }
    grid.sync();
    const int32_t S_new = a.bc->need_slots;
    if (S_new > a.S_cap) {
        // More subjects than the handle has rows for: undo the slot assignment of this batch and apply NOTHING (the host
        // grows the handle and replays the batch; an asynchronous caller gets RAPID_ENOMEM at its next synchronisation).
        for (int64_t i = c0 + t; i < c1; i += PREP_THREADS) {
            if (a.cell_slot[i] != -2) continue;
            const int32_t d = a.dst[i];
            if (a.first_idx[d] == (int32_t)i) {
                if (a.slot_of[d] >= a.S_old) a.slot_of[d] = -1;
                a.first_idx[d] = INT_MAX;
            }
        }
        if (bid == 0 && t == 0) a.bc->overflow = 1;
        return;
    }
    if (bid == 0 && t == 0) a.bc->n_slots = S_new;
    // NOTE: first_idx is reset in P4 (a subject's first cell owner resets it), after every block has used it in P3
    // ---- P4: cell -> slot, cells per slot, distinct subjects ---------------------------------------------------------------------
    {
        int32_t nvalid = 0;
        for (int64_t i = c0 + t; i < c1; i += PREP_THREADS) {
            if (a.cell_slot[i] != -2) continue;
            const int32_t d = a.dst[i], slot = a.slot_of[d];
            if (a.first_idx[d] == (int32_t)i) a.first_idx[d] = INT_MAX;
            a.cell_slot[i] = slot;
            ++nvalid;
            if (atomicExch(&a.touch[slot], a.serial) != a.serial) atomicAdd(&a.bc->n_batch_subj, 1);
            if (a.regroup) atomicAdd(&a.po.seg_cnt[slot], 1);
        }
        int32_t total;
        block_excl_scan(nvalid, warp_sums, &total);
        if (t == 0 && total) atomicAdd(&a.bc->n_valid, total);
    }
    if (!a.regroup) return;
    grid.sync();
    // ---- P5: prefix over slots (touched?, cells) --------------------------------------------------------------------------------
    if (a.wl.has_so && S_new > a.S_old) {
        // Some subject got a slot: refresh "which observers of this subject are subjects themselves" for every slot.  Only
        // subjects with such an observer can ever receive an implicit report (MultiNodeCutDetector.java:147-158), so only
        // they go on the invalidation work list.
        for (int32_t sl = bid * PREP_THREADS + t; sl < S_new; sl += G * PREP_THREADS) {
            const int32_t subject = a.slot_subject[sl];
            bool any = false;
            for (int k = 0; k < a.K; ++k) {
                const int32_t o = a.obs[(size_t)subject * a.K + k];
                const int32_t so = o >= 0 ? a.slot_of[o] : -1;
                a.wl.so_tab[(size_t)sl * SO_STRIDE + k] = so;
                if (so >= 0) any = true;
            }
            const bool old = sl < a.S_old && a.wl.has_so[sl];
            a.wl.has_so[sl] = any ? 1 : 0;
            if (any && !old && sl < a.S_old) {
                // an observer of an OLDER subject joined the dictionary: the subject may sit in the unstable band of any tile
                for (int tile = 0; tile < a.wl.n_tiles; ++tile) worklist_note(a.wl, tile, sl);
            }
        }
    }
    const int32_t sper = ((S_new + G - 1) / G + PREP_THREADS - 1) / PREP_THREADS * PREP_THREADS;
    const int32_t q0 = min(S_new, bid * sper), q1 = min(S_new, q0 + sper);
    {
        int32_t nb = 0, nc = 0;
        for (int32_t sl = q0 + t; sl < q1; sl += PREP_THREADS) {
            const int32_t c = a.po.seg_cnt[sl];
            nb += c > 0 ? 1 : 0;
            nc += c;
        }
        int32_t tb, tc;
        block_excl_scan(nb, warp_sums, &tb);
        block_excl_scan(nc, warp_sums, &tc);
        if (t == 0) { a.blk_a[bid] = tb; a.blk_b[bid] = tc; }
    }
    grid.sync();
    // ---- P6: batch index and segment start of every touched slot ---------------------------------------------------------------
    {
        int32_t base_b = block_sum_before(a.blk_a, bid, warp_sums);
        int32_t base_c = block_sum_before(a.blk_b, bid, warp_sums);
        for (int32_t s0 = q0; s0 < q1; s0 += PREP_THREADS) {
            const int32_t sl = s0 + t;
            const int32_t c = sl < q1 ? a.po.seg_cnt[sl] : 0;
            int32_t tb, tc;
            const int32_t ob = block_excl_scan(c > 0 ? 1 : 0, warp_sums, &tb);
            const int32_t oc = block_excl_scan(c, warp_sums, &tc);
            if (c > 0) {
                const int32_t b = base_b + ob;
                a.po.batch_index[sl] = b;
                a.po.seg_pos[sl] = base_c + oc;
                SubjDesc d;
                d.slot = sl; d.bmask = 0; d.nr = 0; d.any_down = 0; d.tLf = 0; d.tHf = 0;
                d.seg_begin = (uint32_t)(base_c + oc); d.seg_len = (uint32_t)c;
                d.mix1 = 0; d.mix2 = 0; d.pad_ = 0;
                a.po.desc[b] = d;
                a.po.seg_cnt[sl] = 0;                                       // all zero again for the next batch
            }
            base_b += tb; base_c += tc;
        }
    }
    grid.sync();
    // ---- P7: scatter the cell indices into their segment (order inside a segment fixed in P8) --------------------------------------
    for (int64_t i = c0 + t; i < c1; i += PREP_THREADS) {
        const int32_t slot = a.cell_slot[i];
        if (slot < 0) continue;
        const int32_t pos = atomicAdd(&a.po.seg_pos[slot], 1);
        a.po.sidx[pos] = (int32_t)i;
    }
    grid.sync();
    // ---- P8: per subject: arrival order, descriptor --------------------------------------------------------------------------------
    const int32_t Sb = a.bc->n_batch_subj;
    for (int32_t b = bid * PREP_THREADS + t; b < Sb; b += G * PREP_THREADS) {
        SubjDesc d = a.po.desc[b];
        int32_t* seg = a.po.sidx + d.seg_begin;
        const int32_t len = (int32_t)d.seg_len;
        // shell sort (segments are ~10 cells; correct for any length)
        for (int32_t gap = len >> 1; gap > 0; gap >>= 1)
            for (int32_t i = gap; i < len; ++i) {
                const int32_t v = seg[i];
                int32_t j = i;
                for (; j >= gap && seg[j - gap] > v; j -= gap) seg[j] = seg[j - gap];
                seg[j] = v;
            }
        SubjWalk w;
        for (int32_t e = 0; e < len; ++e) {
            const int32_t c = seg[e];
            const int k = a.ring[c];
            const uint8_t st = a.status[c];
            a.po.s_ring[d.seg_begin + e] = (uint8_t)k;
            a.po.s_status[d.seg_begin + e] = st;
            if (st == RAPID_EDGE_DOWN) d.any_down = 1;
            if (!((d.bmask >> k) & 1)) {
                d.bmask |= (uint16_t)(1u << k);
                w.ring[d.nr] = (uint8_t)k;
                w.time[d.nr] = (uint32_t)c + 1u;           // moments are 1-based cell indices (0 = "before the batch")
                ++d.nr;
                if (d.nr == a.L) d.tLf = (uint32_t)c + 1u;
                if (d.nr == a.H) d.tHf = (uint32_t)c + 1u;
            }
        }
        for (int q = d.nr; q < 16; ++q) { w.ring[q] = 0; w.time[q] = 0; }
        const int32_t id = a.slot_subject[d.slot];
        d.mix1 = fp_mix1(id);
        d.mix2 = fp_mix2(id);
        a.po.desc[b] = d;
        a.po.walk[b] = w;
    }
}

int32_t prepare_batch(CD* cd, int64_t cfg, int64_t A, const int32_t* dst_dev, const uint8_t* ring_dev, const uint8_t* status_dev,
                      const int64_t* cfg_dev, const PrepOut* po) {
    cudaStream_t s = cd->stream;
    if (cd->prep_grid_max == 0) {
        int dev = 0, sms = 148, per = 4, coop = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_prepare, PREP_THREADS, 0);
        if (!coop) { set_error("device lacks cooperative launch"); return RAPID_ECUDA; }
        cd->prep_grid_max = std::max(1, sms * std::max(per, 1));
    }
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(cd->prep_grid_max, ceil_div<int64_t>(A, PREP_THREADS * 2)));
    RAPID_CHECK(cd->scan_sums.reserve((size_t)2 * cd->prep_grid_max));
    PrepArgs a;
    a.A = A; a.dst = dst_dev; a.ring = ring_dev; a.status = status_dev; a.cell_cfg = cfg_dev; a.cfg = cfg;
    a.raw = cd->raw ? 1 : 0; a.K = cd->K; a.L = cd->L; a.H = cd->H;
    a.n_members = cd->view->n; a.n_total = cd->view->n + cd->view->nj;
    a.S_old = cd->bucketed ? -1 : cd->S; a.serial = ++cd->batch_serial;
    a.S_cap = cd->bucketed ? (int32_t)std::min<size_t>(cd->S_cap, 0x7fffffff) : INT_MAX;
    a.obs = cd->view->obs.p;
    memset(&a.wl, 0, sizeof(a.wl));
    if (po) a.wl = po->wl;
    a.slot_of = cd->slot_of.p; a.first_idx = cd->first_idx.p; a.slot_subject = cd->slot_subject.p; a.touch = cd->touch.p;
    a.cell_slot = cd->cell_slot.p; a.bc = cd->counts.p;
    a.blk_a = cd->scan_sums.p; a.blk_b = cd->scan_sums.p + cd->prep_grid_max;
    a.regroup = po ? 1 : 0;
    if (po) a.po = *po; else memset(&a.po, 0, sizeof(a.po));
    void* args[] = {(void*)&a};
    RAPID_CUDA(cudaLaunchCooperativeKernel((void*)k_prepare, dim3((unsigned)G), dim3(PREP_THREADS), args, 0, s));
    cd->last_launches += 1;
    return RAPID_OK;
}

}  // namespace rapid
