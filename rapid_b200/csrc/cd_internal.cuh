// Internal layout of the cut-detector handle (shared by cd_core.cu and cd_bucketed.cu).
#pragma once

#include <string>
#include <vector>

#include "common.cuh"

namespace rapid {

// ---- the 16-bit state word per (subject slot, receiver) ------------------------------------------
// bits 0..K-1 : ring r has reported the subject  (reportsPerHost[subject].containsKey(r),
//               MultiNodeCutDetector.java:92-101)
// bit 14      : RAW mode only — emitted by the call in flight (returned list of aggregateForProposal)
// bit 15      : subject was emitted in a proposal (it left `proposal` at :118-121)
// preProposal == { L <= popc < H },  proposal == { popc >= H and bit 15 clear }.
#define CD_BIT_CALL 0x4000u
#define CD_BIT_EMIT 0x8000u

// ---- per-receiver flag word ------------------------------------------------------------------------
#define RF_SEEN_DOWN   1u   // seenLinkDownEvents           (MultiNodeCutDetector.java:88-90)
#define RF_ANNOUNCED   2u   // announcedProposal            (MembershipService.java:318, :335)
#define RF_RULE_GE_H   4u   // announced proposal == { popc >= H } (else == { bit 15 })
#define RF_ANN_NOW     8u   // announced by the batch in flight (votes in rapid_fp_tally_cd)

// Device-resident counters of a handle.  Sweep handles: initialised by the host before every batch and read back after the
// prepare kernel.  Bucketed handles: they never leave the device inside a batch — `n_slots` persists from batch to batch, the
// per-batch fields are reset by the batch's last kernel right after it copied the whole record to a snapshot the host reads at
// its next synchronisation point (no host round trip between the kernels of a batch).
struct BatchCounts {
    int32_t n_slots;       // S after slot assignment (persists across batches)
    int32_t n_valid;       // cells that passed the filter
    int32_t n_batch_subj;  // distinct subjects with at least one valid cell in this batch
    int32_t any_down;      // some valid cell has status DOWN
    int32_t bad_ring;      // index of a cell with ring >= K, or -1 (the cell is dropped, the rest of the batch applies)
    int32_t bad_dst;       // index of a cell with dst outside [0, n + joiners), or -1 (dropped likewise)
    int32_t n_mixed;       // bucketed: receivers needing exact interval resolution
    int32_t n_inval;       // bucketed: receivers that announce only the explicit part (bit-15 marks needed)
    int32_t S_before;      // n_slots when the batch started: slots >= S_before are "fresh" (known-zero state, never read);
                           // between batches S_before == n_slots (the prepare kernel takes the slot count from here)
    int32_t overflow;      // the batch needs more subject slots than the handle holds: NOTHING was applied
    int32_t need_slots;    // ... and this many would do
    int32_t n_times;       // PERMUTED delivery: receivers whose classification needed their own crossing moments
    int32_t mixed_iters;   // fixpoint iterations of the interval analysis
    int32_t n_pairs;       // (tile, subject) pairs on the invalidation work list
    int32_t ticket;        // "last block done" counter of the resolve kernel
    int32_t serial;        // serial of the batch this record describes
    int32_t sticky_bad_ring, sticky_bad_dst, sticky_overflow;   // latched until the host collects them (asynchronous batches)
    // ---- a sequence of batches applied in one pass (see cd_bucketed.cu "sequences")
    int32_t seq_last;      // index of the last non-empty batch of the call in flight (0 for a single batch)
    int32_t seq_down;      // 1-based index of the first batch with a valid DOWN cell, INT_MAX if none
    int32_t seq_abort;     // receivers for which the one-pass treatment is not provably exact: NOTHING was committed
    int32_t mx_first;      // interval analysis: lowest flagged receiver (the reference of the uniform-delivery shortcut), INT_MAX if none
    int32_t mx_left;       // ... flagged receivers that differ from it and take the general passes
    int32_t seq_a1, seq_a2; // ... of which: could have emitted before the last batch / an implicit report would fire inside the prefix
};

struct CD {
    const View* view = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    int K = 0, H = 0, L = 0;
    uint32_t mode = 0;
    bool raw = false, bucketed = false;
    int64_t R = 0, rbegin = 0;
    size_t Rpad = 0;
    int nbuf = 1;                     // 2 = double-buffered rows (bucketed handles)
    int32_t S = 0;                    // slots in use (host mirror; bucketed handles: as of the last synchronisation point)
    int32_t S_before = 0;             // S when the batch in flight started (sweep handles)
    size_t S_cap = 0;
    int64_t ntot_cap = 0;             // capacity of slot_of / first_idx (node + joiner ids)

    DevBuf<uint16_t> masks;           // [S_cap][nbuf][Rpad]
    DevBuf<uint8_t> cur;              // [S_cap] which of the nbuf rows is current
    DevBuf<int32_t> slot_of;          // [ntot_cap] id -> slot, -1 if none
    DevBuf<int32_t> first_idx;        // [ntot_cap] scratch (INT_MAX)
    DevBuf<int32_t> slot_subject;     // [ntot_cap] slot -> id
    DevBuf<int32_t> touch;            // [ntot_cap] slot -> serial of the last batch that had a valid cell for it
    int32_t batch_serial = 0;
    int prep_grid_max = 0;            // co-resident blocks of the cooperative prepare kernel

    DevBuf<int32_t> n_pre;            // [R] updatesInProgress
    DevBuf<int32_t> n_prop;           // [R] proposalCount (sweep handles)
    DevBuf<uint32_t> rflags;          // [R]
    DevBuf<uint64_t> pend_h1, pend_h2;  // [R] fingerprint of `proposal` (>=H, not emitted)   (bucketed handles)
    DevBuf<int32_t> pend_cnt;         // [R]
    DevBuf<uint64_t> out_h1, out_h2;  // [R] outputs of the last batch
    DevBuf<int64_t> batch_off;        // rapid_cd_apply_batches: batch boundaries
    DevBuf<int32_t> out_batch;        // [R] ... and the batch in which each receiver announced
    DevBuf<uint64_t> sq_h1, sq_h2;    // [R] outputs of the announcing batch while a sequence is replayed batch by batch
    DevBuf<int32_t> sq_len;
    // ---- RAPID_CD_LOG: the epoch's filtered cells, for the exact replay of ONE receiver (rapid_cd_num_proposals) -------------
    struct LogRec { int64_t c0, c1; uint32_t flags; uint64_t perm_seed; int64_t blocked_off; };
    bool log_on = false, log_complete = true;
    DevBuf<int32_t> log_slot;         // per logged cell: subject slot, -1 = filtered out
    DevBuf<uint8_t> log_ring, log_status, log_blocked;
    size_t log_cells = 0, log_blocked_bytes = 0;
    std::vector<LogRec> log_batches;
    int32_t seq_merged = 0, seq_replayed = 0;   // sequences served in one pass / replayed batch by batch (diagnostics)
    int32_t seq_refused_a1 = 0, seq_refused_a2 = 0;   // receivers that failed either premise in the last refused attempt
    DevBuf<int32_t> out_len;          // [R]
    DevBuf<uint8_t> out_ann;          // [R]

    // batch staging (device)
    DevBuf<int32_t> c_dst;  DevBuf<uint8_t> c_ring, c_status;  DevBuf<int64_t> c_cfg;
    DevBuf<uint8_t> d_blocked;  DevBuf<uint32_t> d_bitmap;
    PinnedBuf<uint8_t> h_stage;  DevBuf<uint8_t> d_stage;   // host-array path: one pinned blob, one H2D copy
    const uint8_t* cur_ring_dev = nullptr;    // ring / status arrays of the batch in flight (device)
    const uint8_t* cur_status_dev = nullptr;
    DevBuf<int32_t> cell_slot;        // [A] slot or -1
    DevBuf<int32_t> scan_tmp;         // [A]
    DevBuf<int32_t> scan_sums;        // tile totals of the prefix sums
    DevBuf<BatchCounts> counts;       // [1]
    DevBuf<BatchCounts> counts_snap;  // [1] bucketed handles: copy of `counts` taken by the last kernel of a batch
    PinnedBuf<BatchCounts> h_counts;
    cudaEvent_t ev_done = nullptr;    // recorded after the last enqueued operation (other streams wait on it)
    cudaEvent_t ev_t0 = nullptr;      // rapid_cd_timer_start
    bool pending = false;             // an asynchronous batch is in flight: its status has not been collected yet
    int32_t deferred_rc = 0;          // status of asynchronous batches collected since the last rapid_cd_sync
    std::string deferred_msg;
    int32_t est_Sb = 0;               // batch subjects of the previous batch (grid sizing hint only)
    int64_t est_A = 0;
    BatchCounts last;                 // counters of the last collected batch
    int32_t retries = 0;              // batches replayed after growing the subject capacity
    DevBuf<unsigned long long> prep_stamps;   // RAPID_B200_PREP_STAMPS profiling aid
    // bucketed scratch lives in cd_bucketed.cu's own struct hung off here
    void* bucketed_state = nullptr;

    int32_t last_path = 0, last_launches = 0;
    float last_ms = 0.f, last_main_ms = 0.f;
    int64_t last_A = 0;
};

// row pointer helpers (device)
struct RowRef {
    uint16_t* masks;
    const uint8_t* cur;
    size_t Rpad;
    int nbuf;
    __device__ __forceinline__ uint16_t* row(int32_t slot) const {
        return masks + ((size_t)slot * nbuf + (nbuf == 2 ? cur[slot] : 0)) * Rpad;
    }
    __device__ __forceinline__ uint16_t* alt(int32_t slot) const {   // the non-current row (nbuf == 2)
        return masks + ((size_t)slot * nbuf + (cur[slot] ^ 1)) * Rpad;
    }
};

struct DeliveryDev {
    uint32_t flags = 0;
    const uint8_t* blocked = nullptr;
    const uint32_t* bitmap = nullptr;
    int64_t words = 0;
    uint64_t perm_seed = 0;
    int64_t cell_base = 0;            // PERMUTED: a receiver's order key of cell i is a function of i - cell_base (the index inside its batch)
};

// ---- the batch regrouped by subject (built by cd_prepare.cu, consumed by cd_bucketed.cu) -------------------------
struct SubjDesc {                     // 64 bytes, one per subject of the batch
    int32_t slot;
    uint16_t bmask;                   // rings reported in this batch (of a sequence of batches: in its LAST batch)
    uint8_t nr;                       // number of distinct rings in bmask
    uint8_t any_down;
    uint32_t tLf, tHf;                // for a fresh subject (state == pmask): moment of the cell that makes the L-th / H-th distinct ring, 0 if none
    uint32_t seg_begin, seg_len;      // its cells (of the last batch) in the slot-sorted arrays
    uint64_t mix1, mix2;              // fp_mix1 / fp_mix2 of the subject id
    // ---- a sequence of batches in one call (rapid_cd_apply_batches on bucketed handles): everything before the last batch
    uint16_t pmask;                   // rings reported in the PREFIX (batches before the last one); 0 for a single batch
    uint8_t pdown;                    // a DOWN cell in the prefix
    uint8_t pad0_;
    uint32_t f_bLp, f_bHp;            // fresh subject: 1-based prefix batch in which it reaches L / H distinct rings, 0 if it does not
    uint32_t pseg_len;                // prefix cells: they sit right before seg_begin in the sorted arrays
    uint64_t pad1_;
};
static_assert(sizeof(SubjDesc) == 64, "SubjDesc is staged in shared memory as 64-byte records");
struct SubjWalk {                     // first-occurrence ring sequence in arrival order (uniform delivery)
    uint8_t ring[16];
    uint32_t time[16];                // 1-based cell index; in a PREFIX walk (PrepOut::pwalk): 1-based batch index
};

// Invalidation work list of a bucketed handle: the subjects that sit in the unstable band of SOME receiver and have an observer
// that is itself a subject (only those can receive implicit reports, MultiNodeCutDetector.java:147-158), plus, per subject, the
// 1024-receiver tiles in which that is the case.
constexpr int SO_STRIDE = 16;
struct WorkList {
    uint8_t* has_so;                  // [slot] some observer of the subject has a slot
    int32_t* so_tab;                  // [slot][SO_STRIDE] slots of the subject's K observers (-1: not a subject)
    uint8_t* in_tile;                 // [slot][n_tiles] some receiver of the tile left the subject inside the band
    int32_t* listed;                  // [slot] on the list
    int32_t* slots;                   // the list
    int32_t* count;
    int32_t cap, n_tiles;
};
__device__ __forceinline__ void worklist_note(const WorkList& wl, int tile, int32_t slot) {
    wl.in_tile[(size_t)slot * wl.n_tiles + tile] = 1;
    if (*(volatile int32_t*)&wl.listed[slot] == 0 && atomicExch(&wl.listed[slot], 1) == 0) {
        const int32_t at = atomicAdd(wl.count, 1);
        if (at < wl.cap) wl.slots[at] = slot;
    }
}

struct PrepOut {                      // where the prepare kernel writes the regrouped batch
    SubjDesc* desc;
    SubjWalk* walk;
    int32_t* sidx;                    // cell indices grouped by subject, ascending within a subject
    uint8_t* s_ring;
    uint8_t* s_status;
    int32_t* batch_index;             // [slot] -> index of the subject in the batch
    int32_t* seg_cnt;                 // [slot] cells of the subject in this batch (scratch, all zero between batches)
    int32_t* batch_slots;             // [batch index] -> slot
    int32_t* bins;                    // [slot][64] indices of the subject's first 64 cells of this batch (any order)
    int32_t* ovf;                     // [A] cells beyond a subject's bin
    SubjWalk* pwalk;                  // sequences of batches: first-occurrence ring sequence of the prefix, time = batch index + 1
    int32_t* cell_batch;              // sequences of batches: [A] index of the batch of every valid cell
    WorkList wl;                      // invalidation work list (bucketed handles; wl.has_so == nullptr otherwise)
};

// implemented in cd_prepare.cu: filter + slot dictionary (+ regrouping by subject when po != nullptr) in ONE cooperative launch
// batch_off_dev != nullptr: the cells are a SEQUENCE of n_batches batches (cells [batch_off[b], batch_off[b+1]) = batch b) whose
// last non-empty one is `seq_last`; the descriptors then split every subject's cells into prefix and last batch.
int32_t prepare_batch(CD* cd, int64_t cfg, int64_t A, const int32_t* dst_dev, const uint8_t* ring_dev, const uint8_t* status_dev,
                      const int64_t* cfg_dev, const PrepOut* po, const int64_t* batch_off_dev = nullptr, int32_t n_batches = 1,
                      int32_t seq_last = 0);
// implemented in cd_bucketed.cu
int32_t bucketed_prep_buffers(CD* cd, int64_t A, PrepOut* po);

// enqueue only: no host synchronisation.  seq: the cells are a sequence of batches prepared with batch offsets (one pass, checked)
int32_t bucketed_apply(CD* cd, int64_t A, const DeliveryDev& dl, bool seq = false);
void bucketed_destroy(CD* cd);
int32_t bucketed_clear(CD* cd);
int32_t bucketed_clear_sticky(CD* cd);
// Wait for everything enqueued on the handle and collect the outcome of asynchronous batches (host mirrors of the slot count,
// timings, latched errors).  Returns the latched status (and clears it) when take_status is set.
int32_t cd_wait(const CD* cd, bool take_status);

}  // namespace rapid

struct rapid_cd : rapid::CD {};
