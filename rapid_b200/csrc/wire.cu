// Wire-format ingest: serialized protobuf (rapid.proto) -> cell SoA + Endpoint -> id, on the device.
//
// The host only walks the TOP level of a BatchedAlertMessage (one tag + one length per AlertMessage) to find the
// submessage ranges; everything inside them — varints, nested Endpoints, packed / unpacked ring numbers, NodeId,
// unknown fields — is parsed by one thread per AlertMessage.  Endpoints are resolved against an open-addressing
// table over the view's endpoints keyed by the ring-0 key the view already holds (MembershipView.java:579-582),
// with a byte compare on a hit.  Two passes over the messages: parse + count ring numbers, prefix sum, emit cells.
#include <limits.h>

#include <algorithm>
#include <map>
#include <string>
#include <utility>
#include <vector>


#include "common.cuh"
#include "scan.cuh"

namespace rapid {

// ------------------------------------------------------------------ protobuf wire primitives (host + device)
struct Rd {
    const uint8_t* p;
    const uint8_t* end;
    bool ok;
};
RAPID_HD uint64_t rd_varint(Rd& r) {
    uint64_t v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
        if (r.p >= r.end) break;
        const uint8_t b = *r.p++;
        if (shift < 64) v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return v;
    }
    r.ok = false;                       // truncated, or longer than 10 bytes
    return 0;
}
// after a tag: the payload range of a length-delimited field
RAPID_HD bool rd_len(Rd& r, const uint8_t** p, int64_t* len) {
    const uint64_t n = rd_varint(r);
    if (!r.ok || n > (uint64_t)(r.end - r.p)) { r.ok = false; return false; }
    *p = r.p; *len = (int64_t)n;
    r.p += n;
    return true;
}
RAPID_HD void rd_skip(Rd& r, uint32_t wire_type) {
    switch (wire_type) {
        case 0: rd_varint(r); break;
        case 1: if (r.end - r.p < 8) r.ok = false; else r.p += 8; break;
        case 2: { const uint8_t* p; int64_t n; rd_len(r, &p, &n); break; }
        case 5: if (r.end - r.p < 4) r.ok = false; else r.p += 4; break;
        default: r.ok = false;          // groups are not used by rapid.proto
    }
}

struct EpRef {                          // an Endpoint as it sits in the input buffer (rapid.proto:13-17)
    int32_t off, len, port, present;
};
// parse (merge) one Endpoint occurrence
RAPID_HD void parse_endpoint(const uint8_t* base, const uint8_t* p, int64_t n, EpRef* e, bool* ok) {
    Rd r{p, p + n, true};
    e->present = 1;
    while (r.ok && r.p < r.end) {
        const uint64_t tag = rd_varint(r);
        if (!r.ok) break;
        const uint32_t f = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
        if (f == 1 && wt == 2) { const uint8_t* q; int64_t l; if (rd_len(r, &q, &l)) { e->off = (int32_t)(q - base); e->len = (int32_t)l; } }
        else if (f == 2 && wt == 0) e->port = (int32_t)rd_varint(r);
        else if (f == 0) r.ok = false;
        else rd_skip(r, wt);
    }
    if (!r.ok) *ok = false;
}

struct MsgRec {                         // one AlertMessage (rapid.proto:101-110)
    EpRef src, dst;
    int64_t cfg, nid_high, nid_low;
    int32_t status, n_rings, has_nid, meta_off, meta_len, src_id, dst_id, pad_;
};

// ------------------------------------------------------------------ endpoint dictionary
__device__ __forceinline__ uint32_t ep_slot(int64_t key0) { return (uint32_t)(splitmix64((uint64_t)key0) >> 32); }

__global__ void k_wire_table_build(int64_t tot, const int64_t* __restrict__ key0, uint32_t T, int32_t* __restrict__ table) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= tot) return;
    uint32_t pos = ep_slot(key0[id]) & (T - 1);
    for (;;) {
        if (atomicCAS(&table[pos], -1, (int32_t)id) == -1) return;
        pos = (pos + 1) & (T - 1);
    }
}

__device__ int32_t ep_lookup(const uint8_t* __restrict__ buf, const EpRef e, uint32_t T, const int32_t* __restrict__ table,
                             const int64_t* __restrict__ key0, const uint8_t* __restrict__ hb, const int32_t* __restrict__ hoff,
                             const int32_t* __restrict__ hport) {
    if (!e.present) return -1;
    const int64_t k = ring_key(buf + e.off, e.len, e.port, 0);
    uint32_t pos = ep_slot(k) & (T - 1);
    for (uint32_t probes = 0; probes < T; ++probes) {
        const int32_t id = table[pos];
        if (id < 0) return -1;
        if (key0[id] == k && hport[id] == e.port && hoff[id + 1] - hoff[id] == e.len) {
            const uint8_t* a = hb + hoff[id];
            const uint8_t* b = buf + e.off;
            bool same = true;
            for (int32_t i = 0; i < e.len; ++i) if (a[i] != b[i]) { same = false; break; }
            if (same) return id;
        }
        pos = (pos + 1) & (T - 1);
    }
    return -1;
}

struct Dict {
    uint32_t T;
    const int32_t* table;
    const int64_t* key0;
    const uint8_t* hb;
    const int32_t* hoff;
    const int32_t* hport;
};

struct WireScal {
    int32_t bad_msg;        // lowest index of a malformed message, INT_MAX if none
    int32_t n_need;         // UP alerts whose edgeDst is not in the dictionary
    int32_t n_cells, n_dropped, sender_id, bad_vote;
};

// ------------------------------------------------------------------ alert kernels
// pass 1: parse every AlertMessage, resolve its endpoints, count its ring numbers
__global__ void k_wire_parse_alerts(int64_t M, const uint8_t* __restrict__ buf, const int64_t* __restrict__ moff,
                                    const int32_t* __restrict__ mlen, Dict d, MsgRec* __restrict__ rec, int32_t* __restrict__ need,
                                    WireScal* __restrict__ sc, int have_cfg, int64_t cur_cfg) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    MsgRec r;
    memset(&r, 0, sizeof(r));
    const uint8_t* p0 = buf + moff[m];
    Rd rd{p0, p0 + mlen[m], true};
    bool ok = true;
    while (rd.ok && rd.p < rd.end) {
        const uint64_t tag = rd_varint(rd);
        if (!rd.ok) break;
        const uint32_t f = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
        const uint8_t* q; int64_t l;
        if (f == 1 && wt == 2) { if (rd_len(rd, &q, &l)) parse_endpoint(buf, q, l, &r.src, &ok); }
        else if (f == 2 && wt == 2) { if (rd_len(rd, &q, &l)) parse_endpoint(buf, q, l, &r.dst, &ok); }
        else if (f == 3 && wt == 0) r.status = (int32_t)rd_varint(rd);
        else if (f == 4 && wt == 0) r.cfg = (int64_t)rd_varint(rd);
        else if (f == 5 && wt == 0) { rd_varint(rd); ++r.n_rings; }                       // unpacked repeated int32
        else if (f == 5 && wt == 2) {                                                     // packed
            if (rd_len(rd, &q, &l)) { Rd pr{q, q + l, true}; while (pr.ok && pr.p < pr.end) { rd_varint(pr); ++r.n_rings; } if (!pr.ok) ok = false; }
        }
        else if (f == 6 && wt == 2) {                                                     // NodeId {high = 1, low = 2}
            if (rd_len(rd, &q, &l)) {
                r.has_nid = 1;
                Rd nr{q, q + l, true};
                while (nr.ok && nr.p < nr.end) {
                    const uint64_t t2 = rd_varint(nr);
                    if (!nr.ok) break;
                    if ((t2 >> 3) == 1 && (t2 & 7) == 0) r.nid_high = (int64_t)rd_varint(nr);
                    else if ((t2 >> 3) == 2 && (t2 & 7) == 0) r.nid_low = (int64_t)rd_varint(nr);
                    else if ((t2 >> 3) == 0) nr.ok = false;
                    else rd_skip(nr, (uint32_t)(t2 & 7));
                }
                if (!nr.ok) ok = false;
            }
        }
        else if (f == 7 && wt == 2) { if (rd_len(rd, &q, &l)) { r.meta_off = (int32_t)(q - buf); r.meta_len = (int32_t)l; } }
        else if (f == 0) rd.ok = false;
        else rd_skip(rd, wt);
    }
    if (!rd.ok || !ok || r.status < 0 || r.status > 1) { atomicMin(&sc->bad_msg, (int32_t)m); r.n_rings = 0; r.src.present = r.dst.present = 0; }
    r.src_id = ep_lookup(buf, r.src, d.T, d.table, d.key0, d.hb, d.hoff, d.hport);
    r.dst_id = r.dst.present ? ep_lookup(buf, r.dst, d.T, d.table, d.key0, d.hb, d.hoff, d.hport) : -1;
    // a default-valued edgeDst (field absent) is the endpoint {"" , 0}: resolvable like any other
    if (!r.dst.present) { EpRef e{0, 0, 0, 1}; r.dst = e; r.dst_id = ep_lookup(buf, e, d.T, d.table, d.key0, d.hb, d.hoff, d.hport); }
    // UP about an unknown endpoint: a joiner — but only an alert of the CURRENT configuration may introduce one (a stale one is
    // dropped by filterAlertMessages, MembershipService.java:653, before extractJoinerUuidAndMetadata ever sees it)
    const bool nd = r.dst_id < 0 && r.status == 0 && (!have_cfg || r.cfg == cur_cfg);
    need[m] = nd ? 1 : 0;
    if (nd) atomicAdd(&sc->n_need, 1);
    rec[m] = r;
}
// after joiners were registered: resolve what was unknown
__global__ void k_wire_relookup(int64_t M, const uint8_t* __restrict__ buf, Dict d, MsgRec* __restrict__ rec) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    if (rec[m].dst_id < 0) rec[m].dst_id = ep_lookup(buf, rec[m].dst, d.T, d.table, d.key0, d.hb, d.hoff, d.hport);
    if (rec[m].src_id < 0) rec[m].src_id = ep_lookup(buf, rec[m].src, d.T, d.table, d.key0, d.hb, d.hoff, d.hport);
}
__global__ void k_wire_counts(int64_t M, const MsgRec* __restrict__ rec, int32_t* __restrict__ cnt, WireScal* __restrict__ sc) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const bool drop = rec[m].dst_id < 0;
    cnt[m] = drop ? 0 : rec[m].n_rings;
    if (drop) atomicAdd(&sc->n_dropped, 1);
}
// pass 2: one cell per ring number, in message order then ring order
__global__ void k_wire_emit(int64_t M, const uint8_t* __restrict__ buf, const int64_t* __restrict__ moff, const int32_t* __restrict__ mlen,
                            const MsgRec* __restrict__ rec, const int32_t* __restrict__ cnt, const int32_t* __restrict__ pos,
                            int32_t* __restrict__ o_src, int32_t* __restrict__ o_dst, uint8_t* __restrict__ o_ring,
                            uint8_t* __restrict__ o_status, int64_t* __restrict__ o_cfg, WireScal* __restrict__ sc) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    if (m == M - 1) sc->n_cells = pos[m] + cnt[m];
    if (cnt[m] == 0) return;
    const MsgRec r = rec[m];
    int32_t j = pos[m];
    const uint8_t* p0 = buf + moff[m];
    Rd rd{p0, p0 + mlen[m], true};
    while (rd.ok && rd.p < rd.end) {
        const uint64_t tag = rd_varint(rd);
        if (!rd.ok) break;
        const uint32_t f = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
        if (f == 5 && wt == 0) {
            const int32_t ring = (int32_t)rd_varint(rd);
            o_src[j] = r.src_id; o_dst[j] = r.dst_id; o_ring[j] = (uint8_t)(ring < 0 || ring > 255 ? 255 : ring);
            o_status[j] = (uint8_t)r.status; o_cfg[j] = r.cfg; ++j;
        } else if (f == 5 && wt == 2) {
            const uint8_t* q; int64_t l;
            if (rd_len(rd, &q, &l)) {
                Rd pr{q, q + l, true};
                while (pr.ok && pr.p < pr.end) {
                    const int32_t ring = (int32_t)rd_varint(pr);
                    o_src[j] = r.src_id; o_dst[j] = r.dst_id; o_ring[j] = (uint8_t)(ring < 0 || ring > 255 ? 255 : ring);
                    o_status[j] = (uint8_t)r.status; o_cfg[j] = r.cfg; ++j;
                }
            }
        } else rd_skip(rd, wt);
    }
}
__global__ void k_wire_begin(WireScal* sc) { sc->bad_msg = INT_MAX; sc->n_need = 0; sc->n_cells = 0; sc->n_dropped = 0; sc->sender_id = -1; sc->bad_vote = INT_MAX; }
__global__ void k_wire_sender(const uint8_t* __restrict__ buf, EpRef e, Dict d, WireScal* __restrict__ sc) {
    sc->sender_id = ep_lookup(buf, e, d.T, d.table, d.key0, d.hb, d.hoff, d.hport);
}
__global__ void k_wire_msg_fields(int64_t M, const MsgRec* __restrict__ rec, int32_t* dst, uint8_t* status, int32_t* n_rings, int64_t* nh,
                                  int64_t* nl, uint8_t* has, int64_t* moff, int32_t* mlen) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const MsgRec r = rec[m];
    dst[m] = r.dst_id; status[m] = (uint8_t)r.status; n_rings[m] = r.n_rings; nh[m] = r.nid_high; nl[m] = r.nid_low;
    has[m] = (uint8_t)r.has_nid; moff[m] = r.meta_off; mlen[m] = r.meta_len;
}

// ------------------------------------------------------------------ vote kernel: one thread per FastRoundPhase2bMessage
__global__ void k_wire_votes(int64_t n, const uint8_t* __restrict__ buf, const int64_t* __restrict__ off, int unwrap, Dict d,
                             int32_t* __restrict__ sender, int64_t* __restrict__ cfg, uint64_t* __restrict__ h1, uint64_t* __restrict__ h2,
                             int32_t* __restrict__ len, WireScal* __restrict__ sc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = buf + off[i];
    int64_t l = off[i + 1] - off[i];
    bool ok = l >= 0;
    if (ok && unwrap) {                                      // RapidRequest.fastRoundPhase2bMessage = 5
        Rd r{p, p + l, true};
        const uint8_t* q = nullptr; int64_t ql = -1;
        while (r.ok && r.p < r.end) {
            const uint64_t tag = rd_varint(r);
            if (!r.ok) break;
            if ((tag >> 3) == 5 && (tag & 7) == 2) rd_len(r, &q, &ql);
            else if ((tag >> 3) == 0) r.ok = false;
            else {
                if ((tag >> 3) <= 10) { q = nullptr; ql = -1; }                  // another case of the oneof replaces it
                rd_skip(r, (uint32_t)(tag & 7));
            }
        }
        if (!r.ok || ql < 0) ok = false; else { p = q; l = ql; }
    }
    EpRef s{0, 0, 0, 0};
    int64_t c = 0;
    uint64_t a = 0, b = 0;
    int32_t cnt = 0;
    bool unknown = false;
    if (ok) {
        Rd r{p, p + l, true};
        while (r.ok && r.p < r.end) {
            const uint64_t tag = rd_varint(r);
            if (!r.ok) break;
            const uint32_t f = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
            const uint8_t* q; int64_t ql;
            if (f == 1 && wt == 2) { if (rd_len(r, &q, &ql)) parse_endpoint(buf, q, ql, &s, &ok); }
            else if (f == 2 && wt == 0) c = (int64_t)rd_varint(r);
            else if (f == 3 && wt == 2) {
                if (rd_len(r, &q, &ql)) {
                    EpRef e{0, 0, 0, 0};
                    parse_endpoint(buf, q, ql, &e, &ok);
                    const int32_t id = ep_lookup(buf, e, d.T, d.table, d.key0, d.hb, d.hoff, d.hport);
                    if (id >= 0) { a += fp_mix1(id); b += fp_mix2(id); }
                    else {
                        // an endpoint outside the dictionary (a vote of another configuration, say — FastPaxos.java:126-132 drops
                        // those by their configurationId, not by their content): it still gets an identity — its ring-0 key — so that
                        // identical lists keep identical fingerprints and the tally's own filters decide what the vote is worth
                        const uint64_t kk = (uint64_t)ring_key(buf + e.off, e.len, e.port, 0);
                        a += splitmix64(kk ^ 0x554E4B4E4F574E31ULL); b += splitmix64((kk * 0xD6E8FEB86659FD93ULL) ^ 0x554E4B4E4F574E32ULL);
                        unknown = true;
                    }
                    ++cnt;
                }
            }
            else if (f == 0) r.ok = false;
            else rd_skip(r, wt);
        }
        if (!r.ok) ok = false;
    }
    if (!ok) { atomicMin(&sc->bad_msg, (int32_t)i); sender[i] = -1; cfg[i] = 0; h1[i] = 0; h2[i] = 0; len[i] = 0; return; }
    if (unknown) atomicMin(&sc->bad_vote, (int32_t)i);       // reported as a count-free diagnostic only (first such vote)
    sender[i] = ep_lookup(buf, s, d.T, d.table, d.key0, d.hb, d.hoff, d.hport);
    cfg[i] = c; h1[i] = a; h2[i] = b; len[i] = cnt;
}

// ------------------------------------------------------------------ handle
struct Wire {
    rapid_view* view = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    uint64_t table_epoch = 0;
    bool have_cfg = false;            // rapid_wire_set_configuration: only alerts of this configuration register joiners
    int64_t cur_cfg = 0;
    uint32_t T = 0;
    DevBuf<int32_t> table;
    DevBuf<uint8_t> buf;
    DevBuf<int64_t> moff;
    DevBuf<int32_t> mlen, need, cnt, pos, scan_sums;
    DevBuf<MsgRec> rec;
    DevBuf<WireScal> sc;
    PinnedBuf<WireScal> h_sc;
    int64_t M = 0, n_cells = 0;
    DevBuf<int32_t> o_src, o_dst;
    DevBuf<uint8_t> o_ring, o_status;
    DevBuf<int64_t> o_cfg;
    // votes / message-field staging
    DevBuf<int64_t> v_off, v_cfg, t_i64a, t_i64b, t_i64c;
    DevBuf<int32_t> v_sender, v_len, t_i32a, t_i32b, t_i32c;
    DevBuf<uint64_t> v_h1, v_h2;
    DevBuf<uint8_t> t_u8a, t_u8b;
};

static const int TB = 128;
static inline unsigned grid_for(int64_t n) { return (unsigned)ceil_div<int64_t>(n > 0 ? n : 1, TB); }

static int32_t wire_dict(Wire* w, Dict* d) {
    View* v = w->view;
    const int64_t tot = v->n + v->nj;
    if (w->table_epoch != v->epoch || w->T == 0) {
        uint32_t T = 1024;
        while ((int64_t)T < 2 * tot) T <<= 1;
        RAPID_CHECK(w->table.reserve(T));
        w->T = T;
        RAPID_CUDA(cudaStreamSynchronize(v->stream));
        RAPID_CUDA(cudaMemsetAsync(w->table.p, 0xff, (size_t)T * sizeof(int32_t), w->stream));
        if (tot) k_wire_table_build<<<grid_for(tot), TB, 0, w->stream>>>(tot, v->key.p, T, w->table.p);
        RAPID_KERNEL_CHECK();
        w->table_epoch = v->epoch;
    }
    d->T = w->T; d->table = w->table.p; d->key0 = v->key.p; d->hb = v->host_bytes.p; d->hoff = v->host_off.p; d->hport = v->port.p;
    return RAPID_OK;
}

static int32_t wire_read_scal(Wire* w) {
    RAPID_CUDA(cudaMemcpyAsync(w->h_sc.p, w->sc.p, sizeof(WireScal), cudaMemcpyDeviceToHost, w->stream));
    RAPID_CUDA(cudaStreamSynchronize(w->stream));
    return RAPID_OK;
}

// host: find the payload of field `field` (length-delimited) at the top level of [p, p+len); last occurrence wins
static bool host_find_field(const uint8_t* p, int64_t len, uint32_t field, const uint8_t** q, int64_t* ql) {
    Rd r{p, p + len, true};
    *q = nullptr; *ql = -1;
    while (r.ok && r.p < r.end) {
        const uint64_t tag = rd_varint(r);
        if (!r.ok) break;
        if ((tag >> 3) == 0) return false;
        if ((uint32_t)(tag >> 3) == field && (tag & 7) == 2) { if (!rd_len(r, q, ql)) return false; }
        else rd_skip(r, (uint32_t)(tag & 7));
    }
    return r.ok;
}

}  // namespace rapid

using namespace rapid;

struct rapid_wire : rapid::Wire {};

extern "C" {

int32_t rapid_wire_create(rapid_wire** out, rapid_view* v) {
    if (!out || !v) { set_error("NULL argument"); return RAPID_EINVAL; }
    *out = nullptr;
    DeviceGuard g(v->device);
    rapid_wire* w = new rapid_wire();
    w->view = v; w->device = v->device;
    int32_t rc = RAPID_OK;
    do {
        if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&w->ev0) != cudaSuccess ||
            cudaEventCreate(&w->ev1) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "stream", __FILE__, __LINE__); break; }
        if ((rc = w->sc.reserve(1)) || (rc = w->h_sc.reserve(1))) break;
    } while (0);
    if (rc) { rapid_wire_destroy(w); return rc; }
    *out = w;
    return RAPID_OK;
}

int32_t rapid_wire_set_configuration(rapid_wire* w, int64_t cfg_id) {
    if (!w) { set_error("NULL handle"); return RAPID_EINVAL; }
    w->have_cfg = true; w->cur_cfg = cfg_id;
    return RAPID_OK;
}

int32_t rapid_wire_destroy(rapid_wire* w) {
    if (!w) return RAPID_OK;
    DeviceGuard g(w->device);
    if (w->stream) { cudaStreamSynchronize(w->stream); cudaStreamDestroy(w->stream); }
    if (w->ev0) cudaEventDestroy(w->ev0);
    if (w->ev1) cudaEventDestroy(w->ev1);
    delete w;
    return RAPID_OK;
}

int32_t rapid_wire_decode_alerts(rapid_wire* w, const uint8_t* bytes, int64_t len, uint32_t flags, int64_t* n_messages, int64_t* n_cells,
                                 int64_t* n_dropped, int64_t* n_new_joiners, int32_t* sender_id) {
    if (!w || len < 0 || (len && !bytes) || len > 0x7ffffff0LL) { set_error("bad arguments"); return RAPID_EINVAL; }
    DeviceGuard g(w->device);
    cudaStream_t s = w->stream;
    w->M = 0; w->n_cells = 0;
    // ---- top level (host): BatchedAlertMessage { sender = 1; repeated AlertMessage messages = 3 }
    const uint8_t* body = bytes;
    int64_t blen = len;
    if (flags & RAPID_WIRE_REQUEST) {                        // RapidRequest.batchedAlertMessage = 3
        if (!host_find_field(bytes, len, 3, &body, &blen)) { set_error("malformed RapidRequest"); return RAPID_EINVAL; }
        if (blen < 0) { set_error("RapidRequest does not carry a BatchedAlertMessage"); return RAPID_EINVAL; }
    }
    std::vector<int64_t> moff;
    std::vector<int32_t> mlen;
    EpRef sender{0, 0, 0, 0};
    {
        Rd r{body, body + blen, true};
        bool ok = true;
        while (r.ok && r.p < r.end) {
            const uint64_t tag = rd_varint(r);
            if (!r.ok) break;
            const uint32_t f = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
            const uint8_t* q; int64_t l;
            if (f == 3 && wt == 2) { if (rd_len(r, &q, &l)) { moff.push_back(q - bytes); mlen.push_back((int32_t)l); } }
            else if (f == 1 && wt == 2) { if (rd_len(r, &q, &l)) parse_endpoint(bytes, q, l, &sender, &ok); }
            else if (f == 0) r.ok = false;
            else rd_skip(r, wt);
        }
        if (!r.ok || !ok) { set_error("malformed BatchedAlertMessage"); return RAPID_EINVAL; }
    }
    const int64_t M = (int64_t)moff.size();
    RAPID_CUDA(cudaEventRecord(w->ev0, s));
    Dict d;
    RAPID_CHECK(wire_dict(w, &d));
    RAPID_CHECK(w->buf.reserve((size_t)std::max<int64_t>(len, 1)));
    if (len) RAPID_CUDA(cudaMemcpyAsync(w->buf.p, bytes, (size_t)len, cudaMemcpyHostToDevice, s));
    k_wire_begin<<<1, 1, 0, s>>>(w->sc.p);
    if (sender.present) k_wire_sender<<<1, 1, 0, s>>>(w->buf.p, sender, d, w->sc.p);
    RAPID_KERNEL_CHECK();
    int64_t new_joiners = 0;
    if (M > 0) {
        RAPID_CHECK(w->moff.reserve((size_t)M)); RAPID_CHECK(w->mlen.reserve((size_t)M)); RAPID_CHECK(w->need.reserve((size_t)M));
        RAPID_CHECK(w->cnt.reserve((size_t)M)); RAPID_CHECK(w->pos.reserve((size_t)M)); RAPID_CHECK(w->rec.reserve((size_t)M));
        RAPID_CUDA(cudaMemcpyAsync(w->moff.p, moff.data(), (size_t)M * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        RAPID_CUDA(cudaMemcpyAsync(w->mlen.p, mlen.data(), (size_t)M * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        k_wire_parse_alerts<<<grid_for(M), TB, 0, s>>>(M, w->buf.p, w->moff.p, w->mlen.p, d, w->rec.p, w->need.p, w->sc.p, w->have_cfg ? 1 : 0, w->cur_cfg);
        RAPID_KERNEL_CHECK();
        RAPID_CHECK(wire_read_scal(w));
        if (w->h_sc.p->bad_msg != INT_MAX) { set_error("malformed AlertMessage at index %d", w->h_sc.p->bad_msg); return RAPID_EINVAL; }
        if (w->h_sc.p->n_need > 0) {
            // joiners announced by UP alerts: register them in order of first appearance, then resolve again
            std::vector<int32_t> need((size_t)M);
            std::vector<MsgRec> rec((size_t)M);
            RAPID_CUDA(cudaMemcpyAsync(need.data(), w->need.p, (size_t)M * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            RAPID_CUDA(cudaMemcpyAsync(rec.data(), w->rec.p, (size_t)M * sizeof(MsgRec), cudaMemcpyDeviceToHost, s));
            RAPID_CUDA(cudaStreamSynchronize(s));
            std::map<std::pair<std::string, int32_t>, int> seen;
            std::vector<uint8_t> hb;
            std::vector<int32_t> off(1, 0), port;
            for (int64_t m = 0; m < M; ++m) {
                if (!need[(size_t)m]) continue;
                const EpRef& e = rec[(size_t)m].dst;
                std::pair<std::string, int32_t> k(std::string((const char*)bytes + e.off, (size_t)e.len), e.port);
                if (seen.count(k)) continue;
                seen[k] = 1;
                hb.insert(hb.end(), bytes + e.off, bytes + e.off + e.len);
                off.push_back(off.back() + e.len);
                port.push_back(e.port);
            }
            new_joiners = (int64_t)port.size();
            static const uint8_t dummy = 0;
            int32_t first = 0;
            RAPID_CHECK(rapid_view_register_joiners(w->view, new_joiners, hb.empty() ? &dummy : hb.data(), off.data(), port.data(), &first));
            RAPID_CHECK(wire_dict(w, &d));
            k_wire_relookup<<<grid_for(M), TB, 0, s>>>(M, w->buf.p, d, w->rec.p);
            if (sender.present) k_wire_sender<<<1, 1, 0, s>>>(w->buf.p, sender, d, w->sc.p);
            RAPID_KERNEL_CHECK();
        }
        k_wire_counts<<<grid_for(M), TB, 0, s>>>(M, w->rec.p, w->cnt.p, w->sc.p);
        RAPID_KERNEL_CHECK();
        RAPID_CHECK(exclusive_scan_i32_to(w->cnt.p, w->pos.p, M, w->scan_sums, s));
        // every ring number is at least one byte on the wire: len bounds the number of cells
        const size_t cap = (size_t)std::max<int64_t>(len, 1);
        RAPID_CHECK(w->o_src.reserve(cap)); RAPID_CHECK(w->o_dst.reserve(cap)); RAPID_CHECK(w->o_ring.reserve(cap));
        RAPID_CHECK(w->o_status.reserve(cap)); RAPID_CHECK(w->o_cfg.reserve(cap));
        k_wire_emit<<<grid_for(M), TB, 0, s>>>(M, w->buf.p, w->moff.p, w->mlen.p, w->rec.p, w->cnt.p, w->pos.p, w->o_src.p, w->o_dst.p,
                                               w->o_ring.p, w->o_status.p, w->o_cfg.p, w->sc.p);
        RAPID_KERNEL_CHECK();
    }
    RAPID_CUDA(cudaEventRecord(w->ev1, s));
    RAPID_CHECK(wire_read_scal(w));
    cudaEventElapsedTime(&w->last_ms, w->ev0, w->ev1);
    w->M = M; w->n_cells = w->h_sc.p->n_cells;
    if (n_messages) *n_messages = M;
    if (n_cells) *n_cells = w->n_cells;
    if (n_dropped) *n_dropped = w->h_sc.p->n_dropped;
    if (n_new_joiners) *n_new_joiners = new_joiners;
    if (sender_id) *sender_id = w->h_sc.p->sender_id;
    return RAPID_OK;
}

int32_t rapid_wire_cells_dev(const rapid_wire* w, const int32_t** src, const int32_t** dst, const uint8_t** ring, const uint8_t** status,
                             const int64_t** cfg) {
    if (!w) { set_error("NULL handle"); return RAPID_EINVAL; }
    if (src) *src = w->o_src.p;
    if (dst) *dst = w->o_dst.p;
    if (ring) *ring = w->o_ring.p;
    if (status) *status = w->o_status.p;
    if (cfg) *cfg = w->o_cfg.p;
    return RAPID_OK;
}

int32_t rapid_wire_read_cells(const rapid_wire* w, int32_t* src, int32_t* dst, uint8_t* ring, uint8_t* status, int64_t* cfg) {
    if (!w) { set_error("NULL handle"); return RAPID_EINVAL; }
    DeviceGuard g(w->device);
    const size_t n = (size_t)w->n_cells;
    if (n == 0) return RAPID_OK;
    cudaStream_t s = w->stream;
    if (src) RAPID_CUDA(cudaMemcpyAsync(src, w->o_src.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (dst) RAPID_CUDA(cudaMemcpyAsync(dst, w->o_dst.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (ring) RAPID_CUDA(cudaMemcpyAsync(ring, w->o_ring.p, n, cudaMemcpyDeviceToHost, s));
    if (status) RAPID_CUDA(cudaMemcpyAsync(status, w->o_status.p, n, cudaMemcpyDeviceToHost, s));
    if (cfg) RAPID_CUDA(cudaMemcpyAsync(cfg, w->o_cfg.p, n * 8, cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    return RAPID_OK;
}

int32_t rapid_wire_read_messages(const rapid_wire* cw, int32_t* dst, uint8_t* status, int32_t* n_rings, int64_t* node_high, int64_t* node_low,
                                 uint8_t* has_node_id, int64_t* meta_off, int32_t* meta_len) {
    rapid_wire* w = const_cast<rapid_wire*>(cw);             // uses the handle's staging buffers; logically const
    if (!w) { set_error("NULL handle"); return RAPID_EINVAL; }
    DeviceGuard g(w->device);
    const int64_t M = w->M;
    if (M == 0) return RAPID_OK;
    cudaStream_t s = w->stream;
    const size_t m = (size_t)M;
    RAPID_CHECK(w->t_i32a.reserve(m)); RAPID_CHECK(w->t_i32b.reserve(m)); RAPID_CHECK(w->t_i32c.reserve(m));
    RAPID_CHECK(w->t_i64a.reserve(m)); RAPID_CHECK(w->t_i64b.reserve(m)); RAPID_CHECK(w->t_i64c.reserve(m));
    RAPID_CHECK(w->t_u8a.reserve(m)); RAPID_CHECK(w->t_u8b.reserve(m));
    k_wire_msg_fields<<<grid_for(M), TB, 0, s>>>(M, w->rec.p, w->t_i32a.p, w->t_u8a.p, w->t_i32b.p, w->t_i64a.p, w->t_i64b.p, w->t_u8b.p,
                                                 w->t_i64c.p, w->t_i32c.p);
    RAPID_KERNEL_CHECK();
    if (dst) RAPID_CUDA(cudaMemcpyAsync(dst, w->t_i32a.p, m * 4, cudaMemcpyDeviceToHost, s));
    if (status) RAPID_CUDA(cudaMemcpyAsync(status, w->t_u8a.p, m, cudaMemcpyDeviceToHost, s));
    if (n_rings) RAPID_CUDA(cudaMemcpyAsync(n_rings, w->t_i32b.p, m * 4, cudaMemcpyDeviceToHost, s));
    if (node_high) RAPID_CUDA(cudaMemcpyAsync(node_high, w->t_i64a.p, m * 8, cudaMemcpyDeviceToHost, s));
    if (node_low) RAPID_CUDA(cudaMemcpyAsync(node_low, w->t_i64b.p, m * 8, cudaMemcpyDeviceToHost, s));
    if (has_node_id) RAPID_CUDA(cudaMemcpyAsync(has_node_id, w->t_u8b.p, m, cudaMemcpyDeviceToHost, s));
    if (meta_off) RAPID_CUDA(cudaMemcpyAsync(meta_off, w->t_i64c.p, m * 8, cudaMemcpyDeviceToHost, s));
    if (meta_len) RAPID_CUDA(cudaMemcpyAsync(meta_len, w->t_i32c.p, m * 4, cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    return RAPID_OK;
}

int32_t rapid_wire_decode_votes(rapid_wire* w, const uint8_t* bytes, const int64_t* off, int64_t n, uint32_t flags, int32_t* sender,
                                int64_t* vote_cfg, uint64_t* proposal_hash, uint64_t* proposal_hash2, int32_t* proposal_len) {
    if (!w || n < 0 || (n && (!bytes || !off)) || n > 0x7ffffff0LL) { set_error("bad arguments"); return RAPID_EINVAL; }
    if (n == 0) return RAPID_OK;
    for (int64_t i = 0; i < n; ++i)
        if (off[i + 1] < off[i]) { set_error("off must be non-decreasing"); return RAPID_EINVAL; }
    const int64_t len = off[n] - off[0];
    DeviceGuard g(w->device);
    cudaStream_t s = w->stream;
    RAPID_CUDA(cudaEventRecord(w->ev0, s));
    Dict d;
    RAPID_CHECK(wire_dict(w, &d));
    RAPID_CHECK(w->buf.reserve((size_t)std::max<int64_t>(off[n], 1)));
    // keep the caller's offsets valid: copy [0, off[n]) (the prefix before off[0] is never read)
    if (len) RAPID_CUDA(cudaMemcpyAsync(w->buf.p + off[0], bytes + off[0], (size_t)len, cudaMemcpyHostToDevice, s));
    RAPID_CHECK(w->v_off.reserve((size_t)n + 1)); RAPID_CHECK(w->v_sender.reserve((size_t)n)); RAPID_CHECK(w->v_cfg.reserve((size_t)n));
    RAPID_CHECK(w->v_h1.reserve((size_t)n)); RAPID_CHECK(w->v_h2.reserve((size_t)n)); RAPID_CHECK(w->v_len.reserve((size_t)n));
    RAPID_CUDA(cudaMemcpyAsync(w->v_off.p, off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    k_wire_begin<<<1, 1, 0, s>>>(w->sc.p);
    k_wire_votes<<<grid_for(n), TB, 0, s>>>(n, w->buf.p, w->v_off.p, (flags & RAPID_WIRE_REQUEST) ? 1 : 0, d, w->v_sender.p, w->v_cfg.p,
                                            w->v_h1.p, w->v_h2.p, w->v_len.p, w->sc.p);
    RAPID_KERNEL_CHECK();
    RAPID_CUDA(cudaEventRecord(w->ev1, s));
    RAPID_CHECK(wire_read_scal(w));
    cudaEventElapsedTime(&w->last_ms, w->ev0, w->ev1);
    if (w->h_sc.p->bad_msg != INT_MAX) { set_error("malformed FastRoundPhase2bMessage at index %d", w->h_sc.p->bad_msg); return RAPID_EINVAL; }
    // (a vote naming an endpoint outside the dictionary is NOT an error: see k_wire_votes)
    const size_t m = (size_t)n;
    if (sender) RAPID_CUDA(cudaMemcpyAsync(sender, w->v_sender.p, m * 4, cudaMemcpyDeviceToHost, s));
    if (vote_cfg) RAPID_CUDA(cudaMemcpyAsync(vote_cfg, w->v_cfg.p, m * 8, cudaMemcpyDeviceToHost, s));
    if (proposal_hash) RAPID_CUDA(cudaMemcpyAsync(proposal_hash, w->v_h1.p, m * 8, cudaMemcpyDeviceToHost, s));
    if (proposal_hash2) RAPID_CUDA(cudaMemcpyAsync(proposal_hash2, w->v_h2.p, m * 8, cudaMemcpyDeviceToHost, s));
    if (proposal_len) RAPID_CUDA(cudaMemcpyAsync(proposal_len, w->v_len.p, m * 4, cudaMemcpyDeviceToHost, s));
    RAPID_CUDA(cudaStreamSynchronize(s));
    return RAPID_OK;
}

int32_t rapid_wire_last_device_ms(const rapid_wire* w, float* total_ms) {
    if (!w || !total_ms) return RAPID_EINVAL;
    *total_ms = w->last_ms;
    return RAPID_OK;
}

}  // extern "C"
