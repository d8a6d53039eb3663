/*
 * ORACLE — test infrastructure, NOT product code (see rapid_oracle.hpp header).
 *
 * extern "C" veneer over the literal restatement so that tests/ and bench.py's CPU legs can drive
 * it through ctypes.  Endpoints are interned in an `orc_universe` and referred to by int32 tag;
 * the classes underneath stay keyed by Endpoint{hostname bytes, port} like the Java.
 *
 * Status codes mirror include/rapid_b200.h: 0 ok, -1 EINVAL, -2 NOT_IN_RING, -3 ALREADY_IN_RING,
 * -4 UUID_SEEN.
 */
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "rapid_oracle.hpp"
#include "paxos_oracle.hpp"
#include "fd_oracle.hpp"

using namespace oracle;

struct orc_universe {
    std::vector<Endpoint> eps;
    std::unordered_map<Endpoint, int32_t, EndpointHash> index;
    int32_t tagOf(const Endpoint& e) const {
        auto it = index.find(e);
        return it == index.end() ? -1 : it->second;
    }
};

struct orc_view {
    orc_universe* u;
    std::unique_ptr<MembershipView> v;
};

struct orc_cd {
    orc_universe* u;
    std::unique_ptr<MultiNodeCutDetector> cd;
};

struct orc_handler {
    orc_view* view;
    std::unique_ptr<AlertBatchHandler> h;
};

struct orc_fp {
    orc_universe* u;
    std::unique_ptr<FastPaxosTally> fp;
};

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

static int emit_tags(const orc_universe* u, const std::vector<Endpoint>& v, int32_t* out, int32_t cap) {
    int n = 0;
    for (const Endpoint& e : v) {
        if (out && n < cap) out[n] = u->tagOf(e);
        ++n;
    }
    return n;
}

extern "C" {

uint64_t orc_xxh64_bytes(const uint8_t* data, int64_t len, uint64_t seed) { return orc_xxh64(data, (size_t)len, seed); }
uint64_t orc_xxh64_int(int32_t v, uint64_t seed) { return orc_xx_hash_int(v, seed); }
uint64_t orc_xxh64_long(int64_t v, uint64_t seed) { return orc_xx_hash_long(v, seed); }
uint64_t orc_splitmix64(uint64_t x) { return splitmix64(x); }

/* ---------------- universe ---------------- */
orc_universe* orc_universe_create(void) { return new orc_universe(); }
void orc_universe_destroy(orc_universe* u) { delete u; }
int32_t orc_universe_add(orc_universe* u, const uint8_t* host, int32_t len, int32_t port) {
    Endpoint e;
    e.hostname.assign((const char*)host, (size_t)len);
    e.port = port;
    auto it = u->index.find(e);
    if (it != u->index.end()) return it->second;
    const int32_t tag = (int32_t)u->eps.size();
    u->eps.push_back(e);
    u->index.emplace(e, tag);
    return tag;
}
/* bulk: hostnames concatenated, off[n+1] */
int32_t orc_universe_add_bulk(orc_universe* u, int64_t n, const uint8_t* hb, const int32_t* off,
                              const int32_t* port, int32_t* out_tags) {
    for (int64_t i = 0; i < n; ++i)
        out_tags[i] = orc_universe_add(u, hb + off[i], off[i + 1] - off[i], port[i]);
    return 0;
}
int32_t orc_universe_size(const orc_universe* u) { return (int32_t)u->eps.size(); }

/* ---------------- MembershipView ---------------- */
orc_view* orc_view_create(orc_universe* u, int32_t K) {
    if (K <= 0) return nullptr;
    orc_view* v = new orc_view();
    v->u = u;
    v->v.reset(new MembershipView(K));
    return v;
}
orc_view* orc_view_create_bulk(orc_universe* u, int32_t K, const int32_t* tags, int64_t n,
                               const int64_t* id_high, const int64_t* id_low, int64_t n_ids) {
    if (K <= 0) return nullptr;
    std::vector<Endpoint> eps;
    eps.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) eps.push_back(u->eps[(size_t)tags[i]]);
    std::vector<NodeId> ids((size_t)n_ids);
    for (int64_t i = 0; i < n_ids; ++i) { ids[(size_t)i].high = id_high[i]; ids[(size_t)i].low = id_low[i]; }
    orc_view* v = new orc_view();
    v->u = u;
    v->v.reset(new MembershipView(K, ids, eps));
    return v;
}
void orc_view_destroy(orc_view* v) { delete v; }

int32_t orc_view_ring_add(orc_view* v, int32_t tag, int64_t high, int64_t low) {
    try {
        v->v->ringAdd(v->u->eps[(size_t)tag], NodeId{high, low});
    } catch (const UUIDAlreadySeenException&) { return -4; }
    catch (const NodeAlreadyInRingException&) { return -3; }
    return 0;
}
int32_t orc_view_ring_delete(orc_view* v, int32_t tag) {
    try { v->v->ringDelete(v->u->eps[(size_t)tag]); }
    catch (const NodeNotInRingException&) { return -2; }
    return 0;
}
int32_t orc_view_observers(orc_view* v, int32_t tag, int32_t* out, int32_t cap) {
    try { return emit_tags(v->u, v->v->getObserversOf(v->u->eps[(size_t)tag]), out, cap); }
    catch (const NodeNotInRingException&) { return -2; }
}
int32_t orc_view_subjects(orc_view* v, int32_t tag, int32_t* out, int32_t cap) {
    try { return emit_tags(v->u, v->v->getSubjectsOf(v->u->eps[(size_t)tag]), out, cap); }
    catch (const NodeNotInRingException&) { return -2; }
}
int32_t orc_view_expected_observers(orc_view* v, int32_t tag, int32_t* out, int32_t cap) {
    return emit_tags(v->u, v->v->getExpectedObserversOf(v->u->eps[(size_t)tag]), out, cap);
}
int32_t orc_view_ring(orc_view* v, int32_t k, int32_t* out, int32_t cap) {
    return emit_tags(v->u, v->v->getRing(k), out, cap);
}
int32_t orc_view_ring_numbers(orc_view* v, int32_t observer, int32_t subject, int32_t* out, int32_t cap) {
    try {
        const std::vector<int> r = v->v->getRingNumbers(v->u->eps[(size_t)observer], v->u->eps[(size_t)subject]);
        for (size_t i = 0; i < r.size() && (int32_t)i < cap; ++i) out[i] = r[i];
        return (int32_t)r.size();
    } catch (const NodeNotInRingException&) { return -2; }
}
int32_t orc_view_size(orc_view* v) { return v->v->getMembershipSize(); }
int32_t orc_view_is_present(orc_view* v, int32_t tag) { return v->v->isHostPresent(v->u->eps[(size_t)tag]) ? 1 : 0; }
int32_t orc_view_is_safe_to_join(orc_view* v, int32_t tag, int64_t high, int64_t low) {
    return v->v->isSafeToJoin(v->u->eps[(size_t)tag], NodeId{high, low});
}
int64_t orc_view_config_id(orc_view* v) { return v->v->getCurrentConfigurationId(); }
int64_t orc_view_key(orc_view* v, int32_t k, int32_t tag) {
    return v->v->comparator(k).computeHash(v->u->eps[(size_t)tag]);
}
/* all observers / subjects of members, [n][K] in the order of `tags` (bulk helper for generators) */
int32_t orc_view_tables(orc_view* v, const int32_t* tags, int64_t n, int32_t* out_obs, int32_t* out_subj) {
    const int K = v->v->K();
    for (int64_t i = 0; i < n; ++i) {
        const Endpoint& e = v->u->eps[(size_t)tags[i]];
        try {
            const std::vector<Endpoint>& o = v->v->getObserversOf(e);
            const std::vector<Endpoint> s = v->v->getSubjectsOf(e);
            for (int k = 0; k < K; ++k) {
                out_obs[i * K + k] = o.empty() ? -1 : v->u->tagOf(o[(size_t)k]);
                out_subj[i * K + k] = s.empty() ? -1 : v->u->tagOf(s[(size_t)k]);
            }
        } catch (const NodeNotInRingException&) { return -2; }
    }
    return 0;
}

/* ---------------- MultiNodeCutDetector ---------------- */
orc_cd* orc_cd_create(orc_universe* u, int32_t K, int32_t H, int32_t L) {
    try {
        orc_cd* c = new orc_cd();
        c->u = u;
        c->cd.reset(new MultiNodeCutDetector(K, H, L));
        return c;
    } catch (const std::invalid_argument&) { return nullptr; }
}
void orc_cd_destroy(orc_cd* c) { delete c; }
int32_t orc_cd_aggregate(orc_cd* c, int32_t src, int32_t dst, int32_t status, const int32_t* rings,
                         int32_t n_rings, int32_t* out, int32_t cap) {
    AlertMessage m;
    m.edgeSrc = c->u->eps[(size_t)src];
    m.edgeDst = c->u->eps[(size_t)dst];
    m.edgeStatus = status;
    m.ringNumber.assign(rings, rings + n_rings);
    return emit_tags(c->u, c->cd->aggregateForProposal(m), out, cap);
}
int32_t orc_cd_invalidate(orc_cd* c, orc_view* v, int32_t* out, int32_t cap) {
    return emit_tags(c->u, c->cd->invalidateFailingEdges(*v->v), out, cap);
}
int32_t orc_cd_num_proposals(orc_cd* c) { return c->cd->getNumProposals(); }
void orc_cd_clear(orc_cd* c) { c->cd->clear(); }
uint32_t orc_cd_report_mask(orc_cd* c, int32_t tag) { return c->cd->reportMask(c->u->eps[(size_t)tag]); }

/* ---------------- batch handler (MembershipService.handleMessage(BatchedAlertMessage)) ---------------- */
orc_handler* orc_handler_create(orc_view* v, int32_t K, int32_t H, int32_t L) {
    try {
        orc_handler* h = new orc_handler();
        h->view = v;
        h->h.reset(new AlertBatchHandler(v->v.get(), K, H, L));
        return h;
    } catch (const std::invalid_argument&) { return nullptr; }
}
void orc_handler_destroy(orc_handler* h) { delete h; }

static void build_messages(const orc_universe* u, int64_t n_msgs, const int32_t* src, const int32_t* dst,
                           const int32_t* status, const int64_t* cfg, const int32_t* ring_off,
                           const int32_t* rings, std::vector<AlertMessage>& out) {
    out.resize((size_t)n_msgs);
    for (int64_t i = 0; i < n_msgs; ++i) {
        AlertMessage& m = out[(size_t)i];
        m.edgeSrc = u->eps[(size_t)src[i]];
        m.edgeDst = u->eps[(size_t)dst[i]];
        m.edgeStatus = status[i];
        m.configurationId = cfg[i];
        m.ringNumber.assign(rings + ring_off[i], rings + ring_off[i + 1]);
    }
}

int32_t orc_handler_batch(orc_handler* h, int64_t n_msgs, const int32_t* src, const int32_t* dst,
                          const int32_t* status, const int64_t* cfg, const int32_t* ring_off,
                          const int32_t* rings, int32_t* out, int32_t cap) {
    std::vector<AlertMessage> batch;
    build_messages(h->view->u, n_msgs, src, dst, status, cfg, ring_off, rings, batch);
    return emit_tags(h->view->u, h->h->handleBatch(batch), out, cap);
}
int32_t orc_handler_announced(orc_handler* h) { return h->h->announcedProposal() ? 1 : 0; }
void orc_handler_reset(orc_handler* h) { h->h->reset(); }
int32_t orc_handler_num_proposals(orc_handler* h) { return h->h->detector().getNumProposals(); }
uint32_t orc_handler_report_mask(orc_handler* h, int32_t tag) {
    return h->h->detector().reportMask(h->view->u->eps[(size_t)tag]);
}

/* ---------------- FastPaxos fast-round tally ---------------- */
orc_fp* orc_fp_create(orc_universe* u, int64_t cfg, int32_t membership_size) {
    orc_fp* f = new orc_fp();
    f->u = u;
    f->fp.reset(new FastPaxosTally(cfg, membership_size));
    return f;
}
void orc_fp_destroy(orc_fp* f) { delete f; }
int32_t orc_fp_vote(orc_fp* f, int32_t sender, int64_t cfg, const int32_t* tags, int32_t n) {
    std::vector<Endpoint> p;
    p.reserve((size_t)n);
    for (int32_t i = 0; i < n; ++i) p.push_back(f->u->eps[(size_t)tags[i]]);
    return f->fp->handleFastRoundProposal(f->u->eps[(size_t)sender], cfg, p) ? 1 : 0;
}
int32_t orc_fp_decided(orc_fp* f) { return f->fp->decided() ? 1 : 0; }
int32_t orc_fp_votes_received(orc_fp* f) { return f->fp->votesReceived(); }
int32_t orc_fp_decision(orc_fp* f, int32_t* out, int32_t cap) { return emit_tags(f->u, f->fp->decision(), out, cap); }
int32_t orc_fp_votes_for(orc_fp* f, const int32_t* tags, int32_t n) {
    std::vector<Endpoint> p;
    for (int32_t i = 0; i < n; ++i) p.push_back(f->u->eps[(size_t)tags[i]]);
    return f->fp->votesFor(p);
}

/* ---------------- virtual-cluster simulation: R independent AlertBatchHandlers ----------------
 * The cluster-scale driver used for parity at R > 1 and as the timed CPU baseline: R virtual nodes,
 * each a literal AlertBatchHandler, sharing one read-only MembershipView (every process of a
 * configuration holds an identical copy).  Cells are single-ring AlertMessages (a message with r ring
 * numbers == r cells in order, MultiNodeCutDetector.java:79-80).
 *
 * Delivery (same meaning as rapid_delivery in include/rapid_b200.h):
 *   blocked[R]   != NULL: receiver r with blocked[r] != 0 receives nothing
 *   bitmap       != NULL: [A][ceil(R/32)] uint32, bit (r & 31) of word r >> 5 set => cell delivered
 *   permuted     != 0   : receiver r applies its delivered cells in ascending
 *                         splitmix64( splitmix64(perm_seed + receiver_base + r) ^ cell_index )
 */
struct orc_sim {
    orc_view* view;
    int64_t R;
    std::vector<std::unique_ptr<AlertBatchHandler>> nodes;
};

orc_sim* orc_sim_create(orc_view* v, int32_t K, int32_t H, int32_t L, int64_t R) {
    try {
        orc_sim* s = new orc_sim();
        s->view = v;
        s->R = R;
        s->nodes.resize((size_t)R);
        for (int64_t r = 0; r < R; ++r) s->nodes[(size_t)r].reset(new AlertBatchHandler(v->v.get(), K, H, L));
        return s;
    } catch (const std::invalid_argument&) { return nullptr; }
}
void orc_sim_destroy(orc_sim* s) { delete s; }
void orc_sim_reset(orc_sim* s) { for (auto& n : s->nodes) n->reset(); }

/* out_len[R]: proposal length announced by THIS batch (0 if none); out_announced[R]: flag after the batch;
 * out_ids: proposals concatenated in receiver order, canonical (ring-0) order, capacity out_cap
 * (returns total ids written, or -1 if out_cap too small).  seconds_out: wall time of the parallel section. */
int64_t orc_sim_apply_batch(orc_sim* s, int64_t A, const int32_t* src, const int32_t* dst, const uint8_t* ring,
                            const uint8_t* status, const int64_t* cfg,
                            const uint8_t* blocked, const uint32_t* bitmap, int32_t permuted, uint64_t perm_seed,
                            int64_t receiver_base, int32_t n_threads,
                            int32_t* out_len, uint8_t* out_announced, int32_t* out_ids, int64_t out_cap,
                            double* seconds_out) {
    const orc_universe* u = s->view->u;
    MembershipView& view = *s->view->v;
    // One single-ring AlertMessage per cell.
    std::vector<AlertMessage> cells((size_t)A);
    for (int64_t i = 0; i < A; ++i) {
        AlertMessage& m = cells[(size_t)i];
        m.edgeSrc = u->eps[(size_t)src[i]];
        m.edgeDst = u->eps[(size_t)dst[i]];
        m.edgeStatus = status[i];
        m.configurationId = cfg[i];
        m.ringNumber.assign(1, (int32_t)ring[i]);
    }
    // Warm the view's memo caches single-threaded so the parallel section only reads them
    // (the Java view is per-process; sharing one copy read-only is the only liberty taken).
    view.getCurrentConfigurationId();
    for (int64_t i = 0; i < A; ++i) {
        const Endpoint& d = cells[(size_t)i].edgeDst;
        view.getRingZeroComparator().hashOf(d);
        if (view.isHostPresent(d)) view.getObserversOf(d);
        else { for (int k = 0; k < view.K(); ++k) view.comparator(k).hashOf(d); }
    }
    const int64_t R = s->R;
    const int64_t words = (R + 31) / 32;
    std::vector<std::vector<Endpoint>> props((size_t)R);
    const int nt = n_threads < 1 ? 1 : n_threads;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<AlertMessage> mine;
        std::vector<std::pair<uint64_t, int64_t>> order;
        for (;;) {
            const int64_t r0 = next.fetch_add(64);
            if (r0 >= R) break;
            const int64_t r1 = std::min(R, r0 + 64);
            for (int64_t r = r0; r < r1; ++r) {
                if (blocked && blocked[r]) continue;
                const std::vector<AlertMessage>* batch = &cells;
                if (bitmap || permuted) {
                    order.clear();
                    const uint64_t rs = splitmix64(perm_seed + (uint64_t)(receiver_base + r));
                    for (int64_t i = 0; i < A; ++i) {
                        if (bitmap && !((bitmap[(size_t)(i * words + (r >> 5))] >> (r & 31)) & 1u)) continue;
                        order.emplace_back(permuted ? splitmix64(rs ^ (uint64_t)i) : (uint64_t)i, i);
                    }
                    if (permuted) std::sort(order.begin(), order.end());
                    mine.clear();
                    for (const auto& kv : order) mine.push_back(cells[(size_t)kv.second]);
                    batch = &mine;
                }
                props[(size_t)r] = s->nodes[(size_t)r]->handleBatch(*batch);
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (nt == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    int64_t w = 0;
    for (int64_t r = 0; r < R; ++r) {
        const std::vector<Endpoint>& p = props[(size_t)r];
        if (out_len) out_len[r] = (int32_t)p.size();
        if (out_announced) out_announced[r] = s->nodes[(size_t)r]->announcedProposal() ? 1 : 0;
        for (const Endpoint& e : p) {
            if (out_ids) {
                if (w >= out_cap) return -1;
                out_ids[w] = u->tagOf(e);
            }
            ++w;
        }
    }
    return w;
}

uint32_t orc_sim_report_mask(orc_sim* s, int64_t r, int32_t tag) {
    return s->nodes[(size_t)r]->detector().reportMask(s->view->u->eps[(size_t)tag]);
}
int32_t orc_sim_num_proposals(orc_sim* s, int64_t r) { return s->nodes[(size_t)r]->detector().getNumProposals(); }
int32_t orc_sim_updates_in_progress(orc_sim* s, int64_t r) { return s->nodes[(size_t)r]->detector().updatesInProgress(); }

/* Timed CPU baseline for the vote tally: `n_nodes` literal FastPaxosTally instances each receive the
 * same `n_votes` votes (vote v = sender tag, proposal = prop_ids[prop_off[pid[v]] .. prop_off[pid[v]+1]) ).
 * Returns number of instances that decided; seconds_out = wall time. */
int32_t orc_sim_tally(orc_universe* u, int64_t cfg, int32_t membership_size, int32_t n_nodes, int64_t n_votes,
                      const int32_t* sender, const int64_t* vote_cfg, const int32_t* pid,
                      const int32_t* prop_off, const int32_t* prop_ids, int32_t n_threads,
                      int32_t* out_decided_pid, int32_t* out_votes_received, double* seconds_out) {
    // Materialise each vote's endpoint list the way a deserialised FastRoundPhase2bMessage holds it.
    int32_t n_props = 0;
    for (int64_t v = 0; v < n_votes; ++v) n_props = std::max(n_props, pid[v] + 1);
    std::vector<std::vector<Endpoint>> plist((size_t)n_props);
    for (int32_t p = 0; p < n_props; ++p)
        for (int32_t j = prop_off[p]; j < prop_off[p + 1]; ++j) plist[(size_t)p].push_back(u->eps[(size_t)prop_ids[j]]);
    std::vector<int32_t> decided_pid((size_t)n_nodes, -1), received((size_t)n_nodes, 0);
    std::atomic<int32_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const int32_t i = next.fetch_add(1);
            if (i >= n_nodes) break;
            FastPaxosTally fp(cfg, membership_size);
            for (int64_t v = 0; v < n_votes; ++v) {
                if (fp.handleFastRoundProposal(u->eps[(size_t)sender[v]], vote_cfg[v], plist[(size_t)pid[v]]))
                    decided_pid[(size_t)i] = pid[v];
            }
            received[(size_t)i] = fp.votesReceived();
        }
    };
    const int nt = n_threads < 1 ? 1 : n_threads;
    const auto t0 = std::chrono::steady_clock::now();
    if (nt == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    int32_t nd = 0;
    for (int32_t i = 0; i < n_nodes; ++i) {
        if (out_decided_pid) out_decided_pid[i] = decided_pid[(size_t)i];
        if (out_votes_received) out_votes_received[i] = received[(size_t)i];
        if (decided_pid[(size_t)i] >= 0) ++nd;
    }
    return nd;
}

/* ---------------- classic Paxos (Paxos.java) — one literal instance per node ---------------- */
struct orc_px {
    orc_universe* u;
    std::unique_ptr<ClassicPaxos> px;
};
static Value value_of(const orc_universe* u, const int32_t* tags, int32_t n) {
    Value v;
    v.reserve((size_t)(n > 0 ? n : 0));
    for (int32_t i = 0; i < n; ++i) v.push_back(u->eps[(size_t)tags[i]]);
    return v;
}
orc_px* orc_px_create(orc_universe* u, int32_t my_tag, int32_t my_hash, int64_t cfg, int32_t N) {
    orc_px* p = new orc_px();
    p->u = u;
    p->px.reset(new ClassicPaxos(u->eps[(size_t)my_tag], my_hash, cfg, N));
    return p;
}
void orc_px_destroy(orc_px* p) { delete p; }
/* out_rank[2] = crnd */
int32_t orc_px_start_phase1a(orc_px* p, int32_t round, int32_t* out_rank) {
    Phase1aMessage m;
    if (!p->px->startPhase1a(round, &m)) return 0;
    out_rank[0] = m.rank.round; out_rank[1] = m.rank.nodeIndex;
    return 1;
}
/* reply: out_ranks[4] = rnd, vrnd; out_tags/out_len = vval */
int32_t orc_px_phase1a(orc_px* p, int32_t sender, int64_t cfg, int32_t round, int32_t node, int32_t* out_ranks,
                       int32_t* out_tags, int32_t cap, int32_t* out_len) {
    Phase1aMessage m;
    m.sender = p->u->eps[(size_t)sender]; m.configurationId = cfg; m.rank.round = round; m.rank.nodeIndex = node;
    Phase1bMessage r;
    if (!p->px->handlePhase1aMessage(m, &r)) return 0;
    out_ranks[0] = r.rnd.round; out_ranks[1] = r.rnd.nodeIndex; out_ranks[2] = r.vrnd.round; out_ranks[3] = r.vrnd.nodeIndex;
    *out_len = emit_tags(p->u, r.vval, out_tags, cap);
    return 1;
}
/* ranks[4] = rnd, vrnd.  returns 1 iff a Phase2aMessage (rnd = out_rank, vval = out_tags) is broadcast */
int32_t orc_px_phase1b(orc_px* p, int32_t sender, int64_t cfg, const int32_t* ranks, const int32_t* tags, int32_t n,
                       int32_t* out_rank, int32_t* out_tags, int32_t cap, int32_t* out_len) {
    Phase1bMessage m;
    m.sender = p->u->eps[(size_t)sender]; m.configurationId = cfg;
    m.rnd.round = ranks[0]; m.rnd.nodeIndex = ranks[1]; m.vrnd.round = ranks[2]; m.vrnd.nodeIndex = ranks[3];
    m.vval = value_of(p->u, tags, n);
    Phase2aMessage o;
    if (!p->px->handlePhase1bMessage(m, &o)) return 0;
    out_rank[0] = o.rnd.round; out_rank[1] = o.rnd.nodeIndex;
    *out_len = emit_tags(p->u, o.vval, out_tags, cap);
    return 1;
}
/* returns 1 iff a Phase2bMessage (same rnd, same value) is broadcast */
int32_t orc_px_phase2a(orc_px* p, int32_t sender, int64_t cfg, int32_t round, int32_t node, const int32_t* tags, int32_t n) {
    Phase2aMessage m;
    m.sender = p->u->eps[(size_t)sender]; m.configurationId = cfg; m.rnd.round = round; m.rnd.nodeIndex = node;
    m.vval = value_of(p->u, tags, n);
    Phase2bMessage o;
    return p->px->handlePhase2aMessage(m, &o) ? 1 : 0;
}
/* returns 1 iff this message made the node decide */
int32_t orc_px_phase2b(orc_px* p, int32_t sender, int64_t cfg, int32_t round, int32_t node, const int32_t* tags, int32_t n) {
    Phase2bMessage m;
    m.sender = p->u->eps[(size_t)sender]; m.configurationId = cfg; m.rnd.round = round; m.rnd.nodeIndex = node;
    m.endpoints = value_of(p->u, tags, n);
    return p->px->handlePhase2bMessage(m) ? 1 : 0;
}
void orc_px_register_fast_round_vote(orc_px* p, const int32_t* tags, int32_t n) {
    p->px->registerFastRoundVote(value_of(p->u, tags, n));
}
int32_t orc_px_decided(orc_px* p) { return p->px->decided() ? 1 : 0; }
int32_t orc_px_decision(orc_px* p, int32_t* out, int32_t cap) { return emit_tags(p->u, p->px->decision(), out, cap); }
int32_t orc_px_vval(orc_px* p, int32_t* out, int32_t cap) { return emit_tags(p->u, p->px->vval(), out, cap); }
int32_t orc_px_cval(orc_px* p, int32_t* out, int32_t cap) { return emit_tags(p->u, p->px->cval(), out, cap); }
/* out[6] = rnd, vrnd, crnd */
void orc_px_ranks(orc_px* p, int32_t* out) {
    const Rank a = p->px->rnd(), b = p->px->vrnd(), c = p->px->crnd();
    out[0] = a.round; out[1] = a.nodeIndex; out[2] = b.round; out[3] = b.nodeIndex; out[4] = c.round; out[5] = c.nodeIndex;
}
/* selectProposalUsingCoordinatorRule over n messages: vrnd[2n], vval of message i = tags[off[i]..off[i+1]).
 * returns the chosen value's length (tags to out), -1 if the list was empty (IllegalArgumentException). */
int32_t orc_px_coordinator_rule(orc_px* p, int32_t n, const int32_t* vrnd, const int32_t* off, const int32_t* tags,
                                int32_t* out, int32_t cap) {
    std::vector<Phase1bMessage> msgs((size_t)n);
    for (int32_t i = 0; i < n; ++i) {
        msgs[(size_t)i].vrnd.round = vrnd[2 * i]; msgs[(size_t)i].vrnd.nodeIndex = vrnd[2 * i + 1];
        msgs[(size_t)i].vval = value_of(p->u, tags + off[i], off[i + 1] - off[i]);
    }
    try {
        return emit_tags(p->u, p->px->selectProposalUsingCoordinatorRule(msgs), out, cap);
    } catch (const std::invalid_argument&) {
        return -1;
    }
}

/* ---------------- alert generation: one FdNode (K PingPongFailureDetectors) per member ----------------
 * node_flags[tag]: bit0 crashed (answers no probe, runs no detector), bit1 ingress blocked (probes TO it fail),
 * bit2 egress blocked (probes FROM it fail), bit3 bootstrapping (answers BOOTSTRAPPING).
 * edge_fail[tag * K + k] != 0: the probe of `tag`'s k-th detector fails regardless. */
struct orc_fdsim {
    orc_view* view;
    int K;
    std::vector<int32_t> members;                 // tags, in the order the nodes are ticked
    std::vector<std::unique_ptr<FdNode>> nodes;
};
orc_fdsim* orc_fdsim_create(orc_view* v, int32_t K, const int32_t* member_tags, int64_t n) {
    orc_fdsim* s = new orc_fdsim();
    s->view = v; s->K = K;
    for (int64_t i = 0; i < n; ++i) {
        s->members.push_back(member_tags[i]);
        s->nodes.emplace_back(new FdNode(v->v.get(), v->u->eps[(size_t)member_tags[i]]));
    }
    return s;
}
void orc_fdsim_destroy(orc_fdsim* s) { delete s; }
/* one interval for every live node, nodes in member order.  Outputs (capacity cap alerts / cap_rings ring numbers):
 * observer tag, subject tag, ring_off[n_alerts + 1], rings.  Returns the number of alerts (may exceed cap: truncated). */
int64_t orc_fdsim_tick(orc_fdsim* s, const uint8_t* node_flags, const uint8_t* edge_fail, int64_t cfg, int32_t* observer,
                       int32_t* subject, int32_t* ring_off, int32_t* rings, int64_t cap, int64_t cap_rings) {
    std::vector<AlertMessage> out;
    for (size_t i = 0; i < s->nodes.size(); ++i) {
        const int32_t me = s->members[i];
        if (node_flags[me] & 1) continue;                                   // a crashed process runs nothing
        size_t k = 0;
        // the callback sees the detectors in creation order: k counts the probes of this tick
        const std::vector<PingPongFailureDetector>& fds = s->nodes[i]->detectors();
        std::vector<int> probing;                                           // detector indexes that will probe this tick
        for (size_t j = 0; j < fds.size(); ++j)
            if (!(fds[j].failureCount() >= PingPongFailureDetector::FAILURE_THRESHOLD && !fds[j].notified())) probing.push_back((int)j);
        s->nodes[i]->tick(
            [&](const Endpoint&, const Endpoint& subj) -> ProbeOutcome {
                const int j = probing[k++];
                const int32_t st = s->view->u->tagOf(subj);
                if (edge_fail && edge_fail[(size_t)me * (size_t)s->K + (size_t)j]) return PROBE_FAILED;
                if ((node_flags[me] & 4) || (node_flags[st] & 3)) return PROBE_FAILED;
                if (node_flags[st] & 8) return PROBE_BOOTSTRAPPING;
                return PROBE_OK;
            },
            cfg, &out);
    }
    int64_t n = 0, nr = 0;
    if (ring_off && cap > 0) ring_off[0] = 0;
    for (const AlertMessage& m : out) {
        if (n < cap) {
            observer[n] = s->view->u->tagOf(m.edgeSrc);
            subject[n] = s->view->u->tagOf(m.edgeDst);
            for (int32_t r : m.ringNumber) { if (nr < cap_rings) rings[nr] = r; ++nr; }
            ring_off[n + 1] = (int32_t)nr;
        }
        ++n;
    }
    return n;
}
/* state of node i's k-th detector: out[0] = failureCount, out[1] = notified */
void orc_fdsim_state(orc_fdsim* s, int64_t i, int32_t k, int32_t* out) {
    const PingPongFailureDetector& fd = s->nodes[(size_t)i]->detectors()[(size_t)k];
    out[0] = fd.failureCount(); out[1] = fd.notified() ? 1 : 0;
}
int32_t orc_fdsim_num_detectors(orc_fdsim* s, int64_t i) { return (int32_t)s->nodes[(size_t)i]->detectors().size(); }

int32_t orc_hardware_threads(void) { return (int32_t)std::thread::hardware_concurrency(); }

}  // extern "C"
