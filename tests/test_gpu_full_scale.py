"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle cannot finish these sizes in
seconds): every live virtual node announces the SAME proposal, its fingerprint is the fingerprint of the injected cut, the
canonical list is sorted by the ring-0 key, a repeated batch is ignored (announcedProposal), the fast round decides that cut
with exactly quorum votes counted, and the sweep kernel agrees with the bucketed kernels on a slice of the receivers."""
import numpy as np
import pytest

from rapid_b200 import workloads as W
from helpers import fingerprints_from_oracle

pytestmark = pytest.mark.gpu
K, H, L = 10, 9, 4


def _view(rb, n, nj=0):
    hb, off, ports = W.packed_endpoints(0, n + nj)
    v = rb.MembershipView.from_packed(K, hb[: off[n]], off[: n + 1], ports[:n])
    if nj:
        hosts, jports = W.endpoints(n, nj)
        v.registerJoiners(hosts, jports)
    return v


SAMPLE = 256


def _oracle_view(orc, n, nj=0):
    hb, off, ports = W.packed_endpoints(0, n + nj)
    u = orc.Universe()
    tags = u.add_bulk(hb, off, ports)
    hi, lo = W.node_ids(0, n)
    return u, orc.MembershipView(u, K, tags[:n], hi, lo)


def _sampled_oracle_check(orc, rb, oview, cl, res, window_begin, batch, cfg, blocked, perm_seed=None, sim=None):
    """Per-receiver parity AT FULL SCALE: a window of SAMPLE receivers is run through the literal oracle on the same batch and
    compared receiver by receiver (length, both fingerprint words, announced flag, canonical list of one announcer, report
    masks + updatesInProgress of a few receivers that have not announced)."""
    if sim is None:
        sim = orc.ClusterSim(oview, K, H, L, SAMPLE, receiver_base=cl.receiver_begin + window_begin)
    o_len, o_ann, o_ids, o_off = sim.apply_batch(batch.src, batch.dst, batch.ring, batch.status, np.full(len(batch), cfg, np.int64),
                                                 blocked=blocked[window_begin: window_begin + SAMPLE], perm_seed=perm_seed, threads=8)
    sl = slice(window_begin, window_begin + SAMPLE)
    np.testing.assert_array_equal(res.proposal_len[sl], o_len)
    np.testing.assert_array_equal(res.announced[sl], o_ann)
    e1, e2 = fingerprints_from_oracle(rb, o_len, o_ids, o_off)
    np.testing.assert_array_equal(res.proposal_hash[sl], e1)
    np.testing.assert_array_equal(res.proposal_hash2[sl], e2)
    who = np.nonzero(o_len)[0]
    if len(who):
        r = int(who[len(who) // 2])
        assert cl.getProposal(window_begin + r, cap=int(o_len[r]) + 8) == o_ids[o_off[r]: o_off[r + 1]].tolist()
    quiet = np.nonzero(o_ann == 0)[0]
    for r in quiet[:: max(1, len(quiet) // 3)][:3]:
        for subj, m in cl.debugMasks(int(window_begin + r)).items():
            assert sim.reportMask(int(r), int(subj)) == m, "mask of subject %d at receiver %d" % (subj, window_begin + r)
        assert cl.debugCounters(int(window_begin + r))[0] == sim.updatesInProgress(int(r))
    return sim


def _check_converged(rb, v, cl, b, cfg, blocked, perm_seed=None):
    n = v.n
    res = cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, perm_seed=perm_seed)
    live = blocked == 0
    want = rb.proposal_fingerprint(b.expected_cut)
    assert (res.proposal_len[live] == len(b.expected_cut)).all() and (res.proposal_len[~live] == 0).all()
    assert (res.proposal_hash[live] == np.uint64(want[0])).all() and (res.proposal_hash2[live] == np.uint64(want[1])).all()
    assert (res.announced[live] == 1).all() and (res.announced[~live] == 0).all()
    # checksum of checksums: every live node contributed the same fingerprint
    assert int(res.proposal_len.sum()) == int(live.sum()) * len(b.expected_cut)
    # canonical order = ring-0 key order (MembershipService.java:346-348)
    r0 = int(np.nonzero(live)[0][len(np.nonzero(live)[0]) // 2])
    prop = cl.getProposal(r0, cap=len(b.expected_cut) + 8)
    assert sorted(prop) == b.expected_cut.tolist()
    keys0 = v.keys(0)[np.asarray(prop)]
    assert (np.diff(keys0) > 0).all()
    # idempotence: the same batch again is ignored by everyone who announced
    again = cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, perm_seed=perm_seed)
    assert (again.proposal_len == 0).all() and (again.announced == res.announced).all()
    return res, want


def test_c5_one_million_nodes(orc):
    import rapid_b200 as rb
    n = 1_000_000
    nj = n // 200
    v = _view(rb, n, nj)
    obs, _ = v.tables()
    b = W.c5_churn(obs, v.joinerTables(), n, n // 200, nj)
    hi, lo = W.node_ids(0, n)
    cfg = v.getCurrentConfigurationId(hi, lo)
    ring0 = v.getRing(0)
    blocked = W.blocked_by_receiver(b.blocked, ring0, 0, n)
    cl = rb.VirtualCluster(v, H, L, max_subjects=len(b.expected_cut) + 64)
    res, want = _check_converged(rb, v, cl, b, cfg, blocked)
    assert cl.lastPath()[0] == 2                                   # the subject-bucketed uniform kernel served it
    # per-receiver oracle parity on two windows of the million receivers (one of them straddling a 1024-receiver tile edge)
    _, oview = _oracle_view(orc, n, nj)
    assert oview.getCurrentConfigurationId() == cfg
    for w0 in (1024 * 300 - 100, 987_654):
        _sampled_oracle_check(orc, rb, oview, cl, res, w0, b, cfg, blocked)
    # fast round: the decision is that cut, taken at the quorum-th vote
    cl.clear()
    cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, read_outputs=False)
    fp = rb.FastPaxos(cfg, n)
    t = fp.tallyCluster(cl)
    assert t.decided and (t.hash, t.hash2, t.length) == (want[0], want[1], len(b.expected_cut))
    assert t.count == rb.quorum(n) == t.votes_received
    del cl
    # the per-cell sweep kernel on a slice of the receivers agrees bit for bit
    lo_r, cnt = 123_456, 4096
    sw = rb.VirtualCluster(v, H, L, n_receivers=cnt, receiver_begin=lo_r, kernel="sweep", max_subjects=len(b.expected_cut) + 64)
    r2 = sw.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked[lo_r: lo_r + cnt])
    assert (r2.proposal_hash == res.proposal_hash[lo_r: lo_r + cnt]).all()
    assert (r2.proposal_len == res.proposal_len[lo_r: lo_r + cnt]).all()


def test_c3_ten_thousand_nodes_correlated_partition(orc):
    import rapid_b200 as rb
    n = 10_000
    v = _view(rb, n)
    obs, _ = v.tables()
    ring0 = v.getRing(0)
    b = W.c3_correlated_partition(obs, ring0, n, 0.05)
    hi, lo = W.node_ids(0, n)
    cfg = v.getCurrentConfigurationId(hi, lo)
    blocked = W.blocked_by_receiver(b.blocked, ring0, 0, n)
    cl = rb.VirtualCluster(v, H, L)
    res, _ = _check_converged(rb, v, cl, b, cfg, blocked)
    assert cl.debugStats()[1] > 0                                   # the cut came out of invalidateFailingEdges
    _, oview = _oracle_view(orc, n)
    for w0 in (0, 5000 - 128, n - SAMPLE):                          # includes receivers inside and next to the partitioned arc
        _sampled_oracle_check(orc, rb, oview, cl, res, w0, b, cfg, blocked)
    fp = rb.FastPaxos(cfg, n)
    cl.clear()
    cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, read_outputs=False)
    t = fp.tallyCluster(cl)
    assert t.decided and t.length == 500 and t.count == rb.quorum(n)


def test_c4_hundred_thousand_nodes_flip_flop_stream(orc):
    import rapid_b200 as rb
    n = 100_000
    v = _view(rb, n)
    obs, _ = v.tables()
    ring0 = v.getRing(0)
    batches = W.c4_flip_flop_stream(obs, n, 0.01, T=8)
    hi, lo = W.node_ids(0, n)
    cfg = v.getCurrentConfigurationId(hi, lo)
    blocked = W.blocked_by_receiver(batches[0].blocked, ring0, 0, n)
    cl = rb.VirtualCluster(v, H, L)
    fp = rb.FastPaxos(cfg, n)
    want = rb.proposal_fingerprint(batches[-1].expected_cut)
    decided = None
    _, oview = _oracle_view(orc, n)
    windows = (1024 * 40 - 128, 77_777)
    sims = {}
    for b in batches:
        res = cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, perm_seed=b.meta["perm_seed"])
        for w0 in windows:                                          # state carried: the same oracle instances follow the whole stream
            sims[w0] = _sampled_oracle_check(orc, rb, oview, cl, res, w0, b, cfg, blocked, perm_seed=b.meta["perm_seed"], sim=sims.get(w0))
        assert cl.lastPath()[0] == 4                                # per-receiver order, every cell to everyone: uniform kernel, moments on demand
        ann = res.proposal_len > 0
        # whoever announces in a batch announces a subset of the flapping nodes; once everything is in, the whole set
        assert (res.proposal_len[ann] <= len(batches[-1].expected_cut)).all()
        t = fp.tallyCluster(cl)
        if t.decided:
            decided = t
            break
    assert decided is not None and (decided.hash, decided.hash2) == want and decided.count == rb.quorum(n)
