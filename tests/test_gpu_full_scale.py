"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle cannot finish these sizes in
seconds): every live virtual node announces the SAME proposal, its fingerprint is the fingerprint of the injected cut, the
canonical list is sorted by the ring-0 key, a repeated batch is ignored (announcedProposal), the fast round decides that cut
with exactly quorum votes counted, and the sweep kernel agrees with the bucketed kernels on a slice of the receivers."""
import numpy as np
import pytest

from rapid_b200 import workloads as W

pytestmark = pytest.mark.gpu
K, H, L = 10, 9, 4


def _view(rb, n, nj=0):
    hb, off, ports = W.packed_endpoints(0, n + nj)
    v = rb.MembershipView.from_packed(K, hb[: off[n]], off[: n + 1], ports[:n])
    if nj:
        hosts, jports = W.endpoints(n, nj)
        v.registerJoiners(hosts, jports)
    return v


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


def test_c5_one_million_nodes():
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


def test_c3_ten_thousand_nodes_correlated_partition():
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
    _check_converged(rb, v, cl, b, cfg, blocked)
    assert cl.debugStats()[1] > 0                                   # the cut came out of invalidateFailingEdges
    fp = rb.FastPaxos(cfg, n)
    cl.clear()
    cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, read_outputs=False)
    t = fp.tallyCluster(cl)
    assert t.decided and t.length == 500 and t.count == rb.quorum(n)


def test_c4_hundred_thousand_nodes_flip_flop_stream():
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
    for b in batches:
        res = cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, perm_seed=b.meta["perm_seed"])
        assert cl.lastPath()[0] == 4                                # per-receiver order, every cell to everyone: uniform kernel, moments on demand
        ann = res.proposal_len > 0
        # whoever announces in a batch announces a subset of the flapping nodes; once everything is in, the whole set
        assert (res.proposal_len[ann] <= len(batches[-1].expected_cut)).all()
        t = fp.tallyCluster(cl)
        if t.decided:
            decided = t
            break
    assert decided is not None and (decided.hash, decided.hash2) == want and decided.count == rb.quorum(n)
