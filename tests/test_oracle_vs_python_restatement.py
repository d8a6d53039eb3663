"""Two independent readings of the Java must agree: the C++ oracle (what the GPU is checked against) and tests/pyref.py, over
random alert streams — duplicates, stale configurations, UP about members / DOWN about strangers, joiners, bursts that emit in
the middle of a batch, several batches with the announcedProposal gating, clear() — and random vote streams."""
import random

import numpy as np
import pytest

import pyref
from helpers import OracleWorld

K = 10
HL = [(9, 4), (8, 2), (8, 3), (9, 3), (10, 1), (3, 3)]


def same_reports(ma, mb, H):
    """Report masks must agree — except BEYOND the high watermark.  invalidateFailingEdges walks a copy of preProposal in HashSet
    order (MultiNodeCutDetector.java:146-147) and keeps adding implicit reports to a node after it has reached H; when the pass
    emits in the middle (proposal.clear(), :118-119), the nodes still to be visited find nobody in proposal U preProposal any more.
    Which of an emitted node's rings beyond the H-th were recorded therefore depends on an iteration order the reference does not
    define — and nothing can observe it: a count above H never equals L or H again."""
    return ma == mb or (bin(ma).count("1") >= H and bin(mb).count("1") >= H)


def random_stream(rng, w, n, nj, cfg, n_msgs, focus):
    """alerts biased towards `focus` subjects reported by their true observers (so that watermarks are crossed), plus noise"""
    msgs = []
    for _ in range(n_msgs):
        kind = rng.random()
        if kind < 0.70:
            dst = rng.choice(focus)
            if dst < n:
                obs = w.view.getObserversOf(dst)
                status = pyref.DOWN
            else:
                obs = w.view.getExpectedObserversOf(dst)
                status = pyref.UP
            k = rng.randrange(K)
            src = obs[k]
            rings = [r for r in range(K) if obs[r] == src] if rng.random() < 0.5 else [k]
        else:                                                      # noise: any edge, any status, any ring subset
            dst = rng.randrange(n + nj)
            src = rng.randrange(n)
            status = rng.choice([pyref.UP, pyref.DOWN])
            rings = rng.sample(range(K), rng.randint(1, 3))
        c = cfg if rng.random() < 0.93 else cfg + rng.choice([-1, 1, 12345])
        msgs.append((src, dst, status, c, rings))
    return msgs


@pytest.mark.parametrize("seed", range(40))
def test_batch_handler_streams(orc, seed):
    rng = random.Random(1000 + seed)
    n, nj = rng.randint(8, 40), rng.randint(0, 4)
    H, L = HL[seed % len(HL)]
    w = OracleWorld(orc, n, K, n_joiners=nj)
    cfg = w.view.getCurrentConfigurationId()
    a = orc.AlertBatchHandler(w.view, K, H, L)
    b = pyref.PyBatchHandler(w.view, K, H, L)
    focus = rng.sample(range(n + nj), min(n + nj, rng.randint(1, 5)))
    emitted = 0
    for batch in range(rng.randint(2, 8)):
        if rng.random() < 0.15:
            a.reset(); b.reset()
        msgs = random_stream(rng, w, n, nj, cfg, rng.randint(0, 60), focus)
        got_a = a.handleBatch(msgs)
        got_b = b.handleBatch(msgs)
        assert set(got_a) == got_b and len(got_a) == len(got_b), (seed, batch)
        # canonical order of the proposal: ring-0 comparator (MembershipService.java:346-348)
        if got_a:
            key = {t: w.view.key(0, t) for t in got_a}
            assert got_a == sorted(got_a, key=lambda t: key[t])
            emitted += 1
        assert a.announced() == b.announcedProposal
        for t in range(n + nj):
            assert same_reports(a.reportMask(t), b.cd.reportMask(t), H), (seed, batch, t)
    assert emitted >= 0


@pytest.mark.parametrize("seed", range(25))
def test_raw_detector_streams(orc, seed):
    """the bare MultiNodeCutDetector: per-call emission lists (as sets: HashSet iteration order is not specified), the counter,
    explicit invalidation calls at random points"""
    rng = random.Random(2000 + seed)
    n, nj = rng.randint(6, 30), rng.randint(0, 3)
    H, L = HL[seed % len(HL)]
    w = OracleWorld(orc, n, K, n_joiners=nj)
    a = orc.MultiNodeCutDetector(w.u, K, H, L)
    b = pyref.PyCutDetector(K, H, L)
    focus = rng.sample(range(n + nj), min(n + nj, rng.randint(1, 4)))
    for step in range(rng.randint(20, 200)):
        if rng.random() < 0.05:
            ga, gb = a.invalidateFailingEdges(w.view), b.invalidateFailingEdges(w.view)
        elif rng.random() < 0.02:
            a.clear(); b.clear()
            continue
        else:
            (src, dst, status, _, rings), = random_stream(rng, w, n, nj, 0, 1, focus)
            if (status == pyref.DOWN) != (dst < n):                # the service filter would drop it; the raw detector is only ever
                continue                                           # fed consistent alerts (MembershipService.java:644-675)
            ga, gb = a.aggregateForProposal(src, dst, status, rings), b.aggregateForProposal(src, dst, status, rings)
        assert sorted(ga) == sorted(gb), (seed, step)
        assert a.getNumProposals() == b.getNumProposals()


def test_ctor_validation_agrees(orc):
    w = OracleWorld(orc, 5, K)
    for Kx, H, L in [(10, 11, 1), (10, 5, 6), (2, 2, 1), (10, 9, 0), (10, 0, 0), (3, 3, 3), (10, 10, 10)]:
        ok_py = True
        try:
            pyref.PyCutDetector(Kx, H, L)
        except ValueError:
            ok_py = False
        ok_orc = True
        try:
            orc.MultiNodeCutDetector(w.u, Kx, H, L)
        except ValueError:
            ok_orc = False
        assert ok_py == ok_orc, (Kx, H, L)


@pytest.mark.parametrize("seed", range(25))
def test_fast_paxos_vote_streams(orc, seed):
    rng = random.Random(3000 + seed)
    N = rng.randint(1, 60)
    u = orc.Universe()
    tags = [u.add("10.0.0.%d" % (i // 50), 1000 + i) for i in range(N + 5)]
    cfg = 77
    a = orc.FastPaxosTally(u, cfg, N)
    b = pyref.PyFastPaxos(cfg, N)
    props = [sorted(rng.sample(tags, rng.randint(1, 4))) for _ in range(rng.randint(1, 3))]
    if rng.random() < 0.5:
        props.append(list(reversed(props[0])))                     # same endpoints, different order: a DIFFERENT proposal (List.equals)
    weights = [10] + [1] * (len(props) - 1)
    for v in range(3 * N + 5):
        sender = rng.choice(tags)                                  # non-members may vote (FastPaxosWithoutFallbackTests.java:129-148)
        c = cfg if rng.random() < 0.9 else cfg + 1
        p = rng.choices(props, weights)[0]
        da, db = a.handleFastRoundProposal(sender, c, p), b.handleFastRoundProposal(sender, c, p)
        assert da == db, (seed, v)
        assert a.decided() == b.decided and a.votesReceived() == len(b.votesReceived)
        if b.decided:
            assert a.decision() == b.decision
