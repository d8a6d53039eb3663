"""The oracle's MembershipView against tests/pyref.PyMembershipView (python-xxhash, bisect lists): ring orders, observers,
subjects, expected observers of strangers, ring numbers, configuration ids — through random ringAdd / ringDelete sequences with
the UUID / already-in-ring / not-in-ring rules."""
import random

import numpy as np
import pytest

import pyref
from rapid_b200 import workloads as W

xxhash = pytest.importorskip("xxhash")


def make(orc, K, n_total, n_initial):
    hb, off, ports = W.packed_endpoints(0, n_total)
    u = orc.Universe()
    tags = u.add_bulk(hb, off, ports)
    hi, lo = W.node_ids(0, n_total)
    eps = [(int(t), bytes(hb[off[i]: off[i + 1]]), int(ports[i])) for i, t in enumerate(tags)]
    ids = [(int(hi[i]), int(lo[i])) for i in range(n_total)]
    ov = orc.MembershipView(u, K, np.arange(n_initial, dtype=np.int32), hi[:n_initial], lo[:n_initial])
    pv = pyref.PyMembershipView(K, eps[:n_initial], ids[:n_initial])
    for t, h, p in eps[n_initial:]:
        pv.know(t, h, p)
    return ov, pv, ids


def same_view(ov, pv, K, n_total, rng, sample=12):
    assert ov.getMembershipSize() == pv.getMembershipSize()
    for k in range(K):
        assert ov.getRing(k) == pv.getRing(k), k
    assert ov.getCurrentConfigurationId() == pv.getCurrentConfigurationId()
    members = pv.getRing(0)
    for t in rng.sample(members, min(sample, len(members))):
        assert ov.getObserversOf(t) == pv.getObserversOf(t)
        assert ov.getSubjectsOf(t) == pv.getSubjectsOf(t)
        if len(members) > 1:
            o = rng.choice(members)
            assert ov.getRingNumbers(o, t) == pv.getRingNumbers(o, t)
    for t in rng.sample(range(n_total), min(sample, n_total)):
        assert ov.isHostPresent(t) == pv.isHostPresent(t)
        assert ov.getExpectedObserversOf(t) == pv.getExpectedObserversOf(t)       # members and strangers alike (:292-303)


@pytest.mark.parametrize("K,n", [(10, 1), (10, 2), (10, 3), (10, 50), (10, 1000), (3, 40), (14, 200)])
def test_bulk_view(orc, K, n):
    ov, pv, _ = make(orc, K, n + 5, n)
    same_view(ov, pv, K, n + 5, random.Random(n))


@pytest.mark.parametrize("seed", range(12))
def test_random_view_changes(orc, seed):
    rng = random.Random(4000 + seed)
    K = rng.choice([10, 10, 7])
    n_total = rng.randint(5, 80)
    n0 = rng.randint(0, n_total)
    ov, pv, ids = make(orc, K, n_total, n0)
    same_view(ov, pv, K, n_total, rng)
    fresh = 10 ** 6
    for step in range(60):
        t = rng.randrange(n_total)
        if rng.random() < 0.5:
            nid = ids[t] if rng.random() < 0.6 else (fresh + step, step)           # a NodeId seen before is refused even after a delete
            oe = pe = None
            try:
                ov.ringAdd(t, nid)
            except (orc.UUIDAlreadySeenException, orc.NodeAlreadyInRingException) as e:
                oe = type(e).__name__
            try:
                pv.ringAdd(t, nid)
            except KeyError:
                pe = "UUIDAlreadySeenException"
            except ValueError:
                pe = "NodeAlreadyInRingException"
            assert oe == pe, (seed, step)
        else:
            oe = pe = None
            try:
                ov.ringDelete(t)
            except orc.NodeNotInRingException:
                oe = "x"
            try:
                pv.ringDelete(t)
            except LookupError:
                pe = "x"
            assert oe == pe, (seed, step)
        if step % 6 == 0:
            same_view(ov, pv, K, n_total, rng, sample=6)
    same_view(ov, pv, K, n_total, rng)


def test_cached_observers_survive_a_wrap_around_change(orc):
    """The reference caches getObserversOf per node and invalidates, on ringAdd / ringDelete, the cache of `endpoints.lower(node)`
    on every ring (MembershipView.java:133-147, :175-192).  At the FRONT of a ring lower() is null, so the ring's LAST node — whose
    successor wraps around to first() — keeps a cached list that names the old first node.  Both restatements reproduce this
    (they restate the Java, not its intent); a view built afresh from the same members answers from the rings.  The device tables
    (rapid_view_apply_cut -> k_tables) are always recomputed from the rings, i.e. they follow the fresh answer: DESIGN.md §7."""
    K = 10
    n_total = 40
    rng = random.Random(7)
    for attempt in range(200):
        n0 = rng.randint(3, n_total - 1)
        ov, pv, ids = make(orc, K, n_total, n0)
        new = rng.randrange(n0, n_total)
        # a ring on which the stranger would become the first node
        front = [k for k in range(K) if pv._key(k, new) < pv.rings[k][0][0]]
        if not front:
            continue
        k = front[0]
        last = pv.rings[k][-1][1]
        before_o, before_p = list(ov.getObserversOf(last)), list(pv.getObserversOf(last))     # both caches now hold the list
        assert before_o == before_p and before_p[k] == pv.rings[k][0][1]
        ov.ringAdd(new, ids[new]); pv.ringAdd(new, ids[new])
        after_o, after_p = list(ov.getObserversOf(last)), list(pv.getObserversOf(last))
        assert after_o == after_p                                                              # the two readings agree ...
        fresh = pv.computeObserversOf(last)
        assert fresh[k] == new
        if all(pv._lower_no_wrap(kk, new) != last for kk in range(K)):                         # ... on the STALE list, unless another ring invalidated it
            assert after_p[k] == before_p[k] != fresh[k]
            return
    pytest.skip("no wrap-around insertion found")
