"""Port of the reference's CutDetectionTest (rapid/src/test/java/com/vrg/rapid/CutDetectionTest.java,
all 7 tests) against the oracle — these are the pins for oracle::MultiNodeCutDetector."""
import uuid

import pytest

K, H, L = 10, 8, 2
UP, DOWN = 0, 1


def _node_id():
    u = uuid.uuid4().int
    hi, lo = u >> 64, u & (2**64 - 1)
    to_s = lambda x: x - 2**64 if x >= 2**63 else x
    return (to_s(hi), to_s(lo))


def test_cut_detection(orc):                      # CutDetectionTest.java:43-59
    u = orc.Universe()
    wb = orc.MultiNodeCutDetector(u, K, H, L)
    dst = u.add("127.0.0.2", 2)
    for i in range(H - 1):
        ret = wb.aggregateForProposal(u.add("127.0.0.1", i + 1), dst, UP, i)
        assert len(ret) == 0 and wb.getNumProposals() == 0
    ret = wb.aggregateForProposal(u.add("127.0.0.1", H), dst, UP, H - 1)
    assert len(ret) == 1 and wb.getNumProposals() == 1


def test_blocking_one_blocker(orc):               # :62-91
    u = orc.Universe()
    wb = orc.MultiNodeCutDetector(u, K, H, L)
    dst1, dst2 = u.add("127.0.0.2", 2), u.add("127.0.0.3", 2)
    for d in (dst1, dst2):
        for i in range(H - 1):
            assert wb.aggregateForProposal(u.add("127.0.0.1", i + 1), d, UP, i) == []
            assert wb.getNumProposals() == 0
    assert wb.aggregateForProposal(u.add("127.0.0.1", H), dst1, UP, H - 1) == []
    assert wb.getNumProposals() == 0
    ret = wb.aggregateForProposal(u.add("127.0.0.1", H), dst2, UP, H - 1)
    assert len(ret) == 2 and wb.getNumProposals() == 1


def test_blocking_three_blockers(orc):            # :95-137
    u = orc.Universe()
    wb = orc.MultiNodeCutDetector(u, K, H, L)
    d = [u.add("127.0.0.%d" % j, 2) for j in (2, 3, 4)]
    for dst in d:
        for i in range(H - 1):
            assert wb.aggregateForProposal(u.add("127.0.0.1", i + 1), dst, UP, i) == []
            assert wb.getNumProposals() == 0
    src = u.add("127.0.0.1", H)
    assert wb.aggregateForProposal(src, d[0], UP, H - 1) == [] and wb.getNumProposals() == 0
    assert wb.aggregateForProposal(src, d[2], UP, H - 1) == [] and wb.getNumProposals() == 0
    ret = wb.aggregateForProposal(src, d[1], UP, H - 1)
    assert len(ret) == 3 and wb.getNumProposals() == 1


def test_blocking_multiple_blockers_past_h(orc):  # :140-189
    u = orc.Universe()
    wb = orc.MultiNodeCutDetector(u, K, H, L)
    d = [u.add("127.0.0.%d" % j, 2) for j in (2, 3, 4)]
    for dst in d:
        for i in range(H - 1):
            assert wb.aggregateForProposal(u.add("127.0.0.1", i + 1), dst, UP, i) == []
            assert wb.getNumProposals() == 0
    s0, s1 = u.add("127.0.0.1", H), u.add("127.0.0.1", H + 1)
    wb.aggregateForProposal(s0, d[0], UP, H - 1)
    assert wb.aggregateForProposal(s1, d[0], UP, H - 1) == [] and wb.getNumProposals() == 0
    wb.aggregateForProposal(s0, d[2], UP, H - 1)
    assert wb.aggregateForProposal(s1, d[2], UP, H - 1) == [] and wb.getNumProposals() == 0
    ret = wb.aggregateForProposal(s0, d[1], UP, H - 1)
    assert len(ret) == 3 and wb.getNumProposals() == 1


def test_below_l(orc):                            # :192-230
    u = orc.Universe()
    wb = orc.MultiNodeCutDetector(u, K, H, L)
    d = [u.add("127.0.0.%d" % j, 2) for j in (2, 3, 4)]
    for i in range(H - 1):
        assert wb.aggregateForProposal(u.add("127.0.0.1", i + 1), d[0], UP, i) == []
    for i in range(L - 1):
        assert wb.aggregateForProposal(u.add("127.0.0.1", i + 1), d[1], UP, i) == []
    for i in range(H - 1):
        assert wb.aggregateForProposal(u.add("127.0.0.1", i + 1), d[2], UP, i) == []
    assert wb.getNumProposals() == 0
    src = u.add("127.0.0.1", H)
    assert wb.aggregateForProposal(src, d[0], UP, H - 1) == [] and wb.getNumProposals() == 0
    ret = wb.aggregateForProposal(src, d[2], UP, H - 1)
    assert len(ret) == 2 and wb.getNumProposals() == 1


def test_batch(orc):                              # :234-252
    u = orc.Universe()
    wb = orc.MultiNodeCutDetector(u, K, H, L)
    endpoints = [u.add("127.0.0.2", 2 + i) for i in range(3)]
    src = u.add("127.0.0.1", 1)
    proposal = []
    for e in endpoints:
        for ring in range(K):
            proposal += wb.aggregateForProposal(src, e, UP, ring)
    assert len(proposal) == 3


def test_link_invalidation(orc):                  # :255-301
    u = orc.Universe()
    mview = orc.MembershipView(u, K)
    wb = orc.MultiNodeCutDetector(u, K, H, L)
    endpoints = []
    for i in range(30):
        n = u.add("127.0.0.2", 2 + i)
        endpoints.append(n)
        mview.ringAdd(n, _node_id())
    dst = endpoints[0]
    observers = mview.getObserversOf(dst)
    assert len(observers) == K
    for i in range(H - 1):
        assert wb.aggregateForProposal(observers[i], dst, DOWN, i) == []
        assert wb.getNumProposals() == 0
    failed = set()
    for i in range(H - 1, K):
        oo = mview.getObserversOf(observers[i])
        failed.add(observers[i])
        for j in range(K):
            assert wb.aggregateForProposal(oo[j], observers[i], DOWN, j) == []
            assert wb.getNumProposals() == 0
    ret = wb.invalidateFailingEdges(mview)
    assert len(ret) == 4            # a weak pin on the ring hash
    assert wb.getNumProposals() == 1
    for n in ret:
        assert n in failed or n == dst


def test_ctor_validation(orc):                    # MultiNodeCutDetector.java:51-55
    u = orc.Universe()
    for bad in ((10, 11, 2), (10, 8, 9), (2, 2, 1), (10, 8, 0), (10, 0, 0)):
        with pytest.raises(ValueError):
            orc.MultiNodeCutDetector(u, *bad)
    orc.MultiNodeCutDetector(u, 10, 9, 4)
    orc.MultiNodeCutDetector(u, 3, 3, 3)
