"""Port of MembershipViewTest (rapid/src/test/java/com/vrg/rapid/MembershipViewTest.java, 16 tests)
against oracle::MembershipView, plus the seed-answers-with-expected-observers property of
MessagingTest.java:150-182."""
import hashlib
import uuid

import pytest

K = 10


def _rid():
    u = uuid.uuid4().int
    s = lambda x: x - 2**64 if x >= 2**63 else x
    return (s(u >> 64), s(u & (2**64 - 1)))


def _name_id(name: str):
    # UUID.nameUUIDFromBytes == type-3 (MD5) UUID of the raw bytes
    d = bytearray(hashlib.md5(name.encode()).digest())
    d[6] = (d[6] & 0x0F) | 0x30
    d[8] = (d[8] & 0x3F) | 0x80
    v = int.from_bytes(d, "big")
    s = lambda x: x - 2**64 if x >= 2**63 else x
    return (s(v >> 64), s(v & (2**64 - 1)))


def test_one_ring_addition(orc):                               # :43-60
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    a = u.add("127.0.0.1", 123)
    mv.ringAdd(a, _rid())
    for k in range(K):
        assert mv.getRing(k) == [a]


def test_multiple_ring_additions(orc):                         # :65-81
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    for i in range(10):
        mv.ringAdd(u.add("127.0.0.1", i), _rid())
    for k in range(K):
        assert len(mv.getRing(k)) == 10


def test_ring_re_additions(orc):                               # :86-120
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    for i in range(10):
        mv.ringAdd(u.add("127.0.0.1", i), _rid())
    throws = 0
    for i in range(10):
        try:
            mv.ringAdd(u.add("127.0.0.1", i), _rid())
        except orc.NodeAlreadyInRingException:
            throws += 1
    assert throws == 10
    for k in range(K):
        assert len(mv.getRing(k)) == 10


def test_ring_deletions_only(orc):                             # :125-140
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    throws = 0
    for i in range(10):
        try:
            mv.ringDelete(u.add("127.0.0.1", i))
        except orc.NodeNotInRingException:
            throws += 1
    assert throws == 10


def test_ring_additions_and_deletions(orc):                    # :145-162
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    for i in range(10):
        mv.ringAdd(u.add("127.0.0.1", i), _rid())
    for i in range(10):
        mv.ringDelete(u.add("127.0.0.1", i))
    for k in range(K):
        assert mv.getRing(k) == []


def test_monitoring_relationship_edge(orc):                    # :167-195
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    n1 = u.add("127.0.0.1", 1)
    mv.ringAdd(n1, _rid())
    assert mv.getSubjectsOf(n1) == [] and mv.getObserversOf(n1) == []
    n2 = u.add("127.0.0.1", 2)
    with pytest.raises(orc.NodeNotInRingException):
        mv.getSubjectsOf(n2)
    with pytest.raises(orc.NodeNotInRingException):
        mv.getObserversOf(n2)


def test_monitoring_relationship_empty(orc):                   # :200-219
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    n = u.add("127.0.0.1", 1)
    with pytest.raises(orc.NodeNotInRingException):
        mv.getSubjectsOf(n)
    with pytest.raises(orc.NodeNotInRingException):
        mv.getObserversOf(n)


def test_monitoring_relationship_two_nodes(orc):               # :224-238
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    n1, n2 = u.add("127.0.0.1", 1), u.add("127.0.0.1", 2)
    mv.ringAdd(n1, _rid()); mv.ringAdd(n2, _rid())
    assert len(mv.getSubjectsOf(n1)) == K and len(mv.getObserversOf(n1)) == K
    assert set(mv.getSubjectsOf(n1)) == {n2} and set(mv.getObserversOf(n1)) == {n2}


def test_monitoring_relationship_three_nodes_with_delete(orc): # :243-265
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    n = [u.add("127.0.0.1", i) for i in (1, 2, 3)]
    for x in n:
        mv.ringAdd(x, _rid())
    assert len(mv.getSubjectsOf(n[0])) == K and len(mv.getObserversOf(n[0])) == K
    assert len(set(mv.getSubjectsOf(n[0]))) == 2 and len(set(mv.getObserversOf(n[0]))) == 2
    mv.ringDelete(n[1])
    assert len(mv.getSubjectsOf(n[0])) == K and len(mv.getObserversOf(n[0])) == K
    assert set(mv.getSubjectsOf(n[0])) == {n[2]} and set(mv.getObserversOf(n[0])) == {n[2]}


def test_monitoring_relationship_multiple_nodes(orc):          # :270-293
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    nodes = [u.add("127.0.0.1", i) for i in range(1000)]
    for x in nodes:
        mv.ringAdd(x, _rid())
    for x in nodes:
        assert len(mv.getSubjectsOf(x)) == K and len(mv.getObserversOf(x)) == K


def test_monitoring_relationship_bootstrap(orc):               # :298-313
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    n = u.add("127.0.0.1", 1234)
    mv.ringAdd(n, _rid())
    j = u.add("127.0.0.1", 1235)
    assert mv.getExpectedObserversOf(j) == [n] * K


def test_monitoring_relationship_bootstrap_multiple(orc):      # :318-344
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    j = u.add("127.0.0.1", 1233)
    num = 0
    for i in range(20):
        mv.ringAdd(u.add("127.0.0.1", 1234 + i), _rid())
        actual = len(mv.getExpectedObserversOf(j))
        assert num <= actual
        num = actual
    assert K - 3 <= num <= K
    # the Java counts list SIZE (always K once non-empty); the distinct count is the interesting one
    assert K - 3 <= len(set(mv.getExpectedObserversOf(j))) <= K


def test_node_unique_id_no_deletions(orc):                     # :351-398
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    n1 = u.add("127.0.0.1", 1); id1 = _rid()
    mv.ringAdd(n1, id1)
    with pytest.raises(orc.UUIDAlreadySeenException):
        mv.ringAdd(n1, id1)                                   # same host, same id
    with pytest.raises(orc.NodeAlreadyInRingException):
        mv.ringAdd(n1, _rid())                                # same host, different id
    n3 = u.add("127.0.0.1", 2)
    with pytest.raises(orc.UUIDAlreadySeenException):
        mv.ringAdd(n3, id1)                                   # different host, same id
    mv.ringAdd(n3, _rid())
    assert len(mv.getRing(0)) == 2


def test_node_unique_id_with_deletions(orc):                   # :405-434
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    n1 = u.add("127.0.0.1", 1); mv.ringAdd(n1, _rid())
    n2 = u.add("127.0.0.1", 2); id2 = _rid(); mv.ringAdd(n2, id2)
    mv.ringDelete(n2)
    assert len(mv.getRing(0)) == 1
    with pytest.raises(orc.UUIDAlreadySeenException):
        mv.ringAdd(n2, id2)
    mv.ringAdd(n2, _rid())
    assert len(mv.getRing(0)) == 2


def test_node_configuration_change(orc):                       # :442-457
    u = orc.Universe(); mv = orc.MembershipView(u, K)
    seen = set()
    for i in range(1000):
        mv.ringAdd(u.add("127.0.0.1", i), _name_id("127.0.0.1:%d" % i))
        seen.add(mv.getCurrentConfigurationId())
    assert len(seen) == 1000


def test_node_configurations_across_mviews(orc):               # :465-499
    u = orc.Universe(); mv1 = orc.MembershipView(u, K); mv2 = orc.MembershipView(u, K)
    l1, l2 = [], []
    for i in range(1000):
        mv1.ringAdd(u.add("127.0.0.1", i), _name_id("127.0.0.1:%d" % i))
        l1.append(mv1.getCurrentConfigurationId())
    for i in range(999, -1, -1):
        mv2.ringAdd(u.add("127.0.0.1", i), _name_id("127.0.0.1:%d" % i))
        l2.append(mv2.getCurrentConfigurationId())
    for a, b in zip(l1[:-1], l2[:-1]):
        assert a != b
    assert l1[-1] == l2[-1]


def test_ring_numbers_and_bulk_ctor(orc):                      # MembershipView.java:74-89, :397-418
    u = orc.Universe()
    tags = [u.add("10.0.0.%d" % (i // 50), 1000 + i % 50) for i in range(200)]
    ids = [_name_id(str(i)) for i in range(200)]
    bulk = orc.MembershipView(u, K, tags, [i[0] for i in ids], [i[1] for i in ids])
    inc = orc.MembershipView(u, K)
    for t, i in zip(tags, ids):
        inc.ringAdd(t, i)
    assert bulk.getCurrentConfigurationId() == inc.getCurrentConfigurationId()
    for k in range(K):
        assert bulk.getRing(k) == inc.getRing(k)
    for t in tags[:40]:
        obs = bulk.getObserversOf(t)
        for k, o in enumerate(obs):
            # the observer's alert about t carries exactly the rings on which t is o's predecessor
            assert k in bulk.getRingNumbers(o, t)
        assert sorted(set(r for o in set(obs) for r in bulk.getRingNumbers(o, t))) == list(range(K))
