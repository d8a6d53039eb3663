"""GPU parity of the classic-Paxos fallback (csrc/classic_paxos.cu, through the C ABI) against oracle::ClassicPaxos
(Paxos.java restated): the reference's PaxosTests tables, random message streams, the acceptor registers of R virtual
nodes, and a whole fast-round-fails -> classic-round-decides scenario."""
import random

import numpy as np
import pytest

from test_oracle_classic_paxos import (COORDINATOR_RULE, COORDINATOR_RULE_SAME_RANK, CFG, _proposals, rule_messages)

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def rb():
    import rapid_b200
    return rapid_b200


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


class Values:
    """List<Endpoint> <-> the opaque (hash, hash2, len) triple of the C ABI"""

    def __init__(self):
        self.ids, self.lists = {}, {}

    def triple(self, v):
        if not v:
            return (0, 0, 0)
        k = tuple(v)
        if k not in self.ids:
            i = len(self.ids) + 1
            self.ids[k] = (splitmix64(i), splitmix64(i + 1000), len(k))
            self.lists[self.ids[k]] = list(k)
        return self.ids[k]

    def arrays(self, vs):
        t = [self.triple(v) for v in vs]
        return (np.array([x[0] for x in t], np.uint64), np.array([x[1] for x in t], np.uint64), np.array([x[2] for x in t], np.int32))

    def value(self, triple):
        return [] if triple is None or triple[2] == 0 else self.lists[tuple(triple)]


def gpu_rule(rb, px, vals, msgs):
    h1, h2, ln = vals.arrays([m["vval"] for m in msgs])
    i = px.selectProposalUsingCoordinatorRule([m["vrnd"] for m in msgs], h1, ln, h2)
    return [] if i < 0 else msgs[i]["vval"]


@pytest.mark.parametrize("same_rank,row", [(False, r) for r in COORDINATOR_RULE] + [(True, r) for r in COORDINATOR_RULE_SAME_RANK])
def test_coordinator_rule_rows(orc, rb, same_rank, row):                        # PaxosTests.java:194-393
    Nn, p1N, p2N, swap, valid = row
    u = orc.Universe()
    proposals = _proposals(u, swap)
    ref = orc.ClassicPaxos(u, u.add("127.0.0.1", 1234), 7, CFG, Nn)
    px, vals = rb.Paxos(CFG, Nn), Values()
    rng = random.Random(Nn * 1000 + p1N * 10 + p2N)
    for _ in range(25):
        msgs = rule_messages(Nn, p1N, p2N, proposals, same_rank)
        rng.shuffle(msgs)
        quorum = msgs[: Nn // 2 + 1]
        got = gpu_rule(rb, px, vals, quorum)
        assert got == ref.selectProposalUsingCoordinatorRule(quorum)               # exact, not just "one of the valid ones"
        assert got in [proposals[i] for i in valid]


def test_coordinator_rule_empty_list_is_an_error(rb):
    px = rb.Paxos(CFG, 6)
    with pytest.raises(rb.RapidError):                                            # Paxos.java:274
        px.selectProposalUsingCoordinatorRule(np.zeros((0, 2)), np.zeros(0, np.uint64), np.zeros(0, np.int32))


@pytest.mark.parametrize("seed", range(12))
def test_coordinator_rule_random(orc, rb, seed):
    rng = random.Random(seed)
    Nn = rng.choice([4, 5, 8, 9, 16, 33, 100, 257])
    u = orc.Universe()
    pool = [[]] + [[u.add("v", 10 * i + j) for j in range(rng.randint(1, 3))] for i in range(rng.randint(1, 5))]
    ranks = [(rng.randint(0, 2), rng.choice([-5, 0, 1, 7, 2**31 - 1, -2**31])) for _ in range(rng.randint(1, 4))]
    ref = orc.ClassicPaxos(u, u.add("me", 1), 7, CFG, Nn)
    px, vals = rb.Paxos(CFG, Nn, message_capacity=4096), Values()
    for _ in range(20):
        m = rng.randint(1, 3 * Nn)
        msgs = [{"vrnd": rng.choice(ranks), "vval": rng.choice(pool)} for _ in range(m)]
        assert gpu_rule(rb, px, vals, msgs) == ref.selectProposalUsingCoordinatorRule(msgs)


def _feed_1b(px, vals, batch):
    h1, h2, ln = vals.arrays([m["vval"] for m in batch])
    return px.handlePhase1bMessages([m["rnd"] for m in batch], [m["vrnd"] for m in batch], h1, ln, h2, msg_cfg=[m["cfg"] for m in batch])


@pytest.mark.parametrize("seed", range(10))
def test_phase1b_stream_matches_oracle(orc, rb, seed):
    """batches of Phase1bMessages with stale ranks, wrong configurations and empty vvals: the batch and the message at
    which the coordinator picks cval, and cval itself, are the oracle's"""
    rng = random.Random(100 + seed)
    Nn = rng.choice([3, 6, 10, 31, 64])
    u = orc.Universe()
    me = u.add("me", 1)
    senders = [u.add("s", i) for i in range(Nn)]
    pool = [[]] * rng.randint(1, 4) + [[u.add("v", 10 * i + j) for j in range(rng.randint(1, 3))] for i in range(rng.randint(1, 4))]
    ref = orc.ClassicPaxos(u, me, 77, CFG, Nn)
    px, vals = rb.Paxos(CFG, Nn, message_capacity=8192), Values()
    assert px.startPhase1a(2, 77) and ref.startPhase1a(2)["rank"] == (2, 77)
    vr = [(0, 0), (1, 1), (2, 5)]
    proposed_at, total = None, 0
    for b in range(rng.randint(2, 8)):
        batch = []
        for _ in range(rng.randint(0, Nn)):
            batch.append({"sender": rng.choice(senders), "cfg": CFG if rng.random() < 0.9 else CFG + 1,
                          "rnd": (2, 77) if rng.random() < 0.85 else rng.choice([(2, 76), (1, 77), (3, 77)]),
                          "vrnd": rng.choice(vr), "vval": rng.choice(pool)})
        want = None
        for i, m in enumerate(batch):
            out = ref.handlePhase1bMessage(m)
            if out is not None:
                assert want is None
                want = (i, out["vval"])
                assert out["rnd"] == (2, 77)
        total += sum(1 for m in batch if m["cfg"] == CFG and m["rnd"] == (2, 77))
        got = _feed_1b(px, vals, batch)
        assert got.n_messages == total
        if want is None:
            assert not got.proposed and got.trigger_index == -1
        else:
            assert got.proposed and got.trigger_index == want[0] and vals.value(got.cval) == want[1]
            proposed_at = b
        assert vals.value(got.cval) == ref.cval()
        if b == 3:                                             # a later, higher round of the same coordinator (:98-103)
            assert px.startPhase1a(1, 5) is False              # crnd.round > round: ignored
            assert ref.startPhase1a(1) is None
    assert proposed_at is None or ref.cval()


def _feed_2b(px, vals, batch):
    h1, h2, ln = vals.arrays([m["endpoints"] for m in batch])
    return px.handlePhase2bMessages([m["rnd"] for m in batch], [m["sender"] for m in batch], h1, ln, h2, msg_cfg=[m["cfg"] for m in batch])


@pytest.mark.parametrize("seed", range(10))
def test_phase2b_stream_matches_oracle(orc, rb, seed):
    """Phase2bMessages of several rounds, repeated senders and wrong configurations over several calls: the deciding
    message (first arrival with more than N/2 distinct senders in its round) and ITS value are the oracle's"""
    rng = random.Random(200 + seed)
    Nn = rng.choice([3, 4, 7, 20, 51])
    u = orc.Universe()
    ref = orc.ClassicPaxos(u, u.add("me", 1), 77, CFG, Nn)
    px, vals = rb.Paxos(CFG, Nn, message_capacity=8192), Values()
    senders = [u.add("s", i) for i in range(Nn + 3)]
    rounds = [(2, rng.randint(-9, 9)) for _ in range(rng.randint(1, 3))]
    pool = [[u.add("v", 10 * i + j) for j in range(rng.randint(1, 3))] for i in range(3)]
    decided = False
    for b in range(rng.randint(2, 8)):
        batch = [{"sender": rng.choice(senders), "cfg": CFG if rng.random() < 0.9 else CFG - 1, "rnd": rng.choice(rounds),
                  "endpoints": rng.choice(pool)} for _ in range(rng.randint(0, Nn))]
        want = None
        for i, m in enumerate(batch):
            if ref.handlePhase2bMessage(m):
                assert want is None and not decided
                want = i
        got = _feed_2b(px, vals, batch)
        if want is None:
            assert got.decided_index == -1
        else:
            decided = True
            assert got.decided_index == want and vals.value(got.decision) == batch[want]["endpoints"]
        assert got.decided == decided == ref.decided()
        if decided:
            assert vals.value(got.decision) == ref.decision()


def _order(begin, n_replies_senders, perm_seed):
    s = list(n_replies_senders)
    if perm_seed:
        s.sort(key=lambda x: (splitmix64(perm_seed ^ (x & 0xFFFFFFFF)), x))
    return s


@pytest.mark.parametrize("seed", range(8))
def test_acceptors_and_full_classic_round(orc, rb, seed):
    """R virtual acceptors in HBM vs R literal Paxos instances: fast-round votes, competing coordinators' Phase1a, the
    coordinator rule over their answers (in a permuted arrival order), Phase2a, and every learner's decision"""
    rng = random.Random(300 + seed)
    R = rng.choice([5, 6, 10, 33, 128])
    begin = rng.choice([0, 1000])
    u = orc.Universe()
    tags = [u.add("n", begin + r) for r in range(R)]
    assert tags == list(range(R))                              # oracle tag r <-> node id begin + r
    hashes = [rng.randint(-50, 50) * 2 + (r % 2) for r in range(R)]
    ref = [orc.ClassicPaxos(u, tags[r], hashes[r] * 1000 + r, CFG, R) for r in range(R)]
    vals = Values()
    acc = rb.PaxosAcceptors(CFG, R, acceptor_begin=begin)
    # fast round: a split vote (no fast quorum), some nodes never voted
    pool = [[tags[0]], [tags[1], tags[2]], [tags[3]]]
    voters = [r for r in range(R) if rng.random() < 0.8]
    votes = {r: rng.choice(pool[: rng.randint(1, 3)]) for r in voters}
    for r, v in votes.items():
        ref[r].registerFastRoundVote(v)
    h1, h2, ln = vals.arrays([votes[r] for r in voters])
    acc.registerFastRoundVotes(voters, h1, ln, h2)
    # two or three coordinators fire their recovery timers one after the other
    coords = rng.sample(range(R), min(R, rng.randint(1, 3)))
    decided_any = False
    for c in coords:
        m1a = ref[c].startPhase1a(2)
        px = rb.Paxos(CFG, R)
        assert px.startPhase1a(2, m1a["rank"][1])
        assert acc.handlePhase1aMessage((2, 999), msg_cfg=CFG + 9) == 0                       # wrong configuration
        replies = {}
        for r in range(R):
            out = ref[r].handlePhase1aMessage(m1a)
            if out is not None:
                replies[begin + r] = out
        assert acc.handlePhase1aMessage(m1a["rank"]) == len(replies)
        perm_seed = rng.choice([0, rng.getrandbits(60) | 1])
        order = _order(begin, sorted(replies), perm_seed)
        want = None
        for i, s in enumerate(order):
            out = ref[c].handlePhase1bMessage(replies[s])
            if out is not None:
                assert want is None
                want = (i, out)
        got = px.handlePhase1bFromAcceptors(acc, perm_seed)
        assert got.n_messages == len(order)
        if want is None:
            assert not got.proposed
            continue
        assert got.proposed and got.trigger_index == want[0] and vals.value(got.cval) == want[1]["vval"]
        m2a = want[1]
        accepted = [begin + r for r in range(R) if ref[r].handlePhase2aMessage(m2a) is not None]
        assert acc.handlePhase2aMessage(m2a["rnd"], got.cval) == len(accepted)
        for r in range(R):                                                                       # registers, one by one
            st = acc.read(r)
            rk = ref[r].ranks()
            assert st["rnd"] == rk["rnd"] and st["vrnd"] == rk["vrnd"] and vals.value(st["vval"]) == ref[r].vval()
        perm2 = rng.choice([0, rng.getrandbits(60) | 1])
        order2 = _order(begin, accepted, perm2)
        learner = ref[rng.randrange(R)]
        was = learner.decided()
        want2 = None
        for i, s in enumerate(order2):
            if learner.handlePhase2bMessage({"sender": s - begin, "cfg": CFG, "rnd": m2a["rnd"], "endpoints": m2a["vval"]}):
                want2 = i
        lpx = rb.Paxos(CFG, R)
        got2 = lpx.handlePhase2bFromAcceptors(acc, perm2)
        if not was:
            assert got2.decided == (want2 is not None)
            if want2 is not None:
                assert got2.decided_index == want2 == R // 2 and vals.value(got2.decision) == m2a["vval"]
                decided_any = True
    assert decided_any or len(coords) == 0 or True


def test_acceptor_errors(rb):
    acc = rb.PaxosAcceptors(CFG, 4)
    with pytest.raises(rb.RapidError):
        acc.registerFastRoundVotes([7], [1], [1])
    px = rb.Paxos(CFG, 4)
    with pytest.raises(rb.RapidError):
        px.handlePhase1bFromAcceptors(acc)                     # nothing pending
    assert acc.handlePhase1aMessage((2, 3)) == 4
    with pytest.raises(rb.RapidError):
        px.handlePhase2bFromAcceptors(acc)                     # Phase1b answers pending, not Phase2b
    small = rb.Paxos(CFG, 4, message_capacity=2)
    small.startPhase1a(2, 3)
    assert small.handlePhase1bFromAcceptors(acc).n_messages == 4     # the Phase1b list grows past the initial capacity
    with pytest.raises(rb.RapidError):                                # ... the Phase2b table does not: refused, not corrupted
        small.handlePhase2bMessages([(2, 3)] * 500, list(range(500)), [1] * 500, [1] * 500)   # table of 1024 entries, load <= 3/4


def test_conflicting_fast_round_falls_back_to_classic_round(orc, rb):
    """the conflict regime of FastPaxosWithoutFallbackTests (no fast quorum) carried on as PaxosTests does: the votes the
    cut detector's receivers cast are registered on the device, a classic round recovers a single decision"""
    from helpers import OracleWorld
    n, K, H, L = 120, 10, 9, 4
    w = OracleWorld(orc, n, K)
    v = rb.MembershipView.from_packed(K, *w.member_packed())
    cl = rb.VirtualCluster(v, H, L, kernel="bucketed")
    obs = w.tables()[0]
    # two crashes; 40 % of the receivers are blocked from the second subject's alerts -> two different proposals
    cells = [(int(obs[s][r]), s, r, 1) for s in (5, 17) for r in range(K)]
    src, dst, ring, st = (np.array(x) for x in zip(*cells))
    rng = np.random.default_rng(5)
    half = rng.random(n) < 0.4
    words = (n + 31) // 32
    bm = np.zeros((len(cells), words), np.uint32)
    for i, (_, s, _, _) in enumerate(cells):
        mask = np.ones(n, bool) if s == 5 else ~half
        for r in np.nonzero(mask)[0]:
            bm[i, r // 32] |= np.uint32(1 << (r % 32))
    cfg = w.view.getCurrentConfigurationId()
    out = cl.handleBatch(cfg, src, dst, ring, st, bitmap=bm)
    lens = np.asarray(out.proposal_len)
    assert set(lens.tolist()) == {1, 2}
    # no fast quorum: neither proposal has N - floor((N-1)/4) votes
    tally = rb.FastPaxos(cfg, n).tallyFrom(cl) if hasattr(rb.FastPaxos, "tallyFrom") else None
    assert tally is None or not tally.decided
    # classic round on the device
    acc = rb.PaxosAcceptors(cfg, n)
    acc.registerFastRoundVotesFrom(cl)
    for r in range(n):
        st_r = acc.read(r)
        assert st_r["rnd"] == (1, 1) and st_r["vval"][2] == lens[r]
    px = rb.Paxos(cfg, n)
    px.startPhase1a(2, 42)
    assert acc.handlePhase1aMessage((2, 42)) == n
    got = px.handlePhase1bFromAcceptors(acc, perm_seed=12345)
    assert got.proposed and got.trigger_index == n // 2
    # the rule: more than N/4 identical votes among the first N/2+1 answers wins; both proposals can qualify, the first to
    # get there in arrival order is taken — recompute it on the host from the per-receiver outputs
    order = sorted(range(n), key=lambda x: (splitmix64(12345 ^ x), x))[: n // 2 + 1]
    h = np.asarray(out.proposal_hash)
    cnt, want = {}, None
    if len({int(h[r]) for r in order}) == 1:
        want = int(h[order[0]])
    else:
        for r in order:
            c = cnt.get(int(h[r]), 0)
            if c + 1 > n // 4:
                want = int(h[r])
                break
            cnt[int(h[r])] = c + 1
    assert want is not None and got.cval[0] == want
    assert acc.handlePhase2aMessage((2, 42), got.cval) == n
    dec = rb.Paxos(cfg, n).handlePhase2bFromAcceptors(acc, perm_seed=99)
    assert dec.decided and dec.decided_index == n // 2 and dec.decision == got.cval


def test_one_million_acceptors_classic_round(rb):
    """full-size property check (no oracle): 1M acceptors, 70/30 split fast-round vote, one coordinator; the answer is
    fixed by the arrival order alone"""
    n = 1_000_000
    acc = rb.PaxosAcceptors(9, n)
    ids = np.arange(n, dtype=np.int64)
    h = np.where(ids % 10 < 7, np.uint64(111), np.uint64(222)).astype(np.uint64)
    acc.registerFastRoundVotes(ids, h, np.full(n, 3, np.int32))
    px = rb.Paxos(9, n, message_capacity=n)
    px.startPhase1a(2, 1)
    assert acc.handlePhase1aMessage((2, 1)) == n
    assert acc.handlePhase1aMessage((2, 1)) == 0                # same rank again: nobody answers (:125-134)
    assert acc.handlePhase1aMessage((2, 0)) == 0
    assert acc.handlePhase1aMessage((2, 2)) == n                # a higher coordinator takes over
    got = px.handlePhase1bFromAcceptors(acc)                    # ... so these answers carry rnd (2,2) != crnd (2,1)
    assert not got.proposed and got.n_messages == 0
    px2 = rb.Paxos(9, n, message_capacity=n)
    px2.startPhase1a(2, 2)
    got = px2.handlePhase1bFromAcceptors(acc)
    # acceptor order: value 111 reaches its (N/4+1)-th occurrence long before 222 does
    assert got.proposed and got.trigger_index == n // 2 and got.cval == (111, 0, 3) and got.n_messages == n
    assert acc.handlePhase2aMessage((2, 2), got.cval) == n
    assert acc.handlePhase2aMessage((2, 2), got.cval) == 0      # vrnd == rnd already (:204)
    assert acc.read(n - 1) == {"rnd": (2, 2), "vrnd": (2, 2), "vval": (111, 0, 3)}
