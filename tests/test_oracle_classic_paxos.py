"""Port of PaxosTests (rapid/src/test/java/com/vrg/rapid/PaxosTests.java) against oracle::ClassicPaxos — the pins for
SURVEY.md §8 f2: coordinatorRuleTests (19 rows :257-295), coordinatorRuleTestsSameRank (17 rows :362-393), the 8 mixed-value
recovery rows (:176-192) and the three nValues scenarios (:69-136).

The Java harness wires FastPaxos instances through per-node single-thread executors; PaxosNet below is the same
wiring with per-node FIFO inboxes drained in a seeded random interleaving."""
import random

import pytest

CFG = 1
P1_PORTS, P2_PORTS, NOISE_PORTS = (5891, 5821), (5821, 5872), (1, 2)

# (N, p1N, p2N, swap p1/p2, valid proposal indexes)                                    PaxosTests.java:257-295
COORDINATOR_RULE = [
    (6, 4, 2, False, {0}), (6, 5, 1, False, {0}), (6, 6, 0, False, {0}), (9, 6, 3, False, {0, 1}), (9, 7, 2, False, {0}),
    (9, 8, 1, False, {0}), (6, 1, 5, False, {0, 1}), (6, 2, 4, False, {0, 1}), (6, 3, 3, False, {0}), (6, 3, 3, True, {0}),
    (6, 4, 1, False, {0}), (6, 5, 1, False, {0}), (9, 6, 1, False, {0, 1, 2}), (9, 7, 1, False, {0}), (9, 8, 1, False, {0}),
    (6, 1, 2, False, {0, 1, 2}), (6, 2, 1, False, {0, 1, 2}), (6, 3, 0, False, {0}), (6, 3, 0, True, {0}),
]
# same, but p1 and p2 share the highest rank                                             PaxosTests.java:362-393
COORDINATOR_RULE_SAME_RANK = [
    (6, 4, 2, False, {0, 1}), (6, 5, 1, False, {0}), (6, 6, 0, False, {0}), (9, 6, 3, False, {0, 1}), (9, 7, 2, False, {0}),
    (9, 8, 1, False, {0}), (6, 3, 3, False, {0, 1}), (6, 3, 3, True, {0, 1}),
    (6, 4, 1, False, {0, 1}), (6, 5, 0, False, {0}), (9, 6, 1, False, {0, 1, 2}), (9, 7, 1, False, {0}), (9, 8, 1, False, {0}),
    (6, 1, 2, False, {0, 1, 2}), (6, 2, 1, False, {0, 1, 2}), (6, 3, 0, False, {0}), (6, 3, 0, True, {0}),
]
# (N, votes for p2, which values may be decided: "p1" / "p2" / "any")                   PaxosTests.java:176-192
MIXED = [(6, 5, "p2"), (6, 1, "p1"), (6, 4, "any"), (6, 2, "any"), (5, 4, "p2"), (5, 1, "p1"), (10, 4, "any"), (10, 1, "any")]
N_VALUES = [5, 6, 10, 11, 20]                                                           # :128-136


def _proposals(u, swap):
    p1 = [u.add("127.0.0.1", p) for p in P1_PORTS]
    p2 = [u.add("127.0.0.1", p) for p in P2_PORTS]
    noise = [u.add("127.0.0.1", p) for p in NOISE_PORTS]
    return [p2, p1, noise] if swap else [p1, p2, noise]


def rule_messages(N, p1N, p2N, proposals, same_rank):
    """the Phase1bMessage list of PaxosTests.java:203-240 / :318-351"""
    msgs = [{"vrnd": (1, 1), "vval": proposals[0]} for _ in range(p1N)]
    msgs += [{"vrnd": (1, 1) if same_rank else (0, 2**31 - 1), "vval": proposals[1]} for _ in range(p2N)]
    msgs += [{"vrnd": (0, i), "vval": proposals[2]} for i in range(p1N + p2N, N)]
    return msgs


@pytest.mark.parametrize("same_rank,row", [(False, r) for r in COORDINATOR_RULE] + [(True, r) for r in COORDINATOR_RULE_SAME_RANK])
def test_coordinator_rule_rows(orc, same_rank, row):
    N, p1N, p2N, swap, valid = row
    u = orc.Universe()
    proposals = _proposals(u, swap)
    px = orc.ClassicPaxos(u, u.add("127.0.0.1", 1234), 7, CFG, N)
    rng = random.Random(N * 1000 + p1N * 10 + p2N)
    seen = set()
    for _ in range(100):
        msgs = rule_messages(N, p1N, p2N, proposals, same_rank)
        rng.shuffle(msgs)
        chosen = px.selectProposalUsingCoordinatorRule(msgs[: N // 2 + 1])     # a random quorum
        assert chosen in [proposals[i] for i in valid], (chosen, row)
        seen.add(tuple(chosen))
    assert len(seen) >= 1


def test_coordinator_rule_empty_list_and_all_empty_vvals(orc):
    u = orc.Universe()
    px = orc.ClassicPaxos(u, u.add("127.0.0.1", 1234), 7, CFG, 6)
    with pytest.raises(ValueError):                                              # Paxos.java:274 orElseThrow
        px.selectProposalUsingCoordinatorRule([])
    assert px.selectProposalUsingCoordinatorRule([{"vrnd": (0, 0), "vval": []}] * 4) == []   # :318-326


def test_coordinator_rule_counts_in_arrival_order(orc):
    """N/4 + 1 identical votes at the highest rank win as soon as they are reached, in LIST order (:296-307)"""
    u = orc.Universe()
    a, b = [u.add("h", 1)], [u.add("h", 2)]
    px = orc.ClassicPaxos(u, u.add("127.0.0.1", 1234), 7, CFG, 8)                  # N/4 = 2 -> third occurrence wins
    m = lambda v: {"vrnd": (1, 1), "vval": v}
    assert px.selectProposalUsingCoordinatorRule([m(a), m(b), m(b), m(a), m(b), m(a)]) == b
    assert px.selectProposalUsingCoordinatorRule([m(a), m(b), m(b), m(a), m(a), m(b)]) == a
    assert px.selectProposalUsingCoordinatorRule([m(a), m(b), m(b), m(a)]) == a      # nobody reaches 3: first non-empty
    assert px.selectProposalUsingCoordinatorRule([{"vrnd": (0, 5), "vval": b}, m(a)]) == a   # single value at max rank


class PaxosNet:
    """N FastPaxos nodes (fast-round tally + classic Paxos) wired like PaxosTests.createNFastPaxosInstances (:399-414)."""

    def __init__(self, orc, N, seed, drop=()):
        self.u = orc.Universe()
        self.N, self.drop, self.rng = N, set(drop), random.Random(seed)
        self.tags = [self.u.add("127.0.0.1", 1234 + i) for i in range(N)]
        hashes = list(range(100, 100 + N))
        random.Random(seed + 1).shuffle(hashes)                                  # stands in for Endpoint.hashCode()
        self.px = [orc.ClassicPaxos(self.u, self.tags[i], hashes[i], CFG, N) for i in range(N)]
        self.fp = [orc.FastPaxosTally(self.u, CFG, N) for _ in range(N)]
        self.decisions = [None] * N
        self.inbox = [[] for _ in range(N)]

    def broadcast(self, kind, msg):                                              # DirectBroadcaster :420-433
        if kind in self.drop:
            return
        for q in self.inbox:
            q.append((kind, msg))

    def propose(self, i, proposal):                                              # FastPaxos.propose :94-108
        self.px[i].registerFastRoundVote(proposal)
        self.broadcast("fast2b", {"sender": self.tags[i], "cfg": CFG, "endpoints": list(proposal)})

    def start_classic_round(self, i):                                            # FastPaxos.startClassicPaxosRound :193-199
        if self.decisions[i] is None:
            m = self.px[i].startPhase1a(2)
            if m:
                self.broadcast("1a", m)

    def _decide(self, i, value):
        assert self.decisions[i] is None                                         # onDecidedWrapped asserts !decided
        self.decisions[i] = list(value)

    def deliver(self, i, kind, m):                                               # FastPaxos.handleMessages :166-188
        if kind == "fast2b":
            if self.fp[i].handleFastRoundProposal(m["sender"], m["cfg"], m["endpoints"]):
                self._decide(i, self.fp[i].decision())
        elif kind == "1a":
            r = self.px[i].handlePhase1aMessage(m)
            if r:
                self.inbox[self.tags.index(m["sender"])].append(("1b", r))       # client.sendMessage(sender)
        elif kind == "1b":
            r = self.px[i].handlePhase1bMessage(m)
            if r:
                self.broadcast("2a", r)
        elif kind == "2a":
            r = self.px[i].handlePhase2aMessage(m)
            if r:
                self.broadcast("2b", r)
        elif kind == "2b":
            if self.decisions[i] is None and self.px[i].handlePhase2bMessage(m):
                self._decide(i, self.px[i].decision())
            elif self.decisions[i] is not None:
                self.px[i].handlePhase2bMessage(m)

    def run(self):
        while True:
            ready = [i for i in range(self.N) if self.inbox[i]]
            if not ready:
                return
            i = self.rng.choice(ready)
            kind, m = self.inbox[i].pop(0)
            self.deliver(i, kind, m)


@pytest.mark.parametrize("N", N_VALUES)
def test_recovery_for_single_propose(orc, N):                                    # :69-83
    net = PaxosNet(orc, N, seed=N)
    proposal = [net.u.add("172.14.12.3", 1234)]
    net.propose(0, proposal)
    net.run()
    assert all(d is None for d in net.decisions)                                 # one vote is no fast quorum
    net.start_classic_round(0)                                                   # the proposer's recovery timer fires
    net.run()
    assert net.decisions == [proposal] * N


@pytest.mark.parametrize("N", N_VALUES)
def test_recovery_from_fast_round_with_different_proposals(orc, N):              # :88-104
    net = PaxosNet(orc, N, seed=10 + N)
    for i in range(N):
        net.propose(i, [net.tags[i]])
    net.run()
    assert all(d is None for d in net.decisions)
    order = list(range(N))
    random.Random(N).shuffle(order)
    for i in order:                                                              # every node's timer fires
        net.start_classic_round(i)
    net.run()
    d = net.decisions[0]
    assert d is not None and len(d) == 1 and d[0] in net.tags
    assert net.decisions == [d] * N


@pytest.mark.parametrize("N", N_VALUES)
def test_classic_round_after_successful_fast_round(orc, N):                      # :110-126
    net = PaxosNet(orc, N, seed=20 + N, drop=["fast2b"])
    proposal = [net.u.add("127.0.0.1", 1234)]
    for i in range(N):
        net.propose(i, proposal)
    net.run()
    assert all(d is None for d in net.decisions)
    for i in range(N):
        net.start_classic_round(i)
    net.run()
    assert net.decisions == [proposal] * N


@pytest.mark.parametrize("N,p2votes,expect", MIXED)
@pytest.mark.parametrize("seed", range(5))
def test_classic_round_after_successful_fast_round_mixed_values(orc, N, p2votes, expect, seed):   # :140-192
    net = PaxosNet(orc, N, seed=100 * seed + N + p2votes, drop=["fast2b"])
    p1 = [net.u.add("127.0.0.1", p) for p in P1_PORTS]
    p2 = [net.u.add("127.0.0.1", p) for p in P2_PORTS]
    for i in range(N):
        net.propose(i, p1 if i < N - p2votes else p2)
    net.run()
    assert all(d is None for d in net.decisions)
    for i in range(N):
        net.start_classic_round(i)
    net.run()
    d = net.decisions[0]
    assert net.decisions == [d] * N
    if expect == "any":
        assert d in (p1, p2)
    else:
        assert d == (p1 if expect == "p1" else p2)


def test_acceptor_rank_rules(orc):
    u = orc.Universe()
    me, c1, c2 = u.add("a", 1), u.add("c", 1), u.add("c", 2)
    px = orc.ClassicPaxos(u, me, 5, CFG, 5)
    v = [u.add("v", 1)]
    px.registerFastRoundVote(v)
    assert px.ranks()["rnd"] == (1, 1) and px.ranks()["vrnd"] == (1, 1) and px.vval() == v
    assert px.handlePhase1aMessage({"sender": c1, "cfg": CFG + 1, "rank": (2, 9)}) is None      # wrong configuration
    r = px.handlePhase1aMessage({"sender": c1, "cfg": CFG, "rank": (2, 9)})
    assert r == {"sender": me, "cfg": CFG, "rnd": (2, 9), "vrnd": (1, 1), "vval": v}
    assert px.handlePhase1aMessage({"sender": c2, "cfg": CFG, "rank": (2, 9)}) is None          # equal rank: rejected
    assert px.handlePhase1aMessage({"sender": c2, "cfg": CFG, "rank": (2, -3)}) is None         # lower node index
    px.registerFastRoundVote([u.add("v", 2)])                                                    # rnd.round > 1: ignored
    assert px.vval() == v
    w = [u.add("w", 1)]
    assert px.handlePhase2aMessage({"sender": c2, "cfg": CFG, "rnd": (2, 8), "vval": w}) is None  # below rnd
    out = px.handlePhase2aMessage({"sender": c1, "cfg": CFG, "rnd": (2, 9), "vval": w})
    assert out == {"sender": me, "cfg": CFG, "rnd": (2, 9), "endpoints": w}
    assert px.handlePhase2aMessage({"sender": c1, "cfg": CFG, "rnd": (2, 9), "vval": w}) is None  # vrnd == rnd already
    assert px.ranks() == {"rnd": (2, 9), "vrnd": (2, 9), "crnd": (0, 0)}


def test_learner_counts_distinct_senders_per_round(orc):
    u = orc.Universe()
    px = orc.ClassicPaxos(u, u.add("a", 1), 5, CFG, 5)                                            # needs > 2 senders
    s = [u.add("s", i) for i in range(5)]
    v, w = [u.add("v", 1)], [u.add("w", 1)]
    assert not px.handlePhase2bMessage({"sender": s[0], "cfg": CFG, "rnd": (2, 1), "endpoints": v})
    assert not px.handlePhase2bMessage({"sender": s[0], "cfg": CFG, "rnd": (2, 1), "endpoints": v})   # same sender again
    assert not px.handlePhase2bMessage({"sender": s[1], "cfg": CFG, "rnd": (2, 2), "endpoints": w})   # another round
    assert not px.handlePhase2bMessage({"sender": s[1], "cfg": CFG, "rnd": (2, 1), "endpoints": v})
    assert not px.handlePhase2bMessage({"sender": s[2], "cfg": CFG + 1, "rnd": (2, 1), "endpoints": v})  # wrong cfg
    assert px.handlePhase2bMessage({"sender": s[2], "cfg": CFG, "rnd": (2, 1), "endpoints": v})
    assert px.decided() and px.decision() == v
    assert not px.handlePhase2bMessage({"sender": s[3], "cfg": CFG, "rnd": (2, 1), "endpoints": v})   # decided once
