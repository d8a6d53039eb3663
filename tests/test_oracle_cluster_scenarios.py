"""The failure scenarios of ClusterTest (rapid/src/test/java/com/vrg/rapid/ClusterTest.java:213-362) replayed through the oracle
at the level of the path — no sockets, no join protocol:

    StaticFailureDetector ticks (T/StaticFailureDetector.java:41-45: notifier.run() on EVERY tick while the subject is blacklisted)
      -> one BatchedAlertMessage per sender and window (MembershipService.java:472-495, :613-637)
      -> every live node's batch handler (:300-354) with its own delivery order
      -> FastPaxos votes (FastPaxos.java:94-156) -> [classic Paxos fallback, Paxos.java, started by one proposer's recovery timer]
      -> decideViewChange (:385-444: ringDelete of the decided cut, a new configuration, detectors cleared)

repeated until no blacklisted node is left in the view.  The assertions are ClusterTest's own: every surviving node ends with
the same membership, of the expected size (waitAndVerifyAgreement :710-730).  These runs pin the COMPOSITION of the oracle's
pieces (rows a8-a15, f1, f2 of SURVEY.md §8) against outcomes the reference's integration tests hold; the GPU path is checked
against the same pieces in tests/test_gpu_*.py."""
import random

import numpy as np
import pytest

from helpers import OracleWorld
from rapid_b200 import workloads as W

K, H, L = 10, 9, 4                                                 # Cluster.java:72-74
UP, DOWN = 0, 1


class ScenarioCluster:
    def __init__(self, orc, n, seed, n_joiners=0):
        self.orc, self.n, self.rng = orc, n, random.Random(seed)
        self.w = OracleWorld(orc, n, K, n_joiners=n_joiners)
        self.members = list(range(n))
        self.joiner_ids = W.node_ids(n, n_joiners)
        self.rounds, self.fast_decisions, self.classic_decisions = 0, 0, 0

    # one failure-detector interval: every live member's K edge detectors (one per entry of getSubjectsOf, :697-707) fire for the
    # blacklisted subjects; the AlertBatcher ships them as ONE batch per sender
    def sender_batches(self, blacklist, dead, cfg):
        out = {}
        for o in self.members:
            if o in dead:
                continue
            msgs = []
            for s in self.w.view.getSubjectsOf(o):
                if s in blacklist:
                    msgs.append((o, s, DOWN, cfg, self.w.view.getRingNumbers(o, s)))
            if msgs:
                out[o] = msgs
        return out

    # join phase 2 (MembershipService.java:240-281): the joiner asked its K expected observers of THIS configuration; each live one
    # enqueues an UP alert with the ring numbers the joiner named for it
    def join_alerts(self, joiners, dead, cfg, out):
        for j in joiners:
            exp = self.w.view.getExpectedObserversOf(j)
            for o in sorted(set(exp)):
                if o in dead:
                    continue
                out.setdefault(o, []).append((o, j, UP, cfg, [k for k in range(K) if exp[k] == o]))
        return out

    def detection_round(self, blacklist, dead, ticks=2, source=None):
        """-> the decided cut (list of tags) or None if nobody proposed.  `source(cfg)` -> {sender: [AlertMessage]} of one batching
        window; default: the static detector over `blacklist`."""
        orc, view = self.orc, self.w.view
        cfg = view.getCurrentConfigurationId()
        N = view.getMembershipSize()
        live = [m for m in self.members if m not in dead]
        handlers = {m: orc.AlertBatchHandler(view, K, H, L) for m in live}
        proposals = {}
        for _ in range(ticks):
            batches = source(cfg) if source else self.sender_batches(blacklist, dead, cfg)
            for r in live:
                order = list(batches)
                self.rng.shuffle(order)                            # UnicastToAllBroadcaster shuffles; arrival order is per receiver
                for o in order:
                    got = handlers[r].handleBatch(batches[o])      # (ignored once announcedProposal is set, :318-319)
                    if got:
                        assert r not in proposals
                        proposals[r] = got
        if not proposals:
            return None
        # ---- fast round: every proposer broadcasts its vote; every live node tallies all of them in its own order
        votes = list(proposals.items())
        decided = {}
        for r in live:
            fp = orc.FastPaxosTally(self.w.u, cfg, N)
            order = votes[:]
            self.rng.shuffle(order)
            for sender, prop in order:
                if fp.handleFastRoundProposal(sender, cfg, prop):
                    decided[r] = fp.decision()
        if decided:
            assert len(decided) == len(live), "a fast-round decision is reached by everyone who sees all votes"
            self.fast_decisions += 1
        else:
            decided = self.classic_round(live, proposals, cfg, N)
            self.classic_decisions += 1
        vals = {tuple(v) for v in decided.values()}
        assert len(vals) == 1, "agreement"
        return list(vals.pop())

    def classic_round(self, live, proposals, cfg, N):
        """Paxos.java among the live members (the dead ones are acceptors that never answer); one proposer's recovery timer fires
        (FastPaxos.java:193-203)."""
        orc = self.orc
        hashes = list(range(1000, 1000 + len(live)))
        self.rng.shuffle(hashes)
        px = {m: orc.ClassicPaxos(self.w.u, m, hashes[i], cfg, N) for i, m in enumerate(live)}
        for m, prop in proposals.items():
            px[m].registerFastRoundVote(prop)                      # FastPaxos.propose :94-98
        inbox = {m: [] for m in live}
        decided = {}

        def broadcast(kind, msg):
            for m in live:
                inbox[m].append((kind, msg))

        coordinator = self.rng.choice(sorted(proposals))
        m1a = px[coordinator].startPhase1a(2)
        assert m1a is not None
        broadcast("1a", m1a)
        while True:
            ready = [m for m in live if inbox[m]]
            if not ready:
                break
            i = self.rng.choice(ready)
            kind, m = inbox[i].pop(0)
            if kind == "1a":
                r = px[i].handlePhase1aMessage(m)
                if r:
                    inbox[m["sender"]].append(("1b", r))
            elif kind == "1b":
                r = px[i].handlePhase1bMessage(m)
                if r:
                    broadcast("2a", r)
            elif kind == "2a":
                r = px[i].handlePhase2aMessage(m)
                if r:
                    broadcast("2b", r)
            elif kind == "2b":
                if px[i].handlePhase2bMessage(m) and i not in decided:
                    decided[i] = px[i].decision()
        assert set(decided) == set(live), "every live member learns the classic-round decision"
        return decided

    def apply_cut(self, cut):                                      # decideViewChange :385-444
        for t in cut:
            if self.w.view.isHostPresent(t):
                self.w.view.ringDelete(t)
                self.members.remove(t)
            else:                                                  # a joiner: ringAdd with the NodeId its UP alerts carried
                j = t - self.n
                self.w.view.ringAdd(t, (int(self.joiner_ids[0][j]), int(self.joiner_ids[1][j])))
                self.members.append(t)

    def run(self, blacklist, dead, max_rounds=12):
        blacklist, dead = set(blacklist), set(dead)
        while blacklist & set(self.members):
            assert self.rounds < max_rounds, "no convergence"
            before = self.w.view.getCurrentConfigurationId()
            cut = self.detection_round(blacklist, dead)
            self.rounds += 1
            assert cut, "a round in which nobody proposes makes no progress"
            assert set(cut) <= blacklist, "only blacklisted nodes are ever cut"
            self.apply_cut(cut)
            assert self.w.view.getCurrentConfigurationId() != before
        return self.members


def random_hosts(n, count, seed, lo=0):
    return sorted(random.Random(seed).sample(range(lo, n), count))


def test_one_failure_out_of_five_nodes(orc):                                     # ClusterTest.java:212-224
    c = ScenarioCluster(orc, 5, seed=1)
    members = c.run(blacklist=[2], dead=[2])
    assert members == [0, 1, 3, 4] and c.w.view.getMembershipSize() == 4
    assert c.rounds == 1 and c.fast_decisions == 1                               # 4 votes = N - floor((N-1)/4) for N = 5


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_fail_random_quarter_of_nodes(orc, seed):                                # :275-291
    n, f = 50, 12
    failing = random_hosts(n, f, seed)
    c = ScenarioCluster(orc, n, seed)
    members = c.run(blacklist=failing, dead=failing)
    assert members == [m for m in range(n) if m not in failing]
    assert c.w.view.getMembershipSize() == n - f
    # 38 live voters are exactly the fast quorum of N = 50 (50 - floor(49/4)): the first round can decide on the fast path only if
    # EVERY live node proposes the same cut; later rounds (smaller N, same voters) have slack
    assert c.fast_decisions + c.classic_decisions == c.rounds


@pytest.mark.parametrize("seed", [6, 7, 8])
def test_fail_random_third_of_nodes(orc, seed):                                  # :299-315
    n, f = 50, 16
    failing = random_hosts(n, f, seed)
    c = ScenarioCluster(orc, n, seed)
    members = c.run(blacklist=failing, dead=failing)
    assert members == [m for m in range(n) if m not in failing]
    assert c.w.view.getMembershipSize() == n - f
    assert c.classic_decisions >= 1                                              # 34 voters < 38: the first view change needs the fallback


def test_a_third_of_the_nodes_can_block_the_cut(orc):
    """The same scenario with another draw (found by sweeping seeds): the detector is ALLOWED to stall.  A failed node Y whose
    observers on two rings are failed nodes that themselves collect fewer than L reports (most of THEIR observers are dead too)
    stops at H - 1 reports: its two silent observers are neither in proposal nor in preProposal, so invalidateFailingEdges
    (MultiNodeCutDetector.java:147-158) has nothing to add, updatesInProgress never returns to 0 and no live node ever proposes —
    for this configuration.  The reference behaves the same way (its test draws the failed set at random and can time out); what
    must hold is that every live node is stuck on the SAME subject with the SAME reports."""
    n, f, seed = 50, 16, 131
    failing = random_hosts(n, f, seed)
    c = ScenarioCluster(orc, n, seed)
    view = c.w.view
    cfg = view.getCurrentConfigurationId()
    live = [m for m in range(n) if m not in failing]
    handlers = {m: orc.AlertBatchHandler(view, K, H, L) for m in live}
    for tick in range(3):                                          # the static detector keeps firing: nothing changes
        batches = c.sender_batches(set(failing), set(failing), cfg)
        for r in live:
            order = list(batches)
            c.rng.shuffle(order)
            for o in order:
                assert handlers[r].handleBatch(batches[o]) == []
    def seen(r):                                                   # report counts, up to the watermark (beyond it they depend on arrival order)
        return {s: min(H, bin(handlers[r].reportMask(s)).count("1")) for s in failing}
    counts = seen(live[0])
    stuck = [s for s in failing if L <= counts[s] < H]
    dark = [s for s in failing if counts[s] < L]
    assert stuck and dark
    for r in live:                                                 # everyone holds the same reports
        assert seen(r) == counts
    for y in stuck:                                                # every missing ring of a stuck node is observed by a dark node
        obs = view.getObserversOf(y)
        missing = [k for k in range(K) if not (handlers[live[0]].reportMask(y) >> k) & 1]
        assert missing and all(obs[k] in dark for k in missing)
    with pytest.raises(AssertionError, match="nobody proposes"):
        ScenarioCluster(orc, n, seed).run(failing, failing)


@pytest.mark.parametrize("seed", [9, 10])
def test_fail_ten_random_nodes_that_stay_alive(orc, seed):                       # :322-336 (the static detector only: nobody shuts down)
    n, f = 50, 10
    failing = random_hosts(n, f, seed)
    c = ScenarioCluster(orc, n, seed)
    members = c.run(blacklist=failing, dead=[])
    assert members == [m for m in range(n) if m not in failing]
    assert c.w.view.getMembershipSize() == n - f


@pytest.mark.parametrize("seed", [13, 14, 15])
def test_concurrent_node_joins_and_fails(orc, seed):                             # :228-243
    """30 nodes, 5 of them fail while 10 others join: DOWN alerts from the live observers of the failed nodes and UP alerts from
    the joiners' live expected observers travel in the same windows.  A joiner that is not admitted by a view change asks again in
    the next configuration (new expected observers), exactly like a failed node keeps being reported — until the membership is
    the 35 nodes ClusterTest waits for."""
    n, f, nj = 30, 5, 10
    failing = list(range(2, 2 + f))                                              # basePort + 2 .. (the seed node stays)
    joiners = list(range(n, n + nj))
    c = ScenarioCluster(orc, n, seed, n_joiners=nj)
    dead, blacklist = set(failing), set(failing)
    while (blacklist & set(c.members)) or (set(joiners) - set(c.members)):
        assert c.rounds < 12, "no convergence"
        pending = [j for j in joiners if j not in c.members]

        def window(cfg, first=[True]):
            out = c.sender_batches(blacklist, dead, cfg)
            if first[0]:                                                         # a join attempt is one message per observer, not a tick
                c.join_alerts(pending, dead, cfg, out)
                first[0] = False
            return out

        cut = c.detection_round(blacklist, dead, ticks=2, source=window)
        c.rounds += 1
        assert cut and set(cut) <= blacklist | set(joiners)
        c.apply_cut(cut)
    assert sorted(c.members) == sorted([m for m in range(n) if m not in failing] + joiners)
    assert c.w.view.getMembershipSize() == n - f + nj
    for j in joiners:                                                            # identifiersSeen: the same NodeId can never join again
        with pytest.raises(orc.UUIDAlreadySeenException):
            c.w.view.ringDelete(j) or c.w.view.ringAdd(j, (int(c.joiner_ids[0][j - n]), int(c.joiner_ids[1][j - n])))
        c.w.view.ringAdd(j, (int(c.joiner_ids[0][j - n]) ^ 1, int(c.joiner_ids[1][j - n])))


def test_inject_asymmetric_drops(orc):                                           # :342-360
    """Ten nodes drop the first probes they receive (ingress only: their own probes and alerts still flow), the REAL detector
    (PingPongFailureDetector.java:38-121) does the rest: ten failed probes per edge, the notification on the eleventh run — by
    then the subjects answer again, which no longer matters (:71-79) — one AlertMessage per detector, and the cut removes nodes
    that are alive and voting."""
    n, f = 50, 10
    failing = random_hosts(n, f, seed=12, lo=1)
    c = ScenarioCluster(orc, n, seed=12)
    fd = orc.FdSim(c.w.view, K, np.arange(n))
    cfg0 = c.w.view.getCurrentConfigurationId()
    flags = np.zeros(n, np.uint8)
    flags[failing] = orc.FD_INGRESS_BLOCKED
    for _ in range(10):
        assert fd.tick(flags, cfg0) == []                                       # failures are counted, nobody is notified yet
    flags[:] = 0                                                                 # the interceptor has used up its drops

    def window(cfg):
        out = {}
        for o, s_, rings in fd.tick(flags, cfg):
            out.setdefault(o, []).append((o, s_, DOWN, cfg, rings))
        return out

    cut = c.detection_round(set(failing), set(), ticks=2, source=window)        # second window: every detector has notified, no alerts
    assert sorted(cut) == failing
    c.apply_cut(cut)
    assert c.w.view.getMembershipSize() == n - f and c.members == [m for m in range(n) if m not in failing]
    assert c.fast_decisions == 1                                                 # all 50 processes are alive and vote


def test_the_cut_of_a_round_is_what_a_single_detector_computes(orc):
    """Cross-check of the composition: with every observer alive (nobody is dead) and one delivery round, the decided cut equals the
    blacklist, and it is what ONE MultiNodeCutDetector fed the same alerts in sender order returns."""
    n = 50
    failing = random_hosts(n, 10, seed=11)
    c = ScenarioCluster(orc, n, seed=11)
    cfg = c.w.view.getCurrentConfigurationId()
    batches = c.sender_batches(set(failing), set(), cfg)
    cd = orc.MultiNodeCutDetector(c.w.u, K, H, L)
    got = []
    for o in sorted(batches):
        for (src, dst, st, _, rings) in batches[o]:
            got += cd.aggregateForProposal(src, dst, st, rings)
    got += cd.invalidateFailingEdges(c.w.view)
    cut = c.detection_round(set(failing), set(), ticks=1)
    assert sorted(cut) == failing == sorted(set(got))
