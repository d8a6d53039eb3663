"""Paxos.java, both restatements in lockstep: N oracle nodes and N pyref.PyPaxos nodes receive the SAME messages in the same
seeded random interleaving — fast-round votes registered at some nodes, several would-be coordinators (also with equal rounds:
the rank's node index decides), stale configurations, message loss — and every handler must answer with the same message, every
node must hold the same (rnd, vrnd, vval, crnd, cval) and reach the same decision at the same message."""
import random

import pytest

import pyref

CFG = 9


def norm(m):
    if m is None:
        return None
    out = dict(m)
    for k in ("rank", "rnd", "vrnd"):
        if k in out:
            out[k] = tuple(out[k])
    return out


def lockstep_run(orc, seed):
    rng = random.Random(7000 + seed)
    N = rng.choice([1, 2, 3, 4, 5, 6, 7, 10, 11, 20])
    u = orc.Universe()
    tags = [u.add("192.168.1.%d" % i, 4000 + i) for i in range(N)]
    hashes = rng.sample(range(-50, 50), N)                         # Endpoint.hashCode() may be negative
    a = [orc.ClassicPaxos(u, tags[i], hashes[i], CFG, N) for i in range(N)]
    b = [pyref.PyPaxos(tags[i], hashes[i], CFG, N) for i in range(N)]
    values = [sorted(rng.sample(range(100, 120), rng.randint(1, 3))) for _ in range(rng.randint(1, 3))]
    vtags = {}
    def tagv(v):
        return [vtags.setdefault(x, u.add("10.9.9.%d" % x, x)) for x in v]
    values = [tagv(v) for v in values]
    for i in range(N):                                             # some nodes voted in the fast round
        if rng.random() < 0.6:
            v = rng.choice(values)
            a[i].registerFastRoundVote(v); b[i].registerFastRoundVote(v)
    inbox = [[] for _ in range(N)]
    loss = rng.choice([0.0, 0.0, 0.1, 0.3])

    def broadcast(kind, m):
        for q in inbox:
            if rng.random() >= loss:
                q.append((kind, m))

    for c in rng.sample(range(N), rng.randint(1, min(N, 3))):      # recovery timers: a few coordinators, rounds 2 (and sometimes 3)
        r = rng.choice([2, 2, 3])
        ma, mb = a[c].startPhase1a(r), b[c].startPhase1a(r)
        assert norm(ma) == norm(mb)
        if ma:
            if rng.random() < 0.1:
                ma = dict(ma, cfg=CFG + 1)                         # a straggler from another configuration
            broadcast("1a", ma)
    steps = 0
    while any(inbox) and steps < 20000:
        steps += 1
        i = rng.choice([j for j in range(N) if inbox[j]])
        kind, m = inbox[i].pop(0)
        if kind == "1a":
            ra, rb_ = a[i].handlePhase1aMessage(m), b[i].handlePhase1aMessage(m)
            assert norm(ra) == norm(rb_), (seed, steps)
            if ra and rng.random() >= loss:
                inbox[tags.index(m["sender"])].append(("1b", ra))
        elif kind == "1b":
            ra, rb_ = a[i].handlePhase1bMessage(m), b[i].handlePhase1bMessage(m)
            assert norm(ra) == norm(rb_), (seed, steps)
            if ra:
                broadcast("2a", ra)
        elif kind == "2a":
            ra, rb_ = a[i].handlePhase2aMessage(m), b[i].handlePhase2aMessage(m)
            assert norm(ra) == norm(rb_), (seed, steps)
            if ra:
                broadcast("2b", ra)
        else:
            da, db = a[i].handlePhase2bMessage(m), b[i].handlePhase2bMessage(m)
            assert da == db, (seed, steps)
        ranks = a[i].ranks()
        assert (ranks["rnd"], ranks["vrnd"], ranks["crnd"]) == (b[i].rnd, b[i].vrnd, b[i].crnd)
        assert a[i].vval() == b[i].vval and a[i].cval() == b[i].cval
        assert a[i].decided() == b[i].decided
        if b[i].decided:
            assert a[i].decision() == b[i].decision
    decisions = {tuple(x.decision) for x in b if x.decided}
    assert len(decisions) <= 1                                      # agreement
    return sum(1 for x in b if x.decided), N


@pytest.mark.parametrize("seed", range(40))
def test_lockstep_random_runs(orc, seed):
    lockstep_run(orc, seed)


def test_the_runs_are_not_vacuous(orc):
    stats = [lockstep_run(orc, seed) for seed in range(40)]
    assert sum(1 for d, n in stats if d == n) >= 10                 # runs in which EVERY node decided
    assert sum(1 for d, n in stats if 0 < d < n) >= 1               # and runs cut short by message loss
