"""Paxos.selectProposalUsingCoordinatorRule: the oracle (oracle/paxos_oracle.hpp, pinned by the 36 PaxosTests table rows) against a
second restatement (tests/pyref.coordinator_rule) on random Phase1b lists — several ranks, empty vvals, values that differ only in
order, the N/4 threshold met by the first, a later or no value."""
import random

import pytest

import pyref


@pytest.mark.parametrize("seed", range(60))
def test_coordinator_rule_random_lists(orc, seed):
    rng = random.Random(5000 + seed)
    N = rng.choice([1, 2, 3, 4, 5, 6, 8, 9, 10, 16, 20, 33, 50])
    u = orc.Universe()
    tags = [u.add("172.16.0.%d" % i, 7000 + i) for i in range(8)]
    px = orc.ClassicPaxos(u, tags[0], 1, 5, N)
    values = [rng.sample(tags, rng.randint(1, 3)) for _ in range(rng.randint(1, 4))]
    if rng.random() < 0.4:
        values.append(list(reversed(values[0])))
    ranks = [(rng.randint(0, 2), rng.randint(0, 3)) for _ in range(rng.randint(1, 3))]
    for trial in range(20):
        msgs = []
        for _ in range(rng.randint(1, max(2, N + 2))):
            vval = [] if rng.random() < 0.25 else list(rng.choice(values))
            msgs.append({"vrnd": rng.choice(ranks), "vval": vval})
        assert px.selectProposalUsingCoordinatorRule(msgs) == pyref.coordinator_rule(N, msgs), (seed, trial, N, msgs)


def test_coordinator_rule_empty_list(orc):
    u = orc.Universe()
    t = u.add("h", 1)
    with pytest.raises(ValueError):
        orc.ClassicPaxos(u, t, 1, 5, 4).selectProposalUsingCoordinatorRule([])
    with pytest.raises(ValueError):
        pyref.coordinator_rule(4, [])
