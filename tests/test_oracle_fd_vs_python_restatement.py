"""oracle/fd_oracle.hpp (FdSim) against tests/pyref.PyFdCluster: every interval of random scenarios — crashes, ingress / egress
blocks that come and go, bootstrapping subjects, individual probe failures — must raise the same AlertMessages (observer, subject,
ring numbers) in the same order."""
import random

import numpy as np
import pytest

import pyref
from helpers import OracleWorld

K = 10


@pytest.mark.parametrize("seed", range(12))
def test_intervals_of_random_scenarios(orc, seed):
    rng = random.Random(6000 + seed)
    n = rng.randint(2, 60)
    w = OracleWorld(orc, n, K)
    a = orc.FdSim(w.view, K, np.arange(n))
    b = pyref.PyFdCluster(w.view, range(n))
    flags = np.zeros(n, np.uint8)
    total = 0
    for interval in range(rng.randint(45, 70)):
        if rng.random() < 0.25:                                    # the scenario changes
            t = rng.randrange(n)
            flags[t] = rng.choice([0, 0, orc.FD_CRASHED, orc.FD_INGRESS_BLOCKED, orc.FD_EGRESS_BLOCKED, orc.FD_BOOTSTRAPPING,
                                   orc.FD_INGRESS_BLOCKED | orc.FD_BOOTSTRAPPING])
        ef = np.zeros((n, K), np.uint8)
        efd = {}
        if n > 1:
            for _ in range(rng.randint(0, 3)):
                m, j = rng.randrange(n), rng.randrange(K)
                ef[m, j] = 1
                efd[(m, j)] = True
        ga = a.tick(flags, 3, edge_fail=ef)
        gb = b.tick(flags, efd)
        assert ga == [(o, s, r) for o, s, r in gb], (seed, interval)
        total += len(ga)
        # detector state word by word
        for m in rng.sample(range(n), min(n, 5)):
            for j in range(a.numDetectors(m)):
                fc, noted = a.state(m, j)
                assert (fc, noted) == (b.fds[m][j].failureCount, b.fds[m][j].notified)
    assert total >= 0
