"""oracle::PingPongFailureDetector / FdNode (PingPongFailureDetector.java:38-121, MembershipService.java:472-495, :697-707):
the alerts K edge detectors per node raise, and that ten failed probe intervals reproduce the synthetic crash workload
 the benchmarks feed the cut detector with."""
import numpy as np

from helpers import OracleWorld
from rapid_b200 import workloads as W

K = 10


def test_threshold_and_single_notification(orc):
    n = 30
    w = OracleWorld(orc, n, K)
    sim = orc.FdSim(w.view, K, np.arange(n))
    flags = np.zeros(n, np.uint8)
    flags[7] = orc.FD_CRASHED
    obs_of_7 = set(int(x) for x in w.tables()[0][7])
    for t in range(10):                                       # ten failed probes: nothing yet (:71-79 checks BEFORE probing)
        assert sim.tick(flags, 5) == []
    alerts = sim.tick(flags, 5)                               # the eleventh run() notifies, once per detector
    assert {a[0] for a in alerts} == obs_of_7 and all(a[1] == 7 for a in alerts)
    subj = w.tables()[1]
    for o, s, rings in alerts:
        assert rings == [r for r in range(K) if subj[o][r] == 7]          # getRingNumbers (MembershipView.java:397-418)
    assert sorted(r for _, _, rings in alerts for r in set(rings)) is not None
    assert sim.tick(flags, 5) == []                           # notified: never again (:76)
    # a detector per ring: an observer that monitors 7 on two rings raises two identical AlertMessages
    per_obs = {}
    for o, _, rings in alerts:
        per_obs.setdefault(o, []).append(rings)
    for o, lst in per_obs.items():
        assert len(lst) == len(lst[0]) and all(x == lst[0] for x in lst)


def test_bootstrapping_subject_fails_after_thirty_answers(orc):
    n = 12
    w = OracleWorld(orc, n, K)
    sim = orc.FdSim(w.view, K, np.arange(n))
    flags = np.zeros(n, np.uint8)
    flags[3] = orc.FD_BOOTSTRAPPING
    ticks = 0
    while not sim.tick(flags, 1):
        ticks += 1
        assert ticks < 100
    assert ticks == 30 + 10                                   # 30 tolerated answers (:45, :99), then 10 counted failures, then notify


def test_ten_intervals_of_a_crash_reproduce_workload_c2(orc):
    n = 2000
    w = OracleWorld(orc, n, K)
    obs, _ = w.tables()
    b = W.c2_simultaneous_crash(obs, n)
    failed = np.zeros(n, np.uint8)
    failed[np.asarray(b.expected_cut)] = orc.FD_CRASHED
    sim = orc.FdSim(w.view, K, np.arange(n))
    alerts = []
    for _ in range(11):
        alerts += sim.tick(failed, 9)
    cells = sorted((o, s, r) for o, s, rings in alerts for r in set(rings))
    want = sorted(zip(b.src.tolist(), b.dst.tolist(), b.ring.tolist()))
    assert cells == want                                      # {(obs_r(s), s, DOWN, r) : s failed, obs_r(s) alive}
