"""The oracle's R-virtual-node driver (orc_sim_apply_batch — what every GPU parity test at R > 1 compares against) against R
independent pyref.PyBatchHandlers fed cell by cell from plain Python: uniform delivery, blocked receivers, per-receiver delivery
bitmaps, per-receiver PERMUTED order (keys splitmix64(splitmix64(seed + receiver_base + r) ^ cell), include/rapid_b200.h), state
carried over several batches, shards (receiver_base), stale configuration ids."""
import random

import numpy as np
import pytest

import pyref
from helpers import OracleWorld
from test_oracle_vs_python_restatement import same_reports
from rapid_b200 import workloads as W

K = 10
M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def test_splitmix64_is_the_workload_generators(orc):
    xs = np.array([0, 1, 2, 0x5241504944, M64], np.uint64)
    assert [splitmix64(int(x)) for x in xs] == [int(v) for v in W.splitmix64(xs)]
    assert all(orc.splitmix64(int(x)) == splitmix64(int(x)) for x in xs)


@pytest.mark.parametrize("seed", range(24))
def test_cluster_driver_delivery_modes(orc, seed):
    rng = random.Random(8000 + seed)
    n = rng.randint(12, 60)
    H, L = rng.choice([(9, 4), (8, 3), (8, 2)])
    w = OracleWorld(orc, n, K, n_joiners=2)
    cfg = w.view.getCurrentConfigurationId()
    R = rng.randint(1, n)
    base = rng.randint(0, n - R)
    sim = orc.ClusterSim(w.view, K, H, L, R, receiver_base=base)
    py = [pyref.PyBatchHandler(w.view, K, H, L) for _ in range(R)]
    mode = ["uniform", "blocked", "bitmap", "permuted", "permuted+bitmap"][seed % 5]
    failed = rng.sample(range(n), rng.randint(1, 4))
    obs = {s: w.view.getObserversOf(s) for s in failed}
    pending = [(obs[s][k], s, k, pyref.DOWN) for s in failed for k in range(K) if obs[s][k] not in failed]
    pending += [(w.view.getExpectedObserversOf(n)[k], n, k, pyref.UP) for k in range(K)]       # and one joiner
    rng.shuffle(pending)
    n_batches = rng.randint(1, 4)
    for bi in range(n_batches):
        take = pending[bi::n_batches]
        take += [rng.choice(pending) for _ in range(rng.randint(0, 5))]                         # duplicates
        A = len(take)
        src = np.array([c[0] for c in take], np.int32); dst = np.array([c[1] for c in take], np.int32)
        ring = np.array([c[2] for c in take], np.uint8); status = np.array([c[3] for c in take], np.uint8)
        cfgs = np.array([cfg if rng.random() < 0.95 else cfg ^ 1 for _ in take], np.int64)
        blocked = np.array([1 if ("blocked" in mode and rng.random() < 0.2) else 0 for _ in range(R)], np.uint8) if mode == "blocked" else None
        bitmap = None
        if "bitmap" in mode:
            words = (R + 31) // 32
            bitmap = np.zeros((A, words), np.uint32)
            for i in range(A):
                for r in range(R):
                    if rng.random() < 0.8:
                        bitmap[i, r >> 5] |= np.uint32(1 << (r & 31))
        perm_seed = rng.getrandbits(64) if "permuted" in mode else None
        o_len, o_ann, o_ids, o_off = sim.apply_batch(src, dst, ring, status, cfgs, blocked=blocked, bitmap=bitmap, perm_seed=perm_seed)
        for r in range(R):
            if blocked is not None and blocked[r]:
                got = set()
            else:
                cells = [i for i in range(A) if bitmap is None or (int(bitmap[i, r >> 5]) >> (r & 31)) & 1]
                if perm_seed is not None:
                    rs = splitmix64((perm_seed + base + r) & M64)
                    cells.sort(key=lambda i: (splitmix64(rs ^ i), i))
                got = py[r].handleBatch([(int(src[i]), int(dst[i]), int(status[i]), int(cfgs[i]), [int(ring[i])]) for i in cells])
            assert set(o_ids[o_off[r]: o_off[r + 1]].tolist()) == got and o_len[r] == len(got), (seed, bi, r, mode)
            assert bool(o_ann[r]) == py[r].announcedProposal
            assert sim.numProposals(r) == py[r].cd.getNumProposals()
            assert sim.updatesInProgress(r) == py[r].cd.updatesInProgress
            for t in failed + [n]:
                assert same_reports(sim.reportMask(r, t), py[r].cd.reportMask(t), H)
