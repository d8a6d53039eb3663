"""Times BASELINE config 4 (100,000 nodes, 1 % flip-flop alert stream over T = 8 batches, duplicates, every receiver applying
each batch in its own permuted order) on one GPU: the k_apply_generic path.  Not part of the bench.py contract — a profiling
aid whose output is kept under profiles/.   usage: python profiles/bench_generic.py [nodes] [repeats]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rapid_b200 as rb  # noqa: E402
from rapid_b200 import workloads as W  # noqa: E402

K, H, L = 10, 9, 4
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hb, off, ports = W.packed_endpoints(0, n)
view = rb.MembershipView.from_packed(K, hb, off, ports)
obs, _ = view.tables()
ring0 = view.getRing(0)
hi, lo = W.node_ids(0, n)
cfg = view.getCurrentConfigurationId(hi, lo)
batches = W.c4_flip_flop_stream(obs, n, 0.01, T=8)
blocked = W.blocked_by_receiver(batches[0].blocked, ring0, 0, n)
cl = rb.VirtualCluster(view, H, L, max_subjects=len(batches[-1].expected_cut) + 64)
fp = rb.FastPaxos(cfg, n)
want = rb.proposal_fingerprint(batches[-1].expected_cut)
rows = []
for rep in range(reps + 1):
    cl.clear(); fp.reset(cfg)
    tot_ms = main_ms = tally_ms = 0.0
    cells = 0
    decided_at = None
    for t, b in enumerate(batches):
        cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, perm_seed=b.meta["perm_seed"], read_outputs=False)
        a, m = cl.lastDeviceMs()
        tot_ms += a; main_ms += m; cells += len(b)
        r = fp.tallyCluster(cl)
        tally_ms += fp.lastDeviceMs()
        if r.decided:
            decided_at = t
            break
    assert decided_at is not None and (r.hash, r.hash2) == want
    if rep:
        rows.append((tot_ms, main_ms, tally_ms))
tot, main, tal = (float(np.median([x[i] for x in rows])) for i in range(3))
print(json.dumps({"workload": "C4 %d-node flip-flop stream, 8 batches, permuted per-receiver order" % n, "cells": cells,
                  "subjects": len(batches[-1].expected_cut), "decided_at_batch": decided_at, "apply_ms": tot, "k_apply_generic_ms": main,
                  "tally_ms": tal, "cells_per_s": cells / ((tot + tal) * 1e-3), "stats_last_batch": cl.debugStats()}))
