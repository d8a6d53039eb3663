"""Measure alert generation: one failure-detector interval of N x K ping-pong detectors on the device,
quiet (nobody notifies) and raising (1 % of the nodes crashed eleven intervals ago), against the oracle's literal detector
objects on one host core.

    python profiles/bench_fd.py [--nodes 1000000] [--cpu-nodes 20000]
Prints one JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--cpu-nodes", type=int, default=20_000)
    args = ap.parse_args()
    import rapid_b200 as rb
    from rapid_b200 import workloads as W
    n, K = args.nodes, 10
    hb, off, ports = W.packed_endpoints(0, n)
    v = rb.MembershipView.from_packed(K, hb, off, ports)
    fd = rb.EdgeFailureDetectors(v)
    flags = np.zeros(n, np.uint8)
    flags[W.pick_smallest(n, n // 100, 7)] = 1
    quiet_wall, quiet_dev = [], []
    for _ in range(10):
        t0 = time.perf_counter()
        assert fd.tick(flags, 3) == (0, 0)
        quiet_wall.append((time.perf_counter() - t0) * 1e3)
        quiet_dev.append(fd.lastDeviceMs())
    t0 = time.perf_counter()
    na, nc = fd.tick(flags, 3)
    raise_wall, raise_dev = (time.perf_counter() - t0) * 1e3, fd.lastDeviceMs()
    D = n * K
    res = {"nodes": n, "detectors": D,
           "gpu_quiet_interval_wall_ms": min(quiet_wall[2:]), "gpu_quiet_interval_device_ms": min(quiet_dev[2:]),
           "gpu_raising_interval_wall_ms": raise_wall, "gpu_raising_interval_device_ms": raise_dev, "alerts": na, "cells": nc,
           "gpu_detectors_per_s": D / (min(quiet_wall[2:]) * 1e-3),
           "note": "wall = the C-ABI call with host flags (1 B per node copied in); device = CUDA events around copy + kernels"}
    # CPU: the oracle's PingPongFailureDetector objects, one thread
    from helpers import OracleWorld
    from oracle import oracle_py as orc
    orc.build()
    m = args.cpu_nodes
    w = OracleWorld(orc, m, K)
    sim = orc.FdSim(w.view, K, np.arange(m))
    f2 = np.zeros(m, np.uint8)
    f2[W.pick_smallest(m, m // 100, 7)] = 1
    cpu = []
    for _ in range(11):
        t0 = time.perf_counter()
        sim.tick(f2, 3)
        cpu.append((time.perf_counter() - t0) * 1e3)
    res["cpu_oracle"] = {"nodes": m, "detectors": m * K, "quiet_interval_ms": min(cpu[:10]), "raising_interval_ms": cpu[10],
                         "detectors_per_s": m * K / (min(cpu[:10]) * 1e-3)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
