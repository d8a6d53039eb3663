"""Measure the view rows: building the K rings of N endpoints on the device (XXH64 ring keys, K radix
sorts, observer / subject tables), the configuration id, and applying a decided cut (1 % leave + 0.5 % join -> rings rebuilt),
against the oracle's literal MembershipView (K red-black trees, memoised comparator) on one host core at a smaller N.

    python profiles/bench_view.py [--nodes 1000000] [--cpu-nodes 50000]
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
    ap.add_argument("--cpu-nodes", type=int, default=50_000)
    args = ap.parse_args()
    import rapid_b200 as rb
    from rapid_b200 import workloads as W
    n, K = args.nodes, 10
    hb, off, ports = W.packed_endpoints(0, n)
    hi, lo = W.node_ids(0, n)
    build = []
    for _ in range(3):
        t0 = time.perf_counter()
        v = rb.MembershipView.from_packed(K, hb, off, ports)
        build.append((time.perf_counter() - t0) * 1e3)
        if _ < 2:
            v.close() if hasattr(v, "close") else None
    t0 = time.perf_counter()
    cfg = v.getCurrentConfigurationId(hi, lo)
    cfg_ms = (time.perf_counter() - t0) * 1e3
    nj = n // 200
    jh, jp = W.endpoints(n, nj)
    t0 = time.perf_counter()
    v.registerJoiners(jh, jp)
    reg_ms = (time.perf_counter() - t0) * 1e3
    v.setNodeIds(hi, lo)
    jhi, jlo = W.node_ids(n, nj)
    v.setJoinerIds(n, jhi, jlo)
    t0 = time.perf_counter()
    cfg_dev = v.currentConfigurationId()
    cfg_dev_ms = (time.perf_counter() - t0) * 1e3
    assert cfg_dev == cfg
    cut = np.concatenate([W.pick_smallest(n, n // 100, 7), np.arange(n, n + nj)]).astype(np.int32)
    t0 = time.perf_counter()
    v.applyCut(cut, want_map=False)
    cut_ms = (time.perf_counter() - t0) * 1e3
    assert v.getMembershipSize() == n - n // 100 + nj
    t0 = time.perf_counter()
    cfg_after = v.currentConfigurationId()
    cfg_after_ms = (time.perf_counter() - t0) * 1e3
    # a second cut of the same size on the updated view (buffers warm)
    nj2 = n // 200
    jh2, jp2 = W.endpoints(n + nj, nj2)
    first2 = v.registerJoiners(jh2, jp2)[0]
    h2, l2 = W.node_ids(n + nj, nj2)
    v.setJoinerIds(first2, h2, l2)
    m_now = v.getMembershipSize()
    cut2 = np.concatenate([W.pick_smallest(m_now, n // 100, 11), np.arange(first2, first2 + nj2)]).astype(np.int32)
    t0 = time.perf_counter()
    v.applyCut(cut2, want_map=False)
    cut2_ms = (time.perf_counter() - t0) * 1e3
    res = {"nodes": n, "K": K, "gpu_build_ms": min(build), "gpu_build_first_ms": build[0], "gpu_configuration_id_ms": cfg_ms,
           "gpu_register_joiners_ms": reg_ms, "joiners": nj, "gpu_apply_cut_ms": cut_ms, "gpu_apply_cut_second_ms": cut2_ms,
           "cut_size": int(len(cut)), "gpu_configuration_id_device_resident_ms": cfg_dev_ms,
           "gpu_configuration_id_after_cut_ms": cfg_after_ms, "configuration_id_after_cut": int(cfg_after),
           "configuration_id": int(cfg),
           "note": "wall time of the C-ABI calls with host arrays (endpoint bytes copied in, tables stay on the device)"}
    from oracle import oracle_py as orc
    orc.build()
    m = args.cpu_nodes
    hb2, off2, ports2 = W.packed_endpoints(0, m)
    u = orc.Universe()
    tags = u.add_bulk(hb2, off2, ports2)
    h2, l2 = W.node_ids(0, m)
    t0 = time.perf_counter()
    ov = orc.MembershipView(u, K, tags, h2, l2)
    cpu_build = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    ov.getCurrentConfigurationId()
    cpu_cfg = (time.perf_counter() - t0) * 1e3
    dead = W.pick_smallest(m, m // 100, 7)
    t0 = time.perf_counter()
    for d in dead.tolist():
        ov.ringDelete(int(d))
    cpu_cut = (time.perf_counter() - t0) * 1e3
    res["cpu_oracle"] = {"nodes": m, "build_ms": cpu_build, "configuration_id_ms": cpu_cfg, "ring_delete_1pct_ms": cpu_cut,
                         "note": "literal MembershipView: K ordered sets with the memoised hash comparator, one thread"}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
