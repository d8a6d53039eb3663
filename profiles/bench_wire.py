"""Measure the wire-format ingest: one serialized BatchedAlertMessage of the C5 shape (10^5 single-ring
AlertMessages over a 10^6-node view) decoded on the device, steady state, against the protobuf runtime (upb, C) parsing
the same bytes on one host core.

    python profiles/bench_wire.py [--nodes 1000000] [--messages 100000]
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
    ap.add_argument("--messages", type=int, default=100_000)
    args = ap.parse_args()
    import rapid_b200 as rb
    from rapid_b200 import workloads as W
    import wire_proto
    from wire_proto import field, varint
    pb = wire_proto.build()
    n, M, K = args.nodes, args.messages, 10
    hb, off, ports = W.packed_endpoints(0, n)
    view = rb.MembershipView.from_packed(K, hb, off, ports)
    rng = np.random.default_rng(1)
    subj, obs, rings = rng.integers(0, n, M), rng.integers(0, n, M), rng.integers(0, K, M)
    hosts, ports_l = W.endpoints(0, n)
    parts = []
    for o, s, r in zip(obs.tolist(), subj.tolist(), rings.tolist()):
        alert = (field(1, 2, field(1, 2, hosts[o]) + field(2, 0, varint(int(ports_l[o])))) +
                 field(2, 2, field(1, 2, hosts[s]) + field(2, 0, varint(int(ports_l[s])))) +
                 field(3, 0, varint(1)) + field(4, 0, varint(42)) + field(5, 2, varint(r)))
        parts.append(field(3, 2, alert))
    data = b"".join(parts)
    dec = rb.WireDecoder(view)
    wall, dev = [], []
    for _ in range(6):
        t0 = time.perf_counter()
        got = dec.decodeBatchedAlertMessage(data)
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(dec.lastDeviceMs())
    assert got.n_cells == M
    _, dst, _, _, _ = dec.cells()
    assert (dst == subj).all()
    cpu = []
    for _ in range(3):
        t0 = time.perf_counter()
        msg = pb.BatchedAlertMessage.FromString(data)
        cpu.append((time.perf_counter() - t0) * 1e3)
    # the runtime only builds objects; the reference additionally looks every Endpoint up (MembershipService.java:653-664)
    res = {"nodes": n, "messages": M, "bytes": len(data),
           "gpu_wall_ms": min(wall[1:]), "gpu_device_ms": min(dev[1:]), "gpu_first_call_ms": wall[0],
           "gpu_messages_per_s": M / (min(wall[1:]) * 1e-3), "gpu_MB_per_s": len(data) / (min(wall[1:]) * 1e-3) / 1e6,
           "cpu_protobuf_runtime_parse_ms": min(cpu), "cpu_messages_per_s": M / (min(cpu) * 1e-3),
           "cpu_note": "google.protobuf %s (upb) FromString on one core: parse only, no Endpoint -> id lookups, no cell expansion" % __import__("google.protobuf").protobuf.__version__,
           "n_messages_parsed_by_runtime": len(msg.messages)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
