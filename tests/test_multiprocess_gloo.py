"""The N > 1 path on CPU: two processes (torch.distributed, gloo, 127.0.0.1) shard the virtual nodes by ring-0 range,
each computes its own nodes' proposals (with the oracle — this is a test), and the sharded tally protocol
(histogram sum all-reduce + verification max all-reduce, tests/sharding_model.py == csrc/fast_paxos.cu) must reach the
decision a single FastPaxos instance reaches over all votes."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import oracle_py as orc
    from rapid_b200 import workloads as W
    import sharding_model as S
    from rapid_b200._native import lib
    import ctypes as C

    def fingerprint(ids):
        a = np.ascontiguousarray(ids, np.int32); h1 = C.c_uint64(0); h2 = C.c_uint64(0)
        assert lib().rapid_proposal_fingerprint(a.ctypes.data_as(C.c_void_p), len(a), C.byref(h1), C.byref(h2)) == 0
        return h1.value, h2.value

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, K, H, L = 600, 10, 9, 4
    hb, off, ports = W.packed_endpoints(0, n)
    u = orc.Universe(); tags = u.add_bulk(hb, off, ports)
    hi, lo = W.node_ids(0, n)
    v = orc.MembershipView(u, K, tags, hi, lo)
    cfg = v.getCurrentConfigurationId()
    ring0 = np.asarray(v.getRing(0), np.int32)
    b = W.c2_simultaneous_crash(lambda ids: v.tables(ids)[0], n, 0.01)
    begin, R = S.shard_range(n, rank, world)
    # every position is owned by exactly one rank
    owned = torch.zeros(n, dtype=torch.int32); owned[begin:begin + R] = 1
    dist.all_reduce(owned); assert bool((owned == 1).all())
    blocked = W.blocked_by_receiver(b.blocked, ring0, begin, R)
    sim = orc.ClusterSim(v, K, H, L, R, receiver_base=begin)
    o_len, o_ann, o_ids, o_off = sim.apply_batch(b.src, b.dst, b.ring, b.status, np.full(len(b), cfg, np.int64), blocked=blocked)
    if rank == 1:      # make one node on rank 1 dissent, to exercise a second proposal in the histogram
        o_len = o_len.copy(); dissent = int(np.nonzero(o_len)[0][0])
    else:
        dissent = -1
    h1s, h2s, lens = [], [], []
    for r in range(R):
        if o_len[r]:
            ids = o_ids[o_off[r]: o_off[r + 1]]
            if r == dissent: ids = ids[:-1]
            a, c = fingerprint(ids); h1s.append(a); h2s.append(c); lens.append(len(ids))
    hist = torch.from_numpy(S.histogram_of(h1s))
    dist.all_reduce(hist)                                            # THE all-reduce of the path
    Q = n - (n - 1) // 4
    def words_of(bucket):
        w = S.verification_words(h1s, h2s, lens, bucket).view(np.int64).copy()
        # max over uint64 via two int64 halves is overkill here: gloo has no uint64; compare as (hi32, lo32) pairs
        t = torch.from_numpy(np.stack([(w.view(np.uint64) >> np.uint64(32)).astype(np.int64), (w.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)], 1))
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        allw = np.stack([(g[:, 0].numpy().astype(np.uint64) << np.uint64(32)) | g[:, 1].numpy().astype(np.uint64) for g in gathered])
        return allw.max(axis=0)
    decided, dh1, dh2, dlen, dcount = S.decide(hist.numpy(), words_of, Q)
    want = fingerprint(b.expected_cut)
    assert decided and (dh1, dh2, dlen) == (want[0], want[1], len(b.expected_cut)), (decided, dh1, want)
    live = n - int(b.blocked.sum())
    assert dcount == live - 1                                         # everyone but the dissenter
    # the single-all-reduce protocol the library uses first: count-weighted sums, exact division, check word
    sb = torch.from_numpy(S.sum_buffer_of(h1s, h2s, lens, len(h1s)).view(np.int64).copy())
    dist.all_reduce(sb)                                               # int64 sums wrap exactly like the device's uint64
    d2 = S.decide_sum(sb.numpy().view(np.uint64), Q)
    assert d2[0] and not d2[6] and (d2[1], d2[2], d2[3], d2[4], d2[5]) == (want[0], want[1], len(b.expected_cut), live - 1, live), d2
    # single-instance reference: one literal FastPaxosTally over all votes decides the same cut
    if rank == 0:
        fp = orc.FastPaxosTally(u, cfg, n)
        full = orc.ClusterSim(v, K, H, L, n)
        fl, fa, fi, fo = full.apply_batch(b.src, b.dst, b.ring, b.status, np.full(len(b), cfg, np.int64),
                                          blocked=W.blocked_by_receiver(b.blocked, ring0, 0, n))
        for r in range(n):
            if fl[r]: fp.handleFastRoundProposal(int(ring0[r]), cfg, fi[fo[r]: fo[r + 1]].tolist())
        assert fp.decided() and sorted(fp.decision()) == b.expected_cut.tolist()
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_process_sharded_tally(tmp_path, orc):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out[-3000:])
        assert "ok" in out


def test_shard_ranges_partition():
    import sharding_model as S
    for n in (1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [S.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and sum(c for _, c in edges) == n
            for (b0, c0), (b1, _) in zip(edges, edges[1:]):
                assert b0 + c0 == b1
