"""Worker of tests/test_gpu_nccl.py — one process per GPU under torchrun (NCCL).  NOT a pytest module.

Every rank owns a contiguous ring-0 range of the virtual nodes.  Three scenarios, all checked against the oracle:

  stream    BASELINE config 4 scaled to `n` nodes: flip-flop stream, per-receiver permuted order, state carried over 8 batches.
            On EVERY batch a window of >= 256 receivers of this rank is compared, receiver by receiver, with
            orc.ClusterSim(..., receiver_base=begin + window) (proposal length / fingerprints / announced flags, report masks
            and updatesInProgress of a few of them); the sharded tally (ONE NCCL all-reduce) must decide the injected cut with
            the same vote count on every rank, in the batch in which the oracle's nodes announce it.
  dissent   Two proposals in ONE fast round: a per-receiver delivery bitmap keeps the cells about one extra subject `x` from a
            fifth of the receivers, so they announce the cut WITHOUT x.  The majority proposal must win with exactly its own
            votes.  Then the same again with RAPID_B200_FORCE_REFINE=1 (digit-by-digit refinement with max-all-reduces instead
            of the sum buffer): same decision.
  collide   x is chosen so that the two proposals fall into the SAME 12-bit bucket of the all-reduce buffer: the sum check
            fails (ambiguous bucket) and the refinement path runs for real on NCCL.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K, H, L = 10, 9, 4


def main():
    import torch
    import torch.distributed as dist
    import rapid_b200 as rb
    from rapid_b200 import workloads as W
    from oracle import oracle_py as orc

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
    window = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        orc.build()
    dist.barrier()

    hb, off, ports = W.packed_endpoints(0, n)
    view = rb.MembershipView.from_packed(K, hb, off, ports, device=local)
    u = orc.Universe()
    tags = u.add_bulk(hb, off, ports)
    hi, lo = W.node_ids(0, n)
    oview = orc.MembershipView(u, K, tags, hi, lo)
    cfg = view.getCurrentConfigurationId(hi, lo)
    assert cfg == oview.getCurrentConfigurationId()
    ring0 = view.getRing(0)
    assert ring0.tolist() == oview.getRing(0)
    obs, _ = view.tables()
    begin = rank * n // world
    R = (rank + 1) * n // world - begin
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.from_numpy(rb.NcclComm.unique_id()))
    dist.broadcast(uid, 0)
    comm = rb.NcclComm(rank, world, uid.cpu().numpy(), local)
    Q = rb.quorum(n)

    def same_on_all_ranks(vals):
        t = torch.tensor(vals, dtype=torch.int64, device="cuda")
        lo_t, hi_t = t.clone(), t.clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        assert bool((lo_t == hi_t).all()), "ranks disagree: %r" % (vals,)

    def compare_window(cl, sim, w0, res, o):
        o_len, o_ann, o_ids, o_off = o
        sl = slice(w0, w0 + window)
        np.testing.assert_array_equal(res.proposal_len[sl], o_len)
        np.testing.assert_array_equal(res.announced[sl], o_ann)
        for r in range(window):
            if o_len[r]:
                h = rb.proposal_fingerprint(o_ids[o_off[r]: o_off[r + 1]])
                assert (int(res.proposal_hash[w0 + r]), int(res.proposal_hash2[w0 + r])) == h, "receiver %d" % (w0 + r)
        live = np.nonzero(o_ann == 0)[0]
        for r in live[:: max(1, len(live) // 4)][:4]:
            for subj, m in cl.debugMasks(int(w0 + r)).items():
                assert sim.reportMask(int(r), int(subj)) == m
            assert cl.debugCounters(int(w0 + r))[0] == sim.updatesInProgress(int(r))

    # ---------------------------------------------------------------- stream (C4 shape) ----------------------------------------
    batches = W.c4_flip_flop_stream(obs, n, 0.01, T=8)
    cut = batches[-1].expected_cut
    blocked = W.blocked_by_receiver(batches[0].blocked, ring0, begin, R)
    cl = rb.VirtualCluster(view, H, L, n_receivers=R, receiver_begin=begin, max_subjects=len(cut) + 64)
    fp = rb.FastPaxos(cfg, n, sender_capacity=n, device=local)
    w0 = (R - window) // 3
    sim = orc.ClusterSim(oview, K, H, L, window, receiver_base=begin + w0)
    want = rb.proposal_fingerprint(cut)
    decided_at = None
    for bi, b in enumerate(batches):
        o = sim.apply_batch(b.src, b.dst, b.ring, b.status, np.full(len(b), cfg, np.int64), blocked=blocked[w0: w0 + window],
                            perm_seed=b.meta["perm_seed"], threads=4)
        res = cl.handleBatch(cfg, None, b.dst, b.ring, b.status, blocked=blocked, perm_seed=b.meta["perm_seed"])
        assert cl.lastPath()[0] == 4
        compare_window(cl, sim, w0, res, o)
        t = fp.tallyCluster(cl, comm)
        same_on_all_ranks([int(t.decided), t.count, t.votes_received, t.length, t.hash & 0x7FFFFFFFFFFFFFFF])
        if t.decided and decided_at is None:
            decided_at = bi
            assert (t.hash, t.hash2, t.length) == (want[0], want[1], len(cut))
            assert t.count >= Q and t.count == n - len(cut)      # every live node voted for the cut
            # ... and the oracle's nodes announced exactly that cut in this batch
            who = np.nonzero(o[0])[0]
            assert len(who) and sorted(o[2][o[3][who[0]]: o[3][who[0] + 1]].tolist()) == cut.tolist()
    assert decided_at is not None, "the stream did not converge"
    del cl, fp

    # ---------------------------------------------------------------- dissent / collide ----------------------------------------
    base = W.pick_smallest(n, 12, W.SEED + 77)                   # crashed in every receiver's eyes
    h_base = rb.proposal_fingerprint(np.sort(base))
    bucket = lambda h: h >> 52
    base_set = set(base.tolist())
    candidates = [x for x in range(n) if x not in base_set]
    x_far = x_same = None
    for x in candidates:
        hx = rb.proposal_fingerprint(np.sort(np.append(base, x)))
        if bucket(hx[0]) == bucket(h_base[0]):
            x_same = x_same if x_same is not None else x
        elif x_far is None:
            x_far = x
        if x_far is not None and x_same is not None:
            break
    assert x_far is not None
    assert x_same is not None, "no 12-bit bucket collision among %d candidates" % len(candidates)

    def two_proposals(x, force_refine):
        failed = np.append(base, x).astype(np.int32)
        cells = W.crash_cells(obs, np.sort(failed), n)
        src, dst, ring, status = cells["src"], cells["dst"], cells["ring"], cells["status"]
        bl = np.zeros(n, np.uint8)
        bl[failed] = 1
        blocked = W.blocked_by_receiver(bl, ring0, begin, R)
        # receivers whose GLOBAL ring-0 position is a multiple of 5 never hear about x
        gpos = begin + np.arange(R)
        deaf = (gpos % 5 == 0)
        words = (R + 31) // 32
        row_all = np.full(words, 0xFFFFFFFF, np.uint32)
        row_deaf = np.zeros(words, np.uint32)
        hear = np.nonzero(~deaf)[0]
        np.bitwise_or.at(row_deaf, hear >> 5, (np.uint32(1) << (hear & 31).astype(np.uint32)))
        bitmap = np.where((dst == x)[:, None], row_deaf[None, :], row_all[None, :]).astype(np.uint32)
        cl = rb.VirtualCluster(view, H, L, n_receivers=R, receiver_begin=begin)
        fp = rb.FastPaxos(cfg, n, sender_capacity=n, device=local)
        w0 = (R - window) // 2
        sim = orc.ClusterSim(oview, K, H, L, window, receiver_base=begin + w0)
        # the oracle wants the bitmap of its own receivers: columns w0 .. w0 + window
        cols = np.arange(w0, w0 + window)
        bits = (bitmap[:, cols >> 5] >> (cols & 31).astype(np.uint32)) & 1
        wwords = (window + 31) // 32
        obm = np.zeros((len(dst), wwords), np.uint32)
        for j in range(window):
            obm[:, j >> 5] |= (bits[:, j].astype(np.uint32) << np.uint32(j & 31))
        o = sim.apply_batch(src, dst, ring, status, np.full(len(dst), cfg, np.int64), blocked=blocked[w0: w0 + window], bitmap=obm, threads=4)
        res = cl.handleBatch(cfg, src, dst, ring, status, blocked=blocked, bitmap=bitmap)
        compare_window(cl, sim, w0, res, o)
        if force_refine:
            os.environ["RAPID_B200_FORCE_REFINE"] = "1"
        try:
            t = fp.tallyCluster(cl, comm)
        finally:
            os.environ.pop("RAPID_B200_FORCE_REFINE", None)
        same_on_all_ranks([int(t.decided), t.count, t.votes_received, t.length, t.hash & 0x7FFFFFFFFFFFFFFF])
        # expected: live receivers that hear everything vote base + x, the deaf ones vote base
        live_all = np.ones(n, bool)
        live_all[failed] = False
        pos_of = np.empty(n, np.int64)
        pos_of[ring0] = np.arange(n)
        deaf_node = (pos_of % 5 == 0)
        n_major = int((live_all & ~deaf_node).sum())
        n_minor = int((live_all & deaf_node).sum())
        assert t.votes_received == n_major + n_minor
        hmaj = rb.proposal_fingerprint(np.sort(failed))
        if n_major >= Q:
            assert t.decided and (t.hash, t.hash2, t.length) == (hmaj[0], hmaj[1], len(failed)) and t.count == n_major
        else:
            assert not t.decided
        return t

    two_proposals(x_far, False)
    two_proposals(x_far, True)          # refinement path forced
    t = two_proposals(x_same, False)    # refinement path taken because the bucket is ambiguous
    if rank == 0:
        print("nccl worker ok: world=%d n=%d stream decided in batch %d; dissent x=%d, colliding x=%d, majority count %d" % (
            world, n, decided_at, x_far, x_same, t.count), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
