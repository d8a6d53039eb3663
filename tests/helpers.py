"""Shared test plumbing: build the same view in the oracle and (when a GPU is present) in librapid_b200, run the
same alert batches through both, and compare every observable."""
import numpy as np

from rapid_b200 import workloads as W

K_DEFAULT = 10


class OracleWorld:
    """Oracle universe + view over the synthetic endpoints 0..n-1 (+ optional joiners n..): tags == node ids."""

    def __init__(self, orc, n, K=K_DEFAULT, n_joiners=0):
        self.orc = orc
        self.n, self.K = n, K
        self.u = orc.Universe()
        hb, off, ports = W.packed_endpoints(0, n + n_joiners)
        self.packed = (hb, off, ports)
        tags = self.u.add_bulk(hb, off, ports)
        assert (tags == np.arange(n + n_joiners)).all()
        self.id_high, self.id_low = W.node_ids(0, n)
        self.view = orc.MembershipView(self.u, K, np.arange(n, dtype=np.int32), self.id_high, self.id_low)
        self.n_joiners = n_joiners

    def tables(self):
        return self.view.tables(np.arange(self.n, dtype=np.int32))

    def ring0(self):
        return np.asarray(self.view.getRing(0), np.int32)

    def joiner_obs(self):
        return np.asarray([self.view.getExpectedObserversOf(self.n + j) for j in range(self.n_joiners)], np.int32)

    def member_packed(self):
        hb, off, ports = self.packed
        return hb[: off[self.n]], off[: self.n + 1], ports[: self.n]

    def joiner_endpoints(self):
        hosts, ports = W.endpoints(self.n, self.n_joiners)
        return hosts, ports


def fingerprints_from_oracle(rb, out_len, out_ids, off):
    """expected (h1, h2) per receiver from the oracle's proposal lists"""
    R = len(out_len)
    h1 = np.zeros(R, np.uint64)
    h2 = np.zeros(R, np.uint64)
    cache = {}
    for r in range(R):
        if out_len[r]:
            key = out_ids[off[r]: off[r + 1]].tobytes()
            if key not in cache:
                cache[key] = rb.proposal_fingerprint(out_ids[off[r]: off[r + 1]])
            h1[r], h2[r] = cache[key]
    return h1, h2


def compare_batch(rb, orc_world, sim, cluster, cfg, batch_arrays, blocked=None, bitmap=None, perm_seed=None,
                  cell_cfg=None, check_masks=True, sample_proposals=8):
    """Apply one batch to the oracle sim and to the GPU cluster; assert identical observables.
    Returns (oracle out_len, oracle announced)."""
    src, dst, ring, status = batch_arrays
    A = len(dst)
    if cfg is None:     # the oracle filters against the view's real configuration id (MembershipService.java:653)
        cfg = orc_world.view.getCurrentConfigurationId()
    cfgs = np.full(A, cfg, np.int64) if cell_cfg is None else np.asarray(cell_cfg, np.int64)
    o_len, o_ann, o_ids, o_off = sim.apply_batch(src, dst, ring, status, cfgs, blocked=blocked, bitmap=bitmap,
                                                 perm_seed=perm_seed, threads=4)
    res = cluster.handleBatch(cfg, src, dst, ring, status, cell_cfg=cell_cfg, blocked=blocked, bitmap=bitmap,
                              perm_seed=perm_seed)
    np.testing.assert_array_equal(res.proposal_len, o_len)
    np.testing.assert_array_equal(res.announced, o_ann)
    e1, e2 = fingerprints_from_oracle(rb, o_len, o_ids, o_off)
    np.testing.assert_array_equal(res.proposal_hash, e1)
    np.testing.assert_array_equal(res.proposal_hash2, e2)
    R = len(o_len)
    # canonical (ring-0 sorted) proposal lists for a sample of receivers that announced now
    who = np.nonzero(o_len)[0]
    for r in who[:: max(1, len(who) // sample_proposals)][:sample_proposals] if len(who) else []:
        assert cluster.getProposal(int(r)) == o_ids[o_off[r]: o_off[r + 1]].tolist(), "receiver %d" % r
    if check_masks:
        # reportsPerHost of receivers that are still live (state of announced receivers is dead until clear())
        live = np.nonzero(o_ann == 0)[0]
        for r in live[:: max(1, len(live) // 6)][:6] if len(live) else []:
            gm = cluster.debugMasks(int(r))
            for subj, m in gm.items():
                assert sim.reportMask(int(r), int(subj)) == m, "mask of subject %d at receiver %d" % (subj, r)
            npre, _ = cluster.debugCounters(int(r))
            assert npre == sim.updatesInProgress(int(r))
    return o_len, o_ann


def random_batch(rng, n_total, K, n_subjects, n_cells, n_members, dup_frac=0.2):
    """random cells over a few subjects, with duplicates; status consistent with membership"""
    subjects = rng.choice(n_total, size=n_subjects, replace=False).astype(np.int32)
    dst = rng.choice(subjects, size=n_cells).astype(np.int32)
    ring = rng.integers(0, K, size=n_cells).astype(np.uint8)
    src = rng.integers(0, n_members, size=n_cells).astype(np.int32)
    status = np.where(dst < n_members, W.DOWN, W.UP).astype(np.uint8)
    ndup = int(dup_frac * n_cells)
    if ndup and n_cells:
        pick = rng.integers(0, n_cells, size=ndup)
        dst = np.concatenate([dst, dst[pick]])
        ring = np.concatenate([ring, ring[pick]])
        src = np.concatenate([src, rng.integers(0, n_members, size=ndup).astype(np.int32)])
        status = np.concatenate([status, status[pick]])
        order = rng.permutation(len(dst))
        dst, ring, src, status = dst[order], ring[order], src[order], status[order]
    return src, dst, ring, status
