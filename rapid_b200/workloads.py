"""Deterministic synthetic workloads for the cut-detection path (SURVEY.md §8d / BASELINE.md §4).

Pure numpy input generation: endpoints, node ids, crash / partition / churn / flip-flop alert batches.  The ring
topology (observer / expected-observer tables) is supplied by the caller — from the GPU view in bench.py and the
gpu tests, from the oracle in CPU tests — so this module never computes a hash itself.

A batch is a structure-of-arrays of alert *cells*: src, dst (int32 node ids), ring (uint8), status (uint8).
"""
from dataclasses import dataclass, field

import numpy as np

SEED = 0x5241504944  # "RAPID"
UP, DOWN = 0, 1
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """vectorised splitmix64 over uint64 arrays"""
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def endpoints(first, count):
    """hostname = dotted quad of 0x0A000000 + i // 1000, port = 1000 + i % 1000 (several processes per VM)."""
    i = np.arange(first, first + count, dtype=np.int64)
    ip = 0x0A000000 + i // 1000
    ports = (1000 + i % 1000).astype(np.int32)
    uniq, inv = np.unique(ip, return_inverse=True)
    names = np.array([("%d.%d.%d.%d" % ((a >> 24) & 255, (a >> 16) & 255, (a >> 8) & 255, a & 255)).encode()
                      for a in uniq.tolist()], dtype=object)
    hosts = names[inv]
    return hosts.tolist(), ports


def packed_endpoints(first, count):
    hosts, ports = endpoints(first, count)
    lens = np.fromiter((len(h) for h in hosts), dtype=np.int32, count=len(hosts))
    off = np.zeros(count + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    hb = np.frombuffer(b"".join(hosts), dtype=np.uint8).copy() if count else np.zeros(1, np.uint8)
    return hb, off, ports


def node_ids(first, count):
    i = np.arange(first, first + count, dtype=np.uint64)
    hi = splitmix64(np.uint64(2) * i).view(np.int64)
    lo = splitmix64(np.uint64(2) * i + np.uint64(1)).view(np.int64)
    return hi, lo


def pick_smallest(n, count, seed):
    """the `count` ids in [0, n) with the smallest splitmix64(seed ^ i) (ties by id)"""
    keys = splitmix64(np.arange(n, dtype=np.uint64) ^ np.uint64(seed))
    order = np.lexsort((np.arange(n), keys))
    return np.sort(order[:count].astype(np.int32))


def shuffle_cells(cells, seed):
    """cell order = ascending splitmix64(seed ^ cellIndex)"""
    n = len(cells["dst"])
    order = np.argsort(splitmix64(np.arange(n, dtype=np.uint64) ^ np.uint64(seed)), kind="stable")
    return {k: v[order] for k, v in cells.items()}


@dataclass
class Batch:
    src: np.ndarray
    dst: np.ndarray
    ring: np.ndarray
    status: np.ndarray
    blocked: np.ndarray = None          # [N] uint8: receivers (by NODE ID) that get nothing / cast no vote
    expected_cut: np.ndarray = None     # sorted node ids the cluster should converge to (None: no claim)
    meta: dict = field(default_factory=dict)

    def __len__(self):
        return len(self.dst)

    def n_messages(self):
        """AlertMessage count = cells grouped by (src, dst) (the reference's wire unit)"""
        if len(self.dst) == 0:
            return 0
        return len(np.unique(self.src.astype(np.int64) * (1 << 32) + self.dst.astype(np.int64)))


def _mk(src, dst, ring, status):
    return {"src": np.asarray(src, np.int32), "dst": np.asarray(dst, np.int32), "ring": np.asarray(ring, np.uint8),
            "status": np.asarray(status, np.uint8)}


def _rows(obs, ids):
    """observer rows of `ids`: obs is a full [n][K] table or a callable ids -> [len(ids)][K]"""
    ids = np.asarray(ids, np.int32)
    return np.asarray(obs(ids) if callable(obs) else obs[ids], np.int32)


def crash_cells(obs, failed, n, silent=None):
    """DOWN reports (obs_k(s), s, k) for every failed s from every observer that is not itself silent."""
    failed = np.asarray(failed, np.int32)
    rows = _rows(obs, failed)
    K = rows.shape[1]
    src = rows.reshape(-1)
    dst = np.repeat(failed, K)
    ring = np.tile(np.arange(K, dtype=np.uint8), len(failed))
    mute = np.zeros(n, bool)
    mute[failed if silent is None else silent] = True
    keep = ~mute[src]
    return _mk(src[keep], dst[keep], ring[keep], np.full(int(keep.sum()), DOWN, np.uint8))


def c1_single_crash(obs, n=50, seed=SEED):
    """C1: one crashed node, its K observer reports, one batch."""
    failed = pick_smallest(n, 1, seed)
    cells = shuffle_cells(crash_cells(obs, failed, n), seed + 1)
    blocked = np.zeros(n, np.uint8)
    blocked[failed] = 1
    return Batch(**cells, blocked=blocked, expected_cut=failed, meta={"config": "C1", "failed": failed})


def c2_simultaneous_crash(obs, n, frac=0.01, seed=SEED):
    """C2: floor(frac*n) simultaneous crashes; failed nodes send nothing, receive nothing, cast no vote."""
    failed = pick_smallest(n, int(frac * n), seed)
    cells = shuffle_cells(crash_cells(obs, failed, n), seed + 1)
    blocked = np.zeros(n, np.uint8)
    blocked[failed] = 1
    return Batch(**cells, blocked=blocked, expected_cut=failed, meta={"config": "C2", "failed": failed})


def c3_correlated_partition(obs, ring0, n, frac=0.05, seed=SEED):
    """C3: a contiguous arc of floor(frac*n) ring-0 positions is egress-blocked (one-way partition): arc members
    send no alerts and no votes; every outside observer reports them DOWN."""
    count = int(frac * n)
    start = int(splitmix64(np.uint64(seed))) % n
    arc = np.sort(np.asarray(ring0)[(start + np.arange(count)) % n].astype(np.int32))
    cells = shuffle_cells(crash_cells(obs, arc, n), seed + 1)
    blocked = np.zeros(n, np.uint8)
    blocked[arc] = 1
    return Batch(**cells, blocked=blocked, expected_cut=arc, meta={"config": "C3", "failed": arc, "arc_start": start})


def c5_churn(obs, joiner_obs, n, n_leave, n_join, seed=SEED):
    """C5: n_leave crashes (as C2) + n_join joins in one batch.  Joiner j has id n + j; its UP reports come from its K
    expected observers (ring predecessors, `joiner_obs[j]`), except those that crashed."""
    failed = pick_smallest(n, n_leave, seed)
    down = crash_cells(obs, failed, n)
    K = np.asarray(joiner_obs).shape[1]
    jid = n + np.arange(n_join, dtype=np.int32)
    src = np.asarray(joiner_obs, np.int32)[:n_join].reshape(-1)
    dst = np.repeat(jid, K)
    ring = np.tile(np.arange(K, dtype=np.uint8), n_join)
    mute = np.zeros(n, bool)
    mute[failed] = True
    keep = ~mute[src]
    up = _mk(src[keep], dst[keep], ring[keep], np.full(int(keep.sum()), UP, np.uint8))
    cells = {k: np.concatenate([down[k], up[k]]) for k in down}
    cells = shuffle_cells(cells, seed + 1)
    blocked = np.zeros(n, np.uint8)
    blocked[failed] = 1
    cut = np.sort(np.concatenate([failed, jid]))
    return Batch(**cells, blocked=blocked, expected_cut=cut, meta={"config": "C5", "failed": failed, "joiners": jid})


def c4_flip_flop_stream(obs, n, frac=0.01, T=8, seed=SEED):
    """C4: floor(frac*n) nodes flap; ring r of subject s is detected in batch splitmix64(seed ^ s ^ r) % T; every
    batch re-sends each earlier cell with probability 1/4 (the StaticFailureDetector behaviour: duplicates).
    Returns a list of T Batches; receivers apply each in their own permuted order (perm seed = seed + 2 + batch)."""
    failed = pick_smallest(n, int(frac * n), seed)
    base = crash_cells(obs, failed, n)
    phase = (splitmix64(np.uint64(seed) ^ base["dst"].astype(np.uint64) ^ (base["ring"].astype(np.uint64) << np.uint64(32)))
             % np.uint64(T)).astype(np.int64)
    blocked = np.zeros(n, np.uint8)
    blocked[failed] = 1
    out = []
    idx_all = np.arange(len(phase))
    for t in range(T):
        fresh = idx_all[phase == t]
        earlier = idx_all[phase < t]
        coin = splitmix64(np.uint64(seed + 1000 + t) ^ earlier.astype(np.uint64)) % np.uint64(4) == 0
        take = np.concatenate([fresh, earlier[coin]])
        cells = shuffle_cells({k: v[take] for k, v in base.items()}, seed + 1 + t)
        out.append(Batch(**cells, blocked=blocked, expected_cut=failed if t == T - 1 else None,
                         meta={"config": "C4", "t": t, "failed": failed, "perm_seed": seed + 2 + t}))
    return out


def blocked_by_receiver(blocked_by_node, ring0, receiver_begin, n_receivers):
    """delivery.blocked is indexed by receiver = ring-0 position; scenarios mark NODE ids."""
    pos = np.asarray(ring0)[receiver_begin: receiver_begin + n_receivers]
    return np.ascontiguousarray(np.asarray(blocked_by_node, np.uint8)[pos])
