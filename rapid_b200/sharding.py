"""Host-side logic of the multi-GPU path: how the virtual nodes are split across ranks and what the sharded tally
protocol computes.  Pure Python/numpy so that it can be exercised with `gloo` on CPU; the GPU library implements the same
protocol in csrc/fast_paxos.cu (rapid_fp_tally_cd with a communicator)."""
import numpy as np

HIST_BUCKETS = 1 << 16


def shard_range(n, rank, world):
    """contiguous slice of ring-0 positions owned by `rank` ("nodes shard by ring hash")"""
    begin = rank * n // world
    return begin, (rank + 1) * n // world - begin


def histogram_of(h1, counts=None):
    """65 536-bucket histogram over the top 16 bits of the proposal fingerprints voted on this rank"""
    h1 = np.asarray(h1, np.uint64)
    w = np.ones(len(h1), np.int64) if counts is None else np.asarray(counts, np.int64)
    return np.bincount((h1 >> np.uint64(48)).astype(np.int64), weights=w, minlength=HIST_BUCKETS).astype(np.int32)


def verification_words(h1, h2, ln, bucket):
    """(max h1, max ~h1, max h2, max ~h2, max len, max ~len) over this rank's entries in `bucket`; zeros if none"""
    h1 = np.asarray(h1, np.uint64); h2 = np.asarray(h2, np.uint64); ln = np.asarray(ln, np.uint64)
    m = (h1 >> np.uint64(48)).astype(np.int64) == bucket
    if not m.any():
        return np.zeros(6, np.uint64)
    return np.array([h1[m].max(), (~h1[m]).max(), h2[m].max(), (~h2[m]).max(), ln[m].max(), (~ln[m]).max()], np.uint64)


def decide(global_hist, global_words_of, quorum):
    """the decision every rank reaches after the sum all-reduce (global_hist) and the max all-reduce of the winning
    bucket's verification words (global_words_of(bucket)).  Returns (decided, h1, h2, len, count)."""
    cand = np.nonzero(global_hist >= quorum)[0]
    if len(cand) == 0:
        return False, 0, 0, 0, 0
    b = int(cand[0])
    w = global_words_of(b)
    single = w[0] == ~w[1] and w[2] == ~w[3] and w[4] == ~w[5]
    if not single:
        return False, 0, 0, 0, 0          # two proposals share the bucket: the library refines digit by digit
    return True, int(w[0]), int(w[2]), int(w[4]), int(global_hist[b])
