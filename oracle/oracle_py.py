"""ORACLE — test infrastructure, NOT product code.

ctypes binding of oracle/liboracle.so (the literal C++ restatement of the reference's Java hot path,
see oracle/rapid_oracle.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import
this.  Class and method names follow the Java classes they stand for
(rapid/src/main/java/com/vrg/rapid/{MembershipView,MultiNodeCutDetector,FastPaxos}.java).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

UP, DOWN = 0, 1


class NodeNotInRingException(Exception):
    pass


class NodeAlreadyInRingException(Exception):
    pass


class UUIDAlreadySeenException(Exception):
    pass


def build():
    """Compile oracle/liboracle.so with g++ (committed recipe: oracle/Makefile)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    vp, i32, i64, u64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32
    p = C.c_void_p  # raw pointers to numpy buffers
    sig = {
        "orc_xxh64_bytes": (u64, [p, i64, u64]),
        "orc_xxh64_int": (u64, [i32, u64]),
        "orc_xxh64_long": (u64, [i64, u64]),
        "orc_splitmix64": (u64, [u64]),
        "orc_universe_create": (vp, []),
        "orc_universe_destroy": (None, [vp]),
        "orc_universe_add": (i32, [vp, p, i32, i32]),
        "orc_universe_add_bulk": (i32, [vp, i64, p, p, p, p]),
        "orc_universe_size": (i32, [vp]),
        "orc_view_create": (vp, [vp, i32]),
        "orc_view_create_bulk": (vp, [vp, i32, p, i64, p, p, i64]),
        "orc_view_destroy": (None, [vp]),
        "orc_view_ring_add": (i32, [vp, i32, i64, i64]),
        "orc_view_ring_delete": (i32, [vp, i32]),
        "orc_view_observers": (i32, [vp, i32, p, i32]),
        "orc_view_subjects": (i32, [vp, i32, p, i32]),
        "orc_view_expected_observers": (i32, [vp, i32, p, i32]),
        "orc_view_ring": (i32, [vp, i32, p, i32]),
        "orc_view_ring_numbers": (i32, [vp, i32, i32, p, i32]),
        "orc_view_size": (i32, [vp]),
        "orc_view_is_present": (i32, [vp, i32]),
        "orc_view_is_safe_to_join": (i32, [vp, i32, i64, i64]),
        "orc_view_config_id": (i64, [vp]),
        "orc_view_key": (i64, [vp, i32, i32]),
        "orc_view_tables": (i32, [vp, p, i64, p, p]),
        "orc_cd_create": (vp, [vp, i32, i32, i32]),
        "orc_cd_destroy": (None, [vp]),
        "orc_cd_aggregate": (i32, [vp, i32, i32, i32, p, i32, p, i32]),
        "orc_cd_invalidate": (i32, [vp, vp, p, i32]),
        "orc_cd_num_proposals": (i32, [vp]),
        "orc_cd_clear": (None, [vp]),
        "orc_cd_report_mask": (u32, [vp, i32]),
        "orc_handler_create": (vp, [vp, i32, i32, i32]),
        "orc_handler_destroy": (None, [vp]),
        "orc_handler_batch": (i32, [vp, i64, p, p, p, p, p, p, p, i32]),
        "orc_handler_announced": (i32, [vp]),
        "orc_handler_reset": (None, [vp]),
        "orc_handler_num_proposals": (i32, [vp]),
        "orc_handler_report_mask": (u32, [vp, i32]),
        "orc_fp_create": (vp, [vp, i64, i32]),
        "orc_fp_destroy": (None, [vp]),
        "orc_fp_vote": (i32, [vp, i32, i64, p, i32]),
        "orc_fp_decided": (i32, [vp]),
        "orc_fp_votes_received": (i32, [vp]),
        "orc_fp_decision": (i32, [vp, p, i32]),
        "orc_fp_votes_for": (i32, [vp, p, i32]),
        "orc_sim_create": (vp, [vp, i32, i32, i32, i64]),
        "orc_sim_destroy": (None, [vp]),
        "orc_sim_reset": (None, [vp]),
        "orc_sim_apply_batch": (i64, [vp, i64, p, p, p, p, p, p, p, i32, u64, i64, i32, p, p, p, i64, p]),
        "orc_sim_report_mask": (u32, [vp, i64, i32]),
        "orc_sim_num_proposals": (i32, [vp, i64]),
        "orc_sim_updates_in_progress": (i32, [vp, i64]),
        "orc_sim_tally": (i32, [vp, i64, i32, i32, i64, p, p, p, p, p, i32, p, p, p]),
        "orc_px_create": (vp, [vp, i32, i32, i64, i32]),
        "orc_px_destroy": (None, [vp]),
        "orc_px_start_phase1a": (i32, [vp, i32, p]),
        "orc_px_phase1a": (i32, [vp, i32, i64, i32, i32, p, p, i32, p]),
        "orc_px_phase1b": (i32, [vp, i32, i64, p, p, i32, p, p, i32, p]),
        "orc_px_phase2a": (i32, [vp, i32, i64, i32, i32, p, i32]),
        "orc_px_phase2b": (i32, [vp, i32, i64, i32, i32, p, i32]),
        "orc_px_register_fast_round_vote": (None, [vp, p, i32]),
        "orc_px_decided": (i32, [vp]),
        "orc_px_decision": (i32, [vp, p, i32]),
        "orc_px_vval": (i32, [vp, p, i32]),
        "orc_px_cval": (i32, [vp, p, i32]),
        "orc_px_ranks": (None, [vp, p]),
        "orc_px_coordinator_rule": (i32, [vp, i32, p, p, p, p, i32]),
        "orc_fdsim_create": (vp, [vp, i32, p, i64]),
        "orc_fdsim_destroy": (None, [vp]),
        "orc_fdsim_tick": (i64, [vp, p, p, i64, p, p, p, p, i64, i64]),
        "orc_fdsim_state": (None, [vp, i64, i32, p]),
        "orc_fdsim_num_detectors": (i32, [vp, i64]),
        "orc_hardware_threads": (i32, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def xxh64(data: bytes, seed: int = 0) -> int:
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    return lib().orc_xxh64_bytes(_ptr(buf), len(data), seed & 0xFFFFFFFFFFFFFFFF)


def xx_hash_int(v: int, seed: int = 0) -> int:
    return lib().orc_xxh64_int(v, seed & 0xFFFFFFFFFFFFFFFF)


def xx_hash_long(v: int, seed: int = 0) -> int:
    return lib().orc_xxh64_long(v, seed & 0xFFFFFFFFFFFFFFFF)


def splitmix64(x: int) -> int:
    return lib().orc_splitmix64(x & 0xFFFFFFFFFFFFFFFF)


class Universe:
    """Interns Endpoint{hostname bytes, port} -> int32 tag."""

    def __init__(self):
        self.h = lib().orc_universe_create()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_universe_destroy(self.h)
            self.h = None

    def add(self, hostname, port: int) -> int:
        hb = hostname.encode("utf-8") if isinstance(hostname, str) else bytes(hostname)
        buf = np.frombuffer(hb, dtype=np.uint8) if len(hb) else np.zeros(1, np.uint8)
        return lib().orc_universe_add(self.h, _ptr(buf), len(hb), port)

    def add_bulk(self, host_bytes: np.ndarray, host_off: np.ndarray, port: np.ndarray) -> np.ndarray:
        n = len(port)
        out = np.empty(n, np.int32)
        hb = np.ascontiguousarray(host_bytes, np.uint8)
        ho = _i32(host_off)
        po = _i32(port)
        lib().orc_universe_add_bulk(self.h, n, _ptr(hb), _ptr(ho), _ptr(po), _ptr(out))
        return out

    def __len__(self):
        return lib().orc_universe_size(self.h)


class MembershipView:
    """MembershipView.java restated (tags instead of Endpoint objects)."""

    def __init__(self, universe: Universe, K: int, tags=None, id_high=None, id_low=None):
        self.u = universe
        self.K = K
        if tags is None:
            self.h = lib().orc_view_create(universe.h, K)
        else:
            t = _i32(tags)
            hi = np.ascontiguousarray(id_high if id_high is not None else [], np.int64)
            lo = np.ascontiguousarray(id_low if id_low is not None else [], np.int64)
            self.h = lib().orc_view_create_bulk(universe.h, K, _ptr(t), len(t), _ptr(hi), _ptr(lo), len(hi))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_view_destroy(self.h)
            self.h = None

    def ringAdd(self, tag, node_id):
        rc = lib().orc_view_ring_add(self.h, tag, node_id[0], node_id[1])
        if rc == -4:
            raise UUIDAlreadySeenException(tag)
        if rc == -3:
            raise NodeAlreadyInRingException(tag)

    def ringDelete(self, tag):
        if lib().orc_view_ring_delete(self.h, tag) == -2:
            raise NodeNotInRingException(tag)

    def _list(self, fn, tag):
        out = np.empty(self.K, np.int32)
        n = fn(self.h, tag, _ptr(out), self.K)
        if n == -2:
            raise NodeNotInRingException(tag)
        return out[:n].tolist()

    def getObserversOf(self, tag):
        return self._list(lib().orc_view_observers, tag)

    def getSubjectsOf(self, tag):
        return self._list(lib().orc_view_subjects, tag)

    def getExpectedObserversOf(self, tag):
        return self._list(lib().orc_view_expected_observers, tag)

    def getRing(self, k):
        n = self.getMembershipSize()
        out = np.empty(max(n, 1), np.int32)
        m = lib().orc_view_ring(self.h, k, _ptr(out), len(out))
        return out[:m].tolist()

    def getRingNumbers(self, observer, subject):
        out = np.empty(self.K, np.int32)
        n = lib().orc_view_ring_numbers(self.h, observer, subject, _ptr(out), self.K)
        if n == -2:
            raise NodeNotInRingException(observer)
        return out[:n].tolist()

    def getMembershipSize(self):
        return lib().orc_view_size(self.h)

    def isHostPresent(self, tag):
        return bool(lib().orc_view_is_present(self.h, tag))

    def isSafeToJoin(self, tag, node_id):
        return lib().orc_view_is_safe_to_join(self.h, tag, node_id[0], node_id[1])

    def getCurrentConfigurationId(self):
        return lib().orc_view_config_id(self.h)

    def key(self, k, tag):
        return lib().orc_view_key(self.h, k, tag)

    def tables(self, tags):
        t = _i32(tags)
        obs = np.empty((len(t), self.K), np.int32)
        subj = np.empty((len(t), self.K), np.int32)
        rc = lib().orc_view_tables(self.h, _ptr(t), len(t), _ptr(obs), _ptr(subj))
        if rc == -2:
            raise NodeNotInRingException()
        return obs, subj


class MultiNodeCutDetector:
    """MultiNodeCutDetector.java restated."""

    def __init__(self, universe: Universe, K: int, H: int, L: int):
        self.u = universe
        self.h = lib().orc_cd_create(universe.h, K, H, L)
        if not self.h:
            raise ValueError("Arguments do not satisfy K > H >= L >= 0")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cd_destroy(self.h)
            self.h = None

    def aggregateForProposal(self, src, dst, status, rings):
        r = _i32(rings if hasattr(rings, "__len__") else [rings])
        out = np.empty(len(self.u) + 1, np.int32)
        n = lib().orc_cd_aggregate(self.h, src, dst, status, _ptr(r), len(r), _ptr(out), len(out))
        return out[:n].tolist()

    def invalidateFailingEdges(self, view: MembershipView):
        out = np.empty(len(self.u) + 1, np.int32)
        n = lib().orc_cd_invalidate(self.h, view.h, _ptr(out), len(out))
        return out[:n].tolist()

    def getNumProposals(self):
        return lib().orc_cd_num_proposals(self.h)

    def clear(self):
        lib().orc_cd_clear(self.h)

    def reportMask(self, tag):
        return lib().orc_cd_report_mask(self.h, tag)


class AlertBatchHandler:
    """MembershipService.handleMessage(BatchedAlertMessage) restated (MembershipService.java:300-354)."""

    def __init__(self, view: MembershipView, K, H, L):
        self.view = view
        self.h = lib().orc_handler_create(view.h, K, H, L)
        if not self.h:
            raise ValueError("bad K/H/L")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_handler_destroy(self.h)
            self.h = None

    def handleBatch(self, msgs):
        """msgs: list of (src, dst, status, cfg, [rings])"""
        n = len(msgs)
        src = _i32([m[0] for m in msgs])
        dst = _i32([m[1] for m in msgs])
        st = _i32([m[2] for m in msgs])
        cfg = np.ascontiguousarray([m[3] for m in msgs], np.int64)
        off = np.zeros(n + 1, np.int32)
        rings = []
        for i, m in enumerate(msgs):
            rings.extend(m[4])
            off[i + 1] = len(rings)
        rg = _i32(rings if rings else [0])
        out = np.empty(len(self.view.u) + 1, np.int32)
        c = lib().orc_handler_batch(self.h, n, _ptr(src), _ptr(dst), _ptr(st), _ptr(cfg), _ptr(off), _ptr(rg),
                                    _ptr(out), len(out))
        return out[:c].tolist()

    def announced(self):
        return bool(lib().orc_handler_announced(self.h))

    def reset(self):
        lib().orc_handler_reset(self.h)

    def reportMask(self, tag):
        return lib().orc_handler_report_mask(self.h, tag)


class FastPaxosTally:
    """FastPaxos.handleFastRoundProposal restated (FastPaxos.java:125-156)."""

    def __init__(self, universe: Universe, configuration_id: int, membership_size: int):
        self.u = universe
        self.h = lib().orc_fp_create(universe.h, configuration_id, membership_size)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fp_destroy(self.h)
            self.h = None

    def handleFastRoundProposal(self, sender, cfg, endpoints):
        e = _i32(endpoints if len(endpoints) else [0])
        return bool(lib().orc_fp_vote(self.h, sender, cfg, _ptr(e), len(endpoints)))

    def decided(self):
        return bool(lib().orc_fp_decided(self.h))

    def votesReceived(self):
        return lib().orc_fp_votes_received(self.h)

    def decision(self):
        out = np.empty(len(self.u) + 1, np.int32)
        n = lib().orc_fp_decision(self.h, _ptr(out), len(out))
        return out[:n].tolist()


class ClassicPaxos:
    """Paxos.java restated (oracle/paxos_oracle.hpp).  Values are lists of endpoint tags; ranks are (round, nodeIndex).
    Handlers return the outgoing message (or None) instead of handing it to a broadcaster."""

    CAP = 4096

    def __init__(self, universe: Universe, my_tag: int, my_hash: int, configuration_id: int, N: int):
        self.u, self.me, self.cfg, self.N = universe, my_tag, configuration_id, N
        self.h = lib().orc_px_create(universe.h, my_tag, my_hash, configuration_id, N)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_px_destroy(self.h)
            self.h = None

    @staticmethod
    def _tags(v):
        return _i32(v if len(v) else [0])

    def startPhase1a(self, round_):
        """-> Phase1aMessage dict broadcast to all, or None (Paxos.java:98-113)"""
        rk = np.zeros(2, np.int32)
        if not lib().orc_px_start_phase1a(self.h, round_, _ptr(rk)):
            return None
        return {"sender": self.me, "cfg": self.cfg, "rank": (int(rk[0]), int(rk[1]))}

    def handlePhase1aMessage(self, m):
        """-> Phase1bMessage dict sent to m['sender'], or None (:120-151)"""
        rk, out, n = np.zeros(4, np.int32), np.empty(self.CAP, np.int32), C.c_int32(0)
        if not lib().orc_px_phase1a(self.h, m["sender"], m["cfg"], m["rank"][0], m["rank"][1], _ptr(rk), _ptr(out),
                                    self.CAP, C.addressof(n)):
            return None
        return {"sender": self.me, "cfg": self.cfg, "rnd": (int(rk[0]), int(rk[1])), "vrnd": (int(rk[2]), int(rk[3])),
                "vval": out[:n.value].tolist()}

    def handlePhase1bMessage(self, m):
        """-> Phase2aMessage dict broadcast to all, or None (:159-191)"""
        rk = _i32([m["rnd"][0], m["rnd"][1], m["vrnd"][0], m["vrnd"][1]])
        t = self._tags(m["vval"])
        ork, out, n = np.zeros(2, np.int32), np.empty(self.CAP, np.int32), C.c_int32(0)
        if not lib().orc_px_phase1b(self.h, m["sender"], m["cfg"], _ptr(rk), _ptr(t), len(m["vval"]), _ptr(ork), _ptr(out),
                                    self.CAP, C.addressof(n)):
            return None
        return {"sender": self.me, "cfg": self.cfg, "rnd": (int(ork[0]), int(ork[1])), "vval": out[:n.value].tolist()}

    def handlePhase2aMessage(self, m):
        """-> Phase2bMessage dict broadcast to all, or None (:198-216)"""
        t = self._tags(m["vval"])
        if not lib().orc_px_phase2a(self.h, m["sender"], m["cfg"], m["rnd"][0], m["rnd"][1], _ptr(t), len(m["vval"])):
            return None
        return {"sender": self.me, "cfg": self.cfg, "rnd": m["rnd"], "endpoints": list(m["vval"])}

    def handlePhase2bMessage(self, m):
        """-> True iff this message made the node decide (:223-236)"""
        t = self._tags(m["endpoints"])
        return bool(lib().orc_px_phase2b(self.h, m["sender"], m["cfg"], m["rnd"][0], m["rnd"][1], _ptr(t), len(m["endpoints"])))

    def registerFastRoundVote(self, vote):
        t = self._tags(vote)
        lib().orc_px_register_fast_round_vote(self.h, _ptr(t), len(vote))

    def selectProposalUsingCoordinatorRule(self, msgs):
        """msgs: list of dicts with 'vrnd' and 'vval' (:271-328).  Raises ValueError on an empty list."""
        vr = _i32([x for m in msgs for x in m["vrnd"]] or [0])
        off = np.zeros(len(msgs) + 1, np.int32)
        off[1:] = np.cumsum([len(m["vval"]) for m in msgs]) if msgs else []
        tags = _i32([t for m in msgs for t in m["vval"]] or [0])
        out = np.empty(self.CAP, np.int32)
        n = lib().orc_px_coordinator_rule(self.h, len(msgs), _ptr(vr), _ptr(off), _ptr(tags), _ptr(out), self.CAP)
        if n < 0:
            raise ValueError("phase1bMessages was empty")
        return out[:n].tolist()

    def _list(self, fn):
        out = np.empty(self.CAP, np.int32)
        return out[:fn(self.h, _ptr(out), self.CAP)].tolist()

    def decided(self):
        return bool(lib().orc_px_decided(self.h))

    def decision(self):
        return self._list(lib().orc_px_decision)

    def vval(self):
        return self._list(lib().orc_px_vval)

    def cval(self):
        return self._list(lib().orc_px_cval)

    def ranks(self):
        """-> {'rnd': (r, i), 'vrnd': (r, i), 'crnd': (r, i)}"""
        o = np.zeros(6, np.int32)
        lib().orc_px_ranks(self.h, _ptr(o))
        o = o.tolist()
        return {"rnd": (o[0], o[1]), "vrnd": (o[2], o[3]), "crnd": (o[4], o[5])}


FD_CRASHED, FD_INGRESS_BLOCKED, FD_EGRESS_BLOCKED, FD_BOOTSTRAPPING = 1, 2, 4, 8


class FdSim:
    """Alert generation restated (oracle/fd_oracle.hpp): one FdNode — K PingPongFailureDetectors, in getSubjectsOf order —
    per member; tick() = one failure-detector interval of every live node, in member order."""

    def __init__(self, view: MembershipView, K, member_tags):
        self.view, self.K = view, K
        self.members = _i32(member_tags)
        self.h = lib().orc_fdsim_create(view.h, K, _ptr(self.members), len(self.members))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fdsim_destroy(self.h)
            self.h = None

    def tick(self, node_flags, cfg, edge_fail=None):
        """-> list of (observer tag, subject tag, [ring numbers]) in the order the notifiers fired"""
        nf = np.ascontiguousarray(node_flags, np.uint8)
        ef = None if edge_fail is None else np.ascontiguousarray(edge_fail, np.uint8)
        cap = len(self.members) * self.K + 1
        obs, sub = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        off, rings = np.zeros(cap + 1, np.int32), np.zeros(cap * self.K, np.int32)
        n = lib().orc_fdsim_tick(self.h, _ptr(nf), _ptr(ef), cfg, _ptr(obs), _ptr(sub), _ptr(off), _ptr(rings), cap, cap * self.K)
        return [(int(obs[i]), int(sub[i]), rings[off[i]: off[i + 1]].tolist()) for i in range(n)]

    def state(self, i, k):
        out = np.zeros(2, np.int32)
        lib().orc_fdsim_state(self.h, i, k, _ptr(out))
        return int(out[0]), bool(out[1])

    def numDetectors(self, i):
        return lib().orc_fdsim_num_detectors(self.h, i)


class ClusterSim:
    """R literal AlertBatchHandlers over one shared view (parity at R > 1; the timed CPU baseline)."""

    def __init__(self, view: MembershipView, K, H, L, R, receiver_base=0):
        self.view = view
        self.R = R
        self.receiver_base = receiver_base
        self.h = lib().orc_sim_create(view.h, K, H, L, R)
        if not self.h:
            raise ValueError("bad K/H/L")
        self.last_seconds = 0.0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sim_destroy(self.h)
            self.h = None

    def reset(self):
        lib().orc_sim_reset(self.h)

    def apply_batch(self, src, dst, ring, status, cfg, blocked=None, bitmap=None, perm_seed=None, threads=1):
        A = len(dst)
        src = _i32(src)
        dst = _i32(dst)
        ring = np.ascontiguousarray(ring, np.uint8)
        status = np.ascontiguousarray(status, np.uint8)
        cfg = np.ascontiguousarray(cfg, np.int64)
        bl = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
        bm = None if bitmap is None else np.ascontiguousarray(bitmap, np.uint32)
        out_len = np.zeros(self.R, np.int32)
        out_ann = np.zeros(self.R, np.uint8)
        # a proposal can hold subjects reported in earlier batches too
        self._subjects = getattr(self, "_subjects", set()) | set(np.unique(dst).tolist())
        cap = max(1, self.R * max(1, len(self._subjects)))
        out_ids = np.empty(cap, np.int32)
        secs = C.c_double(0.0)
        w = lib().orc_sim_apply_batch(self.h, A, _ptr(src), _ptr(dst), _ptr(ring), _ptr(status), _ptr(cfg),
                                      _ptr(bl), _ptr(bm), 0 if perm_seed is None else 1,
                                      0 if perm_seed is None else (perm_seed & 0xFFFFFFFFFFFFFFFF),
                                      self.receiver_base, threads, _ptr(out_len), _ptr(out_ann), _ptr(out_ids), cap,
                                      C.byref(secs))
        assert w >= 0
        self.last_seconds = secs.value
        off = np.zeros(self.R + 1, np.int64)
        np.cumsum(out_len, out=off[1:])
        return out_len, out_ann, out_ids[:w], off

    def reportMask(self, r, tag):
        return lib().orc_sim_report_mask(self.h, r, tag)

    def numProposals(self, r):
        return lib().orc_sim_num_proposals(self.h, r)

    def updatesInProgress(self, r):
        return lib().orc_sim_updates_in_progress(self.h, r)


def sim_tally(universe, cfg, membership_size, n_nodes, sender, vote_cfg, pid, prop_off, prop_ids, threads=1):
    sender = _i32(sender)
    vote_cfg = np.ascontiguousarray(vote_cfg, np.int64)
    pid = _i32(pid)
    prop_off = _i32(prop_off)
    prop_ids = _i32(prop_ids if len(prop_ids) else [0])
    dec = np.empty(n_nodes, np.int32)
    rec = np.empty(n_nodes, np.int32)
    secs = C.c_double(0.0)
    nd = lib().orc_sim_tally(universe.h, cfg, membership_size, n_nodes, len(sender), _ptr(sender), _ptr(vote_cfg),
                             _ptr(pid), _ptr(prop_off), _ptr(prop_ids), threads, _ptr(dec), _ptr(rec), C.byref(secs))
    return nd, dec, rec, secs.value


def hardware_threads():
    return lib().orc_hardware_threads()
