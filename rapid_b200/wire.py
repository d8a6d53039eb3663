"""Wire-format ingest: serialized rapid.proto messages -> alert cells / votes, decoded on the GPU by
librapid_b200.so (csrc/wire.cu).  This is the step the reference's gRPC server does before
MembershipService.handleMessage(RapidRequest) (MembershipService.java:174): no Python protobuf runtime is involved."""
import ctypes as C

import numpy as np

from . import _native as N


class DecodedAlerts:
    __slots__ = ("n_messages", "n_cells", "n_dropped", "n_new_joiners", "sender")

    def __init__(self, n_messages, n_cells, n_dropped, n_new_joiners, sender):
        self.n_messages, self.n_cells, self.n_dropped, self.n_new_joiners, self.sender = n_messages, n_cells, n_dropped, n_new_joiners, sender

    def __repr__(self):
        return "DecodedAlerts(messages=%d, cells=%d, dropped=%d, new_joiners=%d, sender=%d)" % (
            self.n_messages, self.n_cells, self.n_dropped, self.n_new_joiners, self.sender)


class WireDecoder:
    """Owns the Endpoint{hostname, port} -> id table of a MembershipView on the device."""

    def __init__(self, view):
        self.view = view
        self._h = C.c_void_p()
        N.check(N.lib().rapid_wire_create(C.byref(self._h), view._h))
        self._last = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().rapid_wire_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setConfiguration(self, cfg_id):
        """only UP alerts of this configuration may register joiners from now on (MembershipService.java:653)"""
        N.check(N.lib().rapid_wire_set_configuration(self._h, int(cfg_id)))

    def decodeBatchedAlertMessage(self, data, is_request=False):
        """bytes of a BatchedAlertMessage (or of the RapidRequest carrying it) -> DecodedAlerts; the cells stay on the device"""
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        m, c, d, j, s = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int32(-1)
        N.check(N.lib().rapid_wire_decode_alerts(self._h, N.ptr(buf), len(data), N.WIRE_REQUEST if is_request else 0, C.byref(m),
                                                 C.byref(c), C.byref(d), C.byref(j), C.byref(s)))
        self._last = DecodedAlerts(m.value, c.value, d.value, j.value, s.value)
        return self._last

    def cells(self):
        """host copies of the last decode: (src, dst, ring, status, cfg)"""
        n = self._last.n_cells if self._last else 0
        src, dst = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ring, status, cfg = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int64)
        N.check(N.lib().rapid_wire_read_cells(self._h, N.ptr(src), N.ptr(dst), N.ptr(ring), N.ptr(status), N.ptr(cfg)))
        return src, dst, ring, status, cfg

    def cellsDevice(self):
        """device pointers (src, dst, ring, status, cfg) of the last decode, for rapid_cd_apply_batch_dev"""
        p = [C.c_void_p() for _ in range(5)]
        N.check(N.lib().rapid_wire_cells_dev(self._h, *[C.byref(x) for x in p]))
        return tuple(x.value for x in p)

    def messages(self):
        """per AlertMessage of the last decode: dict of arrays dst, status, n_rings, node_high, node_low, has_node_id,
        meta_off, meta_len"""
        n = self._last.n_messages if self._last else 0
        out = {"dst": np.zeros(n, np.int32), "status": np.zeros(n, np.uint8), "n_rings": np.zeros(n, np.int32),
               "node_high": np.zeros(n, np.int64), "node_low": np.zeros(n, np.int64), "has_node_id": np.zeros(n, np.uint8),
               "meta_off": np.zeros(n, np.int64), "meta_len": np.zeros(n, np.int32)}
        N.check(N.lib().rapid_wire_read_messages(self._h, *[N.ptr(out[k]) for k in ("dst", "status", "n_rings", "node_high", "node_low",
                                                                                   "has_node_id", "meta_off", "meta_len")]))
        return out

    def decodeFastRoundPhase2bMessages(self, messages, is_request=False):
        """list of serialized FastRoundPhase2bMessages -> (sender, cfg, hash, hash2, len) arrays"""
        n = len(messages)
        off = np.zeros(n + 1, np.int64)
        if n:
            off[1:] = np.cumsum([len(m) for m in messages])
        joined = b"".join(bytes(m) for m in messages)
        buf = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(1, np.uint8)
        s, c = np.zeros(n, np.int32), np.zeros(n, np.int64)
        h1, h2, ln = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.int32)
        N.check(N.lib().rapid_wire_decode_votes(self._h, N.ptr(buf), N.ptr(off), n, N.WIRE_REQUEST if is_request else 0, N.ptr(s),
                                                N.ptr(c), N.ptr(h1), N.ptr(h2), N.ptr(ln)))
        return s, c, h1, h2, ln

    def lastDeviceMs(self):
        out = C.c_float(0)
        N.check(N.lib().rapid_wire_last_device_ms(self._h, C.byref(out)))
        return out.value
