"""GPU parity of the wire-format ingest (csrc/wire.cu, through the C ABI) against the protobuf runtime: messages of
rapid.proto serialized by google.protobuf (and hand-rolled bytes for the encodings it never emits) are decoded on the
device and compared field by field with what the runtime parses; then fed to the detector and compared with the same
batch given as arrays."""
import random

import numpy as np
import pytest

import wire_proto
from wire_proto import field, varint
from rapid_b200 import workloads as W

pytestmark = pytest.mark.gpu

K = 10


@pytest.fixture(scope="module")
def rb():
    import rapid_b200
    return rapid_b200


@pytest.fixture(scope="module")
def pb():
    return wire_proto.build()


def make_view(rb, n):
    hb, off, ports = W.packed_endpoints(0, n)
    return rb.MembershipView.from_packed(K, hb, off, ports)


def ep(pb, i):
    hosts, ports = W.endpoints(i, 1)
    return pb.Endpoint(hostname=hosts[0], port=int(ports[0]))


def expect_cells(batch, id_of):
    """what MembershipService would see: one cell per ring number, message order then ring order; DOWN alerts about
    unknown endpoints dropped"""
    cells = []
    for m in batch.messages:
        d = id_of(m.edgeDst)
        if d < 0:
            continue
        for r in m.ringNumber:
            cells.append((id_of(m.edgeSrc), d, r, int(m.edgeStatus), m.configurationId))
    return cells


@pytest.mark.parametrize("seed", range(6))
def test_alert_batches_decode_like_the_protobuf_runtime(rb, pb, seed):
    rng = random.Random(seed)
    n = rng.choice([50, 300, 2000])
    view = make_view(rb, n)
    dec = rb.WireDecoder(view)
    known = {}                                            # (hostname, port) -> id

    def id_of(e):
        return known.get((e.hostname, e.port), -1)

    for i in range(n):
        e = ep(pb, i)
        known[(e.hostname, e.port)] = i
    next_joiner = n
    for _ in range(4):
        batch = pb.BatchedAlertMessage()
        batch.sender.CopyFrom(ep(pb, rng.randrange(n)))
        for _ in range(rng.randint(0, 60)):
            m = batch.messages.add()
            m.edgeSrc.CopyFrom(ep(pb, rng.randrange(n)))
            kind = rng.random()
            if kind < 0.6:                                # DOWN about a member
                m.edgeDst.CopyFrom(ep(pb, rng.randrange(n))); m.edgeStatus = 1
            elif kind < 0.85:                             # UP about a (possibly new) joiner
                m.edgeDst.CopyFrom(ep(pb, n + rng.randrange(40))); m.edgeStatus = 0
                m.nodeId.high, m.nodeId.low = rng.getrandbits(63) - 2**62, rng.getrandbits(63) - 2**62
                if rng.random() < 0.5:
                    m.metadata.metadata["role"] = b"x" * rng.randint(0, 5)
            elif kind < 0.95:                             # DOWN about somebody nobody knows: filtered
                m.edgeDst.hostname, m.edgeDst.port, m.edgeStatus = b"stranger", rng.randrange(100), 1
            else:                                         # UP about a member (the filter will drop it later; ids resolve)
                m.edgeDst.CopyFrom(ep(pb, rng.randrange(n))); m.edgeStatus = 0
            m.configurationId = rng.choice([-7, 0, 12345678901234, -2**63, 2**63 - 1])
            m.ringNumber.extend(rng.sample(range(K), rng.randint(0, 4)))
        # joiners are registered in order of first appearance among the UP alerts
        new = 0
        for m in batch.messages:
            k = (m.edgeDst.hostname, m.edgeDst.port)
            if m.edgeStatus == 0 and k not in known:
                known[k] = next_joiner
                next_joiner += 1
                new += 1
        as_request = rng.random() < 0.5
        data = pb.RapidRequest(batchedAlertMessage=batch).SerializeToString() if as_request else batch.SerializeToString()
        got = dec.decodeBatchedAlertMessage(data, is_request=as_request)
        want = expect_cells(batch, id_of)
        assert got.n_messages == len(batch.messages) and got.n_cells == len(want) and got.n_new_joiners == new
        assert got.n_dropped == sum(1 for m in batch.messages if id_of(m.edgeDst) < 0)
        assert got.sender == id_of(batch.sender)
        assert view.numJoiners() == next_joiner - n
        src, dst, ring, status, cfg = dec.cells()
        assert list(zip(src.tolist(), dst.tolist(), ring.tolist(), status.tolist(), cfg.tolist())) == want
        msgs = dec.messages()
        for i, m in enumerate(batch.messages):
            assert msgs["dst"][i] == id_of(m.edgeDst) and msgs["status"][i] == m.edgeStatus and msgs["n_rings"][i] == len(m.ringNumber)
            assert bool(msgs["has_node_id"][i]) == m.HasField("nodeId")
            assert (msgs["node_high"][i], msgs["node_low"][i]) == (m.nodeId.high, m.nodeId.low)
            if m.HasField("metadata"):
                raw = data[msgs["meta_off"][i]: msgs["meta_off"][i] + msgs["meta_len"][i]]
                assert pb.Metadata.FromString(raw) == m.metadata
            else:
                assert msgs["meta_len"][i] == 0


def test_encodings_the_runtime_never_emits(rb, pb):
    """unpacked ring numbers, unknown fields of every wire type, fields out of order, a repeated (merged) edgeDst, an absent
    edgeSrc, 10-byte negative varints — all legal protobuf a conforming parser accepts"""
    view = make_view(rb, 20)
    dec = rb.WireDecoder(view)
    e3, e5 = ep(pb, 3).SerializeToString(), ep(pb, 5)
    alert = (field(5, 0, varint(7)) +                                  # ringNumber, unpacked
             field(99, 0, varint(1 << 40)) + field(98, 1, b"12345678") + field(97, 5, b"1234") + field(96, 2, b"junk") +
             field(4, 0, varint(-9)) +                                  # configurationId = -9 as a 10-byte varint
             field(2, 2, field(2, 0, varint(e5.port))) +                # edgeDst: port first ...
             field(3, 0, varint(1)) +
             field(5, 2, varint(2) + varint(4)) +                       # ... packed ring numbers after an unpacked one
             field(2, 2, field(1, 2, e5.hostname)) +                    # ... hostname in a SECOND edgeDst occurrence (merge)
             field(5, 0, varint(0)))
    ref = pb.AlertMessage.FromString(alert)                             # the runtime agrees on what this means
    assert list(ref.ringNumber) == [7, 2, 4, 0] and ref.edgeDst == e5 and ref.configurationId == -9 and not ref.HasField("edgeSrc")
    data = field(3, 2, alert) + field(1, 2, e3) + field(50, 0, varint(3)) + field(3, 2, b"")   # + an empty AlertMessage
    got = dec.decodeBatchedAlertMessage(data)
    # the empty AlertMessage is UP about the default endpoint {"", 0} with no ring numbers: a joiner nobody asked for, 0 cells
    assert (got.n_messages, got.n_cells, got.sender, got.n_new_joiners) == (2, 4, 3, 1)
    src, dst, ring, status, cfg = dec.cells()
    assert dst.tolist() == [5] * 4 and ring.tolist() == [7, 2, 4, 0] and status.tolist() == [1] * 4 and cfg.tolist() == [-9] * 4
    assert src.tolist() == [-1] * 4


@pytest.mark.parametrize("cut", [1, 2, 5, 9, 14, 20])
def test_truncated_or_corrupt_bytes_are_refused(rb, pb, cut):
    view = make_view(rb, 20)
    dec = rb.WireDecoder(view)
    b = pb.BatchedAlertMessage()
    b.sender.CopyFrom(ep(pb, 1))
    m = b.messages.add()
    m.edgeSrc.CopyFrom(ep(pb, 2)); m.edgeDst.CopyFrom(ep(pb, 3)); m.edgeStatus = 1; m.configurationId = 5; m.ringNumber.extend([1, 2])
    data = b.SerializeToString()
    assert dec.decodeBatchedAlertMessage(data).n_cells == 2
    bad = data[: len(data) - cut]
    try:
        pb.BatchedAlertMessage.FromString(bad)
        runtime_ok = True
    except Exception:
        runtime_ok = False
    if runtime_ok:
        dec.decodeBatchedAlertMessage(bad)                 # a cut on a field boundary is a shorter, valid message
    else:
        with pytest.raises(rb.RapidError):
            dec.decodeBatchedAlertMessage(bad)
    with pytest.raises(rb.RapidError):
        dec.decodeBatchedAlertMessage(field(3, 2, field(3, 0, varint(2))))           # EdgeStatus 2 does not exist
    with pytest.raises(rb.RapidError):
        dec.decodeBatchedAlertMessage(pb.RapidRequest(probeMessage=pb.ProbeMessage()).SerializeToString(), is_request=True)


def test_decoded_cells_drive_the_detector_like_arrays_do(rb, pb):
    """bytes -> cells on the device -> rapid_cd_apply_batch_dev  ==  the same batch handed over as arrays"""
    import ctypes as C
    from rapid_b200 import _native as Nn
    n = 400
    view = make_view(rb, n)
    obs, _ = view.tables()
    cfg = 77
    failed = [5, 17, 300]
    batch = pb.BatchedAlertMessage()
    batch.sender.CopyFrom(ep(pb, 0))
    cells = []
    for s in failed:
        by_observer = {}
        for r in range(K):
            by_observer.setdefault(int(obs[s][r]), []).append(r)
        for o, rings in by_observer.items():                # one AlertMessage per (observer, subject) edge, all its rings
            m = batch.messages.add()
            m.edgeSrc.CopyFrom(ep(pb, o)); m.edgeDst.CopyFrom(ep(pb, s)); m.edgeStatus = 1; m.configurationId = cfg
            m.ringNumber.extend(rings)
            cells += [(o, s, r) for r in rings]
    a = rb.VirtualCluster(view, 9, 4, kernel="bucketed")
    src, dst, ring = (np.array(x) for x in zip(*cells))
    want = a.handleBatch(cfg, src, dst, ring, np.ones(len(cells), np.uint8))
    dec = rb.WireDecoder(view)
    got = dec.decodeBatchedAlertMessage(batch.SerializeToString())
    assert got.n_cells == len(cells)
    b = rb.VirtualCluster(view, 9, 4, kernel="bucketed")
    p_src, p_dst, p_ring, p_status, p_cfg = dec.cellsDevice()
    Nn.check(Nn.lib().rapid_cd_apply_batch_dev(b._h, cfg, got.n_cells, p_src, p_dst, p_ring, p_status, p_cfg, None))
    res = b.readOutputs()
    np.testing.assert_array_equal(res.proposal_hash, want.proposal_hash)
    np.testing.assert_array_equal(res.proposal_len, want.proposal_len)
    assert set(np.asarray(res.proposal_len).tolist()) == {3}
    del C


@pytest.mark.parametrize("as_request", [False, True])
def test_votes_decode_to_sender_cfg_and_fingerprint(rb, pb, as_request):
    n = 500
    view = make_view(rb, n)
    view.registerJoiners(*W.endpoints(n, 3))
    dec = rb.WireDecoder(view)
    rng = random.Random(9)
    msgs, want = [], []
    for i in range(200):
        v = pb.FastRoundPhase2bMessage()
        s = rng.randrange(n + 10)                           # a few senders nobody knows: FastPaxos never checks (:141)
        v.sender.CopyFrom(ep(pb, s))
        v.configurationId = rng.choice([3, -3, 2**62])
        ids = rng.sample(range(n + 3), rng.randint(0, 12))
        for j in ids:
            v.endpoints.add().CopyFrom(ep(pb, j))
        msgs.append(pb.RapidRequest(fastRoundPhase2bMessage=v).SerializeToString() if as_request else v.SerializeToString())
        h1, h2 = rb.proposal_fingerprint(ids)
        want.append((s if s < n + 3 else -1, v.configurationId, h1, h2, len(ids)))
    s, c, h1, h2, ln = dec.decodeFastRoundPhase2bMessages(msgs, is_request=as_request)
    assert list(zip(s.tolist(), c.tolist(), h1.tolist(), h2.tolist(), ln.tolist())) == want
    # the decoded votes go straight into the tally
    fp = rb.FastPaxos(3, n, sender_capacity=n + 16)
    fp.handleFastRoundProposals(np.where(s >= 0, s, n + 5), h1, h2, ln, vote_cfg=c)
    # A vote naming an endpoint the dictionary does not hold — a delayed vote of an EARLIER configuration about a node that has
    # since left, say — is not an error and does not take the rest of the burst down with it (FastPaxos.java:126-132 drops it by
    # its configurationId): it decodes, identical stranger lists get identical fingerprints, and the tally's filter does the rest.
    def stranger(cfg_id, port):
        m = pb.FastRoundPhase2bMessage()
        m.sender.CopyFrom(ep(pb, 7))
        m.configurationId = cfg_id
        m.endpoints.add().CopyFrom(ep(pb, 1))
        e = m.endpoints.add(); e.hostname = b"nobody"; e.port = port
        return pb.RapidRequest(fastRoundPhase2bMessage=m).SerializeToString() if as_request else m.SerializeToString()
    good = msgs[:5]
    s2, c2, g1, g2, l2 = dec.decodeFastRoundPhase2bMessages([stranger(99, 1), stranger(99, 1), stranger(99, 2)] + good, is_request=as_request)
    assert list(zip(s2[3:].tolist(), c2[3:].tolist(), g1[3:].tolist(), g2[3:].tolist(), l2[3:].tolist())) == want[:5]
    assert l2[:3].tolist() == [2, 2, 2] and c2[:3].tolist() == [99, 99, 99]
    assert (g1[0], g2[0]) == (g1[1], g2[1]) and (g1[0], g2[0]) != (g1[2], g2[2])
    assert (int(g1[0]), int(g2[0])) != rb.proposal_fingerprint([1])
    with pytest.raises(rb.RapidError):
        dec.decodeFastRoundPhase2bMessages([msgs[0][:-1]], is_request=as_request)


def test_one_hundred_thousand_alert_messages(rb, pb):
    """full-size batch (the C5 shape: 10^5 cells as single-ring AlertMessages): decode, spot-check, time"""
    n = 100_000
    view = make_view(rb, n)
    dec = rb.WireDecoder(view)
    rng = np.random.default_rng(1)
    subj = rng.integers(0, n, 100_000)
    obs = rng.integers(0, n, 100_000)
    rings = rng.integers(0, K, 100_000)
    hosts, ports = W.endpoints(0, n)
    parts = []
    for o, s, r in zip(obs.tolist(), subj.tolist(), rings.tolist()):
        alert = (field(1, 2, field(1, 2, hosts[o]) + field(2, 0, varint(int(ports[o])))) +
                 field(2, 2, field(1, 2, hosts[s]) + field(2, 0, varint(int(ports[s])))) +
                 field(3, 0, varint(1)) + field(4, 0, varint(42)) + field(5, 2, varint(r)))
        parts.append(field(3, 2, alert))
    data = b"".join(parts)
    got = dec.decodeBatchedAlertMessage(data)
    assert got.n_messages == got.n_cells == 100_000 and got.n_dropped == 0 and got.n_new_joiners == 0
    src, dst, ring, status, cfg = dec.cells()
    np.testing.assert_array_equal(src, obs)
    np.testing.assert_array_equal(dst, subj)
    np.testing.assert_array_equal(ring, rings)
    assert status.min() == 1 and (cfg == 42).all()
    assert pb.BatchedAlertMessage.FromString(data).messages[77].edgeDst.port == int(ports[subj[77]])
    print("decode of %d bytes: %.3f ms on the device" % (len(data), dec.lastDeviceMs()))


def test_only_alerts_of_the_current_configuration_register_joiners(rb, pb):
    """filterAlertMessages (MembershipService.java:653) drops a stale alert before extractJoinerUuidAndMetadata sees it: with the
    receiver's configuration set on the decoder, an UP alert of another configuration about an unknown endpoint registers nothing
    (and so does not disturb state that hangs off the endpoint dictionary), while one of the current configuration does."""
    n = 300
    view = make_view(rb, n)
    dec = rb.WireDecoder(view)
    dec.setConfiguration(42)

    def up(j, cfg):
        a = pb.AlertMessage()
        a.edgeSrc.CopyFrom(ep(pb, 3)); a.edgeDst.CopyFrom(ep(pb, j)); a.edgeStatus = 0; a.configurationId = cfg
        a.ringNumber.extend([0, 1])
        return a
    b = pb.BatchedAlertMessage()
    b.sender.CopyFrom(ep(pb, 3))
    b.messages.extend([up(n + 1, 41), up(n + 2, 42), up(n + 1, 41)])
    r = dec.decodeBatchedAlertMessage(b.SerializeToString())
    assert view.numJoiners() == 1                           # only the endpoint named by the configuration-42 alert
    src, dst, ring, status, cfg = dec.cells()
    assert dst.tolist() == [n, n] and cfg.tolist() == [42, 42] and r.n_dropped == 2
