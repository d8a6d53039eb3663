"""The test-side protobuf plumbing of the wire-format ingest row (no GPU): the descriptors of tests/wire_proto.py restate
rapid.proto, and the hand-rolled varint / field encoders used for the unusual encodings produce bytes the official
runtime reads back as intended."""
import wire_proto
from wire_proto import field, varint


def test_runtime_round_trip_and_field_numbers():
    pb = wire_proto.build()
    b = pb.BatchedAlertMessage()
    b.sender.hostname, b.sender.port = b"10.0.0.1", 5
    m = b.messages.add()
    m.edgeSrc.hostname, m.edgeSrc.port = b"a", 1
    m.edgeDst.hostname, m.edgeDst.port = b"b", -2
    m.edgeStatus, m.configurationId = 1, -5
    m.ringNumber.extend([0, 3, 9])
    m.nodeId.high = 7
    data = b.SerializeToString()
    # rapid.proto:95-99 (sender = 1, messages = 3), :101-110 (edgeSrc 1 .. metadata 7), :13-17 (hostname 1, port 2)
    alert = (field(1, 2, field(1, 2, b"a") + field(2, 0, varint(1))) + field(2, 2, field(1, 2, b"b") + field(2, 0, varint(-2))) +
             field(3, 0, varint(1)) + field(4, 0, varint(-5)) + field(5, 2, varint(0) + varint(3) + varint(9)) +
             field(6, 2, field(1, 0, varint(7))))
    assert data == field(1, 2, field(1, 2, b"10.0.0.1") + field(2, 0, varint(5))) + field(3, 2, alert)
    assert pb.BatchedAlertMessage.FromString(data) == b
    req = pb.RapidRequest(batchedAlertMessage=b).SerializeToString()
    assert req == field(3, 2, data)                                        # rapid.proto:21-35: oneof case 3


def test_varint_edges():
    assert varint(0) == b"\x00" and varint(127) == b"\x7f" and varint(128) == b"\x80\x01"
    assert len(varint(-1)) == 10 and varint(-1)[-1] == 1                    # negative int32 / int64: ten bytes
    pb = wire_proto.build()
    v = pb.FastRoundPhase2bMessage(configurationId=-(2**63))
    assert v.SerializeToString() == field(2, 0, varint(-(2**63)))
