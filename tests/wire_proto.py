"""The messages of rapid/src/main/proto/rapid.proto that the ingest path touches, built with the protobuf runtime from
descriptors (there is no protoc in the image): Endpoint :13-17, NodeId :48-52, Metadata :178-181, AlertMessage :101-110,
BatchedAlertMessage :95-99, FastRoundPhase2bMessage :105-110 and the RapidRequest oneof cases 3 and 5 (:21-35).
The official runtime is the encoder AND the reference decoder the GPU decoder is compared with."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None, oneof=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if oneof is not None:
        f.oneof_index = oneof
    return f


def build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "rapid_wire_test.proto", "remoting", "proto3"
    m = fd.message_type.add(); m.name = "Endpoint"
    _field(m, "hostname", 1, F.TYPE_BYTES); _field(m, "port", 2, F.TYPE_INT32)
    m = fd.message_type.add(); m.name = "NodeId"
    _field(m, "high", 1, F.TYPE_INT64); _field(m, "low", 2, F.TYPE_INT64)
    m = fd.message_type.add(); m.name = "Metadata"
    e = m.nested_type.add(); e.name = "MetadataEntry"; e.options.map_entry = True
    _field(e, "key", 1, F.TYPE_STRING); _field(e, "value", 2, F.TYPE_BYTES)
    _field(m, "metadata", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".remoting.Metadata.MetadataEntry")
    en = fd.enum_type.add(); en.name = "EdgeStatus"
    v = en.value.add(); v.name, v.number = "UP", 0
    v = en.value.add(); v.name, v.number = "DOWN", 1
    m = fd.message_type.add(); m.name = "AlertMessage"
    _field(m, "edgeSrc", 1, F.TYPE_MESSAGE, type_name=".remoting.Endpoint")
    _field(m, "edgeDst", 2, F.TYPE_MESSAGE, type_name=".remoting.Endpoint")
    _field(m, "edgeStatus", 3, F.TYPE_ENUM, type_name=".remoting.EdgeStatus")
    _field(m, "configurationId", 4, F.TYPE_INT64)
    _field(m, "ringNumber", 5, F.TYPE_INT32, F.LABEL_REPEATED)
    _field(m, "nodeId", 6, F.TYPE_MESSAGE, type_name=".remoting.NodeId")
    _field(m, "metadata", 7, F.TYPE_MESSAGE, type_name=".remoting.Metadata")
    m = fd.message_type.add(); m.name = "BatchedAlertMessage"
    _field(m, "sender", 1, F.TYPE_MESSAGE, type_name=".remoting.Endpoint")
    _field(m, "messages", 3, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".remoting.AlertMessage")
    m = fd.message_type.add(); m.name = "FastRoundPhase2bMessage"
    _field(m, "sender", 1, F.TYPE_MESSAGE, type_name=".remoting.Endpoint")
    _field(m, "configurationId", 2, F.TYPE_INT64)
    _field(m, "endpoints", 3, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".remoting.Endpoint")
    m = fd.message_type.add(); m.name = "ProbeMessage"
    _field(m, "sender", 1, F.TYPE_MESSAGE, type_name=".remoting.Endpoint")
    m = fd.message_type.add(); m.name = "RapidRequest"
    m.oneof_decl.add().name = "content"
    _field(m, "batchedAlertMessage", 3, F.TYPE_MESSAGE, type_name=".remoting.BatchedAlertMessage", oneof=0)
    _field(m, "probeMessage", 4, F.TYPE_MESSAGE, type_name=".remoting.ProbeMessage", oneof=0)
    _field(m, "fastRoundPhase2bMessage", 5, F.TYPE_MESSAGE, type_name=".remoting.FastRoundPhase2bMessage", oneof=0)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)

    class NS:
        pass
    ns = NS()
    for name in ("Endpoint", "NodeId", "Metadata", "AlertMessage", "BatchedAlertMessage", "FastRoundPhase2bMessage", "ProbeMessage",
                 "RapidRequest"):
        setattr(ns, name, message_factory.GetMessageClass(pool.FindMessageTypeByName("remoting." + name)))
    return ns


def varint(v):
    """protobuf base-128 varint of an unsigned (or two's-complement 64-bit) integer"""
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def field(number, wire_type, payload):
    """hand-rolled field: tag + (length +) payload, for encodings the runtime never produces"""
    tag = varint((number << 3) | wire_type)
    return tag + (varint(len(payload)) + payload if wire_type == 2 else payload)
