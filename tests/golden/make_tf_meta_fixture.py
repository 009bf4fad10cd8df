"""Assembles, from the public .proto definitions, the MetaGraphDef `model.ckpt-7.meta` that completes
the hand-made TensorFlow training directory under tests/golden/tf_bundle/ -- WITHOUT touching
hdrnet_b200/checkpoint.py (no import, no shared helper).

Why it exists: the reference stores its model parameters IN THE GRAPH (hdrnet/bin/train.py:60-63:
`tf.add_to_collection('model_params', tf.convert_to_tensor(value, name=key))`) and reads them back by
importing the .meta file and evaluating the collection (hdrnet/utils.py:19-23, hdrnet/bin/run.py:
70-80).  TensorFlow is not installed and the reference ships no checkpoint, so the module's reader
(`read_meta_model_params`) was only ever checked against a builder living in its own test.  This
script is a second, independent statement of the wire format, with what TensorFlow writes around
the nodes of interest (meta_info_def, saver_def, other nodes and collections) present to be skipped.

Messages and field numbers (tensorflow/core/protobuf/meta_graph.proto, framework/graph.proto,
node_def.proto, attr_value.proto, tensor.proto, tensor_shape.proto, types.proto):
  MetaGraphDef   meta_info_def = 1, graph_def = 2, saver_def = 3, collection_def = 4 (map<string, CollectionDef>)
  MetaInfoDef    stripped_op_list = 2 (OpList{op = 1 {name = 1}}), tensorflow_version = 5
  GraphDef       node = 1, versions = 4 (VersionDef{producer = 1})
  NodeDef        name = 1, op = 2, input = 3, attr = 5 (map<string, AttrValue>)
  AttrValue      type = 6, shape = 7, tensor = 8
  TensorProto    dtype = 1, tensor_shape = 2, tensor_content = 4, float_val = 5, int_val = 7,
                 string_val = 8, bool_val = 11       (proto3: repeated scalars are PACKED)
  TensorShapeProto dim = 2 {size = 1}
  CollectionDef  node_list = 1 (NodeList{value = 1}), bytes_list = 2 (BytesList{value = 1})
  DataType       DT_FLOAT = 1, DT_INT32 = 3, DT_STRING = 7, DT_BOOL = 10
  map entries    key = 1, value = 2
How `tf.convert_to_tensor(python value)` (tensor_util.make_tensor_proto) encodes a constant: a
scalar goes to the typed repeated field (int_val / float_val / bool_val / string_val) with an EMPTY
shape message; an array of more than one element goes to tensor_content as raw little-endian bytes.

    python tests/golden/make_tf_meta_fixture.py        # rewrites tests/golden/tf_bundle/model.ckpt-7.meta
"""
import os
import struct

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_bundle", "model.ckpt-7.meta")

WT_VARINT, WT_LEN = 0, 2


def vi(n: int) -> bytes:
    """base-128 varint, least significant group first"""
    groups = []
    while True:
        groups.append(n & 0x7F)
        n >>= 7
        if not n:
            break
    return bytes(g | 0x80 for g in groups[:-1]) + bytes(groups[-1:])


def key(field: int, wire_type: int) -> bytes:
    return vi(field * 8 + wire_type)


def msg(field: int, payload: bytes) -> bytes:
    """length-delimited field (sub-message, string, bytes, packed repeated)"""
    return key(field, WT_LEN) + vi(len(payload)) + payload


def enum(field: int, value: int) -> bytes:
    return key(field, WT_VARINT) + vi(value)


def map_entry(field: int, k: str, value_message: bytes) -> bytes:
    return msg(field, msg(1, k.encode()) + msg(2, value_message))


DT_FLOAT, DT_INT32, DT_STRING, DT_BOOL = 1, 3, 7, 10
SCALAR_SHAPE = msg(2, b"")                                   # TensorProto.tensor_shape = {} (rank 0)


def const_node(name: str, dtype: int, tensor_fields: bytes) -> bytes:
    """NodeDef{name, op: "Const", attr{dtype: type}, attr{value: tensor}} -- attr map in TensorFlow's
    (sorted) order."""
    tensor = enum(1, dtype) + tensor_fields
    return msg(1, msg(1, name.encode()) + msg(2, b"Const")
               + map_entry(5, "dtype", enum(6, dtype))
               + map_entry(5, "value", msg(8, tensor)))


def shape_1d(n: int) -> bytes:
    return msg(2, msg(2, enum(1, n)))                        # tensor_shape{dim{size: n}}


# ---- the graph: train.py's model_params (its argparse defaults except where noted) among other nodes ----
nodes = b""
nodes += const_node("model_name", DT_STRING, SCALAR_SHAPE + msg(8, b"HDRNetPointwiseNNGuide"))       # non-default
nodes += const_node("data_pipeline", DT_STRING, SCALAR_SHAPE + msg(8, b"ImageFilesDataPipeline"))
nodes += const_node("net_input_size", DT_INT32, SCALAR_SHAPE + msg(7, vi(256)))                       # packed int_val
nodes += const_node("output_resolution", DT_INT32, shape_1d(2) + msg(4, struct.pack("<2i", 512, 768)))  # tensor_content
nodes += const_node("batch_norm", DT_BOOL, SCALAR_SHAPE + msg(11, b"\x01"))                           # packed bool_val, True
nodes += const_node("channel_multiplier", DT_INT32, SCALAR_SHAPE + msg(7, vi(1)))
nodes += const_node("guide_complexity", DT_INT32, SCALAR_SHAPE + msg(7, vi(16)))
nodes += const_node("luma_bins", DT_INT32, SCALAR_SHAPE + msg(7, vi(8)))
nodes += const_node("spatial_bin", DT_INT32, SCALAR_SHAPE + msg(7, vi(16)))
# a float scalar parameter (not in train.py's model group; a fork that stores e.g. a learning rate does this)
nodes += const_node("learning_rate", DT_FLOAT, SCALAR_SHAPE + msg(5, struct.pack("<f", 0.0001)))
# a negative int: varints of negative int32 are the 64-bit two's complement, ten bytes
nodes += const_node("crop_offset", DT_INT32, SCALAR_SHAPE + msg(7, vi((1 << 64) - 3)))
# a False flag: the single packed element 0 IS written for a repeated field
nodes += const_node("use_hdrp", DT_BOOL, SCALAR_SHAPE + msg(11, b"\x00"))
# nodes that are NOT parameters: a placeholder with a shape attr, a constant outside the collection,
# an op with inputs
nodes += msg(1, msg(1, b"inference/Placeholder") + msg(2, b"Placeholder")
             + map_entry(5, "dtype", enum(6, DT_FLOAT))
             + map_entry(5, "shape", msg(7, msg(2, enum(1, 1)) + msg(2, enum(1, (1 << 64) - 1))
                                         + msg(2, enum(1, (1 << 64) - 1)) + msg(2, enum(1, 3)))))
nodes += const_node("inference/guide/mul/y", DT_FLOAT, SCALAR_SHAPE + msg(5, struct.pack("<f", 255.0)))
nodes += msg(1, msg(1, b"inference/guide/mul") + msg(2, b"Mul") + msg(3, b"inference/Placeholder")
             + msg(3, b"inference/guide/mul/y") + map_entry(5, "T", enum(6, DT_FLOAT)))
graph_def = nodes + msg(4, enum(1, 24))                      # versions{producer: 24}

meta_info = msg(2, msg(1, msg(1, b"Const")) + msg(1, msg(1, b"Mul")) + msg(1, msg(1, b"Placeholder"))) \
    + msg(5, b"1.1.0")
saver_def = msg(1, b"save/Const:0") + msg(2, b"save/control_dependency:0") + msg(3, b"save/restore_all") \
    + enum(4, 5) + enum(7, 2)                                # max_to_keep = 5, version = V2

PARAM_ORDER = ["model_name", "data_pipeline", "net_input_size", "output_resolution", "batch_norm",
               "channel_multiplier", "guide_complexity", "luma_bins", "spatial_bin", "learning_rate",
               "crop_offset", "use_hdrp"]
node_list = b"".join(msg(1, (n + ":0").encode()) for n in PARAM_ORDER)
collections = map_entry(4, "trainable_variables", msg(2, msg(1, b"\n\x0finference/a:0") + msg(1, b"\n\x0finference/b:0"))) \
    + map_entry(4, "model_params", msg(1, node_list)) \
    + map_entry(4, "summaries", msg(1, msg(1, b"loss:0")))

meta = msg(1, meta_info) + msg(2, graph_def) + msg(3, saver_def) + collections

if __name__ == "__main__":
    with open(OUT, "wb") as f:
        f.write(meta)
    import hashlib
    print(OUT, len(meta), "bytes, sha256", hashlib.sha256(meta).hexdigest())
