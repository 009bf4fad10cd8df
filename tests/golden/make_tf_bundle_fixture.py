"""Assembles, byte by byte from the published formats, the small TensorFlow V2 checkpoint
("tensor bundle") under tests/golden/tf_bundle/ -- WITHOUT touching hdrnet_b200/checkpoint.py.

Why it exists: TensorFlow is not installed and the reference ships no checkpoint, so round 1 could
only test the reader against the module's own writer (missing 7).  This script is a
second, independent statement of the on-disk format, written from the specifications:

  * LevelDB table format (leveldb/doc/table_format.md; TensorFlow's copy is
    tensorflow/core/lib/io/{table_builder,format,block_builder}.cc): data blocks of prefix-
    compressed entries [varint shared | varint non_shared | varint value_len | key tail | value],
    a restart array of uint32 offsets + their count, a 5-byte trailer per block (compression type 0
    = none, masked CRC-32C of contents + type), one index block (last key of each data block ->
    BlockHandle{varint offset, varint size}), an empty metaindex block, and a 48-byte footer
    (metaindex handle, index handle, zero padding to 40 bytes, magic 0xdb4775248b80fb57 LE);
  * tensor bundle (tensorflow/core/util/tensor_bundle/tensor_bundle.cc, protobuf/tensor_bundle.proto):
    key "" -> BundleHeaderProto{num_shards = 1, endianness = LITTLE, version{producer = 1}};
    key <variable name> -> BundleEntryProto{dtype, shape, shard_id, offset, size, crc32c (masked,
    fixed32)}; tensor bytes concatenated in key order in <prefix>.data-00000-of-00001;
  * CRC-32C: reflected polynomial 0x82F63B78, computed here BIT BY BIT (the module uses a table);
    mask(crc) = rotr(crc, 15) + 0xa282ead8.

The protobuf messages are written as literal bytes with the field arithmetic in comments.  Two of
the three variable names share the prefix "inference/" and sit in one restart group, so the reader's
prefix decompression is exercised; a second data block is forced so the index block has two entries.

    python tests/golden/make_tf_bundle_fixture.py        # rewrites the three files
"""
import os
import struct

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_bundle")
PREFIX = "model.ckpt-7"


def crc32c_bitwise(data: bytes) -> int:
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    return crc ^ 0xFFFFFFFF


def mask(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def varint(v: int) -> bytes:
    out = b""
    while v >= 0x80:
        out += bytes([(v & 0x7F) | 0x80])
        v >>= 7
    return out + bytes([v])


def f32(*vals) -> bytes:
    return struct.pack("<%df" % len(vals), *vals)


# ---- the tensors (name -> (dtype code, shape, little-endian bytes)); values are exact in float32 ----
TENSORS = [
    ("global_step", 9, [], struct.pack("<q", 1234567)),                                     # DT_INT64 scalar
    ("inference/coefficients/splat/conv1/biases", 1, [8],
     f32(0.5, -0.25, 1.0, 2.0, -3.5, 0.125, 100.0, -0.0078125)),                           # DT_FLOAT [8]
    ("inference/guide/ccm", 1, [3, 3],
     f32(1.0, 0.0625, -0.0625, 0.03125, 0.96875, 0.0, -0.015625, 0.25, 0.75)),             # DT_FLOAT [3, 3]
]


def bundle_entry(dtype: int, shape, offset: int, size: int, crc_masked: int) -> bytes:
    # BundleEntryProto: 1 dtype (varint) | 2 shape (message) | 3 shard_id | 4 offset | 5 size | 6 crc32c (fixed32)
    shape_msg = b""
    for d in shape:
        dim = b"\x08" + varint(d)                       # TensorShapeProto.Dim: field 1 (size), varint
        shape_msg += b"\x12" + varint(len(dim)) + dim   # TensorShapeProto: field 2 (dim), length-delimited
    out = b"\x08" + varint(dtype)                       # (1 << 3) | 0
    out += b"\x12" + varint(len(shape_msg)) + shape_msg  # (2 << 3) | 2   (present, possibly empty: scalar)
    # shard_id = 0 is proto3's default: not serialised
    if offset:
        out += b"\x20" + varint(offset)                 # (4 << 3) | 0
    out += b"\x28" + varint(size)                       # (5 << 3) | 0
    out += b"\x35" + struct.pack("<I", crc_masked)      # (6 << 3) | 5
    return out


# BundleHeaderProto{num_shards: 1, version{producer: 1}}; endianness LITTLE = 0 is the default (absent)
HEADER = b"\x08\x01" + b"\x1a\x02\x08\x01"


def block(entries, restart_every=16) -> bytes:
    """entries: [(key, value)] sorted.  Returns block contents (without the trailer)."""
    out, restarts, last = b"", [], b""
    for i, (key, value) in enumerate(entries):
        if i % restart_every == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        out += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    return out + struct.pack("<I", len(restarts))


def with_trailer(contents: bytes) -> bytes:
    return contents + b"\x00" + struct.pack("<I", mask(crc32c_bitwise(contents + b"\x00")))


def main():
    os.makedirs(HERE, exist_ok=True)
    data, records = b"", [(b"", HEADER)]
    for name, dtype, shape, raw in TENSORS:
        records.append((name.encode(), bundle_entry(dtype, shape, len(data), len(raw), mask(crc32c_bitwise(raw)))))
        data += raw
    # two data blocks: [header, global_step] and the two "inference/..." variables (prefix-compressed)
    blocks = [records[:2], records[2:]]
    table, index_entries = b"", []
    for ents in blocks:
        contents = block(ents)
        index_entries.append((ents[-1][0], varint(len(table)) + varint(len(contents))))
        table += with_trailer(contents)
    meta = block([])
    meta_handle = varint(len(table)) + varint(len(meta))
    table += with_trailer(meta)
    index = block(index_entries, restart_every=1)       # LevelDB: restart interval 1 in index blocks
    index_handle = varint(len(table)) + varint(len(index))
    table += with_trailer(index)
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + bytes.fromhex("57fb808b247547db")
    table += footer
    with open(os.path.join(HERE, PREFIX + ".index"), "wb") as f:
        f.write(table)
    with open(os.path.join(HERE, PREFIX + ".data-00000-of-00001"), "wb") as f:
        f.write(data)
    with open(os.path.join(HERE, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (PREFIX, PREFIX))
    print("wrote", len(table), "index bytes,", len(data), "data bytes to", HERE)


if __name__ == "__main__":
    main()
