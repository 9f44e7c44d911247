"""TEST INFRASTRUCTURE: writes the HDF5 layout of the reference's item-embedding file (T5/item_encode.py:115-122 with
h5py defaults) from the HDF5 File Format Specification, independently of the reader under test (h5lite.py):
version 0 superblock, version 1 object headers (optionally split by a continuation block), an old-style root group
(version 1 B-tree, local heap, symbol node), a chunked + deflate (+ shuffle) float32 dataset with a one- or two-level chunk
B-tree, and a scalar fixed-length string dataset.  There is no libhdf5 / h5py in the image to produce a real file with.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, data: bytes, align: int = 8) -> int:
        while len(self.b) % align:
            self.b.append(0)
        addr = len(self.b)
        self.b += data
        return addr

    def reserve(self, size: int) -> int:
        return self.alloc(bytes(size))

    def put(self, addr: int, data: bytes):
        self.b[addr:addr + len(data)] = data


def _msg(mtype: int, body: bytes) -> bytes:
    body = body + bytes((-len(body)) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(buf: _Buf, msgs, split_after=None) -> int:
    """Version 1 object header; `split_after`: put the messages after that index into a continuation block."""
    if split_after is None:
        body = b"".join(msgs)
        head = struct.pack("<BxHII4x", 1, len(msgs), 1, len(body))
        return buf.alloc(head + body)
    first, rest = msgs[:split_after], msgs[split_after:]
    cont_body = b"".join(rest)
    cont_addr = buf.alloc(cont_body)
    first_body = b"".join(first) + _msg(0x0010, struct.pack("<QQ", cont_addr, len(cont_body)))
    head = struct.pack("<BxHII4x", 1, len(msgs) + 1, 1, len(first_body))
    return buf.alloc(head + first_body)


def _dataspace(shape) -> bytes:
    return _msg(0x0001, struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape))


def _dtype_f32() -> bytes:
    # class 1 (float) version 1; bit field: little endian, mantissa normalisation = implied (bits 4-5 = 2), sign at bit 31;
    # properties: bit offset 0, precision 32, exponent location 23, exponent size 8, mantissa location 0, size 23, bias 127
    return _msg(0x0003, struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127))


def _dtype_string(n: int) -> bytes:
    return _msg(0x0003, struct.pack("<BBBBI", 0x13, 0x01, 0, 0, n))     # class 3, null-padded ASCII


def _filters(shuffle: bool, level: int = 4) -> bytes:
    descs = []
    if shuffle:
        descs.append(struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\x00" + struct.pack("<I", 4) + bytes(4))
    descs.append(struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\x00" + struct.pack("<I", level) + bytes(4))
    return _msg(0x000B, struct.pack("<BB6x", 1, len(descs)) + b"".join(descs))


def _chunk_node(buf: _Buf, level: int, entries, nd1: int, last_key) -> int:
    """entries: [(size, mask, offsets, child address)], last_key: (size, mask, offsets) terminating key."""
    body = bytearray()
    for size, mask, offs, child in entries:
        body += struct.pack("<II", size, mask) + struct.pack(f"<{nd1}Q", *offs) + struct.pack("<Q", child)
    body += struct.pack("<II", last_key[0], last_key[1]) + struct.pack(f"<{nd1}Q", *last_key[2])
    head = b"TREE" + struct.pack("<BBHQQ", 1, level, len(entries), UNDEF, UNDEF)
    return buf.alloc(head + bytes(body))


def write_item_embs(path: str, embs: np.ndarray, meta_json: bytes, chunk=(64, 32), shuffle=False, two_level=False,
                    split_header=False, base_offset=0):
    embs = np.ascontiguousarray(embs, dtype="<f4")
    n, d = embs.shape
    buf = _Buf()
    buf.reserve(96)                                        # superblock v0 + root symbol table entry, filled in last
    # ---- chunks of 'item_embs'
    leaves = []
    for r0 in range(0, n, chunk[0]):
        for c0 in range(0, d, chunk[1]):
            tile = np.zeros(chunk, dtype="<f4")
            part = embs[r0:r0 + chunk[0], c0:c0 + chunk[1]]
            tile[:part.shape[0], :part.shape[1]] = part
            raw = tile.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 4).T.tobytes()
            comp = zlib.compress(raw, 4)
            leaves.append((len(comp), 0, (r0, c0, 0), buf.alloc(comp)))
    end_key = (0, 0, (((n + chunk[0] - 1) // chunk[0]) * chunk[0], 0, 0))
    if two_level and len(leaves) > 2:
        half = len(leaves) // 2
        left, right = leaves[:half], leaves[half:]
        n0 = _chunk_node(buf, 0, left, 3, right[0][:3])
        n1 = _chunk_node(buf, 0, right, 3, end_key)
        kids = [(left[0][0], left[0][1], left[0][2], n0), (right[0][0], right[0][1], right[0][2], n1)]
        btree = _chunk_node(buf, 1, kids, 3, end_key)
    else:
        btree = _chunk_node(buf, 0, leaves, 3, end_key) if leaves else UNDEF
    layout = _msg(0x0008, struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", btree) + struct.pack("<III", chunk[0], chunk[1], 4))
    fill = _msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0))                                 # fill value v2, undefined
    emb_msgs = [_dataspace((n, d)), _dtype_f32(), fill, layout, _filters(shuffle)]
    emb_hdr = _object_header(buf, emb_msgs, split_after=2 if split_header else None)
    # ---- 'meta': scalar fixed-length string, contiguous
    meta_addr = buf.alloc(meta_json)
    meta_layout = _msg(0x0008, struct.pack("<BB", 3, 1) + struct.pack("<QQ", meta_addr, len(meta_json)))
    meta_hdr = _object_header(buf, [_dataspace(()), _dtype_string(len(meta_json)), meta_layout])
    # ---- root group: local heap, symbol node, B-tree, object header
    names = bytearray(8)                                   # offset 0: the empty name
    offs = {}
    for name in ("item_embs", "meta"):
        offs[name] = len(names)
        names += name.encode() + b"\x00"
        names += bytes((-len(names)) % 8)
    seg_addr = buf.alloc(bytes(names))
    heap = buf.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(names), UNDEF, seg_addr))

    def entry(name_off, hdr):
        return struct.pack("<QQII16x", name_off, hdr, 0, 0)
    snod = buf.alloc(b"SNOD" + struct.pack("<BxH", 1, 2) + entry(offs["item_embs"], emb_hdr) + entry(offs["meta"], meta_hdr))
    gtree = buf.alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, offs["meta"]))
    root_hdr = _object_header(buf, [_msg(0x0011, struct.pack("<QQ", gtree, heap))])
    # ---- superblock
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", base_offset, UNDEF, len(buf.b), UNDEF)      # addresses are relative to the base address
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", gtree, heap)
    assert len(sb) == 96
    buf.put(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(base_offset))                        # a user block: the superblock may start at 512, 1024, …
        f.write(bytes(buf.b))
