"""A small HDF5 reader for the one file layout the reference's pipeline produces.

`T5/item_encode.py:115-122` writes, with h5py defaults (libver "earliest"):
    f.create_dataset('item_embs', data=float32[N, dim], compression='gzip')     # chunked, deflate
    f.create_dataset('meta', data=np.bytes_(json))                              # scalar fixed-length string
and `RQ-VAE/vision_data.py:17-22` reads `f['item_embs'][:]` and `f['meta'][()]`.  h5py / libhdf5 are not part of this
image, so `EmbDataset` falls back to this module when h5py is missing.  It implements exactly what those files use, from
the HDF5 File Format Specification (version 0/1 superblock, version 1 object headers with continuation blocks, old-style
groups: version 1 B-tree + local heap + symbol nodes, dataspace v1/v2, fixed-point / floating-point / string datatypes,
layout v1/v2/v3 compact / contiguous / chunked with a version 1 chunk B-tree, filter pipeline v1/v2 with deflate and
shuffle).
Anything else raises `H5LiteError` naming the feature, so the caller knows to use h5py.

Validation status: written from the specification and tested against files produced by `tests/h5_writer.py` (an
independent writer for the same subset), and against the one libhdf5-written file this image holds (scipy's MATLAB v7.3
test fixture `testhdf5_7.4_GLNX86.mat`: 512-byte user block, version 0 superblock, old-style root group, version 1 object
header, IEEE float64, contiguous layout v2 — `tests/test_h5lite.py::test_reads_a_libhdf5_written_file`).  A libhdf5-written
chunked + deflate file has not been seen here: that part rests on the specification and the independent writer.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5LiteError(RuntimeError):
    pass


class _Dataset:
    def __init__(self, f: "File", header_addr: int):
        self.f = f
        self.msgs = f._object_header(header_addr)
        self.shape = self._dataspace()
        self.dtype, self._is_string = self._datatype()
        self.filters = self._filters()

    # ---- header messages -------------------------------------------------------------------------
    def _one(self, mtype: int, what: str) -> bytes:
        for t, body in self.msgs:
            if t == mtype:
                return body
        raise H5LiteError(f"dataset has no {what} message")

    def _dataspace(self) -> Tuple[int, ...]:
        b = self._one(0x0001, "dataspace")
        version, ndim, flags = b[0], b[1], b[2]
        if version == 1:
            off = 8
        elif version == 2:
            off = 4
            if b[3] == 2:
                raise H5LiteError("null dataspace")
        else:
            raise H5LiteError(f"dataspace message version {version}")
        L = self.f.len_size
        return tuple(int.from_bytes(b[off + i * L: off + (i + 1) * L], "little") for i in range(ndim))

    def _datatype(self):
        b = self._one(0x0003, "datatype")
        cls, version = b[0] & 0x0F, b[0] >> 4
        bits0 = b[1]
        size = struct.unpack_from("<I", b, 4)[0]
        order = ">" if (bits0 & 1) else "<"
        if cls == 1:                                   # floating point
            if size not in (2, 4, 8):
                raise H5LiteError(f"float datatype of {size} bytes")
            return np.dtype(f"{order}f{size}"), False
        if cls == 0:                                   # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), False
        if cls == 3:                                   # fixed-length string
            return np.dtype(f"S{size}"), True
        raise H5LiteError(f"datatype class {cls} (version {version}) is not supported by h5lite")

    def _filters(self) -> List[Tuple[int, List[int]]]:
        body = None
        for t, b in self.msgs:
            if t == 0x000B:
                body = b
        if body is None:
            return []
        version, n = body[0], body[1]
        off = 8 if version == 1 else 2
        if version not in (1, 2):
            raise H5LiteError(f"filter pipeline version {version}")
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", body, off)[0]
            off += 2
            if version == 1 or fid >= 256:
                name_len = struct.unpack_from("<H", body, off)[0]
                off += 2
            else:
                name_len = 0
            _flags, ncd = struct.unpack_from("<HH", body, off)
            off += 4
            if name_len:
                off += (name_len + 7) // 8 * 8 if version == 1 else name_len
            cd = list(struct.unpack_from(f"<{ncd}I", body, off)) if ncd else []
            off += 4 * ncd
            if version == 1 and ncd % 2:
                off += 4
            out.append((fid, cd))
        return out

    # ---- data ------------------------------------------------------------------------------------
    def _unfilter(self, raw: bytes, mask: int) -> bytes:
        for i in range(len(self.filters) - 1, -1, -1):         # reverse order on read
            if mask & (1 << i):
                continue
            fid, cd = self.filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:                                      # shuffle: byte planes → elements
                es = cd[0] if cd else self.dtype.itemsize
                n = len(raw) // es
                a = np.frombuffer(raw, dtype=np.uint8)
                body = a[:n * es].reshape(es, n).T.reshape(-1)
                raw = body.tobytes() + raw[n * es:]
            elif fid == 3:                                      # fletcher32: checksum appended, ignore
                raw = raw[:-4]
            else:
                raise H5LiteError(f"filter id {fid} is not supported by h5lite (deflate, shuffle, fletcher32 are)")
        return raw

    def read_rows(self, lo: int, hi: int) -> np.ndarray:
        """Rows [lo, hi) along the first axis, touching only the bytes / chunks that hold them (a rank of a sharded job
        reads its contiguous item range, not the catalogue)."""
        if not self.shape:
            raise H5LiteError("read_rows needs an array dataset")
        lo, hi = max(0, int(lo)), min(int(hi), self.shape[0])
        return self.read(rows=(lo, max(lo, hi)))

    def read(self, rows=None) -> np.ndarray:
        b = self._one(0x0008, "data layout")
        version = b[0]
        O, Lz = self.f.off_size, self.f.len_size
        count = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        if version in (1, 2):
            # libhdf5 <= 1.6 writers: [version, ndims, class, 5 reserved] [address unless compact] ndims x u32
            # (for chunked storage ndims = rank + 1, the last entry being the element size) [compact: u32 size + data].
            # Re-packed here into the version 3 body the code below reads.
            nd, cls = b[1], b[2]
            pos = 8
            addr_b = b""
            if cls != 0:
                addr_b = b[pos:pos + O]
                pos += O
            dims_b = b[pos:pos + 4 * nd]
            pos += 4 * nd
            if cls == 0:
                size = struct.unpack_from("<I", b, pos)[0]
                b = bytes([3, 0]) + struct.pack("<H", size) + b[pos + 4:pos + 4 + size]
            elif cls == 1:
                b = bytes([3, 1]) + addr_b + (count * self.dtype.itemsize).to_bytes(Lz, "little")
            else:
                b = bytes([3, 2, nd]) + addr_b + dims_b
        elif version != 3:
            raise H5LiteError(f"data layout message version {version} (files written with libver='latest' need h5py)")
        cls = b[1]
        if cls == 0:                                            # compact
            size = struct.unpack_from("<H", b, 2)[0]
            raw = b[4:4 + size]
            arr = np.frombuffer(raw, dtype=self.dtype, count=count)
        elif cls == 1:                                          # contiguous
            addr = int.from_bytes(b[2:2 + O], "little")
            size = int.from_bytes(b[2 + O:2 + O + Lz], "little")
            if rows is not None:
                row_items = int(np.prod(self.shape[1:], dtype=np.int64))
                row_bytes = row_items * self.dtype.itemsize
                nrows = rows[1] - rows[0]
                if addr == UNDEF or nrows == 0:
                    return np.zeros((nrows,) + tuple(self.shape[1:]), dtype=self.dtype)
                raw = self.f._read(addr + rows[0] * row_bytes, nrows * row_bytes)
                return np.frombuffer(raw, dtype=self.dtype, count=nrows * row_items).reshape((nrows,) + tuple(self.shape[1:])).copy()
            if addr == UNDEF:
                arr = np.zeros(count, dtype=self.dtype)
            else:
                arr = np.frombuffer(self.f._read(addr, size), dtype=self.dtype, count=count)
        elif cls == 2:                                          # chunked
            nd1 = b[2]
            btree = int.from_bytes(b[3:3 + O], "little")
            cdims = struct.unpack_from(f"<{nd1}I", b, 3 + O)
            if nd1 - 1 != len(self.shape):
                raise H5LiteError("chunk rank does not match the dataspace")
            r0, r1 = rows if rows is not None else (0, self.shape[0] if self.shape else 0)
            arr = np.zeros((r1 - r0,) + tuple(self.shape[1:]), dtype=self.dtype) if rows is not None else np.zeros(self.shape, dtype=self.dtype)
            if btree != UNDEF:
                for offs, size, mask, addr in self.f._chunk_leaves(btree, nd1):
                    if rows is not None and (offs[0] >= r1 or offs[0] + cdims[0] <= r0):
                        continue                               # chunk holds none of the requested rows: not read, not inflated
                    raw = self._unfilter(self.f._read(addr, size), mask)
                    chunk = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(cdims[:-1]))).reshape(cdims[:-1])
                    sel_d, sel_c = [], []
                    for ax, (o, c, s) in enumerate(zip(offs[:-1], cdims[:-1], self.shape)):
                        n = min(c, s - o)
                        a, z = o, o + n
                        if ax == 0 and rows is not None:
                            a, z = max(o, r0), min(o + n, r1)
                            sel_d.append(slice(a - r0, z - r0))
                        else:
                            sel_d.append(slice(a, z))
                        sel_c.append(slice(a - o, z - o))
                    arr[tuple(sel_d)] = chunk[tuple(sel_c)]
            return arr
        else:
            raise H5LiteError(f"data layout class {cls}")
        arr = arr.reshape(self.shape) if self.shape else arr.reshape(())
        if rows is not None:
            return arr[rows[0]:rows[1]].copy()
        return arr.copy()

    def __getitem__(self, key):
        arr = self.read()
        if self._is_string and arr.shape == ():
            return arr[()] if key == () else arr[key]          # numpy.bytes_, like h5py for a scalar 'S' dataset
        return arr[key]


class File:
    """`with File(path) as f: f['item_embs'][:]; f['meta'][()]` — the two accesses of vision_data.py:18-21."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise H5LiteError("h5lite is read-only")
        self._fh = open(path, "rb")
        self._parse_superblock()
        self._links = self._group_links(self._root_header)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._fh:
            self._fh.close()
            self._fh = None

    def keys(self):
        return list(self._links.keys())

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name: str) -> _Dataset:
        if name not in self._links:
            raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
        return _Dataset(self, self._links[name])

    # ---- low level -------------------------------------------------------------------------------
    def _read(self, addr: int, size: int) -> bytes:
        self._fh.seek(self.base + addr)
        data = self._fh.read(size)
        if len(data) != size:
            raise H5LiteError("truncated file")
        return data

    def _parse_superblock(self):
        pos = 0
        while True:
            self._fh.seek(pos)
            head = self._fh.read(8)
            if head == SIGNATURE:
                break
            if len(head) < 8:
                raise H5LiteError("not an HDF5 file (signature not found)")
            pos = 512 if pos == 0 else pos * 2
        self._fh.seek(pos)
        sb = self._fh.read(128)
        version = sb[8]
        if version not in (0, 1):
            raise H5LiteError(f"superblock version {version}: file was written with libver='latest' (h5py needed)")
        self.off_size, self.len_size = sb[13], sb[14]
        if self.off_size != 8 or self.len_size != 8:
            raise H5LiteError("only 8-byte offsets / lengths are supported")
        p = 24 + (4 if version == 1 else 0)
        self.base = int.from_bytes(sb[p:p + 8], "little")
        p += 4 * 8                                              # base, free-space info, end of file, driver info
        # root group symbol table entry
        self._root_header = int.from_bytes(sb[p + 8:p + 16], "little")
        cache_type = struct.unpack_from("<I", sb, p + 16)[0]
        self._root_cache = None
        if cache_type == 1:
            self._root_cache = (int.from_bytes(sb[p + 24:p + 32], "little"), int.from_bytes(sb[p + 32:p + 40], "little"))

    def _object_header(self, addr: int) -> List[Tuple[int, bytes]]:
        head = self._read(addr, 16)
        if head[:4] == b"OHDR":
            raise H5LiteError("version 2 object header (libver='latest'): h5py needed")
        if head[0] != 1:
            raise H5LiteError(f"object header version {head[0]}")
        nmsgs = struct.unpack_from("<H", head, 2)[0]
        size = struct.unpack_from("<I", head, 8)[0]
        blocks = [(addr + 16, size)]
        msgs: List[Tuple[int, bytes]] = []
        while blocks and len(msgs) < nmsgs:
            baddr, bsize = blocks.pop(0)
            data = self._read(baddr, bsize)
            p = 0
            while p + 8 <= bsize and len(msgs) < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", data, p)
                body = data[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                            # continuation
                    blocks.append((int.from_bytes(body[:8], "little"), int.from_bytes(body[8:16], "little")))
                msgs.append((mtype, body))
        return msgs

    def _group_links(self, header_addr: int) -> Dict[str, int]:
        btree = heap = None
        for t, body in self._object_header(header_addr):
            if t == 0x0011:                                    # symbol table message
                btree, heap = int.from_bytes(body[:8], "little"), int.from_bytes(body[8:16], "little")
            elif t in (0x0002, 0x0006):
                raise H5LiteError("new-style group (link messages): h5py needed")
        if btree is None:
            if self._root_cache is None:
                raise H5LiteError("root group has no symbol table")
            btree, heap = self._root_cache
        h = self._read(heap, 32)
        if h[:4] != b"HEAP":
            raise H5LiteError("bad local heap signature")
        seg_size = int.from_bytes(h[8:16], "little")
        seg_addr = int.from_bytes(h[24:32], "little")
        names = self._read(seg_addr, seg_size)
        links: Dict[str, int] = {}
        self._walk_group_btree(btree, names, links)
        return links

    def _walk_group_btree(self, addr: int, names: bytes, links: Dict[str, int]):
        node = self._read(addr, 24)
        if node[:4] == b"SNOD":
            nsym = struct.unpack_from("<H", node, 6)[0]
            ent = self._read(addr + 8, nsym * 40)
            for i in range(nsym):
                e = ent[i * 40:(i + 1) * 40]
                noff = int.from_bytes(e[:8], "little")
                end = names.index(b"\x00", noff)
                links[names[noff:end].decode("utf-8")] = int.from_bytes(e[8:16], "little")
            return
        if node[:4] != b"TREE" or node[4] != 0:
            raise H5LiteError("bad group B-tree node")
        used = struct.unpack_from("<H", node, 6)[0]
        body = self._read(addr + 24, (2 * used + 1) * 8)
        for i in range(used):
            child = int.from_bytes(body[(2 * i + 1) * 8:(2 * i + 2) * 8], "little")
            self._walk_group_btree(child, names, links)

    def _chunk_leaves(self, addr: int, nd1: int):
        """Yields (offsets[nd1], stored size, filter mask, address) of every chunk below a version 1 chunk B-tree node."""
        node = self._read(addr, 24)
        if node[:4] != b"TREE" or node[4] != 1:
            raise H5LiteError("bad chunk B-tree node")
        level, used = node[5], struct.unpack_from("<H", node, 6)[0]
        key = 8 + 8 * nd1
        body = self._read(addr + 24, used * (key + 8) + key)
        for i in range(used):
            p = i * (key + 8)
            size, mask = struct.unpack_from("<II", body, p)
            offs = struct.unpack_from(f"<{nd1}Q", body, p + 8)
            child = int.from_bytes(body[p + key:p + key + 8], "little")
            if level == 0:
                yield offs, size, mask, child
            else:
                yield from self._chunk_leaves(child, nd1)
