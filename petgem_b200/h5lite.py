"""Minimal HDF5 writer (and reader of what it writes) for the results file of ``Postprocessing.run``.

The reference writes its results with h5py (petgem/postprocessing.py:341-461): groups ``machine`` and ``model``
holding scalar provenance datasets and the receiver responses.  h5py is not a dependency of this package, so
the small subset of the HDF5 file format needed for that schema is produced here directly (HDF5 File Format
Specification v3: version-2 superblock, version-2 object headers with their lookup3 checksums, compact
new-style groups, contiguous datasets).  Supported values: python/numpy scalars and arrays of float64, int64,
complex128 (the compound {r, i} h5py uses), bool (h5py's FALSE/TRUE enum) and strings (fixed-length UTF-8).
``read`` parses the same subset back into nested dicts (used by the tests).  ``write_classic`` produces the same
tree in the CLASSIC layout instead (version-0 superblock, symbol-table groups, version-1 object headers and
datatype encodings) -- byte for byte the structures h5py itself writes by default, and what ``Postprocessing``
uses: ``read_classic`` below, which parses the reference's own h5py-written files, reads it back.

``read_classic`` reads the files h5py writes by default -- the format of PETGEM's INPUT files (receiver positions,
conductivity models: ``preprocessing.py:399-407``, ``tests/data/receiver_pos.h5``): version-0/1 superblock,
symbol-table groups (B-tree + local heap), version-1 object headers with continuation blocks, contiguous or compact
datasets of fixed-point / IEEE float / string / {r, i} compound type.  Chunked or filtered datasets are refused.
"""
from __future__ import annotations

import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_M = 0xFFFFFFFF


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & _M


def lookup3(data: bytes, initval: int = 0) -> int:
    """Bob Jenkins' lookup3 hashlittle, the checksum of HDF5 metadata (H5_checksum_lookup3)."""
    length = len(data)
    a = b = c = (0xDEADBEEF + length + initval) & _M
    i = 0
    while length > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & _M
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & _M
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & _M
        a = (a - c) & _M; a ^= _rot(c, 4); c = (c + b) & _M   # noqa: E702
        b = (b - a) & _M; b ^= _rot(a, 6); a = (a + c) & _M   # noqa: E702
        c = (c - b) & _M; c ^= _rot(b, 8); b = (b + a) & _M   # noqa: E702
        a = (a - c) & _M; a ^= _rot(c, 16); c = (c + b) & _M  # noqa: E702
        b = (b - a) & _M; b ^= _rot(a, 19); a = (a + c) & _M  # noqa: E702
        c = (c - b) & _M; c ^= _rot(b, 4); b = (b + a) & _M   # noqa: E702
        i += 12
        length -= 12
    if length == 0:
        return c
    tail = data[i:] + b"\0" * (12 - length)
    a = (a + int.from_bytes(tail[0:4], "little")) & _M
    b = (b + int.from_bytes(tail[4:8], "little")) & _M
    c = (c + int.from_bytes(tail[8:12], "little")) & _M
    c ^= b; c = (c - _rot(b, 14)) & _M  # noqa: E702
    a ^= c; a = (a - _rot(c, 11)) & _M  # noqa: E702
    b ^= a; b = (b - _rot(a, 25)) & _M  # noqa: E702
    c ^= b; c = (c - _rot(b, 16)) & _M  # noqa: E702
    a ^= c; a = (a - _rot(c, 4)) & _M   # noqa: E702
    b ^= a; b = (b - _rot(a, 14)) & _M  # noqa: E702
    c ^= b; c = (c - _rot(b, 24)) & _M  # noqa: E702
    return c


# ---- datatype messages -------------------------------------------------------------------------
def _dt_float64():
    return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def _dt_int(size, signed=True):
    return struct.pack("<BBBBI", 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, 8 * size)


def _dt_string(n):
    return struct.pack("<BBBBI", 0x13, 0x11, 0, 0, n)  # null-padded, UTF-8


def _dt_complex128():
    body = b"r\0" + struct.pack("<B", 0) + _dt_float64() + b"i\0" + struct.pack("<B", 8) + _dt_float64()
    return struct.pack("<BBBBI", 0x36, 2, 0, 0, 16) + body  # compound, version 3, two members


def _dt_bool():
    base = _dt_int(1)
    return struct.pack("<BBBBI", 0x38, 2, 0, 0, 1) + base + b"FALSE\0TRUE\0" + b"\x00\x01"  # enum, version 3


def _encode(value):
    """python / numpy value -> (datatype message, shape, raw little-endian bytes)."""
    if isinstance(value, (bytes, str)):
        raw = value.encode("utf-8") if isinstance(value, str) else value
        raw = raw or b"\0"
        return _dt_string(len(raw)), (), raw
    a = np.asarray(value)
    if a.dtype.kind in ("U", "S", "O"):
        if a.ndim == 0:
            return _encode(str(a.item()))
        items = [str(v).encode("utf-8") for v in a.reshape(-1)]
        n = max(1, max(len(v) for v in items))
        return _dt_string(n), a.shape, b"".join(v.ljust(n, b"\0") for v in items)
    if a.dtype.kind == "b":
        return _dt_bool(), a.shape, np.ascontiguousarray(a, dtype=np.int8).tobytes()
    if a.dtype.kind in ("i", "u"):
        return _dt_int(8), a.shape, np.ascontiguousarray(a, dtype="<i8").tobytes()
    if a.dtype.kind == "f":
        return _dt_float64(), a.shape, np.ascontiguousarray(a, dtype="<f8").tobytes()
    if a.dtype.kind == "c":
        return _dt_complex128(), a.shape, np.ascontiguousarray(a, dtype="<c16").tobytes()
    raise TypeError("h5lite: unsupported value of dtype %s" % a.dtype)


def _message(mtype, data, flags=0):
    return struct.pack("<BHB", mtype, len(data), flags) + data


def _object_header(messages):
    body = b"".join(messages)
    head = b"OHDR" + struct.pack("<BB", 2, 0x02) + struct.pack("<I", len(body))  # chunk-0 size field: 4 bytes
    blob = head + body
    return blob + struct.pack("<I", lookup3(blob))


class _Writer:
    def __init__(self):
        self.buf = bytearray(b"\0" * 48)  # superblock, filled in at the end

    def _align(self, n=8):
        self.buf.extend(b"\0" * (-len(self.buf) % n))

    def dataset(self, value):
        dt, shape, raw = _encode(value)
        self._align()
        data_addr = len(self.buf)
        self.buf.extend(raw)
        if shape == ():
            space = struct.pack("<BBBB", 2, 0, 0, 0)
        else:
            space = struct.pack("<BBBB", 2, len(shape), 0, 1) + b"".join(struct.pack("<Q", int(d)) for d in shape)
        msgs = [_message(0x01, space), _message(0x03, dt, flags=0x01), _message(0x05, struct.pack("<BB", 3, 0x09)),
                _message(0x08, struct.pack("<BBQQ", 3, 1, data_addr if raw else _UNDEF, len(raw)))]
        self._align()
        addr = len(self.buf)
        self.buf.extend(_object_header(msgs))
        return addr

    def group(self, tree):
        links = []
        for name, value in tree.items():
            child = self.group(value) if isinstance(value, dict) else self.dataset(value)
            nm = str(name).encode("utf-8")
            if len(nm) > 255:
                raise ValueError("h5lite: link name longer than 255 bytes")
            links.append(_message(0x06, struct.pack("<BBBB", 1, 0x10, 1, len(nm)) + nm + struct.pack("<Q", child)))
        msgs = [_message(0x02, struct.pack("<BBQQ", 0, 0, _UNDEF, _UNDEF)), _message(0x0A, struct.pack("<BB", 0, 0))]
        self._align()
        addr = len(self.buf)
        self.buf.extend(_object_header(msgs + links))
        return addr

    def finish(self, root):
        self._align()
        sb = _SIG + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), root)
        self.buf[:48] = sb + struct.pack("<I", lookup3(sb))
        return bytes(self.buf)


def write(path, tree):
    """tree: nested dict; a dict is a group, anything else a dataset (scalar or array)."""
    w = _Writer()
    blob = w.finish(w.group(tree))
    with open(path, "wb") as fh:
        fh.write(blob)


# ---- reader of the same subset -------------------------------------------------------------------
def _parse_dtype(b, o):
    cls, ver = b[o] & 0x0F, b[o] >> 4
    bits0, bits1 = b[o + 1], b[o + 2]
    size = struct.unpack_from("<I", b, o + 4)[0]
    o += 8
    if cls == 0:
        return ("int", size), o + 4
    if cls == 1:
        return ("float", size), o + 12
    if cls == 3:
        return ("str", size), o
    if cls == 6 and ver == 3:
        members = []
        for _ in range(bits0 | (bits1 << 8)):
            e = b.index(b"\0", o)
            name = b[o:e].decode()
            nb = 1 if size < 256 else 2 if size < 65536 else 4
            off = int.from_bytes(b[e + 1:e + 1 + nb], "little")
            mt, o = _parse_dtype(b, e + 1 + nb)
            members.append((name, off, mt))
        return ("compound", size, members), o
    if cls == 8 and ver == 3:
        base, o = _parse_dtype(b, o)
        names = []
        for _ in range(bits0 | (bits1 << 8)):
            e = b.index(b"\0", o)
            names.append(b[o:e].decode())
            o = e + 1
        return ("enum", size, names), o + len(names) * base[1]
    raise ValueError("h5lite.read: datatype class %d version %d" % (cls, ver))


def _read_object(b, addr):
    if b[addr:addr + 4] != b"OHDR" or b[addr + 4] != 2:
        raise ValueError("h5lite.read: not a version-2 object header at %d" % addr)
    flags = b[addr + 5]
    o = addr + 6
    nsz = 1 << (flags & 3)
    size = int.from_bytes(b[o:o + nsz], "little")
    o += nsz
    end = o + size
    if struct.unpack_from("<I", b, end)[0] != lookup3(bytes(b[addr:end])):
        raise ValueError("h5lite.read: object header checksum mismatch at %d" % addr)
    links, dt, shape, layout = {}, None, None, None
    is_group = False
    while o < end:
        mtype, msize, _ = struct.unpack_from("<BHB", b, o)
        d = o + 4
        if mtype == 0x02:
            is_group = True
        elif mtype == 0x06:
            lf = b[d + 1]
            q = d + 2 + (1 if lf & 0x08 else 0) + (8 if lf & 0x04 else 0) + (1 if lf & 0x10 else 0)
            ln = b[q]
            links[b[q + 1:q + 1 + ln].decode()] = struct.unpack_from("<Q", b, q + 1 + ln)[0]
        elif mtype == 0x01:
            rank = b[d + 1]
            shape = tuple(struct.unpack_from("<Q", b, d + 4 + 8 * i)[0] for i in range(rank))
        elif mtype == 0x03:
            dt, _ = _parse_dtype(b, d)
        elif mtype == 0x08:
            layout = struct.unpack_from("<QQ", b, d + 2)
        o = d + msize
    if is_group:
        return {k: _read_object(b, a) for k, a in links.items()}
    raw = bytes(b[layout[0]:layout[0] + layout[1]]) if layout[1] else b""
    kind = dt[0]
    if kind == "str":
        n = dt[1]
        vals = [raw[i:i + n].rstrip(b"\0").decode() for i in range(0, len(raw), n)]
        return vals[0] if shape == () else np.array(vals).reshape(shape)
    np_dt = {"float": "<f8", "int": "<i%d" % dt[1], "compound": "<c16", "enum": "|i1"}[kind]
    a = np.frombuffer(raw, dtype=np_dt).reshape(shape)
    if kind == "enum":
        a = a.astype(bool)
    return a[()] if shape == () else a.copy()


def read(path):
    """Nested dicts of what `write` produced (checks the superblock and object-header checksums)."""
    b = open(path, "rb").read()
    if b[:8] != _SIG or b[8] != 2:
        raise ValueError("h5lite.read: not an HDF5 file with a version-2 superblock")
    if struct.unpack_from("<I", b, 44)[0] != lookup3(b[:44]):
        raise ValueError("h5lite.read: superblock checksum mismatch")
    eof, root = struct.unpack_from("<QQ", b, 28)
    if eof != len(b):
        raise ValueError("h5lite.read: end-of-file address %d != file size %d" % (eof, len(b)))
    return _read_object(b, root)


# ---- reader for the classic layout (what h5py writes by default) -----------------------------------
def _classic_dtype(b, o):
    cls, ver = b[o] & 0x0F, b[o] >> 4
    bits0 = b[o + 1]
    size = struct.unpack_from("<I", b, o + 4)[0]
    if cls == 0:
        return np.dtype(("<" if not bits0 & 1 else ">") + ("i" if bits0 & 0x08 else "u") + str(size))
    if cls == 1:
        return np.dtype(("<" if not bits0 & 1 else ">") + "f" + str(size))
    if cls == 3:
        return np.dtype("S%d" % size)
    if cls == 6:  # {r, i} of two equal floats -> complex
        nmem = b[o + 1] | (b[o + 2] << 8)
        q = o + 8
        members = []
        for _ in range(nmem):
            e = b.index(b"\0", q)
            name = b[q:e].decode()
            if ver < 3:
                q = q + ((e - q) // 8 + 1) * 8       # name padded to a multiple of 8
                off = struct.unpack_from("<I", b, q)[0]
                q += 4 + (28 if ver == 1 else 0)      # v1: dimensionality, permutation, 4 dimension sizes
            else:
                nb = 1 if size < 256 else 2 if size < 65536 else 4
                off = int.from_bytes(b[e + 1:e + 1 + nb], "little")
                q = e + 1 + nb
            mt = _classic_dtype(b, q)
            q += 8 + (12 if mt.kind == "f" else 4)
            members.append((name, off, mt))
        if len(members) == 2 and members[0][2] == members[1][2] and members[0][2].kind == "f":
            return np.dtype("<c%d" % size)
        return np.dtype({"names": [m[0] for m in members], "formats": [m[2] for m in members],
                         "offsets": [m[1] for m in members], "itemsize": size})
    if cls == 8:  # enumeration: h5py's bool is an int8 enum {FALSE, TRUE}
        base = _classic_dtype(b, o + 8)
        nmem = b[o + 1] | (b[o + 2] << 8)
        q = o + 8 + 8 + (12 if base.kind == "f" else 4)
        names = []
        for _ in range(nmem):
            e = b.index(b"\0", q)
            names.append(b[q:e])
            q = q + ((e - q) // 8 + 1) * 8 if ver < 3 else e + 1
        return np.dtype(bool) if names == [b"FALSE", b"TRUE"] and base.itemsize == 1 else base
    raise ValueError("h5lite.read_classic: datatype class %d is not supported" % cls)


def _classic_messages(b, addr):
    """(type, data offset, size) of every message of a version-1 object header, continuation blocks included."""
    if b[addr] != 1:
        raise ValueError("h5lite.read_classic: object header version %d at %d" % (b[addr], addr))
    nmsg = struct.unpack_from("<H", b, addr + 2)[0]
    hsize = struct.unpack_from("<I", b, addr + 8)[0]
    blocks = [(addr + 16, hsize)]
    out = []
    while blocks and len(out) < nmsg:
        o, left = blocks.pop(0)
        end = o + left
        while o + 8 <= end and len(out) < nmsg:
            mtype, msize = struct.unpack_from("<HH", b, o)
            d = o + 8
            if mtype == 0x10:  # continuation
                blocks.append(struct.unpack_from("<QQ", b, d))
            out.append((mtype, d, msize))
            o = d + msize
    return out


def _classic_object(b, addr):
    msgs = _classic_messages(b, addr)
    kinds = {m[0]: m for m in msgs}
    if 0x11 in kinds:  # symbol table message: a group
        btree, heap = struct.unpack_from("<QQ", b, kinds[0x11][1])
        return _classic_group(b, btree, heap)
    if 0x01 not in kinds or 0x03 not in kinds or 0x08 not in kinds:
        raise ValueError("h5lite.read_classic: object at %d is neither a group nor a plain dataset" % addr)
    d = kinds[0x01][1]
    ver, rank = b[d], b[d + 1]
    dims_at = d + (8 if ver == 1 else 4)
    shape = tuple(struct.unpack_from("<Q", b, dims_at + 8 * i)[0] for i in range(rank))
    dt = _classic_dtype(b, kinds[0x03][1])
    L = kinds[0x08][1]
    if b[L] != 3:
        raise ValueError("h5lite.read_classic: data layout version %d" % b[L])
    count = int(np.prod(shape)) if shape else 1
    if b[L + 1] == 1:
        a, nbytes = struct.unpack_from("<QQ", b, L + 2)
        raw = b[a:a + nbytes] if a != _UNDEF else b""
    elif b[L + 1] == 0:
        nbytes = struct.unpack_from("<H", b, L + 2)[0]
        raw = b[L + 4:L + 4 + nbytes]
    else:
        raise ValueError("h5lite.read_classic: chunked datasets are not supported (store the array contiguously or "
                         "convert it to .npy)")
    if 0x0B in kinds:
        raise ValueError("h5lite.read_classic: filtered (compressed) datasets are not supported")
    if len(raw) < count * dt.itemsize:
        raise ValueError("h5lite.read_classic: dataset shorter than its dataspace")
    a = np.frombuffer(raw, dtype=dt, count=count).reshape(shape)
    if dt.kind == "S":
        vals = [v.rstrip(b"\0").decode() for v in a.reshape(-1)]
        return vals[0] if shape == () else np.array(vals).reshape(shape)
    return a[()] if shape == () else a.copy()


def _classic_group(b, btree, heap):
    if b[heap:heap + 4] != b"HEAP":
        raise ValueError("h5lite.read_classic: no local heap at %d" % heap)
    names_at = struct.unpack_from("<Q", b, heap + 24)[0]
    out = {}

    def node(a):
        if b[a:a + 4] == b"TREE":
            level, used = b[a + 5], struct.unpack_from("<H", b, a + 6)[0]
            q = a + 24 + 8  # first child follows key 0
            for _ in range(used):
                child = struct.unpack_from("<Q", b, q)[0]
                node(child)
                q += 16
            del level
        elif b[a:a + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, a + 6)[0]
            for i in range(n):
                e = a + 8 + 40 * i
                noff, oh = struct.unpack_from("<QQ", b, e)
                z = b.index(b"\0", names_at + noff)
                out[b[names_at + noff:z].decode()] = _classic_object(b, oh)
        else:
            raise ValueError("h5lite.read_classic: unexpected node at %d" % a)

    node(btree)
    return out


def read_classic(path):
    """Nested dicts of the groups and datasets of a classic-layout HDF5 file (h5py's default)."""
    b = open(path, "rb").read()
    if b[:8] != _SIG:
        raise ValueError("h5lite.read_classic: %s is not an HDF5 file" % path)
    if b[8] not in (0, 1):
        return read(path)  # version-2 superblock: the subset `write` produces
    if b[13] != 8 or b[14] != 8:
        raise ValueError("h5lite.read_classic: only 8-byte offsets and lengths")
    root = 24 + 32 + (4 if b[8] == 1 else 0)  # root symbol table entry
    oh = struct.unpack_from("<Q", b, root + 8)[0]
    return _classic_object(b, oh)


# ---- writer of the classic layout ------------------------------------------------------------------
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _v1_dt_complex128():
    def member(name, off):
        return _pad8(name + b"\0") + struct.pack("<IB3xI4x4I", off, 0, 0, 0, 0, 0, 0) + _dt_float64()
    return struct.pack("<BBBBI", 0x16, 2, 0, 0, 16) + member(b"r", 0) + member(b"i", 8)


def _v1_dt_bool():
    return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + _dt_int(1) + _pad8(b"FALSE\0") + _pad8(b"TRUE\0") + b"\x00\x01"


def _encode_classic(value):
    dt, shape, raw = _encode(value)
    cls = dt[0] & 0x0F
    if cls == 6:
        dt = _v1_dt_complex128()
    elif cls == 8:
        dt = _v1_dt_bool()
    return dt, shape, raw


def _v1_message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _v1_object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _ClassicWriter:
    INTERNAL_K = 16

    def __init__(self, leaf_k):
        self.leaf_k = leaf_k
        self.buf = bytearray(b"\0" * 96)  # superblock + root symbol table entry, filled in at the end

    def _place(self, blob):
        self.buf.extend(b"\0" * (-len(self.buf) % 8))
        addr = len(self.buf)
        self.buf.extend(blob)
        return addr

    def dataset(self, value):
        dt, shape, raw = _encode_classic(value)
        data_addr = self._place(raw) if raw else _UNDEF
        space = struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)
        msgs = [_v1_message(0x01, space), _v1_message(0x03, dt, flags=0x01),
                _v1_message(0x05, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=0x01),
                _v1_message(0x08, struct.pack("<BBQQ", 3, 1, data_addr, len(raw)))]
        return self._place(_v1_object_header(msgs))

    def group(self, tree):
        """-> (object header address, B-tree address, heap address)"""
        children = []
        for name, value in tree.items():
            nm = str(name).encode("utf-8")
            addr = self.group(value)[0] if isinstance(value, dict) else self.dataset(value)
            children.append((nm, addr))
        children.sort(key=lambda c: c[0])  # symbol table entries are ordered by name (strcmp)
        if len(children) > 2 * self.leaf_k:
            raise ValueError("h5lite.write_classic: group with more entries than one symbol node holds")
        # local heap: offset 0 = the empty name (first B-tree key), then the names, null-terminated, 8-byte padded
        seg = bytearray(b"\0" * 8)
        offs = []
        for nm, _ in children:
            offs.append(len(seg))
            seg.extend(_pad8(nm + b"\0"))
        heap_addr = self._place(b"")
        heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, heap_addr + 32) + bytes(seg)  # free list: none (1)
        self.buf.extend(heap)
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(children)))
        for (nm, addr), off in zip(children, offs):
            snod.extend(struct.pack("<QQII16x", off, addr, 0, 0))
        snod.extend(b"\0" * (8 + 2 * self.leaf_k * 40 - len(snod)))
        snod_addr = self._place(bytes(snod))
        K = self.INTERNAL_K
        tree_node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if children else 0, _UNDEF, _UNDEF))
        if children:
            tree_node.extend(struct.pack("<QQQ", 0, snod_addr, offs[-1]))
        tree_node.extend(b"\0" * (24 + (2 * K + 1) * 8 + 2 * K * 8 - len(tree_node)))
        btree_addr = self._place(bytes(tree_node))
        oh = self._place(_v1_object_header([_v1_message(0x11, struct.pack("<QQ", btree_addr, heap_addr))]))
        return oh, btree_addr, heap_addr

    def finish(self, root):
        oh, btree, heap = root
        self.buf.extend(b"\0" * (-len(self.buf) % 8))
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def _max_entries(tree):
    return max([len(tree)] + [_max_entries(v) for v in tree.values() if isinstance(v, dict)])


def write_classic(path, tree):
    """The same nested dict as `write`, in the classic layout (what h5py writes by default)."""
    w = _ClassicWriter(leaf_k=max(4, (_max_entries(tree) + 1) // 2))
    blob = w.finish(w.group(tree))
    with open(path, "wb") as fh:
        fh.write(blob)
