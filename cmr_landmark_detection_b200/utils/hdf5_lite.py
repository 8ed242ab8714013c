"""Minimal pure-Python HDF5 reader / writer for Keras weight files (`model.h5`).

Why: the reference saves and loads weights with `model.save_weights('model.h5')` / `model.load_weights(...)`
(src/models/predict_model.py:76, src/utils/KerasCallbacks.py:54-61), i.e. through h5py, which is not installed in this
image.  This module restates the subset of the HDF5 file format (HDF5 File Format Specification, version 1.1 / 2.0 of the
format document) that h5py's default settings (`libver='earliest'`) produce for such files and nothing else:

  * superblock version 0 or 1 (optionally behind a user block), 8-byte offsets / lengths;
  * "old-style" groups: symbol-table message -> version-1 B-tree of group nodes -> symbol nodes (SNOD) + local heap;
  * version-1 object headers with continuation blocks;
  * messages: dataspace (v1 / v2), datatype (fixed point, IEEE float, fixed-length string, variable-length string),
    data layout (v1-v3: compact / contiguous, chunked without filters), attribute (v1-v3), symbol table;
  * global heap collections (variable-length strings).

The READER is checked against a file written by the HDF5 library itself (tests/test_hdf5_lite.py: scipy ships a MATLAB
v7.3 file, which is an HDF5 file of exactly this flavour).  The WRITER emits the same structures and is checked by reading
its files back with the reader; no HDF5 library is available here to read them independently, and INTEGRATION.md says so.
Unsupported features (new-style groups of superblock 2/3, filters / compression, other datatypes) raise Hdf5Error."""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(Exception):
    pass


# ======================================================================================================= reader
class _Type:
    def __init__(self, kind: str, size: int, dtype: Optional[np.dtype] = None, base: Optional['_Type'] = None):
        self.kind, self.size, self.dtype, self.base = kind, size, dtype, base


class Node:
    """A group or a dataset of an opened file."""

    def __init__(self, f: 'File', addr: int, name: str):
        self._f, self._addr, self.name = f, addr, name
        self._msgs = f._read_object_header(addr)
        self.attrs: Dict[str, Union[np.ndarray, bytes, str, float, int]] = {}
        self.unreadable_attrs: Dict[str, str] = {}       # attribute name (or '?') -> why it could not be decoded
        for t, body in self._msgs:
            if t == 0x000C:
                try:
                    k, v = f._parse_attribute(body)
                    self.attrs[k] = v
                except (Hdf5Error, struct.error, IndexError, ValueError) as e:
                    # an exotic attribute must not make the weights unreadable; asking for it later says why
                    self.unreadable_attrs[f._attribute_name(body)] = str(e)
        self._children: Optional[Dict[str, int]] = None

    # ---- groups
    @property
    def is_group(self) -> bool:
        return any(t == 0x0011 for t, _ in self._msgs)

    def keys(self) -> List[str]:
        return list(self._links().keys())

    def _links(self) -> Dict[str, int]:
        if self._children is None:
            self._children = {}
            for t, body in self._msgs:
                if t == 0x0011:
                    btree, heap = struct.unpack_from('<QQ', body, 0)
                    self._children = self._f._read_group(btree, heap)
                elif t in (0x0002, 0x0006):
                    raise Hdf5Error('new-style (link message) groups are not supported')
        return self._children

    def __contains__(self, key: str) -> bool:
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str) -> 'Node':
        node = self
        for part in [p for p in path.split('/') if p]:
            links = node._links()
            if part not in links:
                raise KeyError('%s has no member %r' % (node.name or '/', part))
            node = Node(self._f, links[part], (node.name.rstrip('/') + '/' + part))
        return node

    # ---- datasets
    @property
    def shape(self) -> Tuple[int, ...]:
        for t, body in self._msgs:
            if t == 0x0001:
                return self._f._parse_dataspace(body)
        raise Hdf5Error('%s is not a dataset' % self.name)

    def read(self) -> np.ndarray:
        shape, typ, layout = None, None, None
        for t, body in self._msgs:
            if t == 0x0001:
                shape = self._f._parse_dataspace(body)
            elif t == 0x0003:
                typ = self._f._parse_datatype(body)[0]
            elif t == 0x0008:
                layout = body
            elif t == 0x000B:
                raise Hdf5Error('%s: filtered (compressed) datasets are not supported' % self.name)
        if shape is None or typ is None or layout is None:
            raise Hdf5Error('%s is not a dataset' % self.name)
        n = int(np.prod(shape)) if shape else 1
        raw = self._f._read_layout(layout, n * typ.size, shape, typ.size)
        return self._f._decode(raw, typ, shape)


class File(Node):
    """Read-only view of an HDF5 file held in memory:  f = File(path); f.attrs[...]; f['a/b'].read()"""

    def __init__(self, path: str):
        with open(path, 'rb') as fh:
            self._buf = fh.read()
        b = self._buf
        off = 0
        while True:
            if b[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(b):
                raise Hdf5Error('%s: no HDF5 signature' % path)
        ver = b[off + 8]
        if ver not in (0, 1):
            raise Hdf5Error('superblock version %d is not supported (only the default "earliest" format, 0 / 1)' % ver)
        so, sl = b[off + 13], b[off + 14]
        if so != 8 or sl != 8:
            raise Hdf5Error('only 8-byte offsets / lengths are supported (file has %d / %d)' % (so, sl))
        p = off + 24 + (4 if ver == 1 else 0)
        base, _free, _eof, _drv = struct.unpack_from('<QQQQ', b, p)
        p += 32
        # libhdf5 keeps addresses relative to the base address; with a user block and base address 0 in the file they are
        # relative to the superblock; a MATLAB v7.3 file stores its user-block size (512) there
        self._base = base if base not in (0, UNDEF) else off
        _name_off, root_hdr, cache_type = struct.unpack_from('<QQI', b, p)
        self._so = 8
        Node.__init__(self, self, root_hdr, '/')

    # ---- low level
    def _at(self, addr: int) -> int:
        if addr == UNDEF:
            raise Hdf5Error('undefined address')
        return self._base + addr

    def _read_object_header(self, addr: int) -> List[Tuple[int, bytes]]:
        b, p = self._buf, self._at(addr)
        ver = b[p]
        if ver != 1:
            if b[p:p + 4] == b'OHDR':
                raise Hdf5Error('version-2 object headers are not supported')
            raise Hdf5Error('bad object header version %d at %d' % (ver, addr))
        nmsg, _ref, hsize = struct.unpack_from('<HII', b, p + 2)
        blocks = [(p + 16, hsize)]
        out: List[Tuple[int, bytes]] = []
        while blocks and len(out) < nmsg + 64:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end:
                mtype, msize, _flags = struct.unpack_from('<HHB', b, q)
                body = b[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x0010:
                    coff, clen = struct.unpack_from('<QQ', body, 0)
                    blocks.append((self._at(coff), clen))
                elif mtype != 0:
                    out.append((mtype, body))
        return out

    def _heap_string(self, heap_addr: int, offset: int) -> str:
        b, p = self._buf, self._at(heap_addr)
        if b[p:p + 4] != b'HEAP':
            raise Hdf5Error('bad local heap signature')
        _size, _free, data = struct.unpack_from('<QQQ', b, p + 8)
        q = self._at(data) + offset
        e = b.index(b'\x00', q)
        return b[q:e].decode('utf-8')

    def _read_group(self, btree: int, heap: int) -> Dict[str, int]:
        out: Dict[str, int] = {}

        def walk(addr: int):
            b, p = self._buf, self._at(addr)
            if b[p:p + 4] != b'TREE':
                raise Hdf5Error('bad B-tree signature')
            ntype, level, used = struct.unpack_from('<BBH', b, p + 4)
            if ntype != 0:
                raise Hdf5Error('not a group B-tree')
            q = p + 8 + 16
            for i in range(used):
                child = struct.unpack_from('<Q', b, q + 8)[0]      # key i (8) then child i (8)
                q += 16
                if level > 0:
                    walk(child)
                else:
                    s = self._at(child)
                    if b[s:s + 4] != b'SNOD':
                        raise Hdf5Error('bad symbol node signature')
                    nsym = struct.unpack_from('<H', b, s + 6)[0]
                    for k in range(nsym):
                        name_off, hdr = struct.unpack_from('<QQ', b, s + 8 + 40 * k)
                        out[self._heap_string(heap, name_off)] = hdr

        if btree != UNDEF:
            walk(btree)
        return out

    def _parse_dataspace(self, body: bytes) -> Tuple[int, ...]:
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:
                return (0,)
            p = 4
        else:
            raise Hdf5Error('dataspace version %d' % ver)
        return tuple(struct.unpack_from('<%dQ' % rank, body, p)) if rank else ()

    def _parse_datatype(self, body: bytes) -> Tuple[_Type, int]:
        cv = body[0]
        cls, ver = cv & 0x0F, cv >> 4
        bits = body[1] | (body[2] << 8) | (body[3] << 16)
        size = struct.unpack_from('<I', body, 4)[0]
        if cls == 0:
            if bits & 1:
                raise Hdf5Error('big-endian integers are not supported')
            return _Type('int', size, np.dtype('<%s%d' % ('i' if bits & 8 else 'u', size))), 8 + 4
        if cls == 1:
            if bits & 1:
                raise Hdf5Error('big-endian floats are not supported')
            return _Type('float', size, np.dtype('<f%d' % size)), 8 + 12
        if cls == 3:
            return _Type('string', size), 8
        if cls == 9:
            base, used = self._parse_datatype(body[8:])
            return _Type('vlen_string' if (bits & 0x0F) == 1 else 'vlen', size, base=base), 8 + used
        raise Hdf5Error('datatype class %d is not supported' % cls)

    def _read_layout(self, body: bytes, nbytes: int, shape, esize: int) -> bytes:
        b = self._buf
        ver = body[0]
        if ver == 3:
            cls = body[1]
            if cls == 0:
                size = struct.unpack_from('<H', body, 2)[0]
                return body[4:4 + size][:nbytes]
            if cls == 1:
                addr, size = struct.unpack_from('<QQ', body, 2)
                if addr == UNDEF:
                    return b'\x00' * nbytes          # never written: fill value 0
                p = self._at(addr)
                return b[p:p + nbytes]
            if cls == 2:
                ndim = body[2]
                addr = struct.unpack_from('<Q', body, 3)[0]
                cdims = struct.unpack_from('<%dI' % ndim, body, 11)
                return self._read_chunked(addr, cdims, shape, esize)
            raise Hdf5Error('layout class %d' % cls)
        if ver in (1, 2):
            ndim, cls = body[1], body[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from('<Q', body, p)[0]
                p += 8
            dims = struct.unpack_from('<%dI' % ndim, body, p)
            p += 4 * ndim
            if cls == 0:
                size = struct.unpack_from('<I', body, p)[0]
                return body[p + 4:p + 4 + size][:nbytes]
            if cls == 1:
                q = self._at(addr)
                return b[q:q + nbytes]
            return self._read_chunked(addr, dims, shape, esize)
        raise Hdf5Error('data layout version %d' % ver)

    def _read_chunked(self, btree: int, cdims, shape, esize: int) -> bytes:
        """Unfiltered chunked storage: version-1 B-tree of raw-data chunks (node type 1)."""
        rank = len(shape)
        chunk = tuple(int(c) for c in cdims[:rank])
        out = np.zeros(shape, dtype=np.dtype('V%d' % esize))
        b = self._buf

        def walk(addr: int):
            p = self._at(addr)
            if b[p:p + 4] != b'TREE':
                raise Hdf5Error('bad chunk B-tree signature')
            ntype, level, used = struct.unpack_from('<BBH', b, p + 4)
            if ntype != 1:
                raise Hdf5Error('not a chunk B-tree')
            ksize = 8 + 8 * (rank + 1)
            q = p + 24
            for i in range(used):
                csize, fmask = struct.unpack_from('<II', b, q)
                offs = struct.unpack_from('<%dQ' % (rank + 1), b, q + 8)
                child = struct.unpack_from('<Q', b, q + ksize)[0]
                q += ksize + 8
                if level > 0:
                    walk(child)
                    continue
                if fmask:
                    raise Hdf5Error('filtered chunks are not supported')
                c = self._at(child)
                data = np.frombuffer(b, dtype=np.dtype('V%d' % esize), count=int(np.prod(chunk)), offset=c).reshape(chunk)
                sl_out = tuple(slice(o, min(o + cs, s)) for o, cs, s in zip(offs[:rank], chunk, shape))
                sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                out[sl_out] = data[sl_in]

        if btree != UNDEF:
            walk(btree)
        return out.tobytes()

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        b, p = self._buf, self._at(addr)
        if b[p:p + 4] != b'GCOL':
            raise Hdf5Error('bad global heap signature')
        csize = struct.unpack_from('<Q', b, p + 8)[0]
        q, end = p + 16, p + csize
        while q + 16 <= end:
            idx, _ref, _res, size = struct.unpack_from('<HHIQ', b, q)
            if idx == 0:
                break
            if idx == index:
                return b[q + 16:q + 16 + size]
            q += 16 + ((size + 7) & ~7)
        raise Hdf5Error('global heap object %d not found' % index)

    def _decode(self, raw: bytes, typ: _Type, shape):
        n = int(np.prod(shape)) if shape else 1
        if typ.kind in ('int', 'float'):
            a = np.frombuffer(raw, dtype=typ.dtype, count=n).reshape(shape).copy()
            return a
        if typ.kind == 'string':
            a = np.frombuffer(raw, dtype='S%d' % typ.size, count=n).reshape(shape).copy()
            return a
        if typ.kind == 'vlen_string':
            vals = []
            for i in range(n):
                length, addr, idx = struct.unpack_from('<IQI', raw, 16 * i)
                vals.append(self._global_heap_object(addr, idx)[:length] if length else b'')
            a = np.array(vals, dtype=object).reshape(shape)
            return a
        raise Hdf5Error('cannot decode datatype %s' % typ.kind)

    @staticmethod
    def _attribute_name(body: bytes) -> str:
        try:
            ver = body[0]
            name_size = struct.unpack_from('<H', body, 2)[0]
            p = 9 if ver == 3 else 8
            return body[p:p + name_size].split(b'\x00')[0].decode('utf-8', 'replace')
        except Exception:
            return '?'

    def _parse_attribute(self, body: bytes):
        ver = body[0]
        name_size, type_size, space_size = struct.unpack_from('<HHH', body, 2)
        if ver == 1:
            p = 8
            pad = lambda n: (n + 7) & ~7
        elif ver == 2:
            p = 8
            pad = lambda n: n
        elif ver == 3:
            p = 9
            pad = lambda n: n
        else:
            raise Hdf5Error('attribute message version %d' % ver)
        name = body[p:p + name_size].split(b'\x00')[0].decode('utf-8')
        p += pad(name_size)
        typ = self._parse_datatype(body[p:p + type_size])[0]
        p += pad(type_size)
        shape = self._parse_dataspace(body[p:p + space_size])
        p += pad(space_size)
        n = int(np.prod(shape)) if shape else 1
        val = self._decode(body[p:p + n * typ.size], typ, shape) if n else np.zeros(shape)
        if shape == ():
            val = val.reshape(-1)[0]
        return name, val


# ======================================================================================================= writer
def _pad8(b: bytes) -> bytes:
    return b + b'\x00' * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    # version 1: version, rank, flags, 5 reserved bytes, dimension sizes
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', int(d)) for d in shape)


def _datatype_f32() -> bytes:
    # class 1 (floating point), version 1; little endian, pad 0, mantissa normalisation 2 (implied msb), sign bit 31
    bits = 0x20 | (31 << 8)
    head = struct.pack('<B3BI', 0x11, bits & 0xFF, (bits >> 8) & 0xFF, (bits >> 16) & 0xFF, 4)
    # bit offset 0, precision 32, exponent location 23, size 8, mantissa location 0, size 23, bias 127
    return head + struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)


def _datatype_f64() -> bytes:
    bits = 0x20 | (63 << 8)
    head = struct.pack('<B3BI', 0x11, bits & 0xFF, (bits >> 8) & 0xFF, (bits >> 16) & 0xFF, 8)
    return head + struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)


def _datatype_string(n: int) -> bytes:
    # class 3 (string), version 1; null-padded (1), ASCII (0)
    return struct.pack('<B3BI', 0x13, 0x01, 0, 0, max(n, 1))


def _attribute(name: str, value) -> bytes:
    """Attribute message, version 1 (name / datatype / dataspace each padded to 8 bytes)."""
    if isinstance(value, (bytes, str)):
        raw = value.encode('utf-8') if isinstance(value, str) else value
        typ, space, data = _datatype_string(len(raw)), _dataspace(()), raw if raw else b'\x00'
    else:
        arr = np.asarray(value)
        if arr.dtype.kind in ('S', 'U', 'O'):
            items = [(v.encode('utf-8') if isinstance(v, str) else bytes(v)) for v in arr.reshape(-1).tolist()]
            width = max([len(v) for v in items] + [1])
            typ, space = _datatype_string(width), _dataspace(arr.shape)
            data = b''.join(v.ljust(width, b'\x00') for v in items)
        elif arr.size == 0:
            typ, space, data = _datatype_f64(), _dataspace(arr.shape if arr.ndim else (0,)), b''
        else:
            arr = np.ascontiguousarray(arr, dtype='<f4')
            typ, space, data = _datatype_f32(), _dataspace(arr.shape), arr.tobytes()
    nm = name.encode('utf-8') + b'\x00'
    body = struct.pack('<BxHHH', 1, len(nm), len(typ), len(space)) + _pad8(nm) + _pad8(typ) + _pad8(space) + data
    if len(body) > 0xFFF0:
        raise Hdf5Error('attribute %r is too large for one header message (%d bytes)' % (name, len(body)))
    return _msg(0x000C, body)


class _Group:
    def __init__(self):
        self.attrs: List[Tuple[str, object]] = []
        self.children: Dict[str, Union['_Group', np.ndarray]] = {}


class Writer:
    """w = Writer(); w.attr('/', 'backend', b'tensorflow'); w.dataset('/a/a/kernel:0', array); w.save(path)
    Intermediate groups are created on demand, as h5py's create_dataset does for names that contain '/'."""

    def __init__(self):
        self.root = _Group()

    def _group(self, path: str, create: bool = True) -> _Group:
        g = self.root
        for part in [p for p in path.split('/') if p]:
            if part not in g.children:
                if not create:
                    raise KeyError(path)
                g.children[part] = _Group()
            g = g.children[part]
            if not isinstance(g, _Group):
                raise Hdf5Error('%s is a dataset' % path)
        return g

    def group(self, path: str):
        self._group(path)

    def attr(self, path: str, name: str, value):
        self._group(path).attrs.append((name, value))

    def dataset(self, path: str, array: np.ndarray):
        parent, _, leaf = path.rstrip('/').rpartition('/')
        self._group(parent).children[leaf] = np.ascontiguousarray(array, dtype='<f4')

    def save(self, path: str):
        buf = bytearray()
        # the library's default group B-tree geometry: symbol nodes of up to 2 * 4 entries, tree nodes of up to 2 * 16 children
        leaf_k, internal_k = 4, 16

        def alloc(data: bytes) -> int:
            buf.extend(b'\x00' * (-len(buf) % 8))
            addr = len(buf)
            buf.extend(data)
            return addr

        buf.extend(b'\x00' * 96)          # superblock v0 (56 bytes) + root symbol table entry (40 bytes), patched at the end

        def header(messages: List[bytes]) -> int:
            body = b''.join(messages)
            hdr = struct.pack('<BxHII4x', 1, len(messages), 1, len(body))
            return alloc(hdr + body)

        def write_dataset(arr: np.ndarray) -> int:
            data_addr = alloc(arr.tobytes()) if arr.size else UNDEF
            # message order and flags as the HDF5 library writes them (checked against a library-written file):
            # fill value (v1: late allocation, written if set, default value of size 0), datatype, dataspace, layout
            msgs = [_msg(0x0005, struct.pack('<BBBBI', 1, 2, 2, 1, 0), flags=1),
                    _msg(0x0003, _datatype_f32(), flags=1),
                    _msg(0x0001, _dataspace(arr.shape)),
                    _msg(0x0008, struct.pack('<BBQQ', 3, 1, data_addr, arr.nbytes))]      # layout v3, contiguous
            return header(msgs)

        def write_group(g: _Group) -> Tuple[int, int, int]:
            """-> (object header address, B-tree address, local heap address)"""
            entries = []
            for name in g.children:
                c = g.children[name]
                if isinstance(c, _Group):
                    hdr, bt, hp = write_group(c)
                    entries.append((name, hdr, 1, bt, hp))
                else:
                    entries.append((name, write_dataset(c), 0, 0, 0))
            entries.sort(key=lambda e: e[0].encode('utf-8'))
            # local heap data segment: offset 0 holds the empty string (the B-tree's first key), names follow, 8-byte aligned
            seg = bytearray(b'\x00' * 8)
            offs = {}
            for name, *_ in entries:
                offs[name] = len(seg)
                seg.extend(_pad8(name.encode('utf-8') + b'\x00'))
            free_off = len(seg)
            seg.extend(struct.pack('<QQ', 1, 16))          # one free block: next = 1 (last), size 16 (minimum heap free block)
            seg_addr = alloc(bytes(seg))
            heap_addr = alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(seg), free_off, seg_addr))
            def tree_node(level: int, keys: List[int], children: List[int], left: int, right: int) -> bytes:
                node = bytearray(b'TREE' + struct.pack('<BBHQQ', 0, level, len(children), left, right))
                for k, c in zip(keys, children):
                    node.extend(struct.pack('<QQ', k, c))
                if children:
                    node.extend(struct.pack('<Q', keys[-1]))
                node.extend(b'\x00' * (24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8 - len(node)))
                return bytes(node)

            # level 0: symbol nodes of <= 2 * leaf_k entries (sorted by name); a key is the heap offset of the largest name
            # of the child to its left, the very first key the empty string at heap offset 0
            nodes: List[Tuple[int, int]] = []            # (address, heap offset of the largest name below)
            for i in range(0, len(entries), 2 * leaf_k):
                part = entries[i:i + 2 * leaf_k]
                snod = bytearray(b'SNOD' + struct.pack('<BxH', 1, len(part)))
                for name, hdr, ctype, bt, hp in part:
                    scratch = struct.pack('<QQ', bt, hp) if ctype == 1 else b'\x00' * 16
                    snod.extend(struct.pack('<QQI4x', offs[name], hdr, ctype) + scratch)
                snod.extend(b'\x00' * (8 + 2 * leaf_k * 40 - len(snod)))
                nodes.append((alloc(bytes(snod)), offs[part[-1][0]]))
            level = 0
            while True:
                groups = [nodes[i:i + 2 * internal_k] for i in range(0, len(nodes), 2 * internal_k)] or [[]]
                # reserve the addresses first: siblings of one level point at each other
                size = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
                buf.extend(b'\x00' * (-len(buf) % 8))
                addrs = [len(buf) + k * size for k in range(len(groups))]
                first_key = 0
                parents = []
                for k, grp in enumerate(groups):
                    keys = [first_key] + [mx for _, mx in grp]
                    node = tree_node(level, keys, [a for a, _ in grp],
                                     addrs[k - 1] if k > 0 else UNDEF, addrs[k + 1] if k + 1 < len(groups) else UNDEF)
                    assert alloc(node) == addrs[k]
                    if grp:
                        first_key = grp[-1][1]
                        parents.append((addrs[k], grp[-1][1]))
                if len(groups) == 1:
                    tree_addr = addrs[0]
                    break
                nodes, level = parents, level + 1
            msgs = [_msg(0x0011, struct.pack('<QQ', tree_addr, heap_addr), flags=1)] + [_attribute(k, v) for k, v in g.attrs]
            return header(msgs), tree_addr, heap_addr

        root_hdr, root_tree, root_heap = write_group(self.root)
        buf.extend(b'\x00' * (-len(buf) % 8))
        sb = SIGNATURE + struct.pack('<BBBxBBBxHHI', 0, 0, 0, 0, 8, 8, leaf_k, internal_k, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(buf), UNDEF)
        sb += struct.pack('<QQI4xQQ', 0, root_hdr, 1, root_tree, root_heap)
        assert len(sb) == 96
        buf[0:96] = sb
        with open(path, 'wb') as fh:
            fh.write(bytes(buf))


# ======================================================================================================= Keras layout
def save_keras_weights(path: str, layers: List[Tuple[str, List[Tuple[str, np.ndarray]]]], backend: str = 'tensorflow',
                       keras_version: str = '2.4.0'):
    """tf.keras `save_weights(<path>.h5)` layout (hdf5_format.save_weights_to_hdf5_group): root attributes layer_names /
    backend / keras_version; one group per layer with attribute weight_names and datasets <layer>/<weight name>."""
    w = Writer()
    w.attr('/', 'layer_names', np.array([n.encode('utf-8') for n, _ in layers]))
    w.attr('/', 'backend', backend.encode('utf-8'))
    w.attr('/', 'keras_version', keras_version.encode('utf-8'))
    for lname, weights in layers:
        w.group('/' + lname)
        w.attr('/' + lname, 'weight_names', np.array([wn.encode('utf-8') for wn, _ in weights]) if weights else np.zeros((0,)))
        for wn, arr in weights:
            w.dataset('/%s/%s' % (lname, wn), arr)
    w.save(path)


def load_keras_weights(path: str) -> List[Tuple[str, List[Tuple[str, np.ndarray]]]]:
    """Reads a tf.keras weight file the way hdf5_format.load_weights_from_hdf5_group does: layers in `layer_names`
    order, weights in `weight_names` order; layers without weights are dropped.  Accepts full-model files too (weights
    under /model_weights)."""
    f = File(path)
    g: Node = f
    if 'layer_names' not in g.attrs and 'model_weights' in g:
        g = g['model_weights']
    if 'layer_names' not in g.attrs:
        why = g.unreadable_attrs.get('layer_names')
        raise Hdf5Error('%s is not a Keras weight file (%s)' % (path, 'layer_names: ' + why if why else 'no layer_names attribute'))

    def names(v) -> List[str]:
        out = []
        for x in np.asarray(v).reshape(-1).tolist():
            out.append(x.decode('utf-8') if isinstance(x, bytes) else str(x))
        return out

    out = []
    for lname in names(g.attrs['layer_names']):
        lg = g[lname]
        wn = lg.attrs.get('weight_names')
        if wn is None or np.asarray(wn).size == 0:
            continue
        out.append((lname, [(n, lg[n].read()) for n in names(wn)]))
    return out
