"""Minimal pure-Python/numpy reader (and test-support writer) for the HDF5 subset Keras weight files use.

Why: the reference loads its generators with ``Model.load_weights('inference.hdf5')`` (recognition.py:23-26; the files
are written by tools/4_convert_weights_inference.py:50-52 / tools/3_train_pix2pose.py:273-276 through h5py).  h5py is
not part of this image and nothing may be installed, so the drop-in reads the container format itself.  Supported --
what libhdf5 1.8/1.10 writes with its default ("earliest") file-format settings, which is what h5py 2.x/3.x produce for
Keras: superblock v0/v1, version-1 object headers (with continuation blocks), old-style groups (symbol-table message,
v1 B-tree, SNOD nodes, local heap), dataspace v1/v2, little-endian fixed/float datatypes, fixed-length and
variable-length (global heap) strings, data layouts v1-v3 compact / contiguous / chunked WITHOUT filters, attribute
messages v1-v3.  Anything else (new-style groups with fractal heaps, compression filters, superblock v2/v3 files
written with libver='latest') raises ``Hdf5Error`` naming the unsupported feature.

The format description followed is the public "HDF5 File Format Specification Version 2.0" (sections III.A-III.E,
IV.A.1-IV.A.2); the writer below produces the same subset and exists so that the reader and the Keras import can be
tested here (tests/test_hdf5_lite.py) -- this pin is self-consistent only: no file written by libhdf5 itself is
available in this sandbox (stated in DESIGN.md section 7).
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ======================================================================================================================
# reader
class _Datatype:
    def __init__(self, cls, size, dtype=None, vlen_string=False, base=None):
        self.cls, self.size, self.dtype, self.vlen_string, self.base = cls, size, dtype, vlen_string, base


class Dataset:
    def __init__(self, f, name, shape, dt, layout):
        self._f, self.name, self.shape, self._dt, self._layout = f, name, tuple(shape), dt, layout
        self.attrs = {}

    @property
    def dtype(self):
        return self._dt.dtype

    def __array__(self, dtype=None, copy=None):
        a = self[()]
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, key):
        return self._f._read_dataset(self)[key]


class Group:
    def __init__(self, f, name):
        self._f, self.name = f, name
        self._links = {}      # name -> object header address, in B-tree (= name-sorted) order
        self.attrs = {}

    def keys(self):
        return list(self._links.keys())

    def __contains__(self, k):
        try:
            self[k]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError("%s: no member %r" % (node.name, part))
            node = node._f._object(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node


class File(Group):
    """``File(path)`` -- read-only; mimics the slice of the h5py API Keras' weight loading uses
    (``f.attrs['layer_names']``, ``f[layer].attrs['weight_names']``, ``np.asarray(f[layer][weight])``)."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self._buf = fh.read()
        Group.__init__(self, self, "/")
        self._cache = {}
        b = self._buf
        # the superblock sits at offset 0 or, after a user block, at 512, 1024, 2048, ... (HDF5 spec II.A)
        sb = 0
        while not (len(b) >= sb + 96 and b[sb:sb + 8] == SIGNATURE):
            sb = 512 if sb == 0 else sb * 2
            if sb + 96 > len(b):
                raise Hdf5Error("%s: not an HDF5 file (no superblock signature at offset 0, 512, 1024, ...)" % path)
        b = b[sb:]
        ver = b[8]
        if ver in (0, 1):
            so, sl = b[13], b[14]
            if so != 8 or sl != 8:
                raise Hdf5Error("size of offsets/lengths %d/%d not supported (only 8/8)" % (so, sl))
            base_off = 24 if ver == 0 else 28
            self._base = struct.unpack_from("<Q", b, base_off)[0]
            root_entry = base_off + 32
            _, ohdr, cache, _ = struct.unpack_from("<QQII", b, root_entry)
            root = self._object(ohdr, "/")
        elif ver in (2, 3):
            if b[9] != 8 or b[10] != 8:
                raise Hdf5Error("size of offsets/lengths not 8/8")
            self._base = struct.unpack_from("<Q", b, 12)[0]
            ohdr = struct.unpack_from("<Q", b, 36)[0]
            root = self._object(ohdr, "/")
        else:
            raise Hdf5Error("superblock version %d not supported" % ver)
        if not isinstance(root, Group):
            raise Hdf5Error("root object is not a group")
        self._links, self.attrs = root._links, root.attrs

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- low level ------------------------------------------------------------------------------------------------
    def _u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self._buf, off)

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._header_messages(addr)
        attrs, shape, dt, layout, symtab = {}, None, None, None, None
        for mtype, data in msgs:
            if mtype == 0x0011:
                symtab = struct.unpack_from("<QQ", data, 0)
            elif mtype == 0x0001:
                shape = self._dataspace(data)
            elif mtype == 0x0003:
                dt = self._datatype(data, 0)[0]
            elif mtype == 0x0008:
                layout = self._layout(data)
            elif mtype == 0x000B:
                raise Hdf5Error("%s: filter pipeline (compression) is not supported" % name)
            elif mtype == 0x000C:
                k, v = self._attribute(data)
                attrs[k] = v
            elif mtype in (0x0002, 0x0006):
                raise Hdf5Error("%s: new-style group (link info / link messages) is not supported" % name)
        if symtab is not None:
            obj = Group(self, name)
            self._btree_group(symtab[0], symtab[1], obj._links)
        elif layout is not None and dt is not None and shape is not None:
            obj = Dataset(self, name, shape, dt, layout)
        else:
            raise Hdf5Error("%s: object is neither an old-style group nor a dataset" % name)
        obj.attrs = attrs
        self._cache[addr] = obj
        return obj

    def _header_messages(self, addr):
        b = self._buf
        addr += self._base
        if b[addr:addr + 4] == b"OHDR":
            raise Hdf5Error("version-2 object headers (libver='latest') are not supported")
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise Hdf5Error("object header version %d at %d not supported" % (ver, addr))
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, off)
                data = b[off + 8: off + 8 + msize]
                off += 8 + msize
                if mtype == 0x0010:
                    c_off, c_len = struct.unpack_from("<QQ", data, 0)
                    blocks.append((c_off + self._base, c_len))
                out.append((mtype, data))
        return out

    def _btree_group(self, btree, heap, links):
        b = self._buf
        hoff = heap + self._base
        if b[hoff:hoff + 4] != b"HEAP":
            raise Hdf5Error("local heap signature missing at %d" % hoff)
        data_seg = struct.unpack_from("<Q", b, hoff + 24)[0] + self._base

        def name_at(o):
            e = b.index(b"\x00", data_seg + o)
            return b[data_seg + o:e].decode("utf-8")

        def walk(node):
            o = node + self._base
            if b[o:o + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, o + 6)[0]
                for i in range(n):
                    noff, ohdr = struct.unpack_from("<QQ", b, o + 8 + 40 * i)
                    links[name_at(noff)] = ohdr
                return
            if b[o:o + 4] != b"TREE":
                raise Hdf5Error("B-tree node signature missing at %d" % o)
            ntype, _level, used = struct.unpack_from("<BBH", b, o + 4)
            if ntype != 0:
                raise Hdf5Error("group B-tree node of type %d" % ntype)
            p = o + 24
            for i in range(used):
                child = struct.unpack_from("<Q", b, p + 8 + 16 * i)[0]
                walk(child)

        walk(btree)

    @staticmethod
    def _dataspace(d):
        ver, rank = d[0], d[1]
        if ver == 1:
            off = 8
        elif ver == 2:
            if d[3] == 2:
                return None       # null dataspace
            off = 4
        else:
            raise Hdf5Error("dataspace version %d not supported" % ver)
        return struct.unpack_from("<%dQ" % rank, d, off) if rank else ()

    def _datatype(self, d, off):
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", d, off)
        cls = cv & 0x0F
        if cls in (0, 1):
            if b0 & 1:
                raise Hdf5Error("big-endian numeric data is not supported")
            if cls == 1:
                np_dt, plen = {2: "<f2", 4: "<f4", 8: "<f8"}.get(size), 12
            else:
                np_dt, plen = ("<i%d" if (b0 & 8) else "<u%d") % size, 4
            if np_dt is None:
                raise Hdf5Error("float of %d bytes not supported" % size)
            return _Datatype(cls, size, np.dtype(np_dt)), off + 8 + plen
        if cls == 3:
            return _Datatype(cls, size, np.dtype("S%d" % size)), off + 8
        if cls == 9:
            base, end = self._datatype(d, off + 8)
            return _Datatype(cls, size, np.dtype("O"), vlen_string=(b0 & 0x0F) == 1, base=base), end
        raise Hdf5Error("datatype class %d not supported" % cls)

    def _layout(self, d):
        ver = d[0]
        if ver == 3:
            cls = d[1]
            if cls == 0:
                n = struct.unpack_from("<H", d, 2)[0]
                return ("compact", d[4:4 + n])
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", d, 2)
                return ("contiguous", addr, size)
            if cls == 2:
                nd = d[2]
                addr = struct.unpack_from("<Q", d, 3)[0]
                dims = struct.unpack_from("<%dI" % nd, d, 11)
                return ("chunked", addr, dims)
            raise Hdf5Error("data layout class %d not supported" % cls)
        if ver in (1, 2):
            nd, cls = d[1], d[2]
            off = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", d, off)[0]
                off += 8
            dims = struct.unpack_from("<%dI" % nd, d, off)
            off += 4 * nd
            if cls == 0:
                n = struct.unpack_from("<I", d, off)[0]
                return ("compact", d[off + 4: off + 4 + n])
            if cls == 1:
                return ("contiguous", addr, None)
            return ("chunked", addr, dims)
        raise Hdf5Error("data layout version %d not supported" % ver)

    def _vlen_elements(self, raw, n, dt):
        out = []
        for i in range(n):
            length, gaddr, idx = struct.unpack_from("<IQI", raw, 16 * i)
            if length == 0 or gaddr in (0, UNDEF):
                out.append(b"" if dt.vlen_string else np.zeros(0, dt.base.dtype))
                continue
            o = gaddr + self._base
            b = self._buf
            if b[o:o + 4] != b"GCOL":
                raise Hdf5Error("global heap collection signature missing at %d" % o)
            csize = struct.unpack_from("<Q", b, o + 8)[0]
            p, end, found = o + 16, o + csize, None
            while p + 16 <= end:
                oidx, _rc, _, osize = struct.unpack_from("<HHIQ", b, p)
                if oidx == 0:
                    break
                if oidx == idx:
                    found = b[p + 16:p + 16 + osize]
                    break
                p += 16 + _pad8(osize)
            if found is None:
                raise Hdf5Error("global heap object %d not found" % idx)
            out.append(found[:length] if dt.vlen_string else np.frombuffer(found, dt.base.dtype, length))
        return out

    def _decode(self, raw, shape, dt):
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            vals = self._vlen_elements(raw, n, dt)
            if dt.vlen_string:
                vals = [v.decode("utf-8") for v in vals]
            if not shape:
                return vals[0]
            a = np.empty(n, object)
            a[:] = vals
            return a.reshape(shape)
        a = np.frombuffer(raw, dt.dtype, n)
        return a.reshape(shape).copy() if shape else a[0]

    def _attribute(self, d):
        ver = d[0]
        nsz, dsz, ssz = struct.unpack_from("<HHH", d, 2)
        if ver == 1:
            off = 8
            name = d[off:off + nsz].split(b"\x00")[0].decode("utf-8")
            off += _pad8(nsz)
            dt, _ = self._datatype(d, off)
            off += _pad8(dsz)
            shape = self._dataspace(d[off:off + ssz])
            off += _pad8(ssz)
        elif ver in (2, 3):
            off = 8 if ver == 2 else 9
            name = d[off:off + nsz].split(b"\x00")[0].decode("utf-8")
            off += nsz
            dt, _ = self._datatype(d, off)
            off += dsz
            shape = self._dataspace(d[off:off + ssz])
            off += ssz
        else:
            raise Hdf5Error("attribute message version %d not supported" % ver)
        if shape is None:
            return name, None
        return name, self._decode(d[off:], shape, dt)

    def _read_dataset(self, ds):
        lay = ds._layout
        n = int(np.prod(ds.shape)) if ds.shape else 1
        nbytes = n * (16 if ds._dt.cls == 9 else ds._dt.size)
        if lay[0] == "compact":
            raw = lay[1]
        elif lay[0] == "contiguous":
            if lay[1] == UNDEF:
                raw = bytes(nbytes)       # never written: fill value 0
            else:
                raw = self._buf[lay[1] + self._base: lay[1] + self._base + nbytes]
        else:
            raw = self._read_chunked(ds, lay[1], lay[2])
        return self._decode(raw, ds.shape, ds._dt)

    def _read_chunked(self, ds, btree, cdims):
        if ds._dt.cls == 9:
            raise Hdf5Error("chunked variable-length datasets are not supported")
        rank = len(ds.shape)
        chunk = cdims[:rank]
        out = np.zeros(ds.shape, ds._dt.dtype)
        b = self._buf

        def walk(node):
            o = node + self._base
            if b[o:o + 4] != b"TREE":
                raise Hdf5Error("chunk B-tree node signature missing at %d" % o)
            ntype, level, used = struct.unpack_from("<BBH", b, o + 4)
            if ntype != 1:
                raise Hdf5Error("chunk B-tree node of type %d" % ntype)
            ksz = 8 + 8 * (rank + 1)
            p = o + 24
            for i in range(used):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<%dQ" % rank, b, p + 8)
                child = struct.unpack_from("<Q", b, p + ksz)[0]
                if level > 0:
                    walk(child)
                else:
                    if fmask:
                        raise Hdf5Error("filtered chunks are not supported")
                    blk = np.frombuffer(b, ds._dt.dtype, int(np.prod(chunk)), child + self._base).reshape(chunk)
                    sl = tuple(slice(o_, min(o_ + c, s)) for o_, c, s in zip(offs, chunk, ds.shape))
                    out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
                p += ksz + 8

        if btree != UNDEF:
            walk(btree)
        return out.tobytes()


# ======================================================================================================================
# writer (tests / export): superblock v0, version-1 object headers, one SNOD per group, contiguous datasets,
# version-1 attributes with numeric arrays and fixed-length string arrays
class _Node:
    def __init__(self):
        self.children = {}     # name -> _Node | np.ndarray
        self.attrs = {}


class Writer:
    """``w = Writer(); w.create_dataset('conv1/conv1/kernel:0', arr); w.set_attr('', 'layer_names', [b'conv1']); w.save(path)``"""

    def __init__(self):
        self.root = _Node()

    def _node(self, path, create=True):
        n = self.root
        for part in [p for p in path.split("/") if p]:
            if part not in n.children:
                if not create:
                    raise KeyError(path)
                n.children[part] = _Node()
            n = n.children[part]
        return n

    def create_group(self, path):
        return self._node(path)

    def create_dataset(self, path, data):
        parts = [p for p in path.split("/") if p]
        parent = self._node("/".join(parts[:-1]))
        ds = _Node()
        ds.data = np.ascontiguousarray(data)
        parent.children[parts[-1]] = ds
        return ds

    def set_attr(self, path, name, value):
        self._node(path, create=False).attrs[name] = value

    # ---- encoding ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _dt_msg(dtype):
        dtype = np.dtype(dtype)
        if dtype.kind == "f":
            exp, man, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[dtype.itemsize]
            bits = dtype.itemsize * 8
            return struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dtype.itemsize) + struct.pack("<HHBBBBI", 0, bits, man, exp, 0, man, bias)
        if dtype.kind in "iu":
            return struct.pack("<BBBBI", 0x10, 0x08 if dtype.kind == "i" else 0, 0, 0, dtype.itemsize) + struct.pack("<HH", 0, dtype.itemsize * 8)
        if dtype.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0, 0, 0, dtype.itemsize)
        raise Hdf5Error("writer: dtype %s not supported" % dtype)

    @staticmethod
    def _ds_msg(shape):
        return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)

    @classmethod
    def _attr_msg(cls, name, value):
        a = np.asarray(value)
        if a.dtype.kind == "U":
            a = np.char.encode(a, "utf-8")
        nm = name.encode("utf-8") + b"\x00"
        dt, ds = cls._dt_msg(a.dtype), cls._ds_msg(a.shape)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        body += nm.ljust(_pad8(len(nm)), b"\x00") + dt.ljust(_pad8(len(dt)), b"\x00") + ds.ljust(_pad8(len(ds)), b"\x00")
        body += np.ascontiguousarray(a).tobytes()
        return 0x000C, body

    @staticmethod
    def _msg_bytes(msgs):
        body = b""
        for mtype, data in msgs:
            data = data.ljust(_pad8(len(data)), b"\x00")
            body += struct.pack("<HHBBBB", mtype, len(data), 0, 0, 0, 0) + data
        return body

    @classmethod
    def _header(cls, msgs, alloc=None, split=False):
        """Version-1 object header.  ``split``: everything after the first message goes into a continuation block (what
        libhdf5 does when attributes are added to an object whose header chunk is full)."""
        if split and alloc is not None and len(msgs) > 1:
            tail = cls._msg_bytes(msgs[1:])
            caddr = alloc(tail)
            first = cls._msg_bytes([msgs[0], (0x0010, struct.pack("<QQ", caddr, len(tail)))])
            return struct.pack("<BBHII", 1, 0, len(msgs) + 1, 1, len(first)) + b"\x00" * 4 + first
        body = cls._msg_bytes(msgs)
        return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\x00" * 4 + body

    def save(self, path, leaf_k=4, internal_k=16, split_headers=True):
        """Writes the file.  Defaults mimic libhdf5: at most 2 * leaf_k symbols per SNOD, 2 * internal_k children per B-tree
        node (so groups with many members get multi-level B-trees) and attributes in object-header continuation blocks."""
        out = bytearray(96)           # superblock placeholder

        def alloc(blob):
            while len(out) % 8:
                out.append(0)
            addr = len(out)
            out.extend(blob)
            return addr

        def emit(node):
            msgs = [self._attr_msg(k, v) for k, v in node.attrs.items()]
            if hasattr(node, "data"):
                a = node.data
                daddr = alloc(a.tobytes()) if a.size else UNDEF
                msgs = [(0x0001, self._ds_msg(a.shape)), (0x0003, self._dt_msg(a.dtype)),
                        (0x0008, struct.pack("<BBQQ", 3, 1, daddr, a.nbytes))] + msgs
                return alloc(self._header(msgs))
            names = sorted(node.children)
            addrs = [emit(node.children[n]) for n in names]
            heap_data = bytearray(b"\x00" * 8)      # offset 0: the empty string
            offs = []
            for n in names:
                offs.append(len(heap_data))
                e = n.encode("utf-8") + b"\x00"
                heap_data.extend(e.ljust(_pad8(len(e)), b"\x00"))
            data_addr = alloc(bytes(heap_data))
            heap_addr = alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), UNDEF, data_addr))
            # leaves: symbol-table nodes of up to 2 * leaf_k entries, in name order
            cap = 2 * leaf_k
            level = []                                # (address, heap offset of the largest name below)
            for i in range(0, max(len(names), 1), cap):
                part = list(zip(offs[i:i + cap], addrs[i:i + cap]))
                snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
                for o, a in part:
                    snod += struct.pack("<QQII", o, a, 0, 0) + b"\x00" * 16
                snod += b"\x00" * (40 * (cap - len(part)))
                level.append((alloc(snod), part[-1][0] if part else 0))
            depth = 0
            while True:                               # B-tree levels until one node is left
                nodes, fan = [], 2 * internal_k
                for i in range(0, len(level), fan):
                    part = level[i:i + fan]
                    tree = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(part), UNDEF, UNDEF) + struct.pack("<Q", 0)
                    for addr, key in part:
                        tree += struct.pack("<QQ", addr, key)
                    nodes.append((alloc(tree), part[-1][1]))
                level, depth = nodes, depth + 1
                if len(level) == 1:
                    break
            tree_addr = level[0][0]
            node._symtab = (tree_addr, heap_addr)
            hdr = [(0x0011, struct.pack("<QQ", tree_addr, heap_addr))] + msgs
            return alloc(self._header(hdr, alloc, split_headers and len(msgs) > 0))

        root_addr = emit(self.root)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, internal_k, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(out), UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", *self.root._symtab)
        out[:96] = sb
        with open(path, "wb") as fh:
            fh.write(bytes(out))
