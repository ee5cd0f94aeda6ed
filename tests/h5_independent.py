"""An INDEPENDENT reader of the HDF5 subset the signal file uses, written from the HDF5 File Format Specification (version 1.1 /
2.0: superblock version 0, version-1 object headers, symbol-table groups = version-1 B-tree (node type 0) + local heap + symbol
table nodes, dataspace messages version 1 / 2, datatype message, data layout message version 3 with contiguous or chunked storage,
the chunk index a version-1 B-tree (node type 1)).  Pure Python / numpy, sharing no code with csrc/host/h5mini.cpp.

Test infrastructure (tests/test_h5_signal.py): libhdf5 does not exist in this image, so the product's writer cannot be read back
by the real library.  This reader is first checked against a file the real library wrote (scipy's MATLAB v7.3 test file) and
then reads what h5mini writes -- a second implementation of the specification that has to agree with the first one on every
byte it dereferences (signatures, B-tree keys and children, heap offsets, message sizes, chunk addresses)."""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 2**64 - 1


class H5Error(Exception):
    pass


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        off = 0
        while self.b[off:off + 8] != SIG:  # the superblock sits at 0, 512, 1024, ... (user block)
            off = 512 if off == 0 else off * 2
            if off + 8 > len(self.b):
                raise H5Error("no HDF5 signature")
        self.sb = off
        b = self.b
        ver = b[off + 8]
        if ver != 0:
            raise H5Error(f"superblock version {ver} not handled")
        self.so, self.sl = b[off + 13], b[off + 14]  # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets and lengths are handled")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, off + 16)
        base, free, self.eof, drv = struct.unpack_from("<4Q", b, off + 24)
        # addresses are relative to the base address; a file behind a user block stores base 0 and counts from the superblock
        self.base = base if base else off
        # root group symbol table entry
        self.root = self._ste(off + 56)

    # -- primitives
    def _abs(self, a):
        if a == UNDEF:
            raise H5Error("undefined address dereferenced")
        return self.base + a

    def _ste(self, p):
        name_off, hdr, cache = struct.unpack_from("<QQI", self.b, p)
        e = {"name_off": name_off, "header": hdr, "cache": cache}
        if cache == 1:
            e["btree"], e["heap"] = struct.unpack_from("<QQ", self.b, p + 24)
        return e

    def _messages(self, hdr_addr):
        """(type, flags, bytes) of every message of a version-1 object header, continuation blocks followed"""
        p = self._abs(hdr_addr)
        ver, _, nmsg, _ref, size = struct.unpack_from("<BBHII", self.b, p)
        if ver != 1:
            raise H5Error(f"object header version {ver} not handled")
        blocks = [(p + 16, size)]  # the 12-byte prefix is padded to 8-byte alignment
        out = []
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.b, q)
                data = self.b[q + 8:q + 8 + msize]
                out.append((mtype, flags, data))
                if mtype == 0x0010:  # continuation: offset, length
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self._abs(caddr), clen))
                q += 8 + msize
        return out

    def _heap_string(self, heap_addr, offset):
        p = self._abs(heap_addr)
        if self.b[p:p + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        _size, _free, data = struct.unpack_from("<QQQ", self.b, p + 8)
        s = self._abs(data) + offset
        e = self.b.index(b"\0", s)
        return self.b[s:e].decode()

    def _group_entries(self, btree_addr, heap_addr):
        """symbol table entries of a group, in B-tree order"""
        p = self._abs(btree_addr)
        if self.b[p:p + 4] != b"TREE":
            raise H5Error("B-tree signature missing")
        ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
        if ntype != 0:
            raise H5Error("group B-tree expected (node type 0)")
        q = p + 8 + 16  # past the sibling pointers
        out = []
        for i in range(used):
            child = struct.unpack_from("<Q", self.b, q + 8)[0]  # key_i (8), child_i (8)
            q += 16
            if level > 0:
                out += self._group_entries(child, heap_addr)
            else:
                s = self._abs(child)
                if self.b[s:s + 4] != b"SNOD":
                    raise H5Error("symbol table node signature missing")
                n = struct.unpack_from("<H", self.b, s + 6)[0]
                for k in range(n):
                    e = self._ste(s + 8 + 40 * k)
                    e["name"] = self._heap_string(heap_addr, e["name_off"])
                    out.append(e)
        return out

    # -- objects
    def listing(self, entry=None, prefix=""):
        """{path: object header address} of every dataset below the root"""
        entry = entry or self.root
        bt, heap = entry.get("btree"), entry.get("heap")
        if bt is None:
            for mtype, _, data in self._messages(entry["header"]):
                if mtype == 0x0011:
                    bt, heap = struct.unpack_from("<QQ", data, 0)
        if bt is None:
            return {prefix.rstrip("/"): entry["header"]}
        out = {}
        for e in self._group_entries(bt, heap):
            is_group = e["cache"] == 1 or any(m[0] == 0x0011 for m in self._messages(e["header"]))
            if is_group:
                out.update(self.listing(e, prefix + e["name"] + "/"))
            else:
                out[prefix + e["name"]] = e["header"]
        return out

    def _chunks(self, btree_addr, rank):
        """(offsets, address, nbytes) of every chunk, from the version-1 chunk B-tree (node type 1)"""
        if btree_addr == UNDEF:
            return []
        p = self._abs(btree_addr)
        if self.b[p:p + 4] != b"TREE":
            raise H5Error("chunk B-tree signature missing")
        ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
        if ntype != 1:
            raise H5Error("chunk B-tree expected (node type 1)")
        keysize = 8 + 8 * (rank + 1)
        q = p + 24
        out = []
        for i in range(used):
            nbytes, mask = struct.unpack_from("<II", self.b, q)
            offs = struct.unpack_from(f"<{rank + 1}Q", self.b, q + 8)
            child = struct.unpack_from("<Q", self.b, q + keysize)[0]
            q += keysize + 8
            if mask:
                raise H5Error("filtered chunks are not handled")
            if level > 0:
                out += self._chunks(child, rank)
            else:
                out.append((offs[:rank], child, nbytes))
        return out

    def dataset(self, hdr_addr, with_layout=False):
        dims = maxdims = dtype = layout = None
        for mtype, _, d in self._messages(hdr_addr):
            if mtype == 0x0001:  # dataspace
                ver, rank, flags = d[0], d[1], d[2]
                q = 8 if ver == 1 else 4
                dims = struct.unpack_from(f"<{rank}Q", d, q)
                if flags & 1:
                    maxdims = struct.unpack_from(f"<{rank}Q", d, q + 8 * rank)
            elif mtype == 0x0003:  # datatype
                cls, size = d[0] & 15, struct.unpack_from("<I", d, 4)[0]
                if cls == 1 and size == 8:
                    dtype = np.dtype("<f8")
                elif cls == 3:
                    dtype = np.dtype(f"S{size}")
                elif cls == 0:  # fixed point
                    signed = (d[1] >> 3) & 1
                    dtype = np.dtype(("<i" if signed else "<u") + str(size))
                else:
                    raise H5Error(f"datatype class {cls} size {size} not handled")
            elif mtype == 0x0008:  # data layout
                if d[0] in (1, 2):  # libhdf5 1.6: version, dimensionality, class, 5 reserved, address, 4-byte dimension sizes
                    nd, cls = d[1], d[2]
                    addr = struct.unpack_from("<Q", d, 8)[0]
                    sizes = struct.unpack_from(f"<{nd}I", d, 16)
                    if cls == 1:
                        layout = ("contiguous", addr, None)
                    elif cls == 2:
                        layout = ("chunked", addr, sizes + struct.unpack_from("<I", d, 16 + 4 * nd))
                    else:
                        raise H5Error("compact layout not handled")
                    continue
                if d[0] != 3:
                    raise H5Error(f"layout message version {d[0]} not handled")
                if d[1] == 1:
                    addr, size = struct.unpack_from("<QQ", d, 2)
                    layout = ("contiguous", addr, size)
                elif d[1] == 2:
                    nd = d[2]
                    bt = struct.unpack_from("<Q", d, 3)[0]
                    cd = struct.unpack_from(f"<{nd}I", d, 11)
                    layout = ("chunked", bt, cd)
                else:
                    raise H5Error("compact layout not handled")
        if dims is None or dtype is None or layout is None:
            raise H5Error("dataset without dataspace / datatype / layout message")
        n = int(np.prod(dims)) if dims else 1
        if layout[0] == "contiguous":
            raw = b"" if n == 0 or layout[1] == UNDEF else self.b[self._abs(layout[1]):self._abs(layout[1]) + n * dtype.itemsize]
            arr = np.frombuffer(raw, dtype=dtype).reshape(dims) if n else np.zeros(dims, dtype=dtype)
            chunk = None
        else:
            rank = len(dims)
            cd = layout[2]
            if cd[-1] != dtype.itemsize or len(cd) != rank + 1:
                raise H5Error("chunk dimensions do not end in the element size")
            chunk = tuple(cd[:rank])
            arr = np.zeros(dims, dtype=dtype)
            for offs, addr, nbytes in self._chunks(layout[1], rank):
                if nbytes != int(np.prod(chunk)) * dtype.itemsize:
                    raise H5Error("chunk size in the B-tree key differs from the layout message")
                c = np.frombuffer(self.b[self._abs(addr):self._abs(addr) + nbytes], dtype=dtype).reshape(chunk)
                sl = tuple(slice(o, min(o + k, m)) for o, k, m in zip(offs, chunk, dims))
                if any(s.start >= s.stop for s in sl):
                    continue  # a chunk beyond the current extent (after a shrink); none is expected here
                arr[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
        if with_layout:
            return arr, {"maxdims": maxdims, "chunk": chunk}
        return arr


def read(path, with_layout=False):
    f = File(path)
    out, lay = {}, {}
    for name, hdr in f.listing().items():
        r = f.dataset(hdr, with_layout=True)
        out[name], lay[name] = r
    return (out, lay) if with_layout else out
