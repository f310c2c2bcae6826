"""TEST INFRASTRUCTURE -- a minimal, independent HDF5 reader in pure Python (there is no h5py /
libhdf5 in this image).  It understands exactly the subset libhdf5 1.8 emits for the reference's
EMD files (src/rwHdf5.cu) and that fdes_b200/csrc/emd.cpp writes: superblock version 0, version-1
object headers (with continuation blocks), symbol-table groups (v1 B-tree + local heap + SNOD),
contiguous (and compact) datasets of fixed-point / IEEE float / fixed-length string types, version 1-3
attribute messages.  Its reading of a real libhdf5-written file
(ExampleSpecimens/Au_cubeoctahedron_emd/Auparticle.emd of the reference, when present) is what
anchors the writer test: the product's files must parse with the same code."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class Node:
    def __init__(self, name, addr):
        self.name, self.addr = name, addr
        self.attrs, self.children = {}, {}
        self.data = None          # numpy array for datasets
        self.is_group = False
        self.messages = []        # (type, offset, size) for structural checks

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            node = node.children[part]
        return node

    def walk(self, prefix=""):
        yield prefix or "/", self
        for k, c in self.children.items():
            yield from c.walk(prefix + "/" + k)


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        self.sb_version = b[8]
        if self.sb_version != 0:
            raise H5Error(f"superblock version {self.sb_version} not supported")
        self.size_offsets, self.size_lengths = b[13], b[14]
        if (self.size_offsets, self.size_lengths) != (8, 8):
            raise H5Error("only 8-byte offsets/lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        self.base, self.free, self.eof, self.driver = struct.unpack_from("<QQQQ", b, 24)
        if self.eof != len(b):
            raise H5Error(f"end-of-file address {self.eof} != file size {len(b)}")
        name_off, ohdr, cache, _ = struct.unpack_from("<QQII", b, 56)
        self.root = self._object("/", ohdr)

    # ---- object headers ---------------------------------------------------------------------
    def _messages(self, addr):
        b = self.b
        version, _, nmsg, refcount, hsize = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise H5Error(f"object header version {version} at {addr:#x}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if msize % 8:
                    raise H5Error(f"message size {msize} not a multiple of 8 at {pos:#x}")
                if mtype == 0x10:
                    blocks.append(struct.unpack_from("<QQ", b, body))
                out.append((mtype, body, msize))
                pos = body + msize
        if len(out) != nmsg:
            raise H5Error(f"object header at {addr:#x}: {len(out)} of {nmsg} messages found")
        return out

    def _datatype(self, pos):
        b = self.b
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, pos)
        cls, ver = cv & 15, cv >> 4
        if cls == 0:
            signed = bool(b0 & 8)
            return np.dtype(("<" if not b0 & 1 else ">") + ("i" if signed else "u") + str(size)), 8 + 4
        if cls == 1:
            if size == 4:
                exp_loc, exp_size, man_loc, man_size, bias = struct.unpack_from("<BBBBI", b, pos + 12)
                if (exp_loc, exp_size, man_loc, man_size, bias, b1) != (23, 8, 0, 23, 127, 31):
                    raise H5Error("not IEEE float32")
            return np.dtype(("<" if not b0 & 1 else ">") + "f" + str(size)), 8 + 12
        if cls == 3:
            return np.dtype("S" + str(size)), 8
        raise H5Error(f"datatype class {cls} not supported")

    def _dataspace(self, pos):
        b = self.b
        version, rank, flags = struct.unpack_from("<BBB", b, pos)
        if version == 1:
            dims = struct.unpack_from("<" + "Q" * rank, b, pos + 8)
            size = 8 + 8 * rank * (2 if flags & 1 else 1)
        elif version == 2:
            dims = struct.unpack_from("<" + "Q" * rank, b, pos + 4)
            size = 4 + 8 * rank * (2 if flags & 1 else 1)
        else:
            raise H5Error(f"dataspace version {version}")
        return tuple(dims), size, rank == 0

    def _attribute(self, pos):
        b = self.b
        version = b[pos]
        pad = (lambda n: (n + 7) & ~7) if version == 1 else (lambda n: n)
        nsz, tsz, ssz = struct.unpack_from("<HHH", b, pos + 2)
        p = pos + 8 + (1 if version == 3 else 0)
        name = b[p:p + nsz].split(b"\0")[0].decode()
        p += pad(nsz)
        dt, _ = self._datatype(p)
        p += pad(tsz)
        dims, _, scalar = self._dataspace(p)
        p += pad(ssz)
        n = int(np.prod(dims)) if dims else 1
        val = np.frombuffer(b, dt, n, p).copy()
        if dt.kind == "S":
            val = val[0].split(b"\0")[0].decode("latin-1") if n == 1 else [v.split(b"\0")[0].decode("latin-1") for v in val]
        elif scalar:
            val = val[0]
        else:
            val = val.reshape(dims)
        return name, val

    def _object(self, name, addr):
        b = self.b
        node = Node(name, addr)
        msgs = self._messages(addr)
        node.messages = msgs
        dt = dims = layout = None
        for mtype, pos, size in msgs:
            if mtype == 0x11:
                node.is_group = True
                btree, heap = struct.unpack_from("<QQ", b, pos)
                for cname, caddr in self._group_entries(btree, heap):
                    node.children[cname] = self._object(cname, caddr)
            elif mtype == 0x0C:
                k, v = self._attribute(pos)
                node.attrs[k] = v
            elif mtype == 0x03:
                dt, _ = self._datatype(pos)
            elif mtype == 0x01:
                dims, _, _ = self._dataspace(pos)
            elif mtype == 0x08:
                layout = pos
        if dt is not None and dims is not None and layout is not None:
            version, cls = b[layout], b[layout + 1]
            if version != 3:
                raise H5Error(f"layout version {version}")
            n = int(np.prod(dims)) if dims else 1
            if cls == 1:
                daddr, dsize = struct.unpack_from("<QQ", b, layout + 2)
                if dsize != n * dt.itemsize:
                    raise H5Error(f"{name}: layout size {dsize} != {n} x {dt.itemsize}")
                if daddr == UNDEF:
                    node.data = np.zeros(dims, dt)
                else:
                    if daddr + dsize > len(b):
                        raise H5Error(f"{name}: data beyond end of file")
                    node.data = np.frombuffer(b, dt, n, daddr).reshape(dims)
            elif cls == 0:
                (dsize,) = struct.unpack_from("<H", b, layout + 2)
                node.data = np.frombuffer(b, dt, n, layout + 4).reshape(dims)
            else:
                raise H5Error("chunked layout not supported")
        return node

    # ---- groups -----------------------------------------------------------------------------
    def _heap_string(self, heap, off):
        b = self.b
        if b[heap:heap + 4] != b"HEAP":
            raise H5Error(f"no HEAP at {heap:#x}")
        dsize, free, daddr = struct.unpack_from("<QQQ", b, heap + 8)
        if off >= dsize:
            raise H5Error("name offset outside the local heap")
        s = daddr + off
        return b[s:b.index(b"\0", s)].decode()

    def _group_entries(self, btree, heap):
        b = self.b
        if b[btree:btree + 4] != b"TREE":
            raise H5Error(f"no TREE at {btree:#x}")
        ntype, level, used = struct.unpack_from("<BBH", b, btree + 4)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        out = []
        p = btree + 24
        keys = []
        for i in range(used):
            key, child = struct.unpack_from("<QQ", b, p)
            keys.append(key)
            p += 16
            if level > 0:
                out += self._group_entries(child, heap)
                continue
            if b[child:child + 4] != b"SNOD":
                raise H5Error(f"no SNOD at {child:#x}")
            nsym = struct.unpack_from("<H", b, child + 6)[0]
            if nsym > 2 * self.leaf_k:
                raise H5Error("symbol node over-full")
            names = []
            for j in range(nsym):
                noff, oaddr, cache = struct.unpack_from("<QQI", b, child + 8 + 40 * j)
                names.append(self._heap_string(heap, noff))
                out.append((names[-1], oaddr))
            if names != sorted(names):
                raise H5Error("symbol node entries are not sorted by name")
            # the right key of a leaf child is its largest name
            (rkey,) = struct.unpack_from("<Q", b, p)
            if names and self._heap_string(heap, rkey) != names[-1]:
                raise H5Error("B-tree key does not name the largest entry of its child")
        return out


def structure(f: File):
    """{path: (kind, dtype, shape, sorted attr names)} -- for comparing two files' layouts."""
    out = {}
    for path, node in f.root.walk():
        if node.data is not None:
            out[path] = ("dataset", str(node.data.dtype), node.data.shape, tuple(sorted(node.attrs)))
        else:
            out[path] = ("group", None, None, tuple(sorted(node.attrs)))
    return out
