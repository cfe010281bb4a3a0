"""h5py-free reader for the XDMF3 + HDF5 triangle meshes the reference loads.

The reference reads its mesh with DOLFIN's ``XDMFFile(...).read(mesh)``
(/root/reference/flow_solver.py:58-62).  Neither DOLFIN nor h5py exists in this
image, so this module parses the small subset of HDF5 those files use:
superblock v0, v1 object headers, symbol-table groups (TREE/SNOD/HEAP),
contiguous or chunked (layout v3, v1 chunk B-tree) datasets with optional
gzip.  It is host-side input plumbing, not part of the timed hot path.
"""
from __future__ import annotations

import os
import re
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class _H5:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        ver = buf[8]
        if ver != 0:
            raise ValueError(f"unsupported HDF5 superblock version {ver}")
        self.so = buf[13]  # size of offsets
        self.sl = buf[14]  # size of lengths
        if self.so != 8 or self.sl != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        # superblock v0: 8 sig + 8 version bytes + 4 (leaf k, internal k) + 4 flags
        # + base, free-space, eof, driver addresses, then root symbol table entry
        p = 24 + 4 * 8
        self.root = self._sym_entry(p)

    def u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _sym_entry(self, p):
        name_off = self.u(p, 8)
        ohdr = self.u(p + 8, 8)
        cache = self.u(p + 16, 4)
        scratch = self.b[p + 24:p + 40]
        return name_off, ohdr, cache, scratch

    # ---- object header v1 ----
    def messages(self, addr):
        b = self.b
        if b[addr] != 1:
            raise ValueError("only v1 object headers supported")
        nmsg = self.u(addr + 2, 2)
        hsize = self.u(addr + 8, 4)
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype = self.u(p, 2)
                msize = self.u(p + 2, 2)
                body = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self.u(body, 8), self.u(body + 8, 8)))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    # ---- groups ----
    def listdir(self, ohdr):
        btree = heap = None
        for mtype, body, _ in self.messages(ohdr):
            if mtype == 0x11:
                btree, heap = self.u(body, 8), self.u(body + 8, 8)
        if btree is None:
            raise ValueError("group has no symbol table message")
        if self.b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self.u(heap + 24, 8)
        names = {}
        self._walk_group(btree, heap_data, names)
        return names

    def _walk_group(self, node, heap_data, names):
        b = self.b
        if b[node:node + 4] == b"TREE":
            level = b[node + 5]
            n = self.u(node + 6, 2)
            p = node + 24
            p += 8  # key 0
            for _ in range(n):
                child = self.u(p, 8)
                p += 16  # child + next key
                self._walk_group(child, heap_data, names)
        elif b[node:node + 4] == b"SNOD":
            n = self.u(node + 6, 2)
            p = node + 8
            for _ in range(n):
                name_off, ohdr, _, _ = self._sym_entry(p)
                s = heap_data + name_off
                e = b.index(b"\0", s)
                names[b[s:e].decode()] = ohdr
                p += 40
        else:
            raise ValueError("bad group node")

    # ---- datasets ----
    def dataset(self, ohdr):
        shape = dtype = None
        layout = None
        filters = []
        for mtype, body, msize in self.messages(ohdr):
            b = self.b
            if mtype == 0x01:
                ver, rank, flags = b[body], b[body + 1], b[body + 2]
                p = body + (8 if ver == 1 else 4)
                shape = tuple(self.u(p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                cls = b[body] & 0x0F
                bits0 = b[body + 1]
                size = self.u(body + 4, 4)
                order = ">" if (bits0 & 1) else "<"
                if cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{size}")
                else:
                    raise ValueError(f"unsupported datatype class {cls}")
            elif mtype == 0x0B:
                ver = b[body]
                nf = b[body + 1]
                p = body + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self.u(p, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self.u(p + 2, 2)
                        ncd = self.u(p + 6, 2)
                        p += 8 + ((nlen + 7) // 8 * 8 if ver == 1 else nlen)
                    else:
                        ncd = self.u(p + 4, 2)
                        p += 6
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append(fid)
            elif mtype == 0x08:
                ver = b[body]
                if ver != 3:
                    raise ValueError("only layout v3 supported")
                cls = b[body + 1]
                if cls == 1:
                    layout = ("contiguous", self.u(body + 2, 8), self.u(body + 10, 8))
                elif cls == 2:
                    rank = b[body + 2]
                    addr = self.u(body + 3, 8)
                    dims = tuple(self.u(body + 11 + 4 * i, 4) for i in range(rank))
                    layout = ("chunked", addr, dims)
                elif cls == 0:
                    size = self.u(body + 2, 2)
                    layout = ("compact", body + 4, size)
        if shape is None or dtype is None or layout is None:
            raise ValueError("incomplete dataset header")
        for f in filters:
            if f not in (1, 2):  # deflate, shuffle
                raise ValueError(f"unsupported filter {f}")
        if layout[0] in ("contiguous", "compact"):
            _, addr, size = layout
            n = int(np.prod(shape)) * dtype.itemsize
            return np.frombuffer(self.b[addr:addr + n], dtype=dtype).reshape(shape).astype(dtype.newbyteorder("="))
        _, addr, cdims = layout
        chunk_shape = cdims[:-1]  # last entry is the element size
        out = np.zeros(shape, dtype=dtype.newbyteorder("="))
        self._walk_chunks(addr, len(shape), chunk_shape, dtype, filters, out)
        return out

    def _walk_chunks(self, node, rank, cshape, dtype, filters, out):
        b = self.b
        if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
            raise ValueError("bad chunk B-tree node")
        level = b[node + 5]
        n = self.u(node + 6, 2)
        keysize = 8 + 8 * (rank + 1)
        p = node + 24
        for _ in range(n):
            csize = self.u(p, 4)
            fmask = self.u(p + 4, 4)
            offs = tuple(self.u(p + 8 + 8 * i, 8) for i in range(rank))
            child = self.u(p + keysize, 8)
            p += keysize + 8
            if level > 0:
                self._walk_chunks(child, rank, cshape, dtype, filters, out)
                continue
            raw = self.b[child:child + csize]
            for i, f in reversed(list(enumerate(filters))):
                if fmask & (1 << i):
                    continue
                if f == 1:
                    raw = zlib.decompress(raw)
                elif f == 2:
                    a = np.frombuffer(raw, dtype=np.uint8).reshape(dtype.itemsize, -1)
                    raw = a.T.tobytes()
            chunk = np.frombuffer(raw, dtype=dtype).reshape(cshape)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, out.shape))
            sub = tuple(slice(0, s.stop - s.start) for s in sl)
            out[sl] = chunk[sub]


def read_h5(path: str) -> dict:
    """Return {dataset name: ndarray} for every dataset in the root group."""
    with open(path, "rb") as f:
        h = _H5(f.read())
    return {name: h.dataset(ohdr) for name, ohdr in h.listdir(h.root[1]).items()}


def read_xdmf_mesh(path: str):
    """Read a triangle mesh as the reference does (flow_solver.py:58-62).

    ``path`` may be the ``.xdmf`` file (its ``DataItem`` entries name the HDF5
    datasets), the ``.h5`` file itself, or an ``.npz`` with ``coords``/``cells``.
    Returns ``(coords float64 [V,2], cells int32 [C,3])`` in file order.
    """
    if path.endswith(".npz"):
        z = np.load(path)
        return np.ascontiguousarray(z["coords"], dtype=np.float64), np.ascontiguousarray(z["cells"], dtype=np.int32)
    geo, topo = "data0", "data1"
    h5path = path
    if path.endswith(".xdmf"):
        txt = open(path).read()
        g = re.search(r"<Geometry[^>]*>\s*<DataItem[^>]*>([^<]+)</DataItem>", txt)
        t = re.search(r"<Topology[^>]*>\s*<DataItem[^>]*>([^<]+)</DataItem>", txt)
        if not g or not t:
            raise ValueError("XDMF file has no Geometry/Topology DataItem")
        h5name, geo = g.group(1).strip().split(":/")
        topo = t.group(1).strip().split(":/")[1]
        h5path = os.path.join(os.path.dirname(path), h5name)
    d = read_h5(h5path)
    coords = np.ascontiguousarray(d[geo][:, :2], dtype=np.float64)
    cells = np.ascontiguousarray(d[topo], dtype=np.int32)
    return coords, cells
