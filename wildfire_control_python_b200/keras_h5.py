"""Reader for the reference's trained networks.

``Models/<name>`` in the reference tree are Keras ``save_weights`` files (DQN.py:441-443): HDF5 written
by h5py with the library defaults -- superblock version 0, version-1 object headers, groups stored as
symbol tables (v1 B-tree + local heap), datasets with contiguous (or compact) layout, little-endian
IEEE floats, no filters.  Neither h5py nor Keras is available here, so this module reads exactly that
subset of the HDF5 file format (HDF5 File Format Specification, version 1.1 structures) with
``struct`` and NumPy; anything else (chunked / filtered datasets, new-style groups) raises.

    read_keras_weights(path) -> {"dense_1/kernel:0": float32 [in, out], "dense_1/bias:0": float32 [out], ...}

Layer names follow the file (``dense_1`` .. ``dense_4``); kernels are in Keras' [in, out] orientation,
which is what ``agents.DQN.set_weights`` takes.  Because Keras numbers layers globally per process,
files written later in one session can carry names like ``dense_7``: ``canonical_dense_names`` maps
them onto ``dense_1..`` in order.
"""
from __future__ import annotations

import re
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class _File:
    def __init__(self, data: bytes):
        self.b = data
        if data[:8] != _SIG:
            raise H5Error("not an HDF5 file")
        ver = data[8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} is not supported (only the h5py default, 0/1)")
        self.O, self.L = data[13], data[14]  # size of offsets / lengths
        if (self.O, self.L) != (8, 8):
            raise H5Error("only 8-byte offsets and lengths are supported")
        p = 24 + (4 if ver == 1 else 0)  # v1 adds indexed-storage K + reserved
        self.base, _free, _eof, _drv = struct.unpack_from("<4Q", data, p)
        p += 32
        # root group symbol table entry
        _name_off, self.root_header, cache_type, _res = struct.unpack_from("<QQII", data, p)
        self.root_cache = struct.unpack_from("<QQ", data, p + 24) if cache_type == 1 else None

    # -- primitives ------------------------------------------------------------------------------
    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, self.base + off)

    def cstr(self, off):
        end = self.b.index(b"\x00", self.base + off)
        return self.b[self.base + off:end].decode("utf-8")

    # -- object headers (version 1) --------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, bytes)] of the version-1 object header at `addr`, continuations followed."""
        ver, _r, nmsg, _refc, hsize = self.u("BBHII", addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} is not supported")
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = self.u("HHB", p)
                body = self.b[self.base + p + 8: self.base + p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:  # continuation
                    blocks.append(struct.unpack_from("<QQ", body))
                out.append((mtype, mflags, body))
        return out

    # -- groups: symbol table = v1 B-tree of symbol nodes + local heap of names --------------------
    def _heap_data(self, heap_addr):
        if self.b[self.base + heap_addr: self.base + heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        _size, _free, data_addr = self.u("QQQ", heap_addr + 8)
        return data_addr

    def _walk_btree(self, addr, heap_data, out):
        sig = self.b[self.base + addr: self.base + addr + 4]
        if sig == b"TREE":
            ntype, level, used = self.u("BBH", addr + 4)
            if ntype != 0:
                raise H5Error("not a group B-tree")
            p = addr + 8 + 16  # skip sibling pointers
            for k in range(used):
                child = self.u("Q", p + 8)[0]  # key k (8 bytes), then child k
                self._walk_btree(child, heap_data, out)
                p += 16
        elif sig == b"SNOD":
            _ver, _r, nsym = self.u("BBH", addr + 4)
            p = addr + 8
            for _ in range(nsym):
                name_off, header, cache_type, _res = self.u("QQII", p)
                out[self.cstr(heap_data + name_off)] = header
                p += 40
        else:
            raise H5Error(f"unexpected node signature {sig!r}")

    def children(self, header_addr):
        """{name: object header address} for a group, {} for anything else."""
        for mtype, _f, body in self.messages(header_addr):
            if mtype == 0x11:  # symbol table message
                btree, heap = struct.unpack_from("<QQ", body)
                out = {}
                self._walk_btree(btree, self._heap_data(heap), out)
                return out
        return {}

    # -- datasets ----------------------------------------------------------------------------------
    def dataset(self, header_addr):
        """ndarray of a dataset object, or None if the object is not a dataset."""
        shape = dtype = None
        layout = None
        for mtype, _f, body in self.messages(header_addr):
            if mtype == 0x01:  # dataspace
                ver, rank, flags = body[0], body[1], body[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from(f"<{rank}Q", body, p) if rank else ()
            elif mtype == 0x03:  # datatype
                cls, size = body[0] & 15, struct.unpack_from("<I", body, 4)[0]
                big = body[1] & 1
                if cls == 1:
                    dtype = np.dtype((">" if big else "<") + f"f{size}")
                elif cls == 0:
                    signed = (body[1] >> 3) & 1
                    dtype = np.dtype((">" if big else "<") + ("i" if signed else "u") + str(size))
                else:
                    dtype = ("unsupported", cls)
            elif mtype == 0x08:  # data layout
                ver = body[0]
                if ver == 3:
                    lclass = body[1]
                    if lclass == 1:
                        layout = ("contiguous",) + struct.unpack_from("<QQ", body, 2)
                    elif lclass == 0:
                        n = struct.unpack_from("<H", body, 2)[0]
                        layout = ("compact", body[4:4 + n])
                    else:
                        layout = ("chunked",)
                elif ver in (1, 2):
                    rank, lclass = body[1], body[2]
                    if lclass == 1:
                        layout = ("contiguous", struct.unpack_from("<Q", body, 8)[0], None)
                    else:
                        layout = ("chunked",) if lclass == 2 else ("compact_v1",)
                else:
                    raise H5Error(f"data layout message version {ver}")
            elif mtype == 0x0B:
                raise H5Error("filtered (compressed) datasets are not supported")
        if shape is None or dtype is None or layout is None:
            return None
        if isinstance(dtype, tuple):
            raise H5Error(f"datatype class {dtype[1]} is not supported")
        count = int(np.prod(shape)) if shape else 1
        nbytes = count * dtype.itemsize
        if layout[0] == "contiguous":
            addr = layout[1]
            if addr == _UNDEF:
                return np.zeros(shape, dtype.newbyteorder("="))
            raw = self.b[self.base + addr: self.base + addr + nbytes]
        elif layout[0] == "compact":
            raw = layout[1][:nbytes]
        else:
            raise H5Error(f"{layout[0]} dataset layout is not supported")
        if len(raw) != nbytes:
            raise H5Error("truncated dataset")
        return np.frombuffer(raw, dtype=dtype).reshape(shape).astype(dtype.newbyteorder("="))


def read_h5_datasets(path):
    """Every dataset of an (old-style) HDF5 file: {"group/sub/name": ndarray}."""
    with open(path, "rb") as f:
        h5 = _File(f.read())
    out = {}

    def walk(prefix, header, depth=0):
        if depth > 16:
            raise H5Error("group nesting too deep")
        kids = h5.children(header)
        if not kids:
            arr = h5.dataset(header)
            if arr is not None:
                out[prefix] = arr
            return
        for name, addr in kids.items():
            walk(f"{prefix}/{name}" if prefix else name, addr, depth + 1)

    walk("", h5.root_header)
    return out


def read_keras_weights(path):
    """Weights of a Keras ``save_weights`` file: {"<layer>/<weight>": ndarray}.

    Keras stores ``/<layer>/<layer>/<weight>`` (e.g. ``/dense_1/dense_1/kernel:0``); the duplicated group
    level is dropped.  Kernels keep Keras' [in, out] orientation."""
    out = {}
    for key, arr in read_h5_datasets(path).items():
        parts = key.split("/")
        if len(parts) >= 3 and parts[0] == parts[1]:
            parts = parts[1:]
        out["/".join(parts)] = arr
    if not out:
        raise H5Error("no datasets found")
    return out


def canonical_dense_names(weights):
    """Renumber ``dense_K`` layers to ``dense_1..dense_n`` in increasing K (Keras numbers layers per process)."""
    ks = sorted({int(m.group(1)) for k in weights for m in [re.match(r"dense_(\d+)/", k)] if m})
    ren = {k: i + 1 for i, k in enumerate(ks)}
    out = {}
    for key, arr in weights.items():
        m = re.match(r"dense_(\d+)/(.*)", key)
        out[f"dense_{ren[int(m.group(1))]}/{m.group(2)}" if m else key] = arr
    return out
