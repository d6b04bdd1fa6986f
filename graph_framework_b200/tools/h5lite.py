"""Minimal read-only HDF5 (netCDF-4) reader used by the netCDF -> GFBT table converter.

The reference's fixtures (`graph_tests/efit.nc`, `efit_gold.nc`, `vmec.nc`) are
netCDF-4 = HDF5 files and this image has neither libnetcdf, h5py nor netCDF4.
This reader understands exactly what those files use: superblock v0, version-2
object headers ("OHDR"/"OCHK"), link messages (compact or inside fractal-heap
direct blocks), little-endian IEEE f8/f4/i4/i8 datasets with compact,
contiguous or chunked (v1 B-tree, unfiltered) layout.

It is a data-format utility: `gfbt.nc_to_gfbt` uses it to turn an equilibrium file into the flat
container the back end reads.  The compute path never touches HDF5.
"""
import struct
import numpy as np


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert b[8] == 0, "only superblock v0 handled"
        assert b[13] == 8 and b[14] == 8, "8-byte offsets/lengths expected"
        self.ohdr = []
        pos = b.find(b"OHDR")
        while pos >= 0:
            if b[pos + 4] == 2:
                self.ohdr.append(pos)
            pos = b.find(b"OHDR", pos + 4)
        self.ohdr_set = set(self.ohdr)
        self.links = self._scan_links()
        self.datasets = {}
        for name, addr in self.links.items():
            info = self._parse_object(addr)
            if info is not None:
                self.datasets[name] = info

    # -- link messages ---------------------------------------------------
    def _scan_links(self):
        """Find every hard-link message body whose target is an object header."""
        b = self.b
        out = {}
        n = len(b)
        for pos in range(n - 12):
            if b[pos] != 1:
                continue
            flags = b[pos + 1]
            if flags & 0xE0:
                continue
            p = pos + 2
            if flags & 0x08:
                if b[p] != 0:       # hard links only
                    continue
                p += 1
            if flags & 0x04:
                p += 8
            if flags & 0x10:
                p += 1
            lsz = 1 << (flags & 3)
            if p + lsz > n:
                continue
            ln = int.from_bytes(b[p:p + lsz], "little")
            p += lsz
            if ln == 0 or ln > 64 or p + ln + 8 > n:
                continue
            name = b[p:p + ln]
            if not all(32 < c < 127 for c in name):
                continue
            addr = int.from_bytes(b[p + ln:p + ln + 8], "little")
            if addr in self.ohdr_set:
                out[name.decode()] = addr
        return out

    # -- object headers --------------------------------------------------
    def _messages(self, addr):
        b = self.b
        assert b[addr:addr + 4] == b"OHDR"
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        csz = 1 << (flags & 3)
        chunk0 = int.from_bytes(b[p:p + csz], "little")
        p += csz
        todo = [(p, p + chunk0)]
        track = bool(flags & 0x04)
        while todo:
            p, end = todo.pop(0)
            while p + 4 <= end:
                mtype = b[p]
                msize = struct.unpack_from("<H", b, p + 1)[0]
                p += 4
                if track:
                    p += 2
                body = b[p:p + msize]
                p += msize
                if mtype == 0x10:
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    assert b[off:off + 4] == b"OCHK"
                    todo.append((off + 4, off + ln - 4))
                elif mtype != 0:
                    yield mtype, body

    def _parse_object(self, addr):
        shape = None
        dtype = None
        layout = None
        for mtype, body in self._messages(addr):
            if mtype == 0x01:
                ver, rank, fl = body[0], body[1], body[2]
                q = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from("<Q", body, q + 8 * i)[0]
                              for i in range(rank))
            elif mtype == 0x03:
                cls = body[0] & 0x0F
                size = struct.unpack_from("<I", body, 4)[0]
                if cls == 1:
                    dtype = np.dtype("<f%d" % size)
                elif cls == 0:
                    signed = bool(body[1] & 0x08)
                    dtype = np.dtype("<%s%d" % ("i" if signed else "u", size))
                else:
                    dtype = None
            elif mtype == 0x08:
                ver, cls = body[0], body[1]
                assert ver == 3, "layout v%d not handled" % ver
                if cls == 0:
                    sz = struct.unpack_from("<H", body, 2)[0]
                    layout = ("compact", body[4:4 + sz])
                elif cls == 1:
                    a, sz = struct.unpack_from("<QQ", body, 2)
                    layout = ("contiguous", a, sz)
                elif cls == 2:
                    nd = body[2]
                    bt = struct.unpack_from("<Q", body, 3)[0]
                    dims = struct.unpack_from("<%dI" % nd, body, 11)
                    layout = ("chunked", bt, dims)
            elif mtype == 0x0B:
                raise NotImplementedError("filtered dataset")
        if shape is None or dtype is None or layout is None:
            return None
        return shape, dtype, layout

    # -- raw data --------------------------------------------------------
    def _btree_chunks(self, addr, nd):
        b = self.b
        assert b[addr:addr + 4] == b"TREE" and b[addr + 4] == 1
        level = b[addr + 5]
        used = struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 24
        keysz = 8 + 8 * nd
        for _ in range(used):
            csize, fmask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<%dQ" % nd, b, p + 8)
            child = struct.unpack_from("<Q", b, p + keysz)[0]
            p += keysz + 8
            if level == 0:
                assert fmask == 0
                yield offs, csize, child
            else:
                yield from self._btree_chunks(child, nd)

    def read(self, name):
        shape, dtype, layout = self.datasets[name]
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            arr = np.frombuffer(layout[1], dtype=dtype, count=count)
        elif layout[0] == "contiguous":
            a = layout[1]
            if a == 0xFFFFFFFFFFFFFFFF:
                arr = np.zeros(count, dtype)
            else:
                arr = np.frombuffer(self.b, dtype=dtype, count=count, offset=a)
        else:
            bt, dims = layout[1], layout[2]
            nd = len(dims)
            cshape = dims[:-1]
            out = np.zeros(shape, dtype)
            for offs, csize, child in self._btree_chunks(bt, nd):
                chunk = np.frombuffer(self.b, dtype=dtype,
                                      count=int(np.prod(cshape)),
                                      offset=child).reshape(cshape)
                sl = tuple(slice(o, min(o + c, s))
                           for o, c, s in zip(offs[:-1], cshape, shape))
                sub = tuple(slice(0, s.stop - s.start) for s in sl)
                out[sl] = chunk[sub]
            arr = out
        return np.array(arr, dtype=dtype).reshape(shape)

    def names(self):
        return sorted(self.datasets)
