"""GFBT: the flat table container the B200 back end reads.

The reference loads equilibria from netCDF-4 files
(`/root/reference/graph_framework/equilibrium.hpp:1628-1854` make_efit,
`:2424-2640` make_vmec).  libnetcdf/HDF5 are not part of this stack, so the
product reads a trivially simple container instead and `nc_to_gfbt` converts a
netCDF-4 file once (see `h5lite.py`).

Layout (little endian):
    b"GFBT1\\n"
    u32 nvars
    per variable:  u32 name_len | name bytes | u32 rank | u64 dims[rank] | f64 data[prod(dims)]
"""
import struct
import numpy as np

MAGIC = b"GFBT1\n"


def write_gfbt(path, arrays):
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(arrays)))
        for name in sorted(arrays):
            a = np.ascontiguousarray(np.asarray(arrays[name], dtype="<f8"))
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<I", a.ndim))
            for d in a.shape:
                f.write(struct.pack("<Q", d))
            f.write(a.tobytes())


def read_gfbt(path):
    out = {}
    with open(path, "rb") as f:
        b = f.read()
    assert b[:6] == MAGIC, "not a GFBT file: %s" % path
    p = 6
    (n,) = struct.unpack_from("<I", b, p)
    p += 4
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", b, p)
        p += 4
        name = b[p:p + ln].decode()
        p += ln
        (rank,) = struct.unpack_from("<I", b, p)
        p += 4
        dims = struct.unpack_from("<%dQ" % rank, b, p)
        p += 8 * rank
        cnt = int(np.prod(dims)) if rank else 1
        out[name] = np.frombuffer(b, "<f8", cnt, p).reshape(dims).copy()
        p += 8 * cnt
    return out


def nc_to_gfbt(nc_path, out_path):
    """Convert every floating point dataset of a netCDF-4 file."""
    from .h5lite import H5File
    h = H5File(nc_path)
    arrays = {}
    for name in h.names():
        shape, dtype, _ = h.datasets[name]
        if dtype.kind != "f" or dtype.itemsize != 8:
            # netCDF dimension scale (f4 placeholder): keep only its length.
            arrays["dim:" + name] = np.array(float(shape[0]))
            continue
        arrays[name] = h.read(name)
    write_gfbt(out_path, arrays)
    return arrays


def write_trajectory(path, records, names=("t", "w", "x", "y", "z", "kx", "ky", "kz", "residual")):
    """Store RayTracer.trace output ([time, 9, rays]) with one (time, num_rays) variable per quantity,
    the shape of the reference's result files (output.hpp:166: time-unlimited x num_rays)."""
    records = np.asarray(records)
    write_gfbt(path, {name: records[:, i, :] for i, name in enumerate(names)})
