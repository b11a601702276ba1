"""Writer of NetCDF CLASSIC files (CDF-1, or CDF-2 with 64-bit offsets) for initial conditions, numpy only.

The reference's utils/*.py write their initial-condition files through the netCDF4 Python package; this
image has neither libnetcdf nor libhdf5, so synthetic states are written here in the classic format the
reference reads through its HAVE_NETCDF3 branch (source/FieldsInitializer.cc:112-142): fixed-size float or
double variables dimensioned (z, y, x), named `phase`, `quat1`.., `concentration`, `temperature`, plus the
dimension `qlen`.  Format: "The NetCDF Classic Format Specification" (header: magic, numrecs, dim_list,
gatt_list, var_list; big-endian; names and data padded to 4 bytes)."""
import struct

import numpy as np

_NC_DIMENSION, _NC_VARIABLE = 0x0A, 0x0B
_TYPES = {np.dtype("float32"): (5, ">f4"), np.dtype("float64"): (6, ">f8")}


def _name(s):
    b = s.encode()
    return struct.pack(">I", len(b)) + b + b"\0" * ((4 - len(b) % 4) % 4)


def write(filename, variables, extra_dims=None, version=1):
    """variables: dict name -> (nz, ny, nx) float32 / float64 array (all the same shape);
    extra_dims: dict name -> size (e.g. {"qlen": 4}); version 1 (CDF-1) or 2 (64-bit offsets)"""
    if version not in (1, 2):
        raise ValueError("version must be 1 or 2")
    shapes = {tuple(v.shape) for v in variables.values()}
    if len(shapes) != 1 or len(next(iter(shapes))) != 3:
        raise ValueError("all variables must have the same (nz, ny, nx) shape")
    nz, ny, nx = next(iter(shapes))
    dims = [("z", nz), ("y", ny), ("x", nx)] + list((extra_dims or {}).items())
    head = b"CDF" + bytes([version]) + struct.pack(">I", 0)
    head += struct.pack(">II", _NC_DIMENSION, len(dims))
    for n, size in dims:
        head += _name(n) + struct.pack(">I", int(size))
    head += struct.pack(">II", 0, 0)  # no global attributes
    off_fmt = ">I" if version == 1 else ">Q"
    entries, blobs = [], []
    for n, v in variables.items():
        a = np.ascontiguousarray(v)
        if a.dtype not in _TYPES:
            raise ValueError("variable %s: float32 or float64 only" % n)
        code, be = _TYPES[a.dtype]
        data = a.astype(be).tobytes()
        data += b"\0" * ((4 - len(data) % 4) % 4)
        entries.append((n, code, len(data)))
        blobs.append(data)
    # header size: fixed per variable entry
    var_list = struct.pack(">II", _NC_VARIABLE, len(entries)) if entries else struct.pack(">II", 0, 0)
    sizes = [len(_name(n)) + 4 + 3 * 4 + 8 + 4 + 4 + struct.calcsize(off_fmt) for n, _, _ in entries]
    begin = len(head) + len(var_list) + sum(sizes)
    for (n, code, nbytes), blob in zip(entries, blobs):
        var_list += _name(n) + struct.pack(">I", 3) + struct.pack(">III", 0, 1, 2) + struct.pack(">II", 0, 0)
        var_list += struct.pack(">I", code) + struct.pack(">I", nbytes & 0xFFFFFFFF) + struct.pack(off_fmt, begin)
        begin += nbytes
    with open(filename, "wb") as f:
        f.write(head + var_list)
        for blob in blobs:
            f.write(blob)


def write_state(filename, state, qlen=0, dtype=np.float32, version=1):
    """state: dict with phase / quat (qlen, nz, ny, nx) / conc / temperature arrays (ghost-0 SAMRAI order, 2D
    arrays as (1, ny, nx)) -> the variable names FieldsInitializer reads"""
    out = {}

    def arr(a):
        a = np.asarray(a.cpu() if hasattr(a, "cpu") else a)
        return a.reshape((-1,) + a.shape[-2:]).astype(dtype)

    if state.get("phase") is not None:
        out["phase"] = arr(state["phase"])
    if state.get("quat") is not None and qlen > 0:
        q = np.asarray(state["quat"].cpu() if hasattr(state["quat"], "cpu") else state["quat"])
        q = q.reshape((qlen, -1) + q.shape[-2:])
        for m in range(qlen):
            out["quat%d" % (m + 1)] = q[m].astype(dtype)
    if state.get("conc") is not None:
        out["concentration"] = arr(state["conc"])
    if state.get("temperature") is not None:
        out["temperature"] = arr(state["temperature"])
    write(filename, out, extra_dims={"qlen": qlen} if qlen > 0 else None, version=version)
