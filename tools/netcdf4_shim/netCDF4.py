"""Stand-in for the `netCDF4` module (netCDF4-python is not in this image) with the handful of calls the reference's
initial-condition generators make (utils/make_nuclei.py, utils/make4corners.py, tests/ConservedVolume/make_square.py, ...):

    f = Dataset(name, 'w', format='NETCDF4'); f.createDimension('x', nx); v = f.createVariable('phase', 'f', ('z','y','x'));
    v[:, :, :] = array; f.close()

close() writes a NetCDF-4 style container (HDF5: one dataset per variable, one storage-less dataset per dimension) with
tests/hdf5_writer.py, which composes the file from the HDF5 format specification.  Put this directory on PYTHONPATH and the
generators run unmodified:   PYTHONPATH=tools/netcdf4_shim python /path/to/utils/make_nuclei.py ... out.nc
Tooling for tests and demonstrations; nothing under ampe_b200/ imports it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import hdf5_writer  # noqa: E402

_DTYPES = {"f": np.float32, "f4": np.float32, "d": np.float64, "f8": np.float64}


class Variable:
    def __init__(self, dtype, shape):
        self.data = np.zeros(shape, dtype=_DTYPES[dtype])

    def __setitem__(self, key, value):
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]


class Dataset:
    def __init__(self, filename, mode="r", format="NETCDF4"):
        if mode != "w":
            raise NotImplementedError("the stand-in writes files only")
        self.filename, self.dimensions, self.variables = filename, {}, {}

    def createDimension(self, name, size):
        self.dimensions[name] = int(size)

    def createVariable(self, name, dtype, dims):
        v = Variable(dtype, tuple(self.dimensions[d] for d in dims))
        self.variables[name] = v
        return v

    def close(self):
        hdf5_writer.write_hdf5(self.filename, {k: v.data for k, v in self.variables.items()}, dimensions=dict(self.dimensions))
