"""initial conditions from NetCDF classic files (ampe_b200/host/FieldsInitializer.h, mirror
of source/FieldsInitializer.cc) -- host logic, no GPU.  scipy.io.netcdf_file (an independent implementation of
the classic format) writes the fixtures the C++ reader parses and reads back what ampe_b200.netcdf_classic wrote."""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

import parity
from ampe_b200 import host_rhs, netcdf_classic
from ampe_b200.lib import AmpeError


def _scipy_write(path, variables, qlen=0, version=1):
    f = netcdf_file(path, "w", version=version)
    nz, ny, nx = next(iter(variables.values())).shape
    f.createDimension("z", nz)
    f.createDimension("y", ny)
    f.createDimension("x", nx)
    if qlen:
        f.createDimension("qlen", qlen)
    f.history = "test fixture"  # a global attribute the reader has to skip
    for name, a in variables.items():
        v = f.createVariable(name, a.dtype.char, ("z", "y", "x"))
        v.units = "1"  # a variable attribute to skip
        v[:] = a
    f.close()


def _state(name):
    cfg, st = parity.make_case(name)
    y = {k: (None if v is None else v.numpy()) for k, v in st.items()}
    return cfg, y


def _file_vars(cfg, y, dtype):
    out = {"phase": y["phase"].reshape((-1,) + y["phase"].shape[-2:]).astype(dtype)}
    q = y["quat"].reshape((cfg.qlen, -1) + y["quat"].shape[-2:])
    for m in range(cfg.qlen):
        out["quat%d" % (m + 1)] = q[m].astype(dtype)
    if y.get("conc") is not None:
        out["concentration0"] = y["conc"].reshape(out["phase"].shape).astype(dtype)
    if y.get("temperature") is not None:
        out["temperature"] = y["temperature"].reshape(out["phase"].shape).astype(dtype)
    return out


@pytest.mark.parametrize("name,dtype,version", [("auni3d", np.float32, 1), ("auni3d", np.float64, 2),
                                                ("dendrite2d", np.float32, 1), ("auni2d", np.float64, 1)])
def test_reader_against_scipy_written_files(tmp_path, name, dtype, version):
    """float files come back as the float-rounded state, double files bit for bit; `concentration0` is found
    when `concentration` is absent (FieldsInitializer.cc:283-287); attributes are skipped"""
    cfg, y = _state(name)
    path = str(tmp_path / "init.nc")
    _scipy_write(path, _file_vars(cfg, y, dtype), qlen=cfg.qlen, version=version)
    got = host_rhs.read_initial_conditions(path, cfg)
    for k in ("phase", "quat", "conc", "temperature"):
        if y.get(k) is None or (k == "temperature" and not cfg.with_unsteady_temperature):
            continue
        expect = y[k].astype(dtype).astype(np.float64)
        assert got[k] is not None and np.array_equal(got[k].numpy().reshape(expect.shape), expect), k


def test_own_writer_is_read_by_scipy_and_by_the_reader(tmp_path):
    cfg, y = _state("gg3d_hbsm")
    for version in (1, 2):
        path = str(tmp_path / ("own%d.nc" % version))
        netcdf_classic.write_state(path, y, qlen=cfg.qlen, dtype=np.float32, version=version)
        f = netcdf_file(path, "r", mmap=False)
        assert f.dimensions["qlen"] == cfg.qlen and f.dimensions["x"] == cfg.n[0]
        assert np.array_equal(f.variables["phase"][:], y["phase"].astype(np.float32))
        assert np.array_equal(f.variables["quat3"][:], y["quat"][2].astype(np.float32))
        f.close()
        got = host_rhs.read_initial_conditions(path, cfg)
        assert np.array_equal(got["conc"].numpy(), y["conc"].astype(np.float32).astype(np.float64))
        assert np.array_equal(got["quat"].numpy(), y["quat"].astype(np.float32).astype(np.float64))


def test_slab_ranks_read_their_own_planes(tmp_path):
    """two ranks, slab along z: rank r reads planes [r n2, (r+1) n2) of the file (the box of its patch)"""
    cfg, y = _state("auni3d")
    path = str(tmp_path / "init.nc")
    netcdf_classic.write_state(path, y, qlen=cfg.qlen, dtype=np.float64)
    nz = cfg.n[2]
    assert nz % 2 == 0
    for rank in (0, 1):
        cfg_r, _ = parity.make_case("auni3d")
        cfg_r.n[2] = nz // 2
        cfg_r.nranks, cfg_r.rank = 2, rank
        got = host_rhs.read_initial_conditions(path, cfg_r)
        sl = slice(rank * nz // 2, (rank + 1) * nz // 2)
        assert np.array_equal(got["phase"].numpy(), y["phase"][sl])
        assert np.array_equal(got["quat"].numpy(), y["quat"][:, sl])


def test_two_dimensional_run_reads_one_slice(tmp_path):
    """a 2D run on a file with nz > 1 reads slice nz_file / 2 unless slice_index is given (:229-236)"""
    cfg, y = _state("dendrite2d")
    rng = np.random.default_rng(3)
    stack = rng.random((5,) + y["phase"].shape[-2:])
    vars_ = {"phase": stack, "temperature": stack + 1.0}
    for m in range(cfg.qlen):
        vars_["quat%d" % (m + 1)] = stack * (m + 2)
    path = str(tmp_path / "stack.nc")
    netcdf_classic.write(path, vars_, extra_dims={"qlen": cfg.qlen})
    got = host_rhs.read_initial_conditions(path, cfg)
    assert np.array_equal(got["phase"].numpy()[0], stack[2])
    got = host_rhs.read_initial_conditions(path, cfg, slice_index=4)
    assert np.array_equal(got["quat"].numpy()[1, 0], stack[4] * 3)


def test_errors_follow_the_reference(tmp_path):
    cfg, y = _state("auni2d")
    vars_ = _file_vars(cfg, y, np.float32)
    path = str(tmp_path / "a.nc")
    missing = dict(vars_)
    del missing["quat2"]
    _scipy_write(path, missing, qlen=cfg.qlen)
    with pytest.raises(AmpeError, match="Could not read variable 'quat2' from input data"):
        host_rhs.read_initial_conditions(path, cfg)
    _scipy_write(path, vars_, qlen=cfg.qlen + 1)
    with pytest.raises(AmpeError, match="qlen_file=%d, QLEN=%d" % (cfg.qlen + 1, cfg.qlen)):
        host_rhs.read_initial_conditions(path, cfg)
    small = {k: v[:, :, :-2] for k, v in vars_.items()}
    _scipy_write(path, small, qlen=cfg.qlen)
    with pytest.raises(AmpeError, match="Phase input data dimensions are incorrect, nx_file=%d" % (cfg.n[0] - 2)):
        host_rhs.read_initial_conditions(path, cfg)
    with pytest.raises(AmpeError, match="Cannot open file"):
        host_rhs.read_initial_conditions(str(tmp_path / "nope.nc"), cfg)
    with open(path, "wb") as f:
        f.write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(AmpeError, match="corrupt HDF5 superblock"):  # a bare signature; real containers are read (round 2)
        host_rhs.read_initial_conditions(path, cfg)
    with open(path, "wb") as f:
        f.write(b"CDF\x01\0\0")
    with pytest.raises(AmpeError, match="truncated"):
        host_rhs.read_initial_conditions(path, cfg)
