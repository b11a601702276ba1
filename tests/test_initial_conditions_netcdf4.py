"""initial conditions in the NetCDF-4 container -- what the reference's generators write
(utils/make_nuclei.py:438 `format='NETCDF4'`) and AMPE reads through libnetcdf / libhdf5
(source/FieldsInitializer.cc:105-300).  ampe_b200/host/NetCDF4File.h reads the HDF5 subset such files use; host logic,
no GPU.

Two kinds of evidence, kept apart:
  * a REAL HDF5 file: scipy ships one (a MATLAB v7.3 file written by libhdf5 1.6: 512-byte user block, superblock 0,
    symbol-table root group, local heap, version-1 object header, contiguous little-endian doubles);
  * files composed from the format specification by tests/hdf5_writer.py for every structure that file does not have
    (superblock 2, "OHDR" headers, link messages, dense links in a fractal heap, chunk B-trees, the filter pipeline).
    Reader and writer share one reading of the specification: these cases show the reader is consistent with it, not
    that libnetcdf's own output has been read (no such file can be produced in this image).
"""
import os

import numpy as np
import pytest

import hdf5_writer
import parity
from ampe_b200 import host_rhs
from ampe_b200.lib import AmpeError
from test_initial_conditions import _file_vars, _state


def _matlab_sample():
    try:
        import scipy.io.matlab.tests as t
    except ImportError:
        return None
    p = os.path.join(os.path.dirname(t.__file__), "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.exists(p) else None


@pytest.mark.skipif(_matlab_sample() is None, reason="scipy's MATLAB v7.3 sample file is not installed")
def test_real_hdf5_file_written_by_libhdf5():
    """testdouble = 0 : pi/4 : 2 pi, stored by MATLAB as a 9 x 1 dataset of doubles"""
    got = host_rhs.read_hdf5_variable(_matlab_sample(), "testdouble")
    assert got.shape == (9, 1)
    assert np.array_equal(got[:, 0], np.arange(9) * (np.pi / 4))
    with pytest.raises(AmpeError, match="Could not read variable 'nosuch'"):
        host_rhs.read_hdf5_variable(_matlab_sample(), "nosuch")


def test_checksum_of_the_writer_is_jenkins_lookup3():
    """published test values of hashlittle (lookup3.c, driver5)"""
    assert hdf5_writer.lookup3(b"") == 0xDEADBEEF
    assert hdf5_writer.lookup3(b"", 0xDEADBEEF) == 0xBD5B7DDE
    assert hdf5_writer.lookup3(b"Four score and seven years ago") == 0x17770551
    assert hdf5_writer.lookup3(b"Four score and seven years ago", 1) == 0xCD628161


RNG = np.random.default_rng(7)
A = RNG.random((5, 12, 9))

DIALECTS = [
    dict(style="old"),
    dict(style="old", continuation=True, userblock=512),
    dict(style="new"),
    dict(style="new", times=False, track_order=False),
    dict(style="new", continuation=True, userblock=1024),
    dict(style="new", dense=True),
    dict(style="new", dense=True, indirect_root=True),
]


@pytest.mark.parametrize("dialect", DIALECTS, ids=lambda d: "-".join("%s=%s" % kv for kv in d.items()))
def test_group_dialects(tmp_path, dialect):
    """every way the root group of a NetCDF-4 file can list its objects; eleven objects, so the dense forms hold what a
    phase + four quaternion components + composition + temperature file with its three dimensions holds"""
    names = ["phase", "quat1", "quat2", "quat3", "quat4", "concentration0", "temperature", "eta"]
    variables = {n: (A + i).astype(np.float64 if i % 2 else np.float32) for i, n in enumerate(names)}
    path = str(tmp_path / "d.nc")
    hdf5_writer.write_hdf5(path, variables, dimensions={"z": 5, "y": 12, "x": 9}, **dialect)
    for n, a in variables.items():
        got = host_rhs.read_hdf5_variable(path, n)
        assert got.shape == a.shape and np.array_equal(got, a.astype(np.float64)), n
    with pytest.raises(AmpeError, match="Could not read variable 'quat5'"):
        host_rhs.read_hdf5_variable(path, "quat5")
    with pytest.raises(AmpeError, match="no data written"):
        host_rhs.read_hdf5_variable(path, "x")  # a dimension without a variable has no storage


STORAGE = [
    dict(layout="compact"),
    dict(layout="contiguous"),
    dict(layout="chunked", chunks=(5, 12, 9)),
    dict(layout="chunked", chunks=(2, 5, 4)),  # ragged edge chunks in every direction
    dict(layout="chunked", chunks=(2, 5, 4), leaf_fanout=4),  # two-level chunk B-tree
    dict(layout="chunked", chunks=(2, 5, 4), filters=["deflate"]),
    dict(layout="chunked", chunks=(3, 4, 9), filters=["shuffle", "deflate"]),  # zlib=True, shuffle=True of netCDF4-python
    dict(layout="chunked", chunks=(3, 4, 9), filters=["shuffle", "deflate", "fletcher32"]),
    dict(layout="chunked", chunks=(3, 4, 9), filters=["shuffle", "deflate"], skip_deflate_on_first=True),
    dict(layout="chunked", chunks=(1, 12, 9), filters=["fletcher32"], filter_version=1),
]


@pytest.mark.parametrize("style", ["old", "new"])
@pytest.mark.parametrize("dtype", ["<f8", "<f4", ">f8", ">f4"])
@pytest.mark.parametrize("opt", STORAGE, ids=lambda o: "-".join(str(v) for v in o.values()))
def test_storage_layouts_and_filters(tmp_path, style, dtype, opt):
    a = A.astype(dtype)
    path = str(tmp_path / "s.nc")
    hdf5_writer.write_hdf5(path, {"phase": (a, opt), "one_d": (a[0, 0], {}), "two_d": (a[0], {})}, style=style)
    assert np.array_equal(host_rhs.read_hdf5_variable(path, "phase"), a.astype(np.float64))
    assert np.array_equal(host_rhs.read_hdf5_variable(path, "one_d"), a[0, 0].astype(np.float64))
    assert np.array_equal(host_rhs.read_hdf5_variable(path, "two_d"), a[0].astype(np.float64))


@pytest.mark.parametrize("name,dtype,opt", [
    ("auni3d", np.float64, dict(layout="contiguous")),
    ("auni3d", np.float32, dict(layout="chunked", chunks=(3, 5, 7), filters=["shuffle", "deflate"])),
    ("auni2d", np.float64, dict(layout="chunked", chunks=(1, 8, 8), filters=["deflate"])),
    ("dendrite2d", np.float32, dict(layout="contiguous")),
])
def test_initial_conditions_from_a_netcdf4_container(tmp_path, name, dtype, opt):
    """the initial-condition entry point recognises the container by its signature and fills the state vector from it:
    double files bit for bit, float files as the float-rounded state; nine to ten objects in the file, so the root group is dense"""
    cfg, y = _state(name)
    variables = {k: (v, opt) for k, v in _file_vars(cfg, y, dtype).items()}
    nz, ny, nx = next(iter(variables.values()))[0].shape
    path = str(tmp_path / "init.nc")
    hdf5_writer.write_hdf5(path, variables, dimensions={"z": nz, "y": ny, "x": nx, "qlen": cfg.qlen})
    got = host_rhs.read_initial_conditions(path, cfg)
    for k in ("phase", "quat", "conc", "temperature"):
        if y.get(k) is None or (k == "temperature" and not cfg.with_unsteady_temperature):
            continue
        expect = y[k].astype(dtype).astype(np.float64)
        assert got[k] is not None and np.array_equal(got[k].numpy().reshape(expect.shape), expect), k


def test_slab_ranks_and_slices_from_a_chunked_container(tmp_path):
    """every rank reads the planes of its slab (chunks cut by the slab boundary are read by both neighbours); a 2D run
    reads one z-slice of a 3D file (FieldsInitializer.cc:229-236)"""
    cfg, y = _state("auni3d")
    opt = dict(layout="chunked", chunks=(5, 4, 6), filters=["shuffle", "deflate"])
    variables = {k: (v, opt) for k, v in _file_vars(cfg, y, np.float64).items()}
    path = str(tmp_path / "init.nc")
    hdf5_writer.write_hdf5(path, variables, style="new")
    nz = cfg.n[2]
    nranks = 3 if nz % 3 == 0 else 2
    assert nz % nranks == 0
    parts = []
    for r in range(nranks):
        c, _ = parity.make_case("auni3d")
        c.n[2] = nz // nranks
        c.rank, c.nranks = r, nranks
        parts.append(host_rhs.read_initial_conditions(path, c)["phase"].numpy())
    assert np.array_equal(np.concatenate(parts, axis=0).reshape(y["phase"].shape), y["phase"])
    cfg2, y2 = _state("auni2d")
    stack = np.stack([y2["phase"].reshape(cfg2.n[1], cfg2.n[0]) * (1 + k) for k in range(3)])
    fv = {"phase": (stack, opt)}
    q = y2["quat"].reshape(cfg2.qlen, cfg2.n[1], cfg2.n[0])
    for m in range(cfg2.qlen):
        fv["quat%d" % (m + 1)] = (np.stack([q[m]] * 3), opt)
    fv["concentration"] = (np.stack([y2["conc"].reshape(cfg2.n[1], cfg2.n[0])] * 3), opt)
    hdf5_writer.write_hdf5(path, fv, style="old")
    for sl, want in ((-1, 1), (0, 0), (2, 2)):
        got = host_rhs.read_initial_conditions(path, cfg2, slice_index=sl)["phase"].numpy()
        assert np.array_equal(got.reshape(stack[want].shape), stack[want])


def test_errors(tmp_path):
    cfg, y = _state("auni3d")
    variables = _file_vars(cfg, y, np.float64)
    path = str(tmp_path / "bad.nc")
    # the reference's dimension check (FieldsInitializer.cc:685-707)
    wrong = {k: v[:, :, :-1] for k, v in variables.items()}
    hdf5_writer.write_hdf5(path, wrong)
    with pytest.raises(AmpeError, match="Phase input data dimensions are incorrect"):
        host_rhs.read_initial_conditions(path, cfg)
    missing = {k: v for k, v in variables.items() if k != "quat2"}
    hdf5_writer.write_hdf5(path, missing)
    with pytest.raises(AmpeError, match="qlen_file|Could not read variable 'quat2'"):
        host_rhs.read_initial_conditions(path, cfg)
    # truncated file: the data of the last variable is cut off
    hdf5_writer.write_hdf5(path, variables, style="old")
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:200])
    with pytest.raises(AmpeError, match="truncated HDF5 file|corrupt HDF5"):
        host_rhs.read_initial_conditions(path, cfg)
    # structures of the HDF5 1.10 "latest" format are named, not guessed at
    a = np.zeros((2, 2, 2))
    hdf5_writer.write_hdf5(path, {"phase": a}, style="new")
    raw = bytearray(open(path, "rb").read())
    i = raw.index(bytes([0x08, 18, 0, 0]))  # the layout message of the only dataset: type 8, 18 bytes
    raw[i + 6] = 4  # layout version 4 ...
    raw[i + 7] = 2  # ... chunked: an index type of the new format follows
    open(path, "wb").write(bytes(raw))
    with pytest.raises(AmpeError, match="latest format"):
        host_rhs.read_hdf5_variable(path, "phase")


@pytest.mark.skipif(not os.path.exists("/root/reference/utils/make_nuclei.py"), reason="needs the reference's generator (build container only)")
def test_reference_generator_to_container_to_state_vector(tmp_path, monkeypatch):
    """the reference's own utils/make_nuclei.py, run unmodified with the command line of tests/SingleGrainGrowthAuNi/test2d.py:11-15,
    writes through a stand-in `netCDF4.Dataset` whose close() lays the variables out as a NetCDF-4 container (dimension
    datasets z, y, x without storage, one single-precision dataset per variable); the initial-condition entry point turns
    that file into the state vector the committed fixture of the deck holds"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_reference_nuclei as gen
    path = str(tmp_path / "64x64.nc")

    def close(self):
        hdf5_writer.write_hdf5(path, dict(self.vars), dimensions=dict(self.dims))

    monkeypatch.setattr(gen._Dataset, "close", close)
    fields = gen.run(gen.DECKS["single_grain_auni"])
    assert os.path.exists(path) and set(fields) >= {"phase", "concentration0"}
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "ic_single_grain_auni.npz"))
    for k in fields:
        got = host_rhs.read_hdf5_variable(path, k)
        assert got.shape == (1, 64, 64) and np.array_equal(got, golden[k].astype(np.float64)), k


def test_damaged_files_fail_with_a_message(tmp_path):
    """truncations and flipped bytes anywhere in the three group dialects: the reader answers with the data or with an
    error, never with a crash, a hang or an allocation sized by a corrupt field"""
    rng = np.random.default_rng(1)
    a = rng.random((4, 6, 5))
    path = str(tmp_path / "f.nc")
    answered = failed = 0
    for style, kw in (("old", {}), ("new", {}), ("new", dict(dense=True, indirect_root=True))):
        hdf5_writer.write_hdf5(path, {"phase": (a, dict(layout="chunked", chunks=(2, 3, 5), filters=["shuffle", "deflate"])), "q": a},
                               style=style, dimensions={"x": 5}, **kw)
        raw = open(path, "rb").read()
        for trial in range(150):
            b = bytearray(raw)
            if trial % 3 == 0:
                b = b[:rng.integers(0, len(b))]
            else:
                for _ in range(rng.integers(1, 4)):
                    b[rng.integers(0, len(b))] = rng.integers(0, 256)
            open(path, "wb").write(bytes(b))
            for v in ("phase", "q"):
                try:
                    host_rhs.read_hdf5_variable(path, v)
                    answered += 1
                except (AmpeError, ValueError, MemoryError):  # ValueError / MemoryError: numpy refusing a corrupt shape
                    failed += 1
    assert answered > 100 and failed > 100


def test_an_unreadable_object_only_matters_when_it_is_asked_for(tmp_path):
    """a variable stored in a form outside the subset (here: its datatype message turned into a shared-message reference, and
    a layout of the HDF5 1.10 format) does not make the rest of the file unreadable"""
    a = np.arange(24.0).reshape(2, 3, 4)
    path = str(tmp_path / "f.nc")
    hdf5_writer.write_hdf5(path, {"phase": a, "other": a + 1.0}, style="new")
    raw = bytearray(open(path, "rb").read())
    first = raw.index(bytes([0x03, 20, 0, 0]))      # the datatype message of the first dataset written ("phase"): type 3, 20 bytes
    raw[first + 3] |= 2                               # message flag "shared"
    open(path, "wb").write(bytes(raw))
    assert np.array_equal(host_rhs.read_hdf5_variable(path, "other"), a + 1.0)
    with pytest.raises(AmpeError, match="shared header messages"):
        host_rhs.read_hdf5_variable(path, "phase")
    with pytest.raises(AmpeError, match="Could not read variable 'absent'"):
        host_rhs.read_hdf5_variable(path, "absent")
