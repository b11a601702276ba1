"""The oracle must reproduce its committed golden vectors bit for bit (same compiler
flags: -O2 -ffp-contract=off); libm differences between hosts may move the last bits,
so the comparison allows 1e-13 with the parity metric."""
import os

import numpy as np
import pytest
import torch

import parity

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden(name):
    g = np.load(os.path.join(HERE, "golden", "oracle_%s.npz" % name))
    cfg, _ = parity.make_case(name)
    st = {k: (torch.from_numpy(g["in_" + k].copy()) if ("in_" + k) in g else None)
          for k in ("phase", "quat", "conc", "temperature")}
    rot = [g["rot%d" % d] for d in range(cfg.ndim)] if cfg.symmetry_aware else None
    return cfg, st, rot, g


@pytest.mark.parametrize("name", list(parity.SMALL))
def test_oracle_reproduces_golden(name):
    cfg, st, rot, g = load_golden(name)
    outs, extra = parity.run_oracle(cfg, st, fd_flags=(0,), rotations=rot)
    status, yd = outs[0]
    assert status == 0
    for k, v in yd.items():
        if v is None or ("ydot_" + k) not in g:
            continue
        assert parity.rel_err(v, g["ydot_" + k]) < 1e-13, k
    if extra is not None:
        assert parity.rel_err(extra[0], g["cl"]) < 1e-13
        assert parity.rel_err(extra[1], g["ca"]) < 1e-13


@pytest.mark.parametrize("name", ["dendrite2d", "auni2d", "gg3d_hbsm", "auni3d"])
def test_oracle_properties(name):
    """size-independent properties of the RHS (SURVEY.md 8c oracle plan):
    q . ydot_q = 0 (projection, doc/latex/manual/model.tex:406-412) and
    sum ydot_c = 0 under periodic BC (flux form)."""
    cfg, st, rot, g = load_golden(name)
    q = g["in_quat"].reshape(cfg.qlen, -1)
    yq = g["ydot_quat"].reshape(cfg.qlen, -1)
    dot = np.abs((q * yq).sum(0))
    scale = np.abs(yq).max() + 1e-300
    assert dot.max() / scale < 1e-10
    if "ydot_conc" in g:
        yc = g["ydot_conc"]
        assert abs(yc.sum()) / (np.abs(yc).sum() + 1e-300) < 1e-10


def test_fd_flag_lagging_semantics():
    """QuatIntegrator.cc:3183-3189: with lag_quat_sidegrad the face coefficient keeps
    |grad q| of the last fd_flag=0 state; evaluating a perturbed state with fd_flag=1
    must differ from a full re-evaluation, and fd_flag=1 on the SAME state must not."""
    from oracle import pyoracle
    cfg, st = parity.make_case("dendrite2d")
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    _, r0 = o.eval(0.0, y, 0)
    _, r1 = o.eval(0.0, y, 1)
    assert np.array_equal(r0["quat"], r1["quat"])
    y2 = dict(y)
    ang = np.arctan2(y["quat"][1], y["quat"][0]) + 1e-3 * np.sin(np.arange(y["quat"][0].size)).reshape(y["quat"][0].shape)
    y2["quat"] = np.stack([np.cos(ang), np.sin(ang)]).copy()
    _, lag = o.eval(0.0, y2, 1)
    _, full = o.eval(0.0, y2, 0)
    assert not np.array_equal(lag["quat"], full["quat"])
    # phase RHS does not use the lagged data
    assert np.array_equal(lag["phase"], full["phase"])
