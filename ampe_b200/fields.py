"""Deterministic synthetic state fields, following the recipes of
the reference's utils/make_nuclei.py (tanh grain profiles, one orientation per
grain) and benchmarks/PFHub1a/make_initial.py (closed-form spinodal IC).

Arrays are returned in SAMRAI CellData ghost-0 order: torch tensors of shape
(depth, nz, ny, nx) / (nz, ny, nx) whose memory is i-fastest, component-slowest.
Works on CPU or CUDA tensors (device argument)."""
import math

import numpy as np
import torch

SEED = 20240517


def _coords(cfg, device, k0=0, nz_global=None):
    nx, ny = cfg.n[0], cfg.n[1]
    nz = cfg.n[2] if cfg.ndim == 3 else 1
    x = (torch.arange(nx, device=device, dtype=torch.float64) + 0.5)
    y = (torch.arange(ny, device=device, dtype=torch.float64) + 0.5)
    z = (torch.arange(nz, device=device, dtype=torch.float64) + 0.5 + k0)
    return x, y, z


def pfhub1a_conc(cfg, device="cpu"):
    """benchmarks/PFHub1a/make_initial.py:66-87 (c0=0.5, eps=0.01), in double."""
    nx, ny = cfg.n[0], cfg.n[1]
    x = (torch.arange(nx, device=device, dtype=torch.float64) + 0.5) * cfg.dx[0]
    y = (torch.arange(ny, device=device, dtype=torch.float64) + 0.5) * cfg.dx[1]
    X = x[None, :]
    Y = y[:, None]
    t1 = torch.cos(0.105 * X) * torch.cos(0.11 * Y)
    t2 = torch.cos(0.13 * X) * torch.cos(0.087 * Y)
    t3 = torch.cos(0.025 * X - 0.15 * Y) * torch.cos(0.07 * X - 0.02 * Y)
    return (0.5 + 0.01 * (t1 + t2 * t2 + t3)).reshape(1, ny, nx).contiguous()


def grains(cfg, ngrains, radius_cells, delta_cells=3.0, device="cpu", seed=SEED,
           slab=(0, 1)):
    """phi = 1/2 [1 - tanh((r-R)/(sqrt(2) delta))] around `ngrains` seeded centres
    (periodic distance), one random orientation per grain (Voronoi assigned).
    slab=(rank, nranks): this rank's planes of a global domain that is
    nranks x cfg.n[last] planes thick.  Returns phi (nz,ny,nx), q (qlen,nz,ny,nx)."""
    rank, nranks = slab
    D = cfg.ndim
    nx, ny = cfg.n[0], cfg.n[1]
    nzl = cfg.n[2] if D == 3 else 1
    glob = [nx, ny * (nranks if D == 2 else 1), (nzl * nranks) if D == 3 else 1]
    nyl = ny
    rng = np.random.default_rng(seed)
    centres = rng.random((ngrains, 3)) * np.array(glob, dtype=np.float64)
    Q = max(cfg.qlen, 1)
    if cfg.qlen == 4:
        qs = rng.normal(size=(ngrains, 4))
        qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    elif cfg.qlen == 2:
        ang = rng.random(ngrains) * 0.5 * math.pi
        qs = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    else:
        qs = rng.random((ngrains, Q))
    x = torch.arange(nx, device=device, dtype=torch.float64) + 0.5
    y = torch.arange(nyl, device=device, dtype=torch.float64) + 0.5 + (rank * ny if D == 2 else 0)
    z = torch.arange(nzl, device=device, dtype=torch.float64) + 0.5 + (rank * nzl if D == 3 else 0)
    dmin = torch.full((nzl, nyl, nx), 1.0e30, device=device, dtype=torch.float64)
    owner = torch.zeros((nzl, nyl, nx), device=device, dtype=torch.int64)
    for g in range(ngrains):
        dxg = torch.abs(x - centres[g, 0])
        dxg = torch.minimum(dxg, glob[0] - dxg)
        dyg = torch.abs(y - centres[g, 1])
        dyg = torch.minimum(dyg, glob[1] - dyg)
        d2 = dxg[None, None, :] ** 2 + dyg[None, :, None] ** 2
        if D == 3:
            dzg = torch.abs(z - centres[g, 2])
            dzg = torch.minimum(dzg, glob[2] - dzg)
            d2 = d2 + dzg[:, None, None] ** 2
        d2 = d2.expand(nzl, nyl, nx)
        closer = d2 < dmin
        dmin = torch.where(closer, d2, dmin)
        owner = torch.where(closer, torch.full_like(owner, g), owner)
    r = torch.sqrt(dmin)
    phi = 0.5 * (1.0 - torch.tanh((r - radius_cells) / (math.sqrt(2.0) * delta_cells)))
    qtab = torch.tensor(qs, device=device, dtype=torch.float64)
    q = qtab[owner.reshape(-1)].T.reshape(Q, nzl, nyl, nx).contiguous()
    if cfg.qlen == 0:
        q = None
    return phi.contiguous(), q


def smooth_unit(q, passes=1):
    """one periodic 5/7-point smoothing pass + renormalisation (keeps |q|=1)."""
    for _ in range(passes):
        s = 0.5 * q
        dims = [d for d in (1, 2, 3) if q.shape[d] > 1]
        w = 0.5 / (2 * len(dims))
        for d in dims:
            s = s + w * (torch.roll(q, 1, d) + torch.roll(q, -1, d))
        q = s / torch.sqrt((s * s).sum(0, keepdim=True))
    return q.contiguous()


def h_pbg(phi):
    p = phi.clamp(0.0, 1.0)
    return p * p * p * (10.0 - 15.0 * p + 6.0 * p * p)


def smooth_noise(shape, amp, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    nz, ny, nx = shape
    # low-wavenumber cosine mixture (deterministic, periodic)
    ph = torch.rand((6, 3), generator=g, dtype=torch.float64) * 2 * math.pi
    kk = torch.randint(1, 5, (6, 3), generator=g)
    x = torch.arange(nx, device=device, dtype=torch.float64) / nx
    y = torch.arange(ny, device=device, dtype=torch.float64) / ny
    z = torch.arange(nz, device=device, dtype=torch.float64) / max(nz, 1)
    out = torch.zeros(shape, device=device, dtype=torch.float64)
    for t in range(6):
        out = out + (torch.cos(2 * math.pi * int(kk[t, 0]) * x + float(ph[t, 0]))[None, None, :] *
                     torch.cos(2 * math.pi * int(kk[t, 1]) * y + float(ph[t, 1]))[None, :, None] *
                     torch.cos(2 * math.pi * int(kk[t, 2]) * z + float(ph[t, 2]))[:, None, None])
    return amp * out / 6.0


def make_state(name, cfg, device="cpu", slab=(0, 1), seed=SEED):
    """Synthetic y = dict(phase, quat, conc, temperature) for config `name`."""
    D = cfg.ndim
    nx, ny = cfg.n[0], cfg.n[1]
    nz = cfg.n[2] if D == 3 else 1
    st = {"phase": None, "quat": None, "conc": None, "temperature": None}
    if name == "pfhub1a":
        st["conc"] = pfhub1a_conc(cfg, device).reshape(1, ny, nx)
        return st
    if name == "dendrite2d":
        R = 50.0 * nx / 1024.0
        phi, q = grains(cfg, 1, max(R, 4.0), 3.0, device, seed, slab)
        # one nucleus in the middle + a weak orientation modulation so that grad q != 0
        ang = 0.3 + 0.2 * smooth_noise((nz, ny, nx), 1.0, device, seed + 1)
        q = torch.stack([torch.cos(ang), torch.sin(ang)], 0)
        st["phase"], st["quat"] = phi, q.contiguous()
        st["temperature"] = (1.0 * phi + 0.5 * (1.0 - phi)).contiguous()
        return st
    if name in ("auni2d", "auni3d"):
        if name == "auni2d":
            G = max(1, int(round(9 * (nx / 512.0) * (ny / 512.0))))
            R = 22.0 if nx >= 256 else max(3.0, nx / 12.0)
        else:
            G = 8
            R = 20.0 if nx >= 128 else max(3.0, nx / 6.0)
        phi, q = grains(cfg, G, R, 3.0, device, seed, slab)
        q = smooth_unit(q)
        c_in, c_out = 0.096, 0.25
        h = h_pbg(phi)
        conc = c_in * h + c_out * (1.0 - h) + smooth_noise((nz, ny, nx), 1.0e-3, device, seed + 2)
        st["phase"], st["quat"], st["conc"] = phi, q, conc.contiguous()
        return st
    if name == "gg3d_hbsm":
        G = 64 if nx >= 256 else 8
        R = 12.0 if nx >= 128 else max(3.0, nx / 8.0)
        phi, q = grains(cfg, G, R, 3.0, device, seed, slab)
        q = smooth_unit(q)
        c_in, c_out = 0.1, 0.06
        h = h_pbg(phi)
        conc = c_in * h + c_out * (1.0 - h) + smooth_noise((nz, ny, nx), 1.0e-3, device, seed + 2)
        st["phase"], st["quat"], st["conc"] = phi, q, conc.contiguous()
        return st
    raise ValueError(name)
