import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, parity
from ampe_b200 import configs, rhs
from ampe_b200.halo import slab_planes, slab_dim
name = sys.argv[1] if len(sys.argv) > 1 else "pfhub1a"
cfg, st = parity.make_case(name)
cfg.symmetry_aware = 0
ndim = cfg.ndim
y = rhs.to_device(st)
r = rhs.QuatIntegratorRHS(cfg)
kks = cfg.conc_rhs_form in (2, 3)
if kks:
    c0 = y["conc"].reshape(-1).clone(); r.resetRefPhaseConcentrations(c0, c0.clone())
ref = y.like(); r.evaluateRHSFunction(0.0, y, ref, 0); torch.cuda.synchronize()
ns = cfg.n[ndim - 1]; half = ns // 2; ng = r.nghosts(); dim = slab_dim(ndim)
print("ns", ns, "ng", ng, "n", list(cfg.n))
for rank in (0, 1):
    lo_i = rank * half; hi_i = (rank + 1) * half if rank == 0 else ns
    kw = dict(nx=cfg.n[0], ny=cfg.n[1])
    if ndim == 3: kw["nz"] = hi_i - lo_i
    else: kw["ny"] = hi_i - lo_i
    c2 = configs.BUILDERS[name](**kw)
    for d in range(3): c2.dx[d] = cfg.dx[d]
    c2.symmetry_aware = 0; c2.nranks, c2.rank = 2, rank
    print("rank", rank, "n", list(c2.n), "dx", list(c2.dx), list(cfg.dx))
    take = lambda t, idx: t.index_select(dim, torch.tensor([i % ns for i in idx], device=t.device)).contiguous()
    ys = rhs.SolutionVector({k: (None if v is None else slab_planes(v, ndim, slice(lo_i, hi_i)).contiguous()) for k, v in y.items()})
    lo = rhs.SolutionVector({k: (None if v is None else take(v, range(lo_i - ng, lo_i))) for k, v in y.items()})
    hi = rhs.SolutionVector({k: (None if v is None else take(v, range(hi_i, hi_i + ng))) for k, v in y.items()})
    r2 = rhs.QuatIntegratorRHS(c2); r2.setHalo(lo, hi)
    if kks:
        g = take(y["conc"], range(lo_i - ng, hi_i + ng)); r2.setRefPhaseConcentrationsGhosted(g, g.clone())
    out = ys.like(); r2.evaluateRHSFunction(0.0, ys, out, 0); torch.cuda.synchronize()
    for k, v in ref.items():
        if v is None: continue
        e = slab_planes(v, ndim, slice(lo_i, hi_i))
        d = (out[k] - e).abs()
        red = [i for i in range(d.dim()) if i != d.dim() + dim]
        print(k, "max diff per plane", d.amax(dim=red).cpu().numpy())
