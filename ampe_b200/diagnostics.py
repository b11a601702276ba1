"""Combining the per-rank scalar diagnostics of a slab-partitioned run (ampe_scalar_diagnostics works on this
rank's cells): what the reference's HierarchyCellDataOpsReal reductions do with MPI_Allreduce inside
QuatModel::printScalarDiagnostics (QuatModel.cc:2543-2690).  Sums for the integrals, max / min for the extrema,
the ratios recomputed from the global sums."""
import torch
import torch.distributed as dist

_SUM = ("volume", "volume_solid", "integral_concentration", "integral_phase_concentration", "thermal_energy")


def combine_scalar_diagnostics(local, group=None, device=None):
    """local: the dict of QuatIntegratorRHS.printScalarDiagnostics on this rank -> the dict of the whole domain
    (every rank gets it).  Without an initialised process group the input is returned unchanged.
    device: where the reduction buffers live; default = what the group's backend can reduce (the current CUDA
    device for NCCL, the host otherwise)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    sums = torch.tensor([local[k] for k in _SUM] + [local["average_temperature"] * local["volume"]],
                        dtype=torch.float64, device=device)
    mx = torch.tensor([local["max_concentration"], local["max_temperature"], -local["min_temperature"]],
                      dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out = {k: float(sums[i]) for i, k in enumerate(_SUM)}
    vol = out["volume"]
    out["solid_fraction"] = out["volume_solid"] / vol
    out["max_concentration"], out["max_temperature"], out["min_temperature"] = float(mx[0]), float(mx[1]), -float(mx[2])
    out["average_temperature"] = float(sums[len(_SUM)]) / vol
    c0V0 = out["integral_concentration"]
    # Cex = (cphi - c0 vphi) / c0V0 with c0 = c0V0 / vol (QuatModel.cc:2649-2657); 0 without a composition field
    out["cex"] = (out["integral_phase_concentration"] - c0V0 / vol * out["volume_solid"]) / c0V0 if c0V0 != 0.0 else 0.0
    return out
