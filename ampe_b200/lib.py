"""Loader of the C-ABI shared library libampe_b200.so (include/ampe_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is
present, the calls fail loudly."""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMPE_B200_LIB", os.path.join(_HERE, "libampe_b200.so"))

_lib = None


class AmpeError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AmpeError(
            "libampe_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dbl, ci = C.c_void_p, C.c_double, C.c_int
    pcfg, pf = C.POINTER(_abi.RhsConfig), C.POINTER(_abi.RhsFields)
    L.ampe_rhs_create.restype = ci
    L.ampe_rhs_create.argtypes = [pcfg, C.POINTER(vp)]
    L.ampe_rhs_destroy.argtypes = [vp]
    L.ampe_rhs_set_ref_concentrations.restype = ci
    L.ampe_rhs_set_ref_concentrations.argtypes = [vp, vp, vp, vp]
    L.ampe_rhs_set_ref_concentrations_ghosted.restype = ci
    L.ampe_rhs_set_ref_concentrations_ghosted.argtypes = [vp, vp, vp, vp]
    L.ampe_rhs_set_symmetry_rotations.restype = ci
    L.ampe_rhs_set_symmetry_rotations.argtypes = [vp, C.POINTER(vp), vp]
    L.ampe_rhs_set_halo.restype = ci
    L.ampe_rhs_set_halo.argtypes = [vp, pf, pf]
    L.ampe_rhs_nghosts.restype = ci
    L.ampe_rhs_nghosts.argtypes = [vp]
    for name in ("ampe_rhs_eval", "ampe_rhs_eval_interior", "ampe_rhs_eval_boundary"):
        fn = getattr(L, name)
        fn.restype = ci
        fn.argtypes = [vp, dbl, pf, pf, ci, vp]
    L.ampe_rhs_eval_host.restype = ci
    L.ampe_rhs_eval_host.argtypes = [vp, dbl, pf, pf, ci]
    L.ampe_rhs_get_phase_concentrations.restype = ci
    L.ampe_rhs_get_phase_concentrations.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.ampe_rhs_copy_phase_concentrations.restype = ci
    L.ampe_rhs_copy_phase_concentrations.argtypes = [vp, vp, vp, vp]
    L.ampe_rhs_newton_failures.restype = ci
    L.ampe_rhs_newton_failures.argtypes = [vp, vp]
    L.ampe_rhs_last_launch_count.restype = ci
    L.ampe_rhs_last_launch_count.argtypes = [vp]
    # slab ghost-plane exchange (csrc/halo.cu)
    L.ampe_halo_create.restype = ci
    L.ampe_halo_create.argtypes = [vp, ci, ci, C.POINTER(vp)]
    L.ampe_halo_export.restype = ci
    L.ampe_halo_export.argtypes = [vp, vp]
    L.ampe_halo_connect.restype = ci
    L.ampe_halo_connect.argtypes = [vp, vp, vp]
    L.ampe_halo_destroy.restype = ci
    L.ampe_halo_destroy.argtypes = [vp]
    L.ampe_rhs_eval_slab.restype = ci
    L.ampe_rhs_eval_slab.argtypes = [vp, vp, dbl, pf, pf, ci, vp]
    L.ampe_rhs_eval_slab_host.restype = ci
    L.ampe_rhs_eval_slab_host.argtypes = [vp, vp, dbl, pf, pf, ci]
    L.ampe_halo_push.restype = ci
    L.ampe_halo_push.argtypes = [vp, pf, ci, vp]
    L.ampe_halo_wait.restype = ci
    L.ampe_halo_wait.argtypes = [vp, vp]
    L.ampe_rhs_set_ref_concentrations_slab.restype = ci
    L.ampe_rhs_set_ref_concentrations_slab.argtypes = [vp, vp, vp, vp, vp]
    L.ampe_rhs_set_symmetry_rotations_slab.restype = ci
    L.ampe_rhs_set_symmetry_rotations_slab.argtypes = [vp, vp, C.POINTER(vp), vp]
    L.ampe_rhs_compute_symmetry_rotations_slab.restype = ci
    L.ampe_rhs_compute_symmetry_rotations_slab.argtypes = [vp, vp, pf, vp]
    L.ampe_integrate_fixed_slab.restype = ci
    L.ampe_integrate_fixed_slab.argtypes = [vp, vp, pf, pf, pf, dbl, dbl, ci, ci, vp]
    L.ampe_halo_last_launch_count.restype = ci
    L.ampe_halo_last_launch_count.argtypes = [vp]
    L.ampe_rhs_set_kernel_timing.restype = ci
    L.ampe_rhs_set_kernel_timing.argtypes = [vp, ci]
    L.ampe_rhs_last_kernel_ms.restype = ci
    L.ampe_rhs_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    pd = C.POINTER(C.c_double)
    L.ampe_vec_linear_sum.restype = ci
    L.ampe_vec_linear_sum.argtypes = [vp, dbl, pf, dbl, pf, pf, vp]
    L.ampe_vec_scale.restype = ci
    L.ampe_vec_scale.argtypes = [vp, dbl, pf, pf, vp]
    L.ampe_vec_dot.restype = ci
    L.ampe_vec_dot.argtypes = [vp, pf, pf, pd, vp]
    L.ampe_vec_wrms_norm.restype = ci
    L.ampe_vec_wrms_norm.argtypes = [vp, pf, pf, pd, vp]
    L.ampe_vec_max_norm.restype = ci
    L.ampe_vec_max_norm.argtypes = [vp, pf, pd, vp]
    L.ampe_vec_wdot.restype = ci
    L.ampe_vec_wdot.argtypes = [vp, pf, pf, pf, pd, vp]
    L.ampe_vec_error_weights.restype = ci
    L.ampe_vec_error_weights.argtypes = [vp, pf, dbl, dbl, pf, vp]
    L.ampe_normalize_quat.restype = ci
    L.ampe_normalize_quat.argtypes = [vp, pf, vp]
    L.ampe_integrate_fixed.restype = ci
    L.ampe_integrate_fixed.argtypes = [vp, pf, pf, pf, dbl, dbl, ci, ci, vp]
    L.ampe_energy_eval.restype = ci
    L.ampe_energy_eval.argtypes = [vp, pf, pd, vp]
    L.ampe_scalar_diagnostics.restype = ci
    L.ampe_scalar_diagnostics.argtypes = [vp, pf, pd, vp]
    L.ampe_grain_volumes.restype = ci
    L.ampe_grain_volumes.argtypes = [vp, pf, C.c_double, ci, vp, vp, vp, vp]
    L.ampe_grain_numbers.restype = ci
    L.ampe_grain_numbers.argtypes = [vp, vp]
    L.ampe_apply_projection.restype = ci
    L.ampe_apply_projection.argtypes = [vp, pf, pf, pf, vp]
    L.ampe_rhs_compute_symmetry_rotations.restype = ci
    L.ampe_rhs_compute_symmetry_rotations.argtypes = [vp, pf, vp]
    L.ampe_rhs_get_symmetry_rotations.restype = ci
    L.ampe_rhs_get_symmetry_rotations.argtypes = [vp, C.POINTER(vp), vp]
    L.ampe_quat_fundamental.restype = ci
    L.ampe_quat_fundamental.argtypes = [vp, pf, vp]
    # block preconditioners (include/ampe_b200_precond.h)
    pvp = C.POINTER(vp)
    L.ampe_mg_create.restype = ci
    L.ampe_mg_create.argtypes = [ci, C.POINTER(ci), pd, ci, pvp]
    L.ampe_mg_create_multi.restype = ci
    L.ampe_mg_create_multi.argtypes = [ci, C.POINTER(ci), pd, ci, ci, pvp]
    L.ampe_mg_num_components.restype = ci
    L.ampe_mg_num_components.argtypes = [vp]
    L.ampe_mg_destroy.restype = ci
    L.ampe_mg_destroy.argtypes = [vp]
    L.ampe_mg_set_elliptic.restype = ci
    L.ampe_mg_set_elliptic.argtypes = [vp, vp, ci, dbl, vp, ci, dbl, pvp, pvp, ci, dbl, dbl, vp]
    L.ampe_mg_set_quat.restype = ci
    L.ampe_mg_set_quat.argtypes = [vp, dbl, vp, ci, pvp, ci, vp]
    L.ampe_mg_solve.restype = ci
    L.ampe_mg_solve.argtypes = [vp, vp, vp, ci, ci, vp]
    L.ampe_mg_apply.restype = ci
    L.ampe_mg_apply.argtypes = [vp, vp, vp, vp]
    L.ampe_mg_set_zero_slope.restype = ci
    L.ampe_mg_set_zero_slope.argtypes = [vp, vp]
    L.ampe_mg_set_sweeps.restype = ci
    L.ampe_mg_set_sweeps.argtypes = [vp, ci, ci, ci]
    L.ampe_mg_num_levels.restype = ci
    L.ampe_mg_num_levels.argtypes = [vp]
    L.ampe_mg_level_extents.restype = ci
    L.ampe_mg_level_extents.argtypes = [vp, ci, C.POINTER(ci)]
    L.ampe_mg_copy_level.restype = ci
    L.ampe_mg_copy_level.argtypes = [vp, ci, ci, vp, vp]
    L.ampe_mg_last_launch_count.restype = ci
    L.ampe_mg_last_launch_count.argtypes = [vp]
    L.ampe_k_phasefacops_setc.restype = ci
    L.ampe_k_phasefacops_setc.argtypes = [ci, C.POINTER(ci), C.POINTER(ci), vp, ci, vp, ci, dbl, dbl, C.c_char_p,
                                          vp, ci, vp]
    L.ampe_last_error.restype = C.c_char_p
    L.ampe_version.restype = C.c_char_p
    L.ampe_abi_sizeof_config.restype = ci
    if L.ampe_abi_sizeof_config() != C.sizeof(_abi.RhsConfig):
        raise AmpeError("ABI mismatch between _abi.py and libampe_b200.so")
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().ampe_last_error().decode()
        raise AmpeError("%s failed (%d): %s" % (what, rc, msg))
