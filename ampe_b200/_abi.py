"""ctypes mirror of include/ampe_b200.h (struct layouts only, no computation).

Kept field-for-field in sync with the header; tests/test_abi.py checks
sizeof() against the value compiled into the library (ampe_abi_sizeof).
"""
import ctypes as C

AMPE_MAX_TC = 6

AMPE_OK = 0
AMPE_EINVAL = -1
AMPE_ECUDA = -2
AMPE_ENEWTON = -3
AMPE_ENOGPU = -4

FLUX_SIMPLE, FLUX_ISOTROPIC, FLUX_ANISOTROPIC = 0, 1, 2
CONC_NONE, CONC_CAHN_HILLIARD, CONC_KKS, CONC_EBS = 0, 1, 2, 3
FE_NONE, FE_BIASWELL, FE_CALPHAD, FE_QUADRATIC, FE_DELTAT = 0, 1, 2, 3, 4


class CalphadSpecies(C.Structure):
    _fields_ = [
        ("nintervals", C.c_int),
        ("Tc", C.c_double * (AMPE_MAX_TC + 1)),
        ("a", C.c_double * AMPE_MAX_TC),
        ("b", C.c_double * AMPE_MAX_TC),
        ("c", C.c_double * AMPE_MAX_TC),
        ("d2", C.c_double * AMPE_MAX_TC),
        ("d3", C.c_double * AMPE_MAX_TC),
        ("d4", C.c_double * AMPE_MAX_TC),
        ("d7", C.c_double * AMPE_MAX_TC),
        ("dm1", C.c_double * AMPE_MAX_TC),
        ("dm9", C.c_double * AMPE_MAX_TC),
    ]


class CalphadBinary(C.Structure):
    _fields_ = [
        ("g", (CalphadSpecies * 2) * 2),          # [species][phase]
        ("L", ((C.c_double * 2) * 4) * 2),        # [phase][k][2]
        ("qA", ((C.c_double * 2) * 2) * 2),       # [species][phase][2]
        ("qB", ((C.c_double * 2) * 2) * 2),
        ("qAB", (((C.c_double * 2) * 4) * 2) * 2),  # [species][phase][n][2]
        ("nqAB", (C.c_int * 2) * 2),
    ]


class RhsConfig(C.Structure):
    _fields_ = [
        ("ndim", C.c_int),
        ("n", C.c_int * 3),
        ("dx", C.c_double * 3),
        ("qlen", C.c_int),
        ("with_phase", C.c_int),
        ("with_concentration", C.c_int),
        ("with_unsteady_temperature", C.c_int),
        ("evolve_quat", C.c_int),
        ("phase_flux_type", C.c_int),
        ("conc_rhs_form", C.c_int),
        ("free_energy", C.c_int),
        ("symmetry_aware", C.c_int),
        ("lag_quat_sidegrad", C.c_int),
        ("quat_grad_modulus_from_cells", C.c_int),
        ("energy_interp", C.c_char),
        ("conc_interp", C.c_char),
        ("diffusion_interp", C.c_char),
        ("orient_interp1", C.c_char),
        ("orient_interp2", C.c_char),
        ("avg_func", C.c_char),
        ("grad_floor_type", C.c_char),
        ("quat_mobility_func", C.c_char),
        ("conc_avg_func", C.c_char),
        ("epsilon_phase", C.c_double),
        ("epsilon_anisotropy", C.c_double),
        ("knumber", C.c_int),
        ("phi_well_scale", C.c_double),
        ("phi_mobility", C.c_double),
        ("H_parameter", C.c_double),
        ("epsilon_q", C.c_double),
        ("quat_mobility", C.c_double),
        ("min_quat_mobility", C.c_double),
        ("quat_grad_floor", C.c_double),
        ("quat_mobility_alt_scale", C.c_double),
        ("T_uniform", C.c_double),
        ("thermal_diffusivity", C.c_double),
        ("latent_heat", C.c_double),
        ("cp", C.c_double),
        ("meltingT", C.c_double),
        ("bias_well_alpha", C.c_double),
        ("bias_well_gamma", C.c_double),
        ("conc_mobility", C.c_double),
        ("ch_ca", C.c_double),
        ("ch_cb", C.c_double),
        ("ch_well_scale", C.c_double),
        ("ch_kappa", C.c_double),
        ("ch_mobility", C.c_double),
        ("quad_Tref", C.c_double),
        ("quad_A_l", C.c_double),
        ("quad_Ceq_l", C.c_double),
        ("quad_m_l", C.c_double),
        ("quad_A_s", C.c_double),
        ("quad_Ceq_s", C.c_double),
        ("quad_m_s", C.c_double),
        ("D_liquid", C.c_double),
        ("D_solid", C.c_double),
        ("Q0_liquid", C.c_double),
        ("Q0_solid", C.c_double),
        ("vm_liquid", C.c_double),
        ("vm_solid", C.c_double),
        ("newton_max_its", C.c_int),
        ("newton_tol", C.c_double),
        ("newton_alpha", C.c_double),
        ("calphad", CalphadBinary),
        ("nranks", C.c_int),
        ("rank", C.c_int),
        ("zero_slope", C.c_int * 3),
        ("dtemperaturedt", C.c_double),
        ("target_temperature", C.c_double),
    ]


class RhsFields(C.Structure):
    _fields_ = [
        ("phase", C.c_void_p),
        ("quat", C.c_void_p),
        ("conc", C.c_void_p),
        ("temperature", C.c_void_p),
    ]
