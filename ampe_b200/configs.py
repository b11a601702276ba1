"""Parameter records of the five BASELINE.json configurations.

Each builder returns an `_abi.RhsConfig` whose fields follow
QuatModelParameters (reference: source/QuatModelParameters.cc) with the values of
the shipped input decks.  Grid sizes are arguments so that the parity tests can
run the same model on small grids.
"""
import json
import os

from . import _abi

_DATA = os.path.join(os.path.dirname(__file__), "data")


def _ch(s):
    return s.encode("ascii")[:1]


def load_calphad(name="calphadAuNi.json"):
    """thermodynamic_data/calphadAuNi.dat transcribed by tools/make_calphad_json.py"""
    return calphad_from_database(json.load(open(os.path.join(_DATA, name))))


def load_calphad_dat(path):
    """a binary CALPHAD data base in the reference's own format (thermodynamic_data/*.dat: a SAMRAI input file with the blocks
    SpeciesA / SpeciesB {PhaseL, PhaseA {Tc, a, b, c, d2 ...}}, LmixPhaseL / LmixPhaseA {L0..L3}, MobilityParameters) -- what
    ConcentrationModel{Calphad{filename}} names and CALPHADFreeEnergyFunctionsBinary reads"""
    from . import input_deck

    def lists(d):
        return {k: (lists(v) if isinstance(v, dict) else v if isinstance(v, str) else
                    [float(x) for x in (v if isinstance(v, list) else [v])]) for k, v in d.items()}
    return calphad_from_database(lists(input_deck.load(path)))


def calphad_from_database(db):
    out = _abi.CalphadBinary()
    for blk in ("SpeciesA", "SpeciesB", "LmixPhaseL", "LmixPhaseA", "MobilityParameters"):
        if blk not in db:
            raise ValueError("CALPHAD data base: block '%s' is missing (binary liquid / solid-A data bases only)" % blk)
    for si, sp in enumerate(("SpeciesA", "SpeciesB")):
        for pi, ph in enumerate(("PhaseL", "PhaseA")):
            rec = db[sp][ph]
            g = out.g[si][pi]
            tc = rec["Tc"]
            g.nintervals = len(tc) - 1
            assert g.nintervals <= _abi.AMPE_MAX_TC
            for i, v in enumerate(tc):
                g.Tc[i] = v
            for key in ("a", "b", "c", "d2", "d3", "d4", "d7", "dm1", "dm9"):
                vals = rec.get(key, [0.0] * g.nintervals)
                arr = getattr(g, key)
                for i in range(g.nintervals):
                    arr[i] = vals[i]
    for pi, ph in enumerate(("LmixPhaseL", "LmixPhaseA")):
        for k in range(4):
            v = db[ph].get("L%d" % k, [0.0, 0.0])
            out.L[pi][k][0] = v[0]
            out.L[pi][k][1] = v[1]
    mob = db["MobilityParameters"]
    for si in range(2):
        for pi, ph in enumerate(("PhaseL", "PhaseA")):
            rec = mob["Species%d" % si][ph]
            out.qA[si][pi][0], out.qA[si][pi][1] = rec["qA"]
            out.qB[si][pi][0], out.qB[si][pi][1] = rec["qB"]
            n = 0
            for k in range(4):
                # defaults (0, 1): CALPHADMobility.cc initialize()
                v = rec.get("q%dAB" % k, [0.0, 1.0])
                if ("q%dAB" % k) in rec:
                    n = k + 1
                out.qAB[si][pi][k][0] = v[0]
                out.qAB[si][pi][k][1] = v[1]
            out.nqAB[si][pi] = n
    return out


def _base(ndim, n, lo, hi):
    c = _abi.RhsConfig()
    c.ndim = ndim
    for d in range(3):
        c.n[d] = n[d] if d < ndim else 1
        c.dx[d] = (hi[d] - lo[d]) / n[d] if d < ndim else 1.0
    # defaults of QuatModelParameters.cc
    c.lag_quat_sidegrad = 1              # QuatIntegrator.cc:293-298
    c.quat_grad_modulus_from_cells = 1   # quat_grad_modulus_type = "cells" (:692-693)
    c.energy_interp = _ch("p")
    c.conc_interp = _ch("p")
    c.diffusion_interp = _ch("l")        # :1014-1017
    c.orient_interp1 = _ch("q")          # :704-707
    c.orient_interp2 = _ch("c")          # :708-711
    c.avg_func = _ch("h")                # :1039-1041 default "harmonic"
    c.conc_avg_func = _ch("h")
    c.grad_floor_type = _ch("m")         # :683-684
    c.quat_mobility_func = _ch("p")      # :730-731
    c.knumber = 4                        # PhaseFluxStrategyFactory.h:27-28
    c.min_quat_mobility = 1.0e-6         # :651
    c.quat_grad_floor = 1.0e-2           # :671-672
    c.quat_mobility_alt_scale = 1.0
    c.conc_mobility = 1.0                # :291
    c.ch_mobility = 1.0                  # CompositionRHSStrategyFactory.h:83-88
    c.newton_max_its = 20                # Thermo4PFM NewtonSolver defaults
    c.newton_tol = 1.0e-8
    c.newton_alpha = 1.0
    c.cp = 1.0
    c.vm_liquid = c.vm_solid = 1.0e-6
    c.nranks, c.rank = 1, 0
    return c


def pfhub1a(nx=200, ny=200):
    """C1: benchmarks/PFHub1a/2d.input -- Cahn-Hilliard double well, 200 um box.
    M = 5 applied as ConcentrationModel.mobility (tests/CahnHilliard/2d.input)."""
    c = _base(2, (nx, ny), (0.0, 0.0), (200.0, 200.0))
    c.with_concentration = 1
    c.conc_rhs_form = _abi.CONC_CAHN_HILLIARD
    c.T_uniform = 1000.0
    c.ch_ca, c.ch_cb, c.ch_well_scale, c.ch_kappa = 0.3, 0.7, 5.0, 2.0
    c.conc_mobility = 5.0
    return c


def dendrite2d(nx=2048, ny=2048):
    """C2: examples/Dendrite2D/dendrite.input -- KWCcomplex (qlen=2), anisotropic
    phase flux, bias double well, unsteady heat equation, periodic."""
    c = _base(2, (nx, ny), (-4.5, -4.5), (4.5, 4.5))
    c.qlen = 2
    c.with_phase = 1
    c.with_unsteady_temperature = 1
    c.evolve_quat = 1
    c.phase_flux_type = _abi.FLUX_ANISOTROPIC
    c.free_energy = _abi.FE_BIASWELL
    c.epsilon_anisotropy = 0.05
    c.H_parameter = 0.001
    c.epsilon_phase = 0.01
    c.phi_mobility = 3333.3333
    c.quat_mobility = 1.0
    c.epsilon_q = 1.0e3
    c.meltingT = 1.0
    c.cp = 1.0
    c.thermal_diffusivity = 1.0e-8 * 1.0e8   # cm^2/s -> um^2/s (:572-579)
    c.vm_liquid = c.vm_solid = 1.0e-6
    c.latent_heat = 1.0 * (1.0e-6 / 1.0e-6)  # J/mol -> pJ/um^3 (:611-617)
    c.phi_well_scale = 0.015625
    c.bias_well_alpha = 0.9
    c.bias_well_gamma = 10.0
    return c


def auni2d(nx=4096, ny=4096, symmetry=True):
    """C3: examples/AuNi_2D/9grains_AuNi.input -- phi + c + q[4], CALPHAD KKS,
    EBS composition RHS, symmetry-aware quaternions, T = 1450 K uniform."""
    c = _base(2, (nx, ny), (-1.6 * nx / 512, -1.6 * ny / 512), (1.6 * nx / 512, 1.6 * ny / 512))
    _auni_common(c)
    c.symmetry_aware = 1 if symmetry else 0
    return c


def auni3d(nx=1024, ny=1024, nz=128):
    """C5: examples/AuNi_3D/1grain3D_AuNi.input scaled (h kept at 0.8/128 um)."""
    h = 0.8 / 128
    c = _base(3, (nx, ny, nz), (0.0, 0.0, 0.0), (nx * h, ny * h, nz * h))
    _auni_common(c)
    return c


def _auni_common(c):
    c.qlen = 4
    c.with_phase = 1
    c.with_concentration = 1
    c.evolve_quat = 1
    c.phase_flux_type = _abi.FLUX_SIMPLE
    c.conc_rhs_form = _abi.CONC_EBS
    c.free_energy = _abi.FE_CALPHAD
    c.H_parameter = 0.25
    c.epsilon_q = 0.3125
    c.T_uniform = 1450.0
    c.epsilon_phase = 0.25
    c.phi_mobility = 6.4
    c.quat_mobility = 0.64
    c.phi_well_scale = 2.5
    c.energy_interp = _ch("p")
    c.conc_interp = _ch("p")
    c.avg_func = _ch("a")
    c.conc_avg_func = _ch("a")
    c.vm_liquid = c.vm_solid = 7.68e-6
    c.newton_max_its = 50
    c.calphad = load_calphad()


def gg3d_hbsm(nx=512, ny=512, nz=512):
    """C4: examples/GG3D_HBSM with the maintained parameter set of
    tests/TwoGrainsQuadratic/3d.input (H=0.001: quaternions evolve), quadratic
    KKS free energy, KKS composition RHS, T = 873 K, h = 0.05 um."""
    h = 3.2 / 64
    c = _base(3, (nx, ny, nz), (0.0, 0.0, 0.0), (nx * h, ny * h, nz * h))
    c.qlen = 4
    c.with_phase = 1
    c.with_concentration = 1
    c.evolve_quat = 1
    c.phase_flux_type = _abi.FLUX_SIMPLE
    c.conc_rhs_form = _abi.CONC_KKS
    c.free_energy = _abi.FE_QUADRATIC
    c.H_parameter = 0.001
    c.epsilon_q = 0.1
    c.T_uniform = 873.0
    c.epsilon_phase = 0.165
    c.quat_mobility = 200.0
    c.phi_mobility = 200.0
    c.phi_well_scale = 0.4125
    c.energy_interp = _ch("h")
    c.conc_interp = _ch("h")
    c.avg_func = _ch("a")
    c.conc_avg_func = _ch("a")
    c.vm_liquid = c.vm_solid = 1.5e-5
    c.D_solid, c.D_liquid = 1.3e8, 5.6e4
    c.Q0_solid, c.Q0_liquid = 156377.0, 55329.0
    c.quad_Tref = 873.0
    c.quad_A_l = c.quad_A_s = 1.0e4
    c.quad_Ceq_l, c.quad_Ceq_s = 0.05, 0.10
    c.quad_m_l = c.quad_m_s = 0.0
    return c


# ---- the reference's regression decks (tests/*/test*.py): model blocks as the decks give them, with the unit
# conversions of QuatModelParameters.cc (thermal diffusivity cm^2/s -> um^2/s: x 1e8, :576; latent heat and cp
# J/mol -> pJ/um^3: x 1e-6 / V_m, :557, :616)
def dendrite_test2d():
    """tests/Dendrite/2d.input: KWCcomplex (qlen 2, H_parameter unset -> 0: the orientation only feeds the
    anisotropy, evolveQuat() false), anisotropic phase flux, FreeEnergyModel "linear"
    (DeltaTemperatureFreeEnergyStrategy), unsteady heat equation in units of the melting temperature, slope-0
    boundaries on every side, 240 x 240 cells on 160 x 160 um."""
    c = _base(2, (240, 240), (0.0, 0.0), (160.0, 160.0))
    c.qlen = 2
    c.with_phase = 1
    c.with_unsteady_temperature = 1
    c.evolve_quat = 0
    c.phase_flux_type = _abi.FLUX_ANISOTROPIC
    c.free_energy = _abi.FE_DELTAT
    c.epsilon_anisotropy = 0.05
    c.H_parameter = 0.0
    c.epsilon_phase = 2.0
    c.phi_mobility = 0.25
    c.quat_mobility = 1.0
    c.epsilon_q = 0.0
    c.meltingT = 1.0
    vm = 1.0e-6
    c.vm_liquid = c.vm_solid = vm
    c.cp = 17.020371256848037 * (1.0e-6 / vm)
    c.latent_heat = 17.020371256848037 * (1.0e-6 / vm)
    c.thermal_diffusivity = 10.0e-8 * 1.0e8
    c.phi_well_scale = 0.25
    c.energy_interp = _ch("p")
    c.zero_slope[0] = c.zero_slope[1] = 1
    return c


def single_grain_auni_test2d():
    """tests/SingleGrainGrowthAuNi/2d.input: phase + composition (no orientation), CALPHAD KKS, EBS composition
    flux, temperature ramp 1450 K - 200 K/s t (target 1220 K), slope-0 boundaries, 64 x 64 cells on 1.8 x 1.8 um."""
    c = _base(2, (64, 64), (0.0, 0.0), (1.8, 1.8))
    c.qlen = 0
    c.with_phase = 1
    c.with_concentration = 1
    c.evolve_quat = 0
    c.phase_flux_type = _abi.FLUX_SIMPLE
    c.conc_rhs_form = _abi.CONC_EBS
    c.free_energy = _abi.FE_CALPHAD
    c.T_uniform = 1450.0
    c.dtemperaturedt = -200.0
    c.target_temperature = 1220.0
    c.epsilon_phase = 0.25
    c.phi_mobility = 6.4
    c.phi_well_scale = 2.5
    c.energy_interp = _ch("p")
    c.conc_interp = _ch("p")
    c.avg_func = _ch("a")
    c.conc_avg_func = _ch("a")
    c.vm_liquid = c.vm_solid = 7.68e-6
    c.newton_max_its = 50
    c.calphad = load_calphad()
    c.zero_slope[0] = c.zero_slope[1] = 1
    return c


def two_grains_quadratic_test3d():
    """tests/TwoGrainsQuadratic/3d.input: the GG3D_HBSM model block on 64 x 64 x 48 cells (3.2 x 3.2 x 2.4 um,
    periodic: Geometry has no periodic_dimension, PFModel.cc:173-179), temperature ramp 873 K - 20 K/s t."""
    c = gg3d_hbsm(nx=64, ny=64, nz=48)
    c.dtemperaturedt = -20.0
    c.target_temperature = 573.0
    return c


def kks_composition_test2d():
    """tests/KKScomposition/2d.input: the SingleGrainGrowthAuNi set-up with the KKS form of the composition flux
    (rhs_form "kks": D(phi) grad c + ... with constant D_solid 0.125, D_liquid 1224.23 um^2/s) and the CALPHAD free
    energy; Interface{sigma 0.372678, delta 0.035}; temperature ramp 1450 K - 200 K/s t to 1220 K."""
    import math
    c = single_grain_auni_test2d()
    c.conc_rhs_form = _abi.CONC_KKS
    c.D_solid, c.D_liquid = 0.125, 1224.23
    c.Q0_solid = c.Q0_liquid = 0.0
    sigma, delta = 0.372678, 0.035
    c.epsilon_phase = math.sqrt(6.0 * sigma * delta)
    c.phi_well_scale = (3.0 * sigma / delta) / 16.0
    return c


def four_corners_test2d():
    """tests/FourCorners/2d.input: four quarter-disc grains with different orientations (qlen 4 in 2D) growing from the
    corners of a 64 x 64 box (0.128 x 0.128 um) with slope-0 boundaries; phase + evolving quaternions
    (H_parameter 0.884e-3, epsilon_orient 0.0447), FreeEnergyModel "linear" at a constant 975 K (melting point 1000 K,
    latent heat 2e4 J/mol at 1e-5 m^3/mol), Interface{sigma 0.44, delta 2e-3} -> epsilon_phi = sqrt(6 sigma delta),
    well scale 3 sigma / delta / 16 (QuatModelParameters.cc:842-845)."""
    import math
    c = _base(2, (64, 64), (0.0, 0.0), (0.128, 0.128))
    c.qlen = 4
    c.with_phase = 1
    c.evolve_quat = 1
    c.phase_flux_type = _abi.FLUX_SIMPLE
    c.free_energy = _abi.FE_DELTAT
    c.H_parameter = 0.884e-3
    c.epsilon_q = 0.0447
    c.quat_mobility = 1.0
    c.phi_mobility = 1.0
    sigma, delta = 0.44, 2.0e-3
    c.epsilon_phase = math.sqrt(6.0 * sigma * delta)
    c.phi_well_scale = (3.0 * sigma / delta) / 16.0
    c.T_uniform = 975.0
    c.meltingT = 1000.0
    vm = 1.0e-5
    c.vm_liquid = c.vm_solid = vm
    c.latent_heat = 2.0e4 * (1.0e-6 / vm)     # J/mol -> pJ/um^3 (QuatModelParameters.cc:616)
    c.energy_interp = _ch("p")
    c.avg_func = _ch("a")
    c.conc_avg_func = _ch("a")
    c.zero_slope[0] = c.zero_slope[1] = 1
    return c


def solidify_quaternions_test2d():
    """tests/SolidifyQuaternions/2d.input: the FourCorners model with orient_interp_func_type1 = type2 = "q" on 64 x 32
    cells (0.128 x 0.064 um), periodic in x, slope-0 in y; two grains on the lower boundary grow into a liquid whose
    orientation is random cell by cell."""
    c = four_corners_test2d()
    c.n[1] = 32
    c.orient_interp1 = _ch("q")
    c.orient_interp2 = _ch("q")
    c.zero_slope[0] = 0
    return c


def _to3d(c, n, hi):
    """the 3D version of a 2D deck: same model block, one more direction with the same kind of boundary"""
    c.ndim = 3
    for d in range(3):
        c.n[d] = n[d]
        c.dx[d] = hi[d] / n[d]
    c.zero_slope[2] = c.zero_slope[1]
    return c


def dendrite_test3d():
    """tests/Dendrite/3d.input: the 2D deck's model on 60^3 cells (40^3 um), qlen 4 (init_q = 1, 0, 0, 0; H_parameter
    unset: the orientation only feeds the 3D anisotropy, 3d/quatrhs.m4:149-349), slope-0 on all six sides."""
    c = _to3d(dendrite_test2d(), (60, 60, 60), (40.0, 40.0, 40.0))
    c.qlen = 4
    return c


def single_grain_auni_test3d():
    """tests/SingleGrainGrowthAuNi/3d.input: 32^3 cells on 0.9^3 um, otherwise the 2D deck."""
    return _to3d(single_grain_auni_test2d(), (32, 32, 32), (0.9, 0.9, 0.9))


def kks_composition_test3d():
    """tests/KKScomposition/3d.input: 32^3 cells on 0.9^3 um; unlike its 2D sibling it runs rhs_form "ebs" (with the
    Interface{sigma, delta} block of the 2D deck)."""
    c = _to3d(kks_composition_test2d(), (32, 32, 32), (0.9, 0.9, 0.9))
    c.conc_rhs_form = _abi.CONC_EBS
    return c


def four_corners_test3d():
    """tests/FourCorners/3d.input: the 2D deck four cells thick (64 x 64 x 4 on 0.128 x 0.128 x 0.008 um)."""
    return _to3d(four_corners_test2d(), (64, 64, 4), (0.128, 0.128, 0.008))


def two_grains_quadratic_test2d():
    """tests/TwoGrainsQuadratic/2d.input: the 3D deck's model block on 64 x 64 cells (3.2 x 3.2 um), periodic."""
    c = two_grains_quadratic_test3d()
    c.ndim = 2
    c.n[2], c.dx[2] = 1, 1.0
    return c


def solidify_quaternions_test3d():
    """tests/SolidifyQuaternions/3d.input: 48 x 24 x 8 cells (0.096 x 0.048 x 0.016 um), periodic in x and z, slope-0 in y."""
    c = _to3d(solidify_quaternions_test2d(), (48, 24, 8), (0.096, 0.048, 0.016))
    c.zero_slope[0], c.zero_slope[1], c.zero_slope[2] = 0, 1, 0
    return c


def one_grain_quadratic_test(ndim=2):
    """tests/OneGrainQuadratic/{2d,3d}.input: one grain, no orientation; quadratic free energy with rhs_form "ebs" and
    diffusion_type "temperature_dependent" (Arrhenius D of each phase weighted with the phase fraction,
    TbasedCompositionDiffusionStrategy); Interface{sigma 0.1, delta 0.045}; T = 873 K - 20 K/s t (target 573 K);
    64^2 cells on 3.2^2 um / 48^3 on 2.4^3 um, periodic."""
    import math
    if ndim == 2:
        c = _base(2, (64, 64), (0.0, 0.0), (3.2, 3.2))
    else:
        c = _base(3, (48, 48, 48), (0.0, 0.0, 0.0), (2.4, 2.4, 2.4))
    c.qlen = 0
    c.with_phase = 1
    c.with_concentration = 1
    c.evolve_quat = 0
    c.phase_flux_type = _abi.FLUX_SIMPLE
    c.conc_rhs_form = _abi.CONC_EBS
    c.free_energy = _abi.FE_QUADRATIC
    c.T_uniform = 873.0
    c.dtemperaturedt = -20.0
    c.target_temperature = 573.0
    sigma, delta = 0.1, 0.045
    c.epsilon_phase = math.sqrt(6.0 * sigma * delta)
    c.phi_well_scale = (3.0 * sigma / delta) / 16.0
    c.phi_mobility = 200.0
    c.energy_interp = _ch("h")
    c.conc_interp = _ch("h")          # conc_interp_func_type defaults to phi_interp_func_type
    c.avg_func = _ch("a")
    c.conc_avg_func = _ch("a")
    c.vm_liquid = c.vm_solid = 1.5e-5
    c.D_solid, c.D_liquid = 1.3e8, 5.6e4
    c.Q0_solid, c.Q0_liquid = 156377.0, 55329.0
    c.quad_Tref = 1000.0
    c.quad_A_l = c.quad_A_s = 1.0e4
    c.quad_Ceq_l, c.quad_Ceq_s = 0.05, 0.10
    c.quad_m_l = c.quad_m_s = 0.0
    return c


BUILDERS = {
    "pfhub1a": pfhub1a,
    "dendrite2d": dendrite2d,
    "auni2d": auni2d,
    "gg3d_hbsm": gg3d_hbsm,
    "auni3d": auni3d,
}
