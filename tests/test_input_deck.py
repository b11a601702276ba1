"""The caller side of the path: AMPE input decks -> ampe_rhs_config (ampe_b200/input_deck.py, mirror of
QuatModelParameters::readModelParameters, source/QuatModelParameters.cc:811-1117 and the readers it calls).  Host logic,
no GPU.  Three layers: the SAMRAI input syntax; the keys, defaults, deprecated spellings and unit conversions on decks
written here; and -- where /root/reference exists (the build container) -- every deck the reference ships: the ones whose
model is built must give the hand-written configuration of ampe_b200/configs.py field by field (those configurations
are what the regression decks were integrated with, so this closes the loop deck file -> record -> acceptance number),
all others must be refused by name."""
import ctypes as C
import glob
import math
import os

import pytest

from ampe_b200 import _abi, configs, input_deck
from ampe_b200.input_deck import DeckError

REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reads the reference's own decks (build container only)")


# ---- syntax ---------------------------------------------------------------------------------------------------------
def test_samrai_input_syntax():
    db = input_deck.parse('''
       // a comment
       end_time = 1.e-2   /* block
          comment */ max_timesteps=300
       flag = TRUE  other = FALSE
       name = "a // string { with } punctuation"
       Outer { inner_value = 6.6/16.   // expression
               Inner { v = 1., 2 , 3.5e0 }
               list = "slope", "0." }
       expr = 2 * (3 + 4) ^ 2
       expr2 = 6.6 / 16.
       box = [(0,0),(63,31)]
       neg = -1.
    ''')
    assert db["end_time"] == 0.01 and db["max_timesteps"] == 300 and isinstance(db["max_timesteps"], int)
    assert db["flag"] is True and db["other"] is False
    assert db["name"] == "a // string { with } punctuation"
    assert db["Outer"]["inner_value"] == 6.6 / 16.0
    assert db["Outer"]["Inner"]["v"] == [1.0, 2, 3.5]
    assert db["Outer"]["list"] == ["slope", "0."]
    assert db["expr"] == 98 and db["expr2"] == 6.6 / 16.0 and db["neg"] == -1.0
    assert db["box"] == ((0, 0), (63, 31))


@pytest.mark.parametrize("text,msg", [("A { x = 1", "unbalanced '{'"), ("x = 1 }", "unbalanced '}'"), ("x = ", "without a value"),
                                      ("x 1", "expected '=' or '{'"), ("x = __import__", "not a number"),
                                      ("x = /* open", "unterminated")])
def test_syntax_errors(text, msg):
    with pytest.raises(DeckError, match=msg):
        input_deck.parse(text)


# ---- keys, defaults, conversions --------------------------------------------------------------------------------------
GEOMETRY = 'Geometry { coarsest_level_resolution = 32, 16  x_lo = 0., 0.  x_up = 3.2, 0.8  periodic_dimension = 1, 0 }\n'


def _cfg(model, extra=""):
    return input_deck.rhs_config(input_deck.parse(GEOMETRY + extra + "ModelParameters {\n" + model + "\n}\n"))


def test_defaults_of_a_minimal_deck():
    c = _cfg('epsilon_phi = 0.25  phi_well_scale = 2.5  phi_mobility = 6.4  temperature = 1000.')
    assert (c.ndim, list(c.n), c.dx[0], c.dx[1]) == (2, [32, 16, 1], 0.1, 0.05)
    assert list(c.zero_slope)[:2] == [0, 1]                      # non-periodic direction: slope 0
    assert c.with_phase == 1 and c.with_concentration == 0 and c.with_unsteady_temperature == 0
    assert c.qlen == 0 and c.evolve_quat == 0                    # H_parameter defaults to -1: no orientation (:830)
    assert c.phase_flux_type == _abi.FLUX_SIMPLE and c.free_energy == _abi.FE_NONE
    assert (c.energy_interp, c.conc_interp, c.diffusion_interp, c.avg_func) == (b"p", b"p", b"l", b"h")
    assert c.lag_quat_sidegrad == 1 and c.symmetry_aware == 0
    assert c.T_uniform == 1000.0 and c.dtemperaturedt == 0.0


def test_interface_block_and_deprecated_spellings():
    c = _cfg('Interface { sigma = 0.1  delta = 0.045 }  phi_mobility = 1.  T_parameter = 900.')
    assert c.epsilon_phase == math.sqrt(6.0 * 0.1 * 0.045) and c.phi_well_scale == (3.0 * 0.1 / 0.045) / 16.0   # :842-845
    assert c.T_uniform == 900.0
    with pytest.raises(DeckError, match="sigma and delta"):
        _cfg('Interface { sigma = 0.1 }  phi_mobility = 1.  temperature = 900.')
    c = _cfg('epsilon_parameter = 0.3  scale_energy_well = 1.5  energy_interp_func_type = "harmonic"  PhaseMobility { value = 2. }'
             '  temperature0 = 800.')
    assert (c.epsilon_phase, c.phi_well_scale, c.energy_interp, c.conc_interp, c.phi_mobility) == (0.3, 1.5, b"h", b"h", 2.0)
    with pytest.raises(DeckError, match="phi_mobility"):
        _cfg('epsilon_phi = 0.25  temperature = 1000.')


def test_orientation_block():
    model = ('epsilon_phi = 0.25 phi_well_scale = 2.5 phi_mobility = 6.4 temperature = 1450. H_parameter = 0.25 '
             'epsilon_q = 0.3125 tau_quat = 4. orient_grad_floor_type = "tanh" quat_grad_floor = 1.e-3 '
             'orient_mobility_func_type = "inv" max_orient_mobility = 1.e3 orient_interp_func_type = "pbg"')
    c = _cfg(model, 'Symmetry { enabled = TRUE }\nIntegrator { lag_quat_sidegrad = FALSE }\n')
    assert c.qlen == 4 and c.evolve_quat == 1 and c.H_parameter == 0.25 and c.epsilon_q == 0.3125
    assert c.quat_mobility == 0.25 and c.min_quat_mobility == 1.0e-6                      # 1 / tau_quat (:641-644)
    assert (c.grad_floor_type, c.quat_grad_floor, c.quat_mobility_func, c.quat_mobility_alt_scale) == (b"t", 1.0e-3, b"i", 1.0e3)
    assert (c.orient_interp1, c.orient_interp2) == (b"p", b"c")
    assert c.symmetry_aware == 1 and c.lag_quat_sidegrad == 0
    assert _cfg(model, 'model_type = "KWCcomplex"\n').qlen == 2                             # AMPE.cc:94-110
    with pytest.raises(DeckError, match="quaternion mobility not specified"):
        _cfg('epsilon_phi = 0.25 phi_mobility = 1. temperature = 1. H_parameter = 0.1 epsilon_q = 1.')
    with pytest.raises(DeckError, match="Invalid model_type"):
        _cfg(model, 'model_type = "Other"\n')
    # anisotropy needs the quaternions: H_parameter -1 becomes 0, the orientation is carried but frozen (:876)
    c = _cfg('epsilon_phi = 2. phi_well_scale = 0.25 phi_mobility = 0.25 temperature = 1. epsilon_anisotropy = 0.05', 'model_type = "KWCcomplex"\n')
    assert (c.qlen, c.evolve_quat, c.phase_flux_type, c.epsilon_anisotropy) == (2, 0, _abi.FLUX_ANISOTROPIC, 0.05)


def test_heat_equation_units():
    c = _cfg('''epsilon_phi = 2. phi_well_scale = 0.25 phi_mobility = 0.25 molar_volume = 2.e-6  H_parameter = 0.5
                orient_mobility = 1. epsilon_orient = 1.
                Temperature { type = "heat" equation_type = "unsteady" meltingT = 4.
                              cp { SpeciesA { a = 10. } } thermal_diffusivity = 3.e-8 latent_heat = 20. }
                FreeEnergyModel { type = "linear" }''')
    assert c.with_unsteady_temperature == 1 and c.free_energy == _abi.FE_DELTAT
    assert c.cp == 10.0 * (1.0e-6 / 2.0e-6) * 4.0                # J/mol/K -> pJ/um^3/K (:555-559), times the melting point (:566-573)
    assert c.latent_heat == 20.0 * (1.0e-6 / 2.0e-6)             # :611-617
    assert c.thermal_diffusivity == 3.0e-8 * 1.0e8               # cm^2/s -> um^2/s (:574-579)
    assert c.H_parameter == 0.5 * 4.0                            # rescaled with the temperature (:603-605)
    with pytest.raises(DeckError, match="steady heat equation"):
        _cfg('epsilon_phi = 2. phi_mobility = 1. Temperature { type = "heat" cp { SpeciesA { a = 1. } } thermal_diffusivity = 1. }')


def test_concentration_block():
    base = 'epsilon_phi = 0.165 phi_well_scale = 0.4125 phi_mobility = 200. temperature = 873. avg_func_type = "arithmetic" '
    quad = 'Quadratic { T_ref = 1000. A_liquid = 1.e4 A_solid = 2.e4 Ceq_liquid = 0.05 Ceq_solid = 0.1 m_liquid = 0. m_solid = 1. }'
    c = _cfg(base + 'ConcentrationModel { model = "quadratic" molar_volume_liquid = 1.e-5 molar_volume_solid_A = 1.5e-5 '
             'D_liquid = 5.6e4 D_solid = 1.3e8 Q0_solid_A = 7. ' + quad + ' NewtonSolver { max_its = 7 tol = 1.e-9 alpha = 0.5 } }')
    assert c.with_concentration == 1 and c.conc_rhs_form == _abi.CONC_KKS and c.free_energy == _abi.FE_QUADRATIC   # rhs_form defaults to "kks"
    assert (c.vm_liquid, c.vm_solid) == (1.0e-5, 1.5e-5)                                      # :138-163 via ConcentrationModel
    assert (c.D_liquid, c.D_solid, c.Q0_liquid, c.Q0_solid) == (5.6e4, 1.3e8, 0.0, 7.0)         # temperature_dependent by default
    assert (c.quad_Tref, c.quad_A_s, c.quad_m_s, c.quad_Ceq_l) == (1000.0, 2.0e4, 1.0, 0.05)
    assert (c.newton_max_its, c.newton_tol, c.newton_alpha) == (7, 1.0e-9, 0.5)
    assert c.conc_avg_func == b"a"                                                               # inherits avg_func_type (:292-293)
    c = _cfg(base + 'ConcentrationModel { model = "quadratic" rhs_form = "ebs" molar_volume = 1.e-5 ' + quad + ' avg_func_type = "harmonic" }')
    assert c.conc_rhs_form == _abi.CONC_EBS and c.D_liquid == 0.0 and c.conc_avg_func == b"h"   # EBS: composition_dependent by default
    with pytest.raises(DeckError, match="'D_liquid' is required"):
        _cfg(base + 'ConcentrationModel { model = "quadratic" molar_volume = 1.e-5 ' + quad + ' }')
    with pytest.raises(DeckError, match="'T_ref' is required"):
        _cfg(base + 'ConcentrationModel { model = "quadratic" rhs_form = "ebs" molar_volume = 1.e-5 Quadratic { A_liquid = 1. } }')
    c = _cfg('temperature = 1000. ConcentrationModel { model = "cahn_hilliard" rhs_form = "cahn_hilliard" diffusion_type = "mobility" '
             'mobility = 5. CahnHilliard { ca = 0.3 cb = 0.7 well_scale = 5. kappa = 2. } }')
    assert c.with_phase == 0 and c.conc_rhs_form == _abi.CONC_CAHN_HILLIARD and c.conc_mobility == 5.0
    assert (c.ch_ca, c.ch_cb, c.ch_well_scale, c.ch_kappa) == (0.3, 0.7, 5.0, 2.0)
    c = _cfg(base + 'ConcentrationModel { model = "calphad" rhs_form = "ebs" molar_volume = 7.68e-6 Calphad { filename = "calphadAuNi.dat" } }')
    assert c.free_energy == _abi.FE_CALPHAD and bytes(c.calphad) == bytes(configs.load_calphad())


@pytest.mark.parametrize("model,msg", [
    ('three_phases = TRUE', "three_phases"), ('norderp = 3', "order parameters"), ('MovingFrame { velocity = 1. }', "MovingFrame"),
    ('epsilon_phi = 1. phi_mobility = 1. Temperature { type = "frozen" }', "frozen"),
    ('epsilon_phi = 1. phi_mobility = 1. temperature = 1. ConcentrationModel { model = "dilute" }', "dilute"),
    ('epsilon_phi = 1. phi_mobility = 1. temperature = 1. BoundaryConditions { Phase { boundary_2 = "value", "1." } }', "boundary condition"),
    ('epsilon_phi = 1. phi_mobility = 1. temperature = 1. BoundaryConditions { Phase { boundary_3 = "slope", "1.e-4" } }', "boundary condition"),
])
def test_models_outside_the_path_are_refused_by_name(model, msg):
    with pytest.raises(DeckError, match=msg):
        _cfg(model)


def test_periodic_faces_ignore_their_boundary_entries():
    c = _cfg('epsilon_phi = 1. phi_mobility = 1. temperature = 1. BoundaryConditions { Phase { boundary_0 = "value", "1." boundary_2 = "slope", "0" } }')
    assert list(c.zero_slope)[:2] == [0, 1]


def test_run_parameters():
    db = input_deck.parse('end_time = 0.3 max_delta_cycles = 500 Integrator { atol = 1.e-5 } ScalarDiagnostics { interval = 0.02 '
                          'interval_type = "time" } InitialConditions { filename = "64x64.nc" init_t = 0.7 init_q = 1., 0. }')
    r = input_deck.run_parameters(db)
    assert (r["end_time"], r["max_timesteps"], r["atol"], r["rtol"]) == (0.3, 500, 1.0e-5, 1.0e-2 * 1.0e-5)
    assert (r["scalar_diagnostics_interval"], r["scalar_diagnostics_interval_type"]) == (0.02, "time")
    assert (r["initial_conditions_file"], r["init_t"], r["init_q"], r["slice_index"]) == ("64x64.nc", 0.7, [1.0, 0.0], -1)


# ---- the reference's own decks ----------------------------------------------------------------------------------------
DECKS = [
    ("tests/Dendrite/2d.input", configs.dendrite_test2d, ()),
    ("tests/Dendrite/3d.input", configs.dendrite_test3d, ()),
    ("tests/SingleGrainGrowthAuNi/2d.input", configs.single_grain_auni_test2d, ()),
    ("tests/SingleGrainGrowthAuNi/3d.input", configs.single_grain_auni_test3d, ()),
    ("tests/TwoGrainsQuadratic/2d.input", configs.two_grains_quadratic_test2d, ()),
    ("tests/TwoGrainsQuadratic/3d.input", configs.two_grains_quadratic_test3d, ()),
    ("tests/KKScomposition/2d.input", configs.kks_composition_test2d, ()),
    # the 3D deck switches to rhs_form "ebs": the constant diffusivities the hand-written record inherits from 2D are not read
    ("tests/KKScomposition/3d.input", configs.kks_composition_test3d, ("D_liquid", "D_solid")),
    ("tests/FourCorners/2d.input", configs.four_corners_test2d, ()),
    ("tests/FourCorners/3d.input", configs.four_corners_test3d, ()),
    ("tests/SolidifyQuaternions/2d.input", configs.solidify_quaternions_test2d, ()),
    ("tests/SolidifyQuaternions/3d.input", configs.solidify_quaternions_test3d, ()),
    ("tests/OneGrainQuadratic/2d.input", lambda: configs.one_grain_quadratic_test(2), ()),
    ("tests/OneGrainQuadratic/3d.input", lambda: configs.one_grain_quadratic_test(3), ()),
    # the bench workloads are the examples' model blocks on the grids BASELINE.json names
    ("examples/Dendrite2D/dendrite.input", configs.dendrite2d, ("n", "dx")),
    # ... at the uniform 1450 K BASELINE.json quotes (the decks go on to cool at 200 K/s; the same record at t = 0)
    ("examples/AuNi_2D/9grains_AuNi.input", configs.auni2d, ("n", "dx", "dtemperaturedt", "target_temperature")),
    ("examples/AuNi_3D/1grain3D_AuNi.input", configs.auni3d, ("n", "dx", "dtemperaturedt", "target_temperature")),
    # PFHub 1a asks for M = 5; the deck puts it in CahnHilliard{mobility}, which readCahnHilliard (:427-435) does not read (the
    # maintained tests/CahnHilliard deck gives ConcentrationModel{mobility = 5}); NewtonSolver{} is inert without a KKS model
    ("benchmarks/PFHub1a/2d.input", configs.pfhub1a, ("n", "dx", "conc_mobility", "newton_max_its")),
]


def _plain(v):
    if isinstance(v, C.Array):
        return [_plain(x) for x in v]
    if isinstance(v, C.Structure):
        return bytes(v)
    return v


@needs_reference
@pytest.mark.parametrize("deck,builder,differs", DECKS, ids=[d[0] for d in DECKS])
def test_reference_decks_give_the_hand_written_records(deck, builder, differs):
    got, want = input_deck.rhs_config(input_deck.load(os.path.join(REF, deck))), builder()
    for name, _ in _abi.RhsConfig._fields_:
        if name in differs:
            continue
        a, b = _plain(getattr(got, name)), _plain(getattr(want, name))
        if name == "dx":   # (x_up - x_lo) / n against n h / n written by hand: two roundings
            assert all(abs(x - y) <= 4.0e-16 * abs(y) for x, y in zip(a, b)), (name, a, b)
        else:
            assert a == b, (name, a, b)


@needs_reference
def test_pfhub1a_deck_as_the_reference_reads_it():
    c = input_deck.rhs_config(input_deck.load(os.path.join(REF, "benchmarks/PFHub1a/2d.input")))
    assert list(c.n)[:2] == [128, 128] and c.dx[0] == 200.0 / 128 and c.conc_mobility == 1.0


@needs_reference
def test_every_deck_of_the_reference_is_configured_or_refused_by_name():
    decks = sorted(glob.glob(REF + "/tests/*/*.input") + glob.glob(REF + "/examples/*/*.input") + glob.glob(REF + "/benchmarks/*/*.input"))
    assert len(decks) > 100
    configured, refused = [], {}
    for p in decks:
        try:
            db = input_deck.load(p)                 # every file of the reference parses
        except DeckError as e:
            raise AssertionError("%s: %s" % (p, e))
        try:
            input_deck.rhs_config(db)
            input_deck.run_parameters(db)
            configured.append(os.path.relpath(p, REF))
        except DeckError as e:
            refused[os.path.relpath(p, REF)] = str(e)
    for deck, _, _ in DECKS:
        assert deck in configured
    assert "tests/ConservedVolume/2d.input" in configured     # the model is built; its integration is the open item of DESIGN.md 4
    assert "three_phases" in refused["tests/3Ph2Sl/2d.input"]
    assert "MovingFrame" in refused["tests/AlCuMovingFrame/2d.input"]
    assert "dilute" in refused["tests/AlCu/2d.input"] or "antitrapping" in refused["tests/AlCu/2d.input"]
    assert "boundary condition" in refused["tests/PlanarFront/2d.input"]
    assert "T_ref" in refused["examples/GG3D_HBSM/gg3d_hbsm.input"]   # a stale example: QuadraticFreeEnergyStrategy.cc:56 needs the key too


def _samrai_text(db, indent=0):
    out = []
    for k, v in db.items():
        pad = " " * indent
        if isinstance(v, dict):
            out.append("%s%s {\n%s%s}\n" % (pad, k, _samrai_text(v, indent + 3), pad))
        elif isinstance(v, str):
            out.append('%s%s = "%s"\n' % (pad, k, v))
        else:
            out.append("%s%s = %s   // %d value(s)\n" % (pad, k, ", ".join(repr(float(x)) for x in v), len(v)))
    return "".join(out)


def test_calphad_data_base_in_the_reference_format(tmp_path):
    """ConcentrationModel{Calphad{filename}}: a data base file in the reference's format next to the deck is read as it is (the
    packaged transcription only stands in when the file is not there, as in the reference's tests, which link it into the run
    directory)"""
    import json
    db = json.load(open(os.path.join(os.path.dirname(configs.__file__), "data", "calphadAuNi.json")))
    (tmp_path / "mydb.dat").write_text(_samrai_text(db))
    assert bytes(configs.load_calphad_dat(str(tmp_path / "mydb.dat"))) == bytes(configs.load_calphad())
    db["LmixPhaseL"]["L0"][0] += 1.0   # a different data base must give a different record
    (tmp_path / "other.dat").write_text(_samrai_text(db))
    deck = GEOMETRY + '''ModelParameters { epsilon_phi = 0.25 phi_well_scale = 2.5 phi_mobility = 6.4 temperature = 1450.
        ConcentrationModel { model = "calphad" rhs_form = "ebs" molar_volume = 7.68e-6 Calphad { filename = "other.dat" } } }'''
    c = input_deck.rhs_config(input_deck.parse(deck), deck_dir=str(tmp_path))
    assert c.calphad.L[0][0][0] == configs.load_calphad().L[0][0][0] + 1.0
    with pytest.raises(DeckError, match="other.dat"):       # neither the file nor a packaged transcription of that name
        input_deck.rhs_config(input_deck.parse(deck), deck_dir=str(tmp_path / "nowhere"))
    del db["MobilityParameters"]
    (tmp_path / "other.dat").write_text(_samrai_text(db))
    with pytest.raises(DeckError, match="MobilityParameters"):
        input_deck.rhs_config(input_deck.parse(deck), deck_dir=str(tmp_path))


@needs_reference
def test_reference_thermodynamic_data_files():
    """thermodynamic_data/calphadAuNi.dat read directly = the packaged transcription; the binary Cu-Ni data base reads; the ternary
    ones are refused by what they lack"""
    assert bytes(configs.load_calphad_dat(REF + "/thermodynamic_data/calphadAuNi.dat")) == bytes(configs.load_calphad())
    assert configs.load_calphad_dat(REF + "/thermodynamic_data/calphadCuNi.dat").g[0][0].nintervals >= 1
    with pytest.raises(ValueError, match="LmixPhaseL"):
        configs.load_calphad_dat(REF + "/thermodynamic_data/calphadMoNbTa.dat")


def test_damaged_decks_fail_with_a_deck_error():
    """random edits of a valid deck: a configuration or a DeckError, nothing else"""
    import random
    rng = random.Random(5)
    base = GEOMETRY + '''model_type = "Quat"  end_time = 1.  Symmetry { enabled = TRUE }
        ModelParameters { H_parameter = 0.25 epsilon_q = 0.3125 orient_mobility = 0.64 Interface { sigma = 0.1 delta = 0.045 } phi_mobility = 6.4
           Temperature { type = "scalar" temperature = 1450. dtemperaturedt = -200. }
           ConcentrationModel { model = "quadratic" rhs_form = "kks" molar_volume = 1.e-5 D_liquid = 1. D_solid = 2.
              Quadratic { T_ref = 1000. A_liquid = 1.e4 A_solid = 1.e4 Ceq_liquid = 0.05 Ceq_solid = 0.1 m_liquid = 0. m_solid = 0. } }
           BoundaryConditions { Phase { boundary_2 = "slope", "0" boundary_3 = "slope", "0" } } }'''
    assert input_deck.rhs_config(input_deck.parse(base)).qlen == 4
    alphabet = '{}=,"/* \n0123456789.eE-+abcTRUEFALSE'
    ok = refused = 0
    for trial in range(1500):
        t = list(base)
        for _ in range(rng.randint(1, 4)):
            i = rng.randrange(len(t))
            op = rng.random()
            if op < 0.4:
                t[i] = rng.choice(alphabet)
            elif op < 0.7:
                del t[i]
            else:
                t.insert(i, rng.choice(alphabet))
        try:
            db = input_deck.parse("".join(t))
            input_deck.rhs_config(db)
            input_deck.run_parameters(db)
            ok += 1
        except DeckError:
            refused += 1
    assert ok > 100 and refused > 100
