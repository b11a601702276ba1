"""AMPE input decks -> the configuration record of the fused right-hand side (host logic, no GPU).

The caller side of the path: AMPE's driver reads a SAMRAI input database (`2d.input`) and
`QuatModelParameters::readModelParameters` (source/QuatModelParameters.cc:811-1117, with readMolarVolumes :138-163,
readConcDB :175-402, readTemperatureModel :461-626, initializeOrientation :628-766, readPhaseMobility :1119-1178,
readFreeEnergies :1180-1203) turns its `ModelParameters` block into the numbers the Strategy classes are built from;
a few more keys are read next to it (model_type -> quaternion length, source/AMPE.cc:94-110; Geometry and
periodic_dimension, PFModel.cc:170-181; Symmetry{enabled}, QuatModel.cc:506-512; NewtonSolver{}, QuatModel.cc:353-357;
Integrator{lag_quat_sidegrad}, QuatIntegrator.cc:297-298; scalar temperature and its ramp,
TemperatureStrategyFactory.cc:55-63, ScalarTemperatureStrategy.cc:27-30; the phase-flux choice,
PhaseFluxStrategyFactory.h:18-40; the free-energy choice, FreeEnergyStrategyFactory.h:37-260).

`parse` reads the SAMRAI input syntax (nested `Name { key = value, value ... }` blocks, `//` and `/* */` comments,
quoted strings, TRUE / FALSE, numbers and arithmetic expressions); `rhs_config` applies the same keys, defaults, deprecated
spellings and unit conversions as the routines above and returns an `ampe_rhs_config` (ampe_b200._abi.RhsConfig), so the
decks the reference ships run unmodified.  What the fused path does not build (three phases, several order parameters,
ternary alloys, dilute / linear / sintering models, Dirichlet or non-zero slope boundaries, moving frames, ...) is
refused by name, never approximated.  tests/test_input_deck.py holds every deck of the reference whose model is built
against the hand-written configuration of the same deck in ampe_b200/configs.py, field by field.
"""
import ast
import math
import operator
import os
import re

from . import _abi
from . import configs as _configs


class DeckError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# SAMRAI input syntax
_TOKEN = re.compile(r'\s*(?:(?P<str>"(?:[^"\\]|\\.)*")|(?P<box>\[[^\]]*\])|(?P<punct>[{}=,])|(?P<word>[^\s{}=,"\[]+))')
_ARITH = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
          ast.Pow: operator.pow, ast.USub: operator.neg, ast.UAdd: operator.pos}
_FUNCS = {"sqrt": math.sqrt, "exp": math.exp, "log": math.log, "sin": math.sin, "cos": math.cos, "abs": abs}


def _strip_comments(text):
    out, i, n = [], 0, len(text)
    while i < n:
        ch = text[i]
        if ch == '"':
            j = i + 1
            while j < n and text[j] != '"':
                j += 2 if text[j] == "\\" else 1
            out.append(text[i:j + 1])
            i = j + 1
        elif text.startswith("//", i):
            while i < n and text[i] != "\n":
                i += 1
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            if j < 0:
                raise DeckError("unterminated /* comment")
            out.append(" ")
            i = j + 2
        else:
            out.append(ch)
            i += 1
    return "".join(out)


def _number(expr):
    """a SAMRAI arithmetic expression (numbers, + - * / ^, parentheses, a few functions)"""
    def ev(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.BinOp) and type(node.op) in _ARITH:
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, ast.Pow) and (abs(b) > 64 or abs(a) > 1.0e6):
                raise DeckError("not a number: %r" % expr)
            return _ARITH[type(node.op)](a, b)
        if isinstance(node, ast.UnaryOp) and type(node.op) in _ARITH:
            return _ARITH[type(node.op)](ev(node.operand))
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS:
            return _FUNCS[node.func.id](*[ev(a) for a in node.args])
        raise DeckError("not a number: %r" % expr)
    s = expr.replace("^", "**")
    s = re.sub(r"(?<![\w.])(\d+)\.(?![\d\w])", r"\1.0", s)        # "10." -> "10.0"
    s = re.sub(r"(\d)\.([eE][-+]?\d)", r"\1.0\2", s)              # "1.e-5" -> "1.0e-5"
    if len(s) > 200 or s.count("**") > 2:
        raise DeckError("not a number: %r" % expr[:40])
    try:
        v = ev(ast.parse(s, mode="eval").body)
    except (SyntaxError, ZeroDivisionError, OverflowError, ValueError, TypeError, RecursionError, MemoryError):
        raise DeckError("not a number: %r" % expr)
    if isinstance(v, complex):
        raise DeckError("not a number: %r" % expr)
    return v


def _value(tok):
    kind, text = tok
    if kind == "str":
        return text[1:-1]
    if kind == "box":  # a SAMRAI box literal [(0,0),(63,63)]: lower and upper cell index
        corners = re.findall(r"\(([^)]*)\)", text)
        if len(corners) != 2:
            raise DeckError("not a box: %r" % text)
        return tuple(tuple(int(v) for v in c.split(",")) for c in corners)
    up = text.upper()
    if up in ("TRUE", "FALSE"):
        return up == "TRUE"
    return _number(text)


def parse(text):
    """SAMRAI input text -> nested dict (blocks are dicts, `a = 1, 2` is a list, single values are scalars)"""
    toks = []
    pos, text = 0, _strip_comments(text)
    while True:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip():
                raise DeckError("cannot read the deck near %r" % text[pos:pos + 30])
            break
        pos = m.end()
        toks.append((m.lastgroup, m.group(m.lastgroup)))
    # words separated by blanks inside one expression ("6.6 / 16.") are glued back together
    i = 0

    def block(closing):
        nonlocal i
        db = {}
        while i < len(toks):
            kind, t = toks[i]
            if kind == "punct" and t == "}":
                if not closing:
                    raise DeckError("unbalanced '}'")
                i += 1
                return db
            if kind != "word":
                raise DeckError("expected a key, found %r" % t)
            key = t
            i += 1
            if i >= len(toks):
                raise DeckError("key %r without a value" % key)
            kind, t = toks[i]
            if (kind, t) == ("punct", "{"):
                i += 1
                db[key] = block(True)
                continue
            if (kind, t) != ("punct", "="):
                raise DeckError("expected '=' or '{' after %r" % key)
            i += 1
            vals = []
            while True:
                parts = []
                while i < len(toks) and toks[i][0] != "punct" and not (
                        toks[i][0] == "word" and parts and i + 1 < len(toks) and toks[i + 1] in (("punct", "="), ("punct", "{"))):
                    parts.append(toks[i])
                    i += 1
                    if parts[-1][0] in ("str", "box"):
                        break
                if not parts:
                    raise DeckError("key %r without a value" % key)
                vals.append(_value(parts[0]) if len(parts) == 1 else _number(" ".join(p[1] for p in parts)))
                if i < len(toks) and toks[i] == ("punct", ","):
                    i += 1
                    continue
                break
            db[key] = vals[0] if len(vals) == 1 else vals
        if closing:
            raise DeckError("unbalanced '{'")
        return db

    return block(False)


def load(path):
    with open(path) as f:
        return parse(f.read())


# ---------------------------------------------------------------------------------------------------------------------
# tbox::Database look-alikes
def _get(db, key, default=None, required=False, where=""):
    if key in db and not isinstance(db[key], dict):
        return db[key]
    if required:
        raise DeckError("key '%s' is required%s" % (key, where and " in " + where))
    return default


def _first(db, keys, default=None):
    """the current spelling of a key or one of its deprecated ones (printDeprecated)"""
    for k in keys:
        if k in db and not isinstance(db[k], dict):
            return db[k]
    return default


def _block(db, key):
    v = db.get(key)
    return v if isinstance(v, dict) else None


def _unsupported(what):
    raise DeckError("%s: not built in the fused right-hand side (DESIGN.md 8)" % what)


def _ch(s):
    return s[0].lower().encode()


def _zero_slope(model, periodic, ndim):
    """BoundaryConditions{Phase / Conc / Quat / Temperature {boundary_k = "slope", "0"}}: the fused path knows periodic
    and slope-0 faces; every field has to ask for the same thing on a non-periodic direction"""
    bc = _block(model, "BoundaryConditions") or {}
    for field, b in bc.items():
        if not isinstance(b, dict):
            continue
        for k, v in b.items():
            if not k.startswith("boundary_"):
                continue
            face = int(k[len("boundary_"):])
            if face >= 2 * ndim or periodic[face // 2]:
                continue
            kind, val = (v + ["0"])[:2] if isinstance(v, list) else (v, "0")
            if str(kind) != "slope" or float(val) != 0.0:
                _unsupported("boundary condition %s %s = %r" % (field, k, v))
    return [0 if periodic[d] else 1 for d in range(ndim)]


def _deck_errors(fn):
    """a missing required key or a value of the wrong kind is an input error with the key's name, as tbox::Database reports it"""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        try:
            return fn(*args, **kw)
        except KeyError as e:
            raise DeckError("key %s is required" % e)
        except (TypeError, ValueError, IndexError, ZeroDivisionError, OverflowError, AttributeError) as e:
            if isinstance(e, DeckError):
                raise
            raise DeckError("a value in the deck has the wrong type or range (%s)" % e)
    return wrapped


@_deck_errors
def rhs_config(db, ndim=None, deck_dir=None):
    """the `ampe_rhs_config` of a deck (see the module docstring for the routines mirrored); deck_dir: where files the deck names
    (the CALPHAD data base) are looked for after the working directory"""
    geo = _block(db, "Geometry")
    if geo is None:
        raise DeckError("block 'Geometry' is required")
    res = geo.get("coarsest_level_resolution")
    if res is None:
        # CartesianGeometry{domain_boxes = [(0,0),(63,63)]} decks are not among the reference's tests of built models
        raise DeckError("Geometry{coarsest_level_resolution} is required")
    res = res if isinstance(res, list) else [res]
    ndim = ndim or len(res)
    if ndim not in (2, 3):
        _unsupported("%d-dimensional runs" % ndim)
    lo, hi = geo["x_lo"], geo["x_up"]
    lo, hi = (lo if isinstance(lo, list) else [lo]), (hi if isinstance(hi, list) else [hi])
    c = _configs._base(ndim, tuple(int(r) for r in res[:ndim]), tuple(lo[:ndim]), tuple(hi[:ndim]))
    per = geo.get("periodic_dimension", [1] * ndim)      # PFModel.cc:171-179: periodic unless said otherwise
    per = [int(p) for p in (per if isinstance(per, list) else [per])][:ndim]

    amr = _block(db, "Amr")
    if amr is not None and amr.get("enabled", True) and int(amr.get("max_levels", 1)) > 1:
        _unsupported("adaptive mesh refinement (Amr{max_levels > 1})")

    model = _block(db, "ModelParameters")
    if model is None:
        raise DeckError("block 'ModelParameters' is required")
    if model.get("three_phases", False):
        _unsupported("three_phases")
    if int(model.get("norderp", 1)) != 1:
        _unsupported("several order parameters (norderp)")
    for blk in ("MovingFrame", "RigidBody"):
        if _block(model, blk) is not None:
            _unsupported(blk)

    # quaternion length: the model type of the run (AMPE.cc:94-110)
    model_type = db.get("model_type", "Quat")
    if model_type not in ("Quat", "KWC", "KWCcomplex"):
        raise DeckError("Invalid model_type")
    qlen_model = {"Quat": 4, "KWC": 1, "KWCcomplex": 2}[model_type]

    H = float(model.get("H_parameter", -1.0))            # :830: negative turns the orientation terms off
    # interface energy (:833-872)
    itf = _block(model, "Interface")
    with_phase = True
    if itf is not None:
        if "sigma" in itf:
            if "delta" not in itf:
                raise DeckError("Interface: sigma and delta  needed together!")
            sigma, delta = float(itf["sigma"]), float(itf["delta"])
            c.epsilon_phase = math.sqrt(6.0 * sigma * delta)
            c.phi_well_scale = (3.0 * sigma / delta) / 16.0
        else:
            c.epsilon_phase = float(_get(itf, "epsilon_phi", required=True, where="Interface"))
            c.phi_well_scale = float(_get(itf, "phi_well_scale", required=True, where="Interface"))
    else:
        eps = _first(model, ("epsilon_phi", "epsilon_phase", "epsilon_parameter"))
        if eps is None:
            with_phase = False
        else:
            c.epsilon_phase = float(eps)
            c.phi_well_scale = float(_first(model, ("phi_well_scale", "scale_energy_well"), 0.0))
    c.with_phase = 1 if with_phase else 0
    eps_aniso = float(model.get("epsilon_anisotropy", -1.0))
    if eps_aniso > 0.0 and H < 0.0:
        H = 0.0                                          # :876: anisotropy needs the quaternions
    well = str(_first(model, ("phi_well_func_type", "energy_well_func_type"), "double"))
    if well[0] not in "sd":
        raise DeckError("Error: invalid value for phi_well_func_type")
    if well[0] != "d":
        _unsupported("phi_well_func_type = %r" % well)
    conc_db = _block(model, "ConcentrationModel")
    bias_alpha = bias_gamma = None
    if conc_db is None:
        if "bias_well_alpha" in model and float(model["bias_well_alpha"]) > 0.0:
            bias_alpha = float(model["bias_well_alpha"])
        if "bias_well_gamma" in model:
            bias_gamma = float(model["bias_well_gamma"])

    # molar volumes (:138-163): ModelParameters first, else ConcentrationModel
    def molar(d):
        if "molar_volume" in d:
            return float(d["molar_volume"]), float(d["molar_volume"])
        if "molar_volume_solid_A" in d:
            return float(_get(d, "molar_volume_liquid", required=True)), float(d["molar_volume_solid_A"])
        return None
    vm = molar(model) or (molar(conc_db) if conc_db is not None else None)
    if vm is not None:
        c.vm_liquid, c.vm_solid = vm

    # interpolation and averaging (:907-1046)
    e_interp = str(_first(model, ("energy_interp_func_type", "phi_interp_func_type"), "pbg"))
    if e_interp[0].lower() not in "lphu":
        raise DeckError("Error: invalid energy_interp_func_type!!!")
    c.energy_interp = _ch(e_interp)
    c_interp = str(model.get("conc_interp_func_type", e_interp))
    if c_interp[0].lower() not in "lph":
        raise DeckError("Error: invalid conc_interp_func_type!!!")
    c.conc_interp = _ch(c_interp)
    d_interp = str(model.get("diffusion_interp_func_type", "linear"))
    if d_interp[0] not in "lLpPb":
        raise DeckError("Error: invalid diffusion_interp_type!!!")
    c.diffusion_interp = _ch(d_interp)
    avg = str(model.get("avg_func_type", "harmonic"))
    if avg[0] not in "ah":
        raise DeckError("Error: invalid value for avg_func_type")
    c.avg_func = _ch(avg)
    if str(model.get("stencil_type", "normal")) not in ("normal", "isotropic"):
        _unsupported("stencil_type = %r" % model["stencil_type"])

    # orientation (:628-766); with_orientation() is H >= 0, evolveQuat() is H > 0
    c.H_parameter = max(H, 0.0)
    c.evolve_quat = 1 if H > 0.0 else 0
    c.qlen = qlen_model if H >= 0.0 else 0
    if H > 0.0:
        mob = _first(model, ("orient_mobility", "quat_mobility"))
        if mob is None and "tau_quat" in model:
            mob = 1.0 / float(model["tau_quat"])
        if mob is None:
            raise DeckError("Error: quaternion mobility not specified")
        c.quat_mobility = float(mob)
        c.min_quat_mobility = float(_first(model, ("min_orient_mobility", "min_quat_mobility"), 1.0e-6))
        eq = _first(model, ("epsilon_orient", "epsilon_q", "epsilon_quat"))
        if eq is None:
            raise DeckError("Error: epsilon_quat not specified")
        c.epsilon_q = float(eq)
        c.quat_grad_floor = float(_first(model, ("orient_grad_floor", "quat_grad_floor"), 1.0e-2))
        c.grad_floor_type = _ch(str(model.get("orient_grad_floor_type", "max")))
        c.quat_grad_modulus_from_cells = 1 if str(model.get("quat_grad_modulus_type", "cells"))[0] == "c" else 0
        i1 = _first(model, ("diff_interp_func_type", "orient_interp_func_type", "orient_interp_func_type1"), "quadratic")
        if str(i1)[0] not in "qwplts3c":
            raise DeckError("Error: invalid value for orient_interp_func_type1")
        c.orient_interp1 = _ch(str(i1))
        c.orient_interp2 = _ch(str(model.get("orient_interp_func_type2", "constant")))
        mf = str(_first(model, ("quat_mobility_func_type", "orient_mobility_func_type"), "pbg"))
        if mf[0] not in "pie":
            raise DeckError("Error: invalid value for orient_mobility_func_type")
        c.quat_mobility_func = _ch(mf)
        if mf[0] == "i":
            c.quat_mobility_alt_scale = float(_first(model, ("max_orient_mobility", "max_quat_mobility"), 1.0e6))
        elif mf[0] == "e":
            c.quat_mobility_alt_scale = float(_first(model, ("exp_scale_orient_mobility", "exp_scale_quat_mobility"), 1.0e6))
    elif H == 0.0:
        # the hand-written decks carry these two inert numbers for a frozen orientation field
        c.quat_mobility = float(_first(model, ("orient_mobility", "quat_mobility"), 1.0))

    # temperature (:461-626; TemperatureStrategyFactory.cc:55-63; ScalarTemperatureStrategy.cc:27-30)
    tdb = _block(model, "Temperature")
    # (a Temperature block without `type` is an input error in the reference; the shipped examples/AuNi_* decks predate the
    # key and mean the scalar model, which is what they get here)
    ttype = str(tdb.get("type", "scalar")) if tdb is not None else str(model.get("temperature_type", "scalar"))
    if tdb is None:
        tdb = model
    ttype = ttype[0].lower() + ttype[1:]
    if ttype not in ("scalar", "frozen", "gaussian", "constant", "heat", "file"):
        raise DeckError("Error: invalid value for temperature_type")
    c.meltingT = max(float(tdb.get("meltingT", -1.0)), 0.0)     # :491 keeps -1 for "not given"; the record keeps 0
    rescale = -1.0
    if ttype == "scalar":
        t0 = _first(tdb, ("temperature", "temperature0", "T_parameter"))
        if t0 is None:
            raise DeckError("key 'temperature' is required in Temperature")
        c.T_uniform = float(t0)
        c.dtemperaturedt = float(tdb.get("dtemperaturedt", 0.0))
        c.target_temperature = float(tdb.get("target_temperature", 0.0))
    elif ttype == "heat":
        method = str(tdb.get("equation_type", "steady"))
        if method == "steady":
            _unsupported("steady heat equation")
        c.with_unsteady_temperature = 1
        cpdb = _block(tdb, "cp")
        if cpdb is None or _block(cpdb, "SpeciesA") is None:
            raise DeckError("Temperature{cp{SpeciesA{a}}} is required")
        if len(_block(cpdb, "SpeciesA")) > 1:
            _unsupported("temperature-dependent heat capacity (cp b / dm2)")
        c.cp = float(_block(cpdb, "SpeciesA")["a"]) * (1.0e-6 / c.vm_liquid)          # :555-559
        if c.meltingT > 0.0:
            rescale = c.meltingT                                                         # :511-517
            c.cp *= rescale                                                              # :566-573
            c.H_parameter *= rescale                                                     # :603-605
        c.thermal_diffusivity = float(_get(tdb, "thermal_diffusivity", required=True)) * 1.0e8   # :574-579
    else:
        _unsupported("Temperature{type = %r}" % ttype)
    if "latent_heat" in tdb:
        c.latent_heat = float(tdb["latent_heat"]) * (1.0e-6 / c.vm_liquid)              # :611-617

    # free energy (:1180-1203 and FreeEnergyStrategyFactory.h)
    fdb = _block(model, "FreeEnergyModel") or model
    fe_type = str(fdb.get("type", "none"))
    c.free_energy = _abi.FE_NONE

    # composition (:175-402)
    c.conc_avg_func = c.avg_func
    if conc_db is not None:
        c.with_concentration = 1
        if int(conc_db.get("nspecies", 2)) != 2:
            _unsupported("ternary alloys (nspecies)")
        cmodel = str(conc_db.get("model", "undefined"))
        if cmodel not in ("calphad", "quadratic", "linear", "independent", "dilute", "cahn_hilliard", "wang_sintering"):
            raise DeckError("Error: unknown concentration model in QuatModelParameters")
        if cmodel not in ("calphad", "quadratic", "cahn_hilliard"):
            _unsupported("ConcentrationModel{model = %r}" % cmodel)
        rhs = str(conc_db.get("rhs_form", "kks"))
        if rhs == "kks":
            c.conc_rhs_form = _abi.CONC_KKS
        elif rhs == "ebs":
            c.conc_rhs_form = _abi.CONC_EBS
            if str(conc_db.get("ebs_stencil", "regular")) != "regular":
                _unsupported("ebs_stencil = %r" % conc_db["ebs_stencil"])
        elif rhs == "cahn_hilliard":
            c.conc_rhs_form = _abi.CONC_CAHN_HILLIARD
            ch = _block(conc_db, "CahnHilliard")
            if ch is None:
                raise DeckError("block 'CahnHilliard' is required")
            c.ch_ca, c.ch_cb = float(ch["ca"]), float(ch["cb"])
            c.ch_well_scale, c.ch_kappa = float(ch["well_scale"]), float(ch["kappa"])
        else:
            _unsupported("ConcentrationModel{rhs_form = %r}" % rhs)
        default_diff = "composition_dependent" if rhs == "ebs" else "temperature_dependent"
        diff = str(conc_db.get("diffusion_type", default_diff))
        if diff == "temperature_dependent" and rhs != "cahn_hilliard":
            c.D_liquid = float(_get(conc_db, "D_liquid", required=True, where="ConcentrationModel"))
            c.Q0_liquid = float(conc_db.get("Q0_liquid", 0.0))
            ds = _first(conc_db, ("D_solid_A", "D_solid"))
            if ds is None:
                raise DeckError("key 'D_solid' is required in ConcentrationModel")
            c.D_solid = float(ds)
            c.Q0_solid = float(_first(conc_db, ("Q0_solid_A", "Q0_solid"), 0.0))
        elif diff == "mobility" and rhs != "cahn_hilliard":
            _unsupported("diffusion_type = \"mobility\"")
        c.conc_mobility = float(conc_db.get("mobility", 1.0))
        cavg = str(conc_db.get("avg_func_type", avg))
        if cavg[0] not in "ah":
            raise DeckError("Error: invalid value for avg_func_type")
        c.conc_avg_func = _ch(cavg)
        for key in ("gradT_Q0", "antitrapping", "gc"):
            if conc_db.get(key):
                _unsupported("ConcentrationModel{%s}" % key)
        if str(conc_db.get("partition_coeff", "none")) != "none":
            _unsupported("partition coefficients")
        if cmodel == "calphad":
            c.free_energy = _abi.FE_CALPHAD
            cal = _block(conc_db, "Calphad")
            if cal is None:
                raise DeckError("block 'Calphad' is required")
            fname = str(_get(cal, "filename", required=True, where="Calphad"))
            found = [p for p in (fname, os.path.join(deck_dir or ".", fname)) if os.path.isfile(p)]
            try:
                if found:   # the data base file itself, in the reference's format
                    c.calphad = _configs.load_calphad_dat(found[0])
                else:       # not there (the reference's tests link it into the run directory): the packaged transcription
                    c.calphad = _configs.load_calphad(os.path.splitext(os.path.basename(fname))[0] + ".json")
            except OSError:
                _unsupported("CALPHAD data base %r: the file is not there and ampe_b200/data only holds calphadAuNi" % fname)
            except (ValueError, KeyError, AssertionError) as e:
                _unsupported("CALPHAD data base %r (%s)" % (fname, e))
        elif cmodel == "quadratic":
            c.free_energy = _abi.FE_QUADRATIC
            q = _block(conc_db, "Quadratic")
            if q is None:
                raise DeckError("block 'Quadratic' is required")
            need = lambda k: float(_get(q, k, required=True, where="Quadratic"))  # QuadraticFreeEnergyStrategy.cc:56-64
            c.quad_Tref = need("T_ref")
            c.quad_A_l, c.quad_Ceq_l, c.quad_m_l = need("A_liquid"), need("Ceq_liquid"), need("m_liquid")
            c.quad_A_s, c.quad_Ceq_s, c.quad_m_s = need("A_solid"), need("Ceq_solid"), need("m_solid")
        elif cmodel == "cahn_hilliard":
            pass
        else:
            _unsupported("ConcentrationModel{model = %r}" % cmodel)
        newton = _block(conc_db, "NewtonSolver") or {}
        c.newton_max_its = int(newton.get("max_its", c.newton_max_its))
        c.newton_tol = float(newton.get("tol", c.newton_tol))
        c.newton_alpha = float(newton.get("alpha", c.newton_alpha))
    elif c.with_unsteady_temperature and bias_alpha is not None:
        c.free_energy = _abi.FE_BIASWELL
    if conc_db is None or c.free_energy == _abi.FE_NONE:
        if bias_alpha is not None and c.free_energy == _abi.FE_NONE and c.with_unsteady_temperature:
            c.free_energy = _abi.FE_BIASWELL
        if fe_type[0] == "l":                                                             # "pure element free energy"
            c.free_energy = _abi.FE_DELTAT
        elif fe_type[0] == "s":
            _unsupported("FreeEnergyModel{type = \"scalar\"}")
    if bias_alpha is not None:
        c.bias_well_alpha = bias_alpha
        if bias_gamma is not None:
            c.bias_well_gamma = bias_gamma * (rescale if rescale > 0.0 else 1.0)

    # phase mobility (:1108-1116, :1119-1178)
    if with_phase:
        pm = _block(model, "PhaseMobility")
        if pm is not None:
            if str(pm.get("type", "scalar")) != "scalar":
                _unsupported("PhaseMobility{type = %r}" % pm.get("type"))
            if "value" not in pm:
                raise DeckError("Error: phi_mobility not specified")
            c.phi_mobility = float(pm["value"])
        else:
            c.phi_mobility = float(_get(model, "phi_mobility", required=True, where="ModelParameters"))
        if float(model.get("q0_phi_mobility", 0.0)) != 0.0:
            _unsupported("q0_phi_mobility")

    # phase flux (PhaseFluxStrategyFactory.h:18-40)
    if eps_aniso >= 0.0:
        c.phase_flux_type = _abi.FLUX_ANISOTROPIC
        c.epsilon_anisotropy = eps_aniso
    elif str(model.get("stencil_type", "normal")) == "isotropic":
        c.phase_flux_type = _abi.FLUX_ISOTROPIC
    else:
        c.phase_flux_type = _abi.FLUX_SIMPLE

    symm = _block(db, "Symmetry")
    c.symmetry_aware = 1 if (symm is not None and "enabled" in symm and bool(symm["enabled"])) else 0
    integ = _block(db, "Integrator") or {}
    c.lag_quat_sidegrad = 1 if integ.get("lag_quat_sidegrad", True) else 0
    zs = _zero_slope(model, per, ndim)
    for d in range(ndim):
        c.zero_slope[d] = zs[d]
    return c


@_deck_errors
def run_parameters(db):
    """what the driver around the integrator reads (PFModel.cc:351-399, :625-640; QuatIntegrator.cc:285-289;
    EventInterval): end time, step limit, tolerances, output interval, initial-condition file and uniform fields"""
    integ = _block(db, "Integrator") or {}
    atol = float(integ.get("tolerance", integ.get("atol", 3.0e-4)))   # QuatIntegrator.cc:285-289
    ic = _block(db, "InitialConditions") or {}
    sd = _block(db, "ScalarDiagnostics") or {}
    init_q = ic.get("init_q")
    return {
        "end_time": float(db["end_time"]) if "end_time" in db else None,
        "max_timesteps": int(_first(db, ("max_delta_cycles", "max_cycles", "max_timesteps"), 2 ** 31 - 1)),
        "atol": atol,
        "rtol": float(integ.get("rtol", 1.0e-2 * atol)),
        "scalar_diagnostics_interval": float(sd["interval"]) if "interval" in sd else None,
        "scalar_diagnostics_interval_type": str(sd.get("interval_type", "step")),
        "initial_conditions_file": ic.get("filename"),
        "init_t": float(ic["init_t"]) if "init_t" in ic else None,
        "init_q": [float(v) for v in (init_q if isinstance(init_q, list) else [init_q])] if init_q is not None else None,
        "init_c": ic.get("init_c"),
        "slice_index": int(ic.get("slice_index", -1)),
    }


def describe(cfg):
    """the record as a plain dict (the CALPHAD block summarised)"""
    import ctypes as C
    out = {}
    for name, _ in _abi.RhsConfig._fields_:
        v = getattr(cfg, name)
        if isinstance(v, C.Array):
            v = list(v)
        elif isinstance(v, C.Structure):
            v = "<%d bytes>" % C.sizeof(v)
        elif isinstance(v, bytes):
            v = v.decode()
        out[name] = v
    return out


if __name__ == "__main__":   # python -m ampe_b200.input_deck deck.input: what the fused path is configured with, or why it refuses
    import json
    import sys
    for path in sys.argv[1:]:
        try:
            db = load(path)
            print(json.dumps({"deck": path, "config": describe(rhs_config(db, deck_dir=os.path.dirname(os.path.abspath(path)))),
                              "run": run_parameters(db)}, indent=1))
        except DeckError as e:
            print(json.dumps({"deck": path, "refused": str(e)}))
