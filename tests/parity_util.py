"""Shared helpers of the parity tests: the named small configurations (all are the reference's own
samples with JSON-key overrides), the per-field error measure, and the reference runner.

Tolerance (BASELINE.json north_star): density, h, pressure, acceleration and du/dt agree per
particle to a relative 1e-10 in FP64.  "Relative" needs a floor for quantities that are sums with
cancellation (the pressure force on a particle of a uniform lattice is ~1e-12 while its terms are
~10): the reference differs from ITSELF there at the 1e-14 absolute level when only the summation
order changes (tree vs EXHAUSTIVE_SEARCH build).  The floor used is the
particle's natural scale of the quantity: c^2/h for accelerations, c^3/h for du/dt.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sphcode_b200 import sample_params, make_sample  # noqa: E402

RTOL = 1e-10

# name -> (sample, overrides)
CONFIGS = {
    "shock_tube_c1": ("shock_tube", dict(N=50)),                                    # BASELINE configs[0], as shipped
    "shock_tube_tdav": ("shock_tube", dict(N=30, useTimeDependentAV=True)),
    "khi_disph_ac": ("khi", dict(N=64, SPHType="disph", useArtificialConductivity=True)),   # configs[1] physics
    "khi_ssph": ("khi", dict(N=48)),
    "khi_ssph_nobalsara": ("khi", dict(N=32, useBalsaraSwitch=False)),
    "gresho_gsph2": ("gresho_chan_vortex", dict(N=64, SPHType="gsph", use2ndOrderGSPH=True)),  # configs[2] physics
    "gresho_gsph1": ("gresho_chan_vortex", dict(N=32, SPHType="gsph", use2ndOrderGSPH=False)),
    "gresho_ssph": ("gresho_chan_vortex", dict(N=32)),
    "pairing_cubic": ("pairing_instability", dict(N=32)),
    "hydrostatic_disph": ("hydrostatic", dict(N=32)),
    "evrard_c4": ("evrard", dict(N=20)),                                             # configs[3] physics
    "evrard_n30": ("evrard", dict(N=30)),
    "evrard_ssph_cubic": ("evrard", dict(N=16, SPHType="ssph", kernel="cubic_spline")),
    "evrard_gsph": ("evrard", dict(N=16, SPHType="gsph")),
    "evrard_noiter": ("evrard", dict(N=16, iterativeSmoothingLength=False)),
    "evrard_leaf1": ("evrard", dict(N=12, leafParticleNumber=1)),
    "evrard_shallow": ("evrard", dict(N=16, maxTreeLevel=2)),                        # leaves of several hundred particles
    "khi_shallow": ("khi", dict(N=32, maxTreeLevel=3, SPHType="disph")),
    "khi_gravity_periodic": ("khi", dict(N=32, SPHType="disph", useGravity=True, leafParticleNumber=8)),   # periodic box + tree gravity
}


def make_case(name):
    sample, over = CONFIGS[name]
    p = sample_params(sample, **over)
    return p, make_sample(p)


def vnorm(a):
    a = np.asarray(a)
    return np.sqrt((a * a).sum(axis=-1)) if a.ndim > 1 else np.abs(a)


def field_errors(got, ref):
    """Worst per-particle relative error of every state member; `got`, `ref` are SPHParticle arrays."""
    h, c = ref["sml"], ref["sound"]
    with np.errstate(divide="ignore", invalid="ignore"):
        a_scale = np.where(h > 0, c * c / h, 0.0)
        e_scale = np.where(h > 0, c * c * c / h, 0.0)
    out = {}
    for f in ("sml", "dens", "pres", "gradh", "alpha", "sound", "ene", "ene_p", "phi", "mass"):
        d = np.abs(got[f] - ref[f])
        out[f] = float(np.max(d / np.maximum(np.abs(ref[f]), 1e-300)))
    # the Balsara switch |div v| / (|div v| + |rot v| + 1e-4 c/h) lives in [0, 1]; where div v is pure
    # cancellation noise (uniform lattice) only its absolute value is meaningful
    out["balsara"] = float(np.max(np.abs(got["balsara"] - ref["balsara"])))
    # ... and its sensitivity: b = |div| / den with den = |div| + |rot| + 1e-4 c/h >= 1e-4 (c/h) / (1 - b), so an error of
    # rtol x (c/h) in div v (the bar of every other velocity-derived quantity) moves b by up to rtol x 1e4 (1 - b).  Where the
    # flow is uniform (b ~ 0: KHI away from the shear layers) the switch amplifies the 1e-14 velocity differences of a
    # second step by 1e4.  balsara_sens = |db| / (1 + 1e4 (1 - b)) is held to rtol where the plain difference cannot be.
    out["balsara_sens"] = float(np.max(np.abs(got["balsara"] - ref["balsara"]) / (1.0 + 1e4 * (1.0 - np.minimum(ref["balsara"], 1.0)))))
    for f, floor in (("acc", a_scale), ("vel", c), ("vel_p", c)):
        d = vnorm(got[f] - ref[f])
        out[f] = float(np.max(d / np.maximum(vnorm(ref[f]) + floor, 1e-300)))
    d = np.abs(got["dene"] - ref["dene"])
    out["dene"] = float(np.max(d / np.maximum(np.abs(ref["dene"]) + e_scale, 1e-300)))
    d = vnorm(got["pos"] - ref["pos"])
    out["pos"] = float(np.max(d / np.maximum(h, 1e-300)))
    out["neighbor"] = int(np.sum(got["neighbor"] != ref["neighbor"]))
    out["id"] = int(np.sum(got["id"] != ref["id"]))
    return out


def boundary_ties(ref, params, idx, tol=1e-12):
    """For particles idx: number of j whose minimum-image distance sits within tol (relative) of the
    support radius h_i.  SPHParticle::neighbor counts `r < h_i`; on exact lattices (1-D shock tube:
    h converges to 2 dx) a neighbour can sit ON the support boundary, where W = 0 and the count is
    decided by the last bit of h, i.e. by the reference compiler's -ffast-math code generation."""
    pos = ref["pos"]
    out = []
    for i in idx:
        d = pos[i] - pos
        if params["periodic"]:
            L = np.asarray(params["rangeMax"]) - np.asarray(params["rangeMin"])
            d = d - L * np.round(d / L)
        r = np.sqrt((d * d).sum(axis=1))
        out.append(int(np.sum(np.abs(r - ref["sml"][i]) <= tol * ref["sml"][i])))
    return np.array(out)


def assert_fields(got, ref, fields, rtol=RTOL, what="", params=None):
    e = field_errors(got, ref)
    if params is not None and "neighbor" in fields and e["neighbor"]:
        idx = np.nonzero(got["neighbor"] != ref["neighbor"])[0]
        ties = boundary_ties(ref, params, idx)
        if np.all(np.abs(got["neighbor"][idx].astype(int) - ref["neighbor"][idx]) <= ties):
            e["neighbor_boundary_ties"] = e["neighbor"]
            e["neighbor"] = 0
    bad = {f: e[f] for f in fields if (e[f] > (0 if f in ("neighbor", "id") else rtol) or not np.isfinite(e[f]))}
    assert not bad, f"{what}: fields beyond rtol={rtol}: {bad} (all: {e})"
    return e


PRE_FIELDS = ("sml", "dens", "pres", "gradh", "balsara", "alpha", "neighbor")
FORCE_FIELDS = ("acc", "dene", "phi")
STEP_FIELDS = PRE_FIELDS + FORCE_FIELDS + ("pos", "vel", "vel_p", "ene", "ene_p", "sound")


def lists_equal(a, b):
    (oa, ia), (ob, ib) = a, b
    return np.array_equal(oa, ob) and np.array_equal(ia, ib)


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")
