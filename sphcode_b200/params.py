"""SPHParameters: the reference's JSON parameter surface (README.md:122-152, parsed in
src/solver.cpp:155-299) with the same keys, defaults and error conditions."""
import json
import re

# key -> default (None = required), src/solver.cpp:193-298
_DEFAULTS = {
    "outputDirectory": None,
    "startTime": 0.0,
    "endTime": None,
    "outputTime": None,            # (end - start) / 100
    "energyTime": None,            # outputTime
    "SPHType": "ssph",
    "cflSound": 0.3,
    "cflForce": 0.125,
    "avAlpha": 1.0,
    "useBalsaraSwitch": True,
    "useTimeDependentAV": False,
    "alphaMax": 2.0,
    "alphaMin": 0.1,
    "epsilonAV": 0.2,
    "useArtificialConductivity": False,
    "alphaAC": 1.0,
    "maxTreeLevel": 20,
    "leafParticleNumber": 1,
    "neighborNumber": 32,
    "gamma": None,
    "kernel": "cubic_spline",
    "iterativeSmoothingLength": True,
    "periodic": False,
    "rangeMax": [],
    "rangeMin": [],
    "useGravity": False,
    "G": 1.0,
    "theta": 0.5,
    "use2ndOrderGSPH": True,
}

# sample name -> (DIM, default N), src/solver.cpp:164-187 and src/sample/*.cpp
SAMPLES = {
    "shock_tube": (1, 100),
    "gresho_chan_vortex": (2, 64),
    "pairing_instability": (2, 64),
    "hydrostatic": (2, 32),
    "khi": (2, 128),
    "evrard": (3, 20),
}


class SPHParameterError(ValueError):
    pass


def resolve(user, dim=None):
    """Apply the reference's defaults and checks to a dict of JSON keys."""
    p = dict(_DEFAULTS)
    for k, v in user.items():
        p[k] = v
    for k in ("endTime", "gamma"):
        if p[k] is None:
            raise SPHParameterError(f"No such node ({k})")        # boost ptree_bad_path text
    if p["outputDirectory"] is None:
        p["outputDirectory"] = "results"
    if p["endTime"] < p["startTime"]:
        raise SPHParameterError("endTime < startTime")             # src/solver.cpp:198
    if p["outputTime"] is None:
        p["outputTime"] = (p["endTime"] - p["startTime"]) / 100
    if p["energyTime"] is None:
        p["energyTime"] = p["outputTime"]
    if p["SPHType"] not in ("ssph", "disph", "gsph"):
        raise SPHParameterError("Unknown SPH type")                # src/solver.cpp:213
    if p["useTimeDependentAV"] and p["alphaMax"] < p["alphaMin"]:
        raise SPHParameterError("alphaMax < alphaMin")             # src/solver.cpp:228
    if p["kernel"] not in ("cubic_spline", "wendland"):
        raise SPHParameterError("kernel is unknown.")              # src/solver.cpp:255
    if p["periodic"] and dim is not None:
        if len(p["rangeMax"]) != dim or len(p["rangeMin"]) != dim:
            raise SPHParameterError("rangeMax != DIM")             # src/solver.cpp:264,277
    if dim == 1 and p["kernel"] == "wendland":
        raise SPHParameterError("Wendland C4 is not defined for DIM == 1")   # wendland_kernel.hpp:25-28
    return p


def load_json(path):
    """The sample JSONs write gamma with 36 digits; Python floats parse them like std::stod."""
    with open(path) as f:
        return json.load(f)


# The shipped sample parameter files (sample/<name>/<name>.json), restated as dicts so that the
# package does not need the reference tree at run time.
SHIPPED = {
    "shock_tube": {
        "outputDirectory": "sample/shock_tube/results", "endTime": 0.2, "avAlpha": 1.0,
        "neighborNumber": 4, "gamma": 1.4, "kernel": "cubic_spline", "N": 50, "periodic": True,
        "iterativeSmoothingLength": True, "rangeMax": [1.5], "rangeMin": [-0.5],
        "useTimeDependentAV": False, "SPHType": "ssph"},
    "gresho_chan_vortex": {
        "outputDirectory": "sample/gresho_chan_vortex/results", "endTime": 1.0, "avAlpha": 1.0,
        "neighborNumber": 32, "useBalsaraSwitch": True, "leafParticleNumber": 32,
        "gamma": 1.66666666666666666666666666666666667, "kernel": "wendland", "N": 64,
        "periodic": True, "rangeMax": [0.5, 0.5], "rangeMin": [-0.5, -0.5], "SPHType": "ssph"},
    "hydrostatic": {
        "outputDirectory": "sample/hydrostatic/results", "endTime": 8.0, "avAlpha": 1.0,
        "neighborNumber": 32, "useBalsaraSwitch": True, "leafParticleNumber": 16,
        "gamma": 1.66666666666666666666666666666666667, "kernel": "wendland", "N": 32,
        "periodic": True, "rangeMax": [0.5, 0.5], "rangeMin": [-0.5, -0.5], "SPHType": "disph"},
    "khi": {
        "outputDirectory": "sample/khi/results", "endTime": 3.0, "outputTime": 0.1, "avAlpha": 1.0,
        "neighborNumber": 32, "useBalsaraSwitch": True, "useTimeDependentAV": True,
        "useArtificialConductivity": False, "leafParticleNumber": 16,
        "gamma": 1.66666666666666666666666666666666667, "kernel": "wendland", "N": 256,
        "periodic": True, "rangeMax": [1.0, 1.0], "rangeMin": [0.0, 0.0], "SPHType": "ssph"},
    "pairing_instability": {
        "outputDirectory": "sample/pairing_instability/results", "endTime": 1.0, "avAlpha": 1.0,
        "neighborNumber": 32, "leafParticleNumber": 32,
        "gamma": 1.66666666666666666666666666666666667, "kernel": "cubic_spline", "N": 64,
        "periodic": True, "rangeMax": [0.5, 0.5], "rangeMin": [-0.5, -0.5]},
    "evrard": {
        "outputDirectory": "sample/evrard/results", "endTime": 3.0, "avAlpha": 1.0,
        "neighborNumber": 32, "useBalsaraSwitch": True, "useTimeDependentAV": True,
        "useArtificialConductivity": False, "leafParticleNumber": 32,
        "gamma": 1.66666666666666666666666666666666667, "kernel": "wendland", "N": 30,
        "periodic": False, "useGravity": True, "SPHType": "disph"},
}


def sample_params(name, **overrides):
    """Resolved parameters of a shipped sample, with JSON-key overrides (e.g. N=124, SPHType="disph")."""
    if name not in SHIPPED:
        raise SPHParameterError("unknown sample type.")            # src/solver.cpp:491
    user = dict(SHIPPED[name])
    user.update(overrides)
    dim, n_default = SAMPLES[name]
    p = resolve(user, dim)
    p.setdefault("N", n_default)
    p["sample"] = name
    p["DIM"] = dim
    return p
