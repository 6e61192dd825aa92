"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsphref_d*.so, built by
`make -C oracle ref` from the sources under /root/reference; only possible in the build container).

Per case: the reference's state after Solver::initialize (src/solver.cpp:353-414) and after each of
two Solver::integrate steps (417-429), dt and h_per_v_sig, the energy sums (src/output.cpp:72-83),
and the exact neighbour sets of the EXHAUSTIVE_SEARCH build (src/exhaustive_search.cpp:11-42) for
the post-initialize smoothing lengths, gather and symmetric.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from sphcode_b200 import sample_params, make_sample  # noqa: E402
from oracle.refsim import RefSim, build  # noqa: E402

# small versions of the BASELINE configs + one case per kernel / SPH type / DIM combination
GOLDEN = {
    "shock_tube_c1": ("shock_tube", dict(N=50)),
    "khi_disph_ac": ("khi", dict(N=32, SPHType="disph", useArtificialConductivity=True)),
    "khi_ssph": ("khi", dict(N=32)),
    "gresho_gsph2": ("gresho_chan_vortex", dict(N=32, SPHType="gsph", use2ndOrderGSPH=True)),
    "pairing_cubic": ("pairing_instability", dict(N=24)),
    "evrard_c4": ("evrard", dict(N=12)),
    "evrard_ssph_cubic": ("evrard", dict(N=10, SPHType="ssph", kernel="cubic_spline")),
    "evrard_gsph": ("evrard", dict(N=10, SPHType="gsph")),
}
STEPS = 2


def main():
    build("ref")
    for name, (sample, over) in GOLDEN.items():
        p = sample_params(sample, **over)
        parts = make_sample(p)
        dim = p["DIM"]
        ref = RefSim(p, parts, dim, "tree")
        ref.initialize()
        out = {"ic": parts, "state0": ref.particles, "hpvs0": ref.h_per_v_sig, "energy0": ref.energy()}
        if p["SPHType"] == "gsph":
            for nm in ["grad_density", "grad_pressure"] + [f"grad_velocity_{k}" for k in range(dim)]:
                out["g0_" + nm] = ref.vector_array(nm)
        ex = RefSim(p, out["state0"], dim, "exhaustive")
        og, ig = ex.neighbor_lists(symmetric=False)
        os_, is_ = ex.neighbor_lists(symmetric=True)
        out.update(nl_gather_off=og, nl_gather_ids=ig, nl_sym_off=os_, nl_sym_ids=is_)
        for s in range(1, STEPS + 1):
            dt = ref.integrate()
            out[f"state{s}"] = ref.particles
            out[f"dt{s}"] = dt
            out[f"hpvs{s}"] = ref.h_per_v_sig
            out[f"energy{s}"] = ref.energy()
        out["params_json"] = np.array(repr(sorted((k, v) for k, v in p.items())))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, len(parts), "particles ->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
