"""CPU tests (-m "not gpu"): pin the oracle.

* the plain-C port oracle/sph_oracle.c against the committed golden vectors generated from the
  UNMODIFIED reference (tests/golden/make_golden.py) — state after initialize and two integrate
  steps, dt, h_per_v_sig, energies, exact neighbour sets (tree search and brute force);
* the port against the unmodified reference run live (oracle/_ref/*.so) when that library exists
  (it does wherever /root/reference is mounted);
* the reference's only known-answer material for this path: the kernel-derivative self-test of
  test/kernel_test/kernel_test.cpp (second-order agreement of dw / dhw with central differences).
"""
import glob
import os
import sys

import numpy as np
import pytest

import parity_util as U
from parity_util import RTOL

sys.path.insert(0, U.GOLDEN_DIR)
from make_golden import GOLDEN  # noqa: E402
from sphcode_b200 import sample_params  # noqa: E402
from oracle import refsim  # noqa: E402
from oracle.refsim import RefSim  # noqa: E402


@pytest.fixture(scope="session", autouse=True)
def _build_port():
    refsim.build("port")


GOLD = sorted(n for n in (os.path.basename(f)[:-4] for f in glob.glob(U.golden_path("*"))) if n in GOLDEN)


def _gp(name):
    sample, over = GOLDEN[name]
    return sample_params(sample, **over)


@pytest.mark.parametrize("name", GOLD)
def test_port_matches_golden(name):
    g = np.load(U.golden_path(name))
    p = _gp(name)
    sim = RefSim(p, g["ic"], p["DIM"], "port")
    sim.initialize()
    U.assert_fields(sim.particles, g["state0"], U.PRE_FIELDS + U.FORCE_FIELDS, what=f"{name} initialize", params=p)
    assert abs(sim.h_per_v_sig - float(g["hpvs0"])) <= RTOL * float(g["hpvs0"])
    np.testing.assert_allclose(sim.energy(), g["energy0"], rtol=1e-9, atol=1e-14)
    for s in (1, 2):
        dt = sim.integrate()
        assert abs(dt - float(g[f"dt{s}"])) <= RTOL * float(g[f"dt{s}"])
        U.assert_fields(sim.particles, g[f"state{s}"], U.STEP_FIELDS, what=f"{name} step {s}", params=p)


@pytest.mark.parametrize("name", GOLD)
def test_port_neighbor_sets_match_golden(name):
    g = np.load(U.golden_path(name))
    p = _gp(name)
    sim = RefSim(p, g["state0"], p["DIM"], "port")
    for sym, ko, ki in ((False, "nl_gather_off", "nl_gather_ids"), (True, "nl_sym_off", "nl_sym_ids")):
        off, ids = sim.neighbor_lists(symmetric=sym, exhaustive=True)
        assert np.array_equal(off, g[ko]) and np.array_equal(ids, g[ki]), (name, sym)
    # the tree search returns the same gather sets (src/bhtree.cpp:251-261 vs exhaustive_search.cpp:22-33)
    sim.make_tree()
    off, ids = sim.neighbor_lists(symmetric=False)
    assert np.array_equal(off, g["nl_gather_off"]) and np.array_equal(ids, g["nl_gather_ids"])


@pytest.mark.parametrize("name", sorted(U.CONFIGS))
def test_port_matches_live_reference(name):
    p, parts = U.make_case(name)
    if not refsim.available(p["DIM"], "tree"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = RefSim(p, parts, p["DIM"], "tree")
    port = RefSim(p, parts, p["DIM"], "port")
    ref.initialize(); port.initialize()
    U.assert_fields(port.particles, ref.particles, U.PRE_FIELDS + U.FORCE_FIELDS, what=f"{name} initialize", params=p)
    for s in range(2):
        a, b = ref.integrate(), port.integrate()
        assert abs(a - b) <= RTOL * a
        U.assert_fields(port.particles, ref.particles, U.STEP_FIELDS, what=f"{name} step {s + 1}", params=p)
    np.testing.assert_allclose(port.energy(), ref.energy(), rtol=1e-9, atol=1e-14)


def test_reference_tree_vs_exhaustive_noise_floor():
    """The reference against ITSELF (tree vs EXHAUSTIVE_SEARCH build): the summation-order noise that
    defines what 'relative 1e-10' can mean."""
    p, parts = U.make_case("evrard_c4")
    if not (refsim.available(3, "tree") and refsim.available(3, "exhaustive")):
        pytest.skip("oracle/_ref not built")
    a, b = RefSim(p, parts, 3, "tree"), RefSim(p, parts, 3, "exhaustive")
    for s in (a, b):
        s.init_state(); s.make_tree(); s.pre(); s.fluid()
    e = U.assert_fields(a.particles, b.particles, U.PRE_FIELDS + ("acc", "dene"), what="tree vs exhaustive", params=p)
    assert e["sml"] == 0.0 and e["dens"] == 0.0


@pytest.mark.parametrize("dim,kernel", [(1, "cubic_spline"), (2, "cubic_spline"), (3, "cubic_spline"), (2, "wendland"), (3, "wendland")])
def test_kernel_derivatives_known_answer(dim, kernel):
    """test/kernel_test/kernel_test.cpp:8-85: mean |dw - central difference of w| and
    |dhw - central difference in h| fall by 4x per doubling of n (second order)."""
    name = {1: "shock_tube", 2: "khi", 3: "evrard"}[dim]
    p = sample_params(name, kernel=kernel, N=4 if dim > 1 else 2)
    from sphcode_b200 import make_sample
    sim = RefSim(p, make_sample(p), dim, "port")
    errs = []
    for n in (100, 200, 400):
        dx = 1.0 / n
        e_dw = e_dh = 0.0
        for i in range(1, n):
            r = i * dx
            rij = np.zeros(dim); rij[0] = r
            w, dhw, dw = sim.kernel_eval(rij, 1.0)
            rp = rij.copy(); rp[0] += dx * 0.5
            rm = rij.copy(); rm[0] -= dx * 0.5
            wp, wm = sim.kernel_eval(rp, 1.0)[0], sim.kernel_eval(rm, 1.0)[0]
            e_dw += abs(dw[0] - (wp - wm) / dx)
            hp, hm = sim.kernel_eval(rij, 1.0 + dx * 0.5)[0], sim.kernel_eval(rij, 1.0 - dx * 0.5)[0]
            e_dh += abs(dhw - (hp - hm) / dx)
        errs.append((e_dw / n, e_dh / n))
    for k in (0, 1):
        assert 3.5 < errs[0][k] / errs[1][k] < 4.5 and 3.5 < errs[1][k] / errs[2][k] < 4.5, errs
    if dim == 3 and refsim.available(3, "tree"):
        ref = RefSim(p, make_sample(p), 3, "tree")
        for r in (0.0, 0.1, 0.37, 0.77, 0.99, 1.2):
            rij = np.array([r * 0.6, r * 0.0, r * 0.8])
            a, b = sim.kernel_eval(rij, 0.9), ref.kernel_eval(rij, 0.9)
            np.testing.assert_allclose(np.hstack([a[0], a[1], a[2]]), np.hstack([b[0], b[1], b[2]]), rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("name", ["shock_tube", "khi", "evrard"])
def test_port_energy_history_matches_golden(name):
    """The port tracks the unmodified reference's energy history over 60 steps (tests/golden/
    make_energy_golden.py; evrard runs through maximum compression)."""
    from make_energy_golden import ENERGY_CASES, STEPS, history
    from sphcode_b200 import make_sample
    g = np.load(U.golden_path("energy_histories"))
    sample, over = ENERGY_CASES[name]
    p = sample_params(sample, **over)
    sim = RefSim(p, make_sample(p), p["DIM"], "port")
    sim.initialize()
    e, dts = history(sim, STEPS)
    ge, gdt = g[name + "_energy"], g[name + "_dt"]
    assert np.abs(e - ge).max() <= 1e-9 * np.abs(ge).max()
    assert (np.abs(dts - gdt) / gdt).max() <= 1e-9


@pytest.mark.parametrize("name", ["evrard_c4", "khi_disph_ac", "shock_tube_c1"])
def test_port_interaction_counters_are_consistent(name):
    """The port counts the interactions of the reference ALGORITHM (the numerators of the algorithmic-FLOP model,
    SURVEY.md 8d).  Self-consistency: the neighbour counter equals the sum of SPHParticle::neighbor, force pairs
    come in (i, j) / (j, i) twins, every Newton iteration evaluates at least the particle itself, and with gravity
    every particle visits the root and meets itself."""
    p, parts = U.make_case(name)
    sim = RefSim(p, parts, p["DIM"], "port")
    sim.initialize()
    sim.counters()
    sim.integrate()
    k = sim.counters()
    n = len(parts)
    assert k["pre_neighbors"] == int(sim.particles["neighbor"].sum())
    assert k["force_pairs"] % 2 == 0 and k["force_pairs"] > 0
    assert k["pre_candidates"] >= k["pre_neighbors"] >= n
    assert k["newton_iters"] >= n and k["newton_evals"] >= k["newton_iters"]
    if p["useGravity"]:
        assert k["grav_node_visits"] >= n and k["grav_pp"] >= n and k["grav_pc"] > 0
    else:
        assert k["grav_node_visits"] == k["grav_pp"] == k["grav_pc"] == 0
    assert sim.counters() == dict.fromkeys(RefSim.COUNTER_NAMES, 0)          # reset


def test_reference_subsample_mode_equals_the_full_pass():
    """oracle/ref_driver.cpp ref_set_active: the unmodified modules on particles 0..k-1 against all sources must give,
    bit for bit, what the full pass gives for those particles (this is what the 16 M parity check relies on), and
    ref_direct_gravity must equal the EXHAUSTIVE_SEARCH GravityForce."""
    from oracle import refsim
    if not (refsim.available(3, "tree") and refsim.available(3, "exhaustive")):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import parity_util as U
    p, parts = U.make_case("evrard_c4")
    n = len(parts)
    rng = np.random.default_rng(3)
    parts = parts[rng.permutation(n)]
    parts["id"] = np.arange(n)                  # the reference identifies particles by id == index (src/bhtree.cpp:257)
    full = refsim.RefSim(p, parts, 3, "tree")
    full.initialize()
    F = full.particles
    k = 300
    sub = refsim.RefSim(p, parts, 3, "tree")
    sub.init_state(); sub.make_tree(); sub.set_active(k); sub.pre()
    S = sub.particles
    for f in U.PRE_FIELDS:
        assert np.array_equal(S[f][:k], F[f][:k]), f
    state = F.copy()
    for f in ("acc", "dene", "phi"):
        state[f] = 0
    sub.set_active(0)
    sub.particles = state
    sub.make_tree(); sub.set_kernel(); sub.set_active(k)
    sub.fluid(); sub.gravity()
    S = sub.particles
    for f in U.FORCE_FIELDS:
        assert np.array_equal(S[f][:k], F[f][:k]), f
    assert not S["phi"][k:].any()
    fd, phid = sub.direct_gravity(64)
    ex = refsim.RefSim(p, state, 3, "exhaustive")
    ex.gravity()
    E = ex.particles
    np.testing.assert_allclose(fd, E["acc"][:64], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(phid, E["phi"][:64], rtol=1e-13)
