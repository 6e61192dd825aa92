"""CPU tests (-m "not gpu") of the host side: the C-ABI library loads and exports every symbol that
include/sphb.h declares (no compute calls without a GPU), the parameter surface keeps the
reference's keys / defaults / errors, the sample generators, and the N>1 host logic under gloo."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import parity_util as U
from sphcode_b200 import params as P
from sphcode_b200 import samples as S
from sphcode_b200 import lib


@pytest.fixture(scope="session")
def sphb_lib():
    lib.build()                     # nvcc cross-compiles without a GPU
    return ctypes.CDLL(lib.LIB_PATH)


def test_cabi_exports_every_declared_symbol(sphb_lib):
    hdr = open(os.path.join(U.ROOT, "include", "sphb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(sphb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(sphb_lib, name), f"libsphb.so lacks {name}"
    assert sorted(lib.SYMBOLS) == declared, "sphcode_b200/lib.py binds a different set than include/sphb.h declares"
    for dim, size in ((1, 144), (2, 176), (3, 208)):        # sizeof(SPHParticle), include/particle.hpp:8-33
        sphb_lib.sphb_sizeof_particle.restype = ctypes.c_size_t
        assert sphb_lib.sphb_sizeof_particle(dim) == size == S.particle_dtype(dim).itemsize


def test_no_cpu_fallback(sphb_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.SphbError, match="no CUDA device"):
        lib.Context(P.sample_params("shock_tube"), 1)


def test_params_struct_layout_matches_header():
    # sphb_params in include/sphb.h: 2 ints, 3 doubles, 2 ints, 3 doubles, 2 ints, 1 double, 4 ints, 1 double,
    # 2 ints, 6 doubles, 2 doubles, 2 ints
    assert ctypes.sizeof(lib.SphbParams) == 8 + 24 + 8 + 24 + 8 + 8 + 16 + 8 + 8 + 48 + 16 + 8


def test_parameter_defaults_and_errors():
    p = P.resolve({"endTime": 1.0, "gamma": 1.4})
    assert p["SPHType"] == "ssph" and p["cflSound"] == 0.3 and p["cflForce"] == 0.125          # src/solver.cpp:205-219
    assert p["neighborNumber"] == 32 and p["leafParticleNumber"] == 1 and p["maxTreeLevel"] == 20
    assert p["outputTime"] == pytest.approx(0.01) and p["energyTime"] == p["outputTime"]
    assert p["iterativeSmoothingLength"] is True and p["theta"] == 0.5 and p["G"] == 1.0
    with pytest.raises(P.SPHParameterError):
        P.resolve({"gamma": 1.4})
    with pytest.raises(P.SPHParameterError, match="Unknown SPH type"):
        P.resolve({"endTime": 1, "gamma": 1.4, "SPHType": "xsph"})
    with pytest.raises(P.SPHParameterError, match="kernel is unknown"):
        P.resolve({"endTime": 1, "gamma": 1.4, "kernel": "gauss"})
    with pytest.raises(P.SPHParameterError, match="alphaMax < alphaMin"):
        P.resolve({"endTime": 1, "gamma": 1.4, "useTimeDependentAV": True, "alphaMax": 0.05})
    with pytest.raises(P.SPHParameterError, match="rangeMax != DIM"):
        P.resolve({"endTime": 1, "gamma": 1.4, "periodic": True, "rangeMax": [1.0], "rangeMin": [0.0]}, dim=2)
    with pytest.raises(P.SPHParameterError):
        P.sample_params("no_such_sample")
    assert set(P.SAMPLES) == {"shock_tube", "gresho_chan_vortex", "pairing_instability", "hydrostatic", "khi", "evrard"}


def test_sample_generators():
    st = S.shock_tube(50, 1.4)                                   # src/sample/shock_tube.cpp:18-51
    assert len(st) == 500 and np.isclose(st["pos"][0, 0], -0.5 + 0.00125)
    assert np.count_nonzero(st["dens"] == 1.0) == 400 and np.count_nonzero(st["dens"] == 0.25) == 100
    assert np.allclose(st["mass"], 0.0025)
    k = S.khi(64, 5.0 / 3.0)                                     # src/sample/khi.cpp:18-73
    assert len(k) == 64 * 64 * 3 // 4 and set(np.unique(k["dens"])) == {1.0, 2.0}
    ev = S.evrard(20, 5.0 / 3.0)                                 # src/sample/evrard.cpp:19-63
    r = np.sqrt((ev["pos"] ** 2).sum(axis=1))
    assert len(ev) == 4224 and r.max() <= 1.0 and np.isclose(ev["mass"].sum(), 1.0)
    assert np.allclose(ev["dens"], 1.0 / (2 * np.pi * r))
    g = S.gresho_chan_vortex(32, 5.0 / 3.0)
    assert len(g) == 1024 and np.isclose(g["mass"].sum(), 1.0)
    pi1, pi2 = S.pairing_instability(16, 5.0 / 3.0), S.pairing_instability(16, 5.0 / 3.0)
    assert np.array_equal(pi1["pos"], pi2["pos"])                # mt19937(1): deterministic
    h = S.hydrostatic(16, 5.0 / 3.0)
    assert set(np.unique(h["dens"])) == {1.0, 4.0}


def _dd_bookkeeping(keys_by_rank, split):
    """Host side of the multi-GPU migration (migrate_t / set_offsets in sphcode_b200/csrc/sphb_api.cu), restated: rank d owns
    the keys in [split[d], split[d + 1]); every rank counts its leavers per destination, the count matrix is all-gathered,
    and every rank derives ALL ranks' new counts, the offsets of the global tree order, its sort length (arrivals first
    fill the leavers' slots) and the source offsets of the blocks it pulls."""
    W = len(keys_by_rank)
    mat = np.zeros((W, W), dtype=np.int64)
    for r, k in enumerate(keys_by_rank):
        dest = np.searchsorted(np.asarray(split[1:], dtype=np.uint64), k, side="right")
        for d in range(W):
            if d != r:
                mat[r, d] = int(np.sum(dest == d))
    n_old = np.array([len(k) for k in keys_by_rank])
    n_new = n_old - mat.sum(axis=1) + mat.sum(axis=0)
    off = np.concatenate([[0], np.cumsum(n_new)])
    n_sort = n_old + np.maximum(0, mat.sum(axis=0) - mat.sum(axis=1))
    src_off = np.array([[mat[s, :r].sum() for s in range(W)] for r in range(W)])      # [receiver][sender]
    return mat, n_new, off, n_sort, src_off


@pytest.mark.parametrize("n,world", [(500, 2), (998592, 8), (33, 4), (4096, 3)])
def test_domain_decomposition_bookkeeping(n, world):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 60, size=n, dtype=np.uint64)
    allk = np.sort(keys)
    split = [0] + [int(allk[n * q // world]) for q in range(1, world)]                 # k_next_splitters
    parts = np.array_split(rng.permutation(keys), world)                                # an arbitrary initial split
    mat, n_new, off, n_sort, src_off = _dd_bookkeeping(parts, split)
    assert n_new.sum() == n and off[-1] == n and np.all(n_new >= 0)
    assert abs(int(n_new.max()) - int(n_new.min())) <= 2                               # exact quantile splitters balance the ranks
    # after the move every rank holds exactly the keys of its range, and the ranges tile the sorted order
    for r in range(world):
        lo = split[r]
        hi = split[r + 1] if r + 1 < world else 1 << 62
        mine = allk[(allk >= lo) & (allk < hi)]
        assert len(mine) == n_new[r]
        assert np.array_equal(mine, allk[off[r]:off[r + 1]])
        assert n_sort[r] >= n_new[r] and n_sort[r] <= max(len(parts[r]), n_new[r])
        for snd in range(world):
            assert src_off[r][snd] + mat[snd, r] <= mat[snd].sum()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the unique-id hand-off of bench.py: rank 0 creates 128 bytes, everyone receives the same bytes
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(uid, 0)
    # migration bookkeeping across real processes: every rank only knows its own keys, the leaver counts are all-gathered
    # (ncclAllGather in libsphb) and every rank must derive the same global picture
    n = 4000
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << 60, size=n, dtype=np.uint64)
    allk = np.sort(keys)
    split = [0] + [int(allk[n * k // world]) for k in range(1, world)]
    parts = np.array_split(rng.permutation(keys), world)
    mine = parts[rank]
    dest = np.searchsorted(np.asarray(split[1:], dtype=np.uint64), mine, side="right")
    row = torch.tensor([int(np.sum(dest == d)) if d != rank else 0 for d in range(world)], dtype=torch.int64)
    rows = [torch.zeros_like(row) for _ in range(world)]
    dist.all_gather(rows, row)
    mat = torch.stack(rows).numpy()
    ref_mat, n_new, off, n_sort, _ = _dd_bookkeeping(parts, split)
    # dt / h_per_v_sig: min over the ranks' minima; energies: sum (all-reduce in libsphb)
    dtm = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(dtm, op=dist.ReduceOp.MIN)
    es = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(es)
    q.put((rank, bytes(uid.numpy().tobytes()), bool(np.array_equal(mat, ref_mat)), float(dtm), float(es) == n,
           int(len(mine) - mat[rank].sum() + mat[:, rank].sum()) == int(n_new[rank])))
    dist.destroy_process_group()


def test_multi_rank_host_logic_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29731
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p_ in procs:
        p_.join(timeout=60)
    assert all(r[1] == bytes(range(128)) for r in res)
    assert all(r[2] for r in res) and all(r[3] == 1.0 for r in res) and all(r[4] and r[5] for r in res)


def _sph_gpu():
    import os
    import subprocess
    exe = os.path.join(U.ROOT, "sphcode_b200", "host", "sph_gpu")
    if not os.path.exists(exe):
        r = subprocess.run(["make", "-C", os.path.dirname(exe), "sph_gpu"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    return exe


@pytest.mark.parametrize("name,N", [("shock_tube", 50), ("khi", 32), ("gresho_chan_vortex", 24), ("pairing_instability", 16),
                                    ("hydrostatic", 16), ("evrard", 14)])
def test_cli_initial_conditions_match_generators(name, N, tmp_path):
    """sph_gpu (C++ host: JSON reader, sample registry, generators) builds the same particle
    set as the Python restatement of src/sample/*.cpp that the golden vectors were made from: bit-exact, except
    evrard where libm's pow and numpy's differ in the last bit."""
    import subprocess
    from sphcode_b200 import sample_params, make_sample
    from sphcode_b200.samples import particle_dtype
    out = str(tmp_path / "ic.bin")
    r = subprocess.run([_sph_gpu(), name, "--set", f"N={N}", "--dump-ic", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p = sample_params(name, N=N)
    ref = make_sample(p)
    got = np.fromfile(out, dtype=particle_dtype(p["DIM"]))
    assert len(got) == len(ref)
    for f in ref.dtype.names:
        if f == "next":
            continue
        if name == "evrard":
            np.testing.assert_allclose(got[f], ref[f], rtol=2e-15, atol=0)
        else:
            assert np.array_equal(got[f], ref[f]), f


def test_cli_parameter_errors(tmp_path):
    """Error texts and exit status of Solver::read_parameterfile (src/solver.cpp:155-299) / exception_handler."""
    import json
    import subprocess
    exe = _sph_gpu()
    ic = str(tmp_path / "ic.bin")

    def run(*args):
        r = subprocess.run([exe, *args, "--dump-ic", ic], capture_output=True, text=True)
        return r.returncode, r.stderr

    assert run("evrard", "--set", "SPHType=foo") == (1, "error: Unknown SPH type\n")
    assert run("evrard", "--set", "kernel=gauss") == (1, "error: kernel is unknown.\n")
    assert run("khi", "--set", "rangeMax=[1.0]") == (1, "error: rangeMax != DIM\n")
    assert run("khi", "--set", "endTime=-1") == (1, "error: endTime < startTime\n")
    assert run("khi", "--set", "useTimeDependentAV=true", "--set", "alphaMax=0.01") == (1, "error: alphaMax < alphaMin\n")
    rc, err = run(str(tmp_path / "missing.json"))
    assert rc == 1 and "cannot open file" in err
    # a parameter file of the reference has no sample: "unknown sample type." (src/solver.cpp:491); with the key it runs
    pf = tmp_path / "p.json"
    pf.write_text(json.dumps({"outputDirectory": str(tmp_path), "endTime": 0.1, "gamma": 1.4}))
    assert run(str(pf)) == (1, "error: unknown sample type.\n")
    pf.write_text(json.dumps({"outputDirectory": str(tmp_path), "endTime": 0.1, "gamma": 1.4, "sample": "shock_tube", "N": 20,
                              "periodic": True, "rangeMax": [1.5], "rangeMin": [-0.5], "neighborNumber": 4}))
    assert run(str(pf))[0] == 0
    pf.write_text(json.dumps({"outputDirectory": str(tmp_path), "endTime": 0.1, "sample": "shock_tube"}))
    assert run(str(pf)) == (1, "error: No such node (gamma)\n")


def test_cli_json_reader_accepts_what_the_reference_accepts(tmp_path):
    """The parameter reader of sph_gpu (own JSON reader, no Boost): numbers in any JSON spelling and quoted (the
    reference goes through std::stod on the node text), booleans as true/false, unknown keys ignored, arrays for the
    periodic range, `--set` overriding the file."""
    import subprocess
    from sphcode_b200 import sample_params, make_sample
    from sphcode_b200.samples import particle_dtype
    exe = _sph_gpu()
    pf = tmp_path / "p.json"
    pf.write_text("""{
        "sample" : "khi", "outputDirectory" : "%s", "unknownKey" : {"ignored": "no"} ,
        "endTime":1e-1, "gamma" : "1.4", "N": 16, "periodic":true,
        "rangeMax" : [ 1.0 , 1.0 ], "rangeMin":[0,0],
        "neighborNumber":32,"kernel":"wendland", "SPHType" : "disph"
    }""" % tmp_path)
    ic = str(tmp_path / "ic.bin")
    r = subprocess.run([exe, str(pf), "--dump-ic", ic], capture_output=True, text=True)
    # nested objects are not part of the reference's parameter surface: rejected with a parse error, not a crash
    assert r.returncode == 1 and "json:" in r.stderr
    pf.write_text(pf.read_text().replace('"unknownKey" : {"ignored": "no"} ,', '"unknownKey" : "ignored",'))
    r = subprocess.run([exe, str(pf), "--set", "N=24", "--dump-ic", ic], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(ic, dtype=particle_dtype(2))
    ref = make_sample(sample_params("khi", N=24, gamma=1.4))
    assert len(got) == len(ref) and np.array_equal(got["pos"], ref["pos"]) and np.array_equal(got["ene"], ref["ene"])


# ---------------------------------------------------------------------------------------------------------------------
# f-3 / f-4 pinned to the reference itself: the whole unmodified Solver (src/solver.cpp, src/sample/*.cpp) built against
# the Boost stand-in (oracle/Makefile `stock`), driven through oracle/ref_solver_driver.cpp.  Needs /root/reference at
# build time and for the shipped JSON files, i.e. runs in the build container only.
# ---------------------------------------------------------------------------------------------------------------------
REF_ROOT = "/root/reference"
SAMPLE_DIM = {k: v[0] for k, v in P.SAMPLES.items()}
GENERATOR_FIELDS = ("pos", "vel", "mass", "dens", "ene", "pres", "id")       # what src/sample/*.cpp set (hydrostatic: no vel)


def _ref_solver_or_skip(dim):
    from oracle import refsim
    if not os.path.exists(refsim.solver_lib_path(dim)) or not os.path.isdir(REF_ROOT):
        pytest.skip("oracle/_ref/libsphsolver_d*.so not built (needs /root/reference: `make -C oracle stock`)")
    return refsim


def _reference_workdir(tmp_path, name, edits=None, drop=()):
    """A directory laid out like the reference root as far as Solver needs it: sample/<name>/<name>.json = the SHIPPED
    file with `edits`, outputDirectory inside the tmp dir (the reference creates it, src/logger.cpp:24-37)."""
    import json
    d = tmp_path / "refroot"
    (d / "sample" / name).mkdir(parents=True, exist_ok=True)
    j = json.load(open(os.path.join(REF_ROOT, "sample", name, name + ".json")))
    j.update(edits or {})
    for k in drop:
        j.pop(k, None)
    j["outputDirectory"] = str(d / "results")
    (d / "sample" / name / (name + ".json")).write_text(json.dumps(j))
    return d, j


def _dump_params(exe, arg, cwd, tmp_path, extra=()):
    import subprocess
    out = str(tmp_path / "params.txt")
    r = subprocess.run([exe, arg, *extra, "--dump-params", out], capture_output=True, text=True, cwd=cwd)
    if r.returncode:
        return None, (r.stdout + r.stderr)
    kv = {}
    for line in open(out):
        k, _, v = line.rstrip("\n").partition(" ")
        kv[k] = v
    return kv, ""


@pytest.mark.parametrize("name,N", [("shock_tube", 50), ("shock_tube", 37), ("khi", 32), ("khi", 50), ("gresho_chan_vortex", 24),
                                    ("pairing_instability", 16), ("hydrostatic", 16), ("hydrostatic", 24), ("evrard", 14), ("evrard", 30)])
def test_generators_pinned_to_the_reference(name, N, tmp_path, monkeypatch):
    """src/sample/*.cpp through the unmodified Solver::make_initial_condition against (a) sph_gpu --dump-ic (C++ host,
    parallel two-pass fill) and (b) sphcode_b200.samples (numpy, what the golden vectors were made from): same particle
    count and order, every member the generator sets equal — bit for bit, except where the reference's -ffast-math pow /
    sin / exp / divisions differ from libm / numpy in the last bits (tolerance 1e-14 of the member's magnitude, reported)."""
    import subprocess
    dim = SAMPLE_DIM[name]
    refsim = _ref_solver_or_skip(dim)
    d, j = _reference_workdir(tmp_path, name, {"N": N})
    monkeypatch.chdir(d)
    s = refsim.RefSolver(name, dim)
    assert s.params.n_side == N
    ref = s.initial_condition()
    ic = str(tmp_path / "ic.bin")
    r = subprocess.run([_sph_gpu(), name, "--dump-ic", ic], capture_output=True, text=True, cwd=d)       # reads the same sample/<name>/<name>.json
    assert r.returncode == 0, r.stdout + r.stderr
    cpp = np.fromfile(ic, dtype=S.particle_dtype(dim))
    py = S.make_sample(P.sample_params(name, N=N))
    fields = [f for f in GENERATOR_FIELDS if not (name == "hydrostatic" and f == "vel")]
    for what, got in (("sph_gpu", cpp), ("samples.py", py)):
        assert len(got) == len(ref), (what, len(got), len(ref))
        worst = 0.0
        for f in fields:
            a, b = got[f], ref[f]
            if np.array_equal(a, b):
                continue
            assert a.dtype.kind == "f", (what, f)
            # relative to the member's magnitude over the set: sin(4 pi x) near a zero crossing (khi's vy) carries the
            # rounding of its argument, an absolute 2e-16 * |4 pi x| on an amplitude of 0.1
            err = np.abs(a - b).max() / np.abs(b).max()
            worst = max(worst, err)
            assert err <= 1e-14, (what, f, err)
        print(f"{name} N={N} {what}: n={len(ref)} worst relative difference {worst:.1e}")


_PARAM_CASES = [
    ("shock_tube", {}, ()), ("gresho_chan_vortex", {}, ()), ("pairing_instability", {}, ()), ("hydrostatic", {}, ()), ("khi", {}, ()), ("evrard", {}, ()),
    ("evrard", {"N": 124, "theta": 0.7, "G": 2.5, "energyTime": 0.25, "startTime": 0.5, "cflSound": 0.25, "cflForce": 0.1}, ()),
    ("khi", {"SPHType": "disph", "useArtificialConductivity": True, "alphaAC": 0.5, "N": 1152, "alphaMax": 3.0, "alphaMin": 0.2, "epsilonAV": 0.3}, ()),
    ("gresho_chan_vortex", {"SPHType": "gsph", "use2ndOrderGSPH": False, "N": 2048, "maxTreeLevel": 12, "iterativeSmoothingLength": False}, ()),
    ("shock_tube", {"useTimeDependentAV": True, "avAlpha": 2.0}, ("iterativeSmoothingLength", "kernel", "N")),      # defaults of dropped keys
    ("pairing_instability", {"useBalsaraSwitch": False, "outputTime": 0.125}, ("leafParticleNumber", "neighborNumber")),
]


@pytest.mark.parametrize("name,edits,drop", _PARAM_CASES)
def test_parameter_reader_pinned_to_the_reference(name, edits, drop, tmp_path, monkeypatch):
    """Solver::read_parameterfile (src/solver.cpp:155-299, Boost property_tree) against the Boost-free readers: every
    member of SPHParameters, the time block, N and the output directory as parsed by the unmodified reference equal
    what sph_gpu --dump-params (C++) and sphcode_b200.params.resolve (Python) make of the same file."""
    dim = SAMPLE_DIM[name]
    refsim = _ref_solver_or_skip(dim)
    d, j = _reference_workdir(tmp_path, name, edits, drop)
    monkeypatch.chdir(d)
    rp = refsim.RefSolver(name, dim).params
    kv, err = _dump_params(_sph_gpu(), name, d, tmp_path)
    assert kv is not None, err
    py = P.resolve({k: v for k, v in j.items() if k != "N"}, dim)
    sp = lib.to_sphb_params(py)
    n_py = j.get("N", P.SAMPLES[name][1])
    ref = {
        "outputDirectory": rp.output_dir.decode(), "startTime": rp.t_start, "endTime": rp.t_end, "outputTime": rp.t_output, "energyTime": rp.t_energy,
        "N": rp.n_side, "sph_type": rp.sph_type, "kernel": rp.kernel, "cfl_sound": rp.cfl_sound, "cfl_force": rp.cfl_force, "av_alpha": rp.av_alpha,
        "use_balsara_switch": rp.use_balsara, "use_time_dependent_av": rp.use_tdav, "use_ac": rp.use_ac,
        "max_tree_level": rp.max_tree_level, "leaf_particle_num": rp.leaf_particle_num, "neighbor_number": rp.neighbor_number,
        "iterative_sml": rp.iterative_sml, "gamma": rp.gamma, "periodic": rp.periodic, "use_gravity": rp.use_gravity,
    }
    # members the reference only reads when their switch is on (it leaves them unset otherwise, src/solver.cpp:222-235,287-297)
    if rp.use_tdav:
        ref.update(alpha_max=rp.alpha_max, alpha_min=rp.alpha_min, epsilon_av=rp.epsilon_av)
    if rp.use_ac:
        ref["alpha_ac"] = rp.alpha_ac
    if rp.use_gravity:
        ref.update(G=rp.G, theta=rp.theta)
    if rp.sph_type == 2:
        ref["gsph_2nd_order"] = rp.gsph_2nd_order
    if rp.periodic:
        for k in range(dim):
            ref[f"range_max{k}"] = rp.range_max[k]
            ref[f"range_min{k}"] = rp.range_min[k]
    pyv = {"outputDirectory": py["outputDirectory"], "startTime": py["startTime"], "endTime": py["endTime"], "outputTime": py["outputTime"],
           "energyTime": py["energyTime"], "N": n_py}
    for k in ref:
        want = ref[k]
        got_cpp = kv[k]
        if isinstance(want, str):
            assert got_cpp == want, (k, got_cpp, want)
        elif isinstance(want, float):
            assert float(got_cpp) == want, (k, got_cpp, want)
        else:
            assert int(got_cpp) == int(want), (k, got_cpp, want)
        if k in pyv:
            got_py = pyv[k]
        elif k.startswith("range_m"):
            got_py = getattr(sp, k[:9])[int(k[9:])]
        else:
            got_py = getattr(sp, k)
        assert (got_py == want) if isinstance(want, (str, float)) else (int(got_py) == int(want)), ("python", k, got_py, want)


@pytest.mark.parametrize("name,edits,drop,text", [
    ("evrard", {"endTime": -1.0}, (), "endTime < startTime"),
    ("khi", {"SPHType": "xsph"}, (), "Unknown SPH type"),
    ("khi", {"kernel": "gauss"}, (), "kernel is unknown."),
    ("khi", {"alphaMax": 0.05}, (), "alphaMax < alphaMin"),
    ("khi", {"rangeMax": [1.0]}, (), "rangeMax != DIM"),
    ("evrard", {}, ("gamma",), "No such node (gamma)"),
    ("evrard", {}, ("endTime",), "No such node (endTime)"),
])
def test_parameter_errors_pinned_to_the_reference(name, edits, drop, text, tmp_path, monkeypatch):
    """The same bad files are refused by the reference, by sph_gpu and by params.resolve with the reference's message."""
    dim = SAMPLE_DIM[name]
    refsim = _ref_solver_or_skip(dim)
    d, j = _reference_workdir(tmp_path, name, edits, drop)
    monkeypatch.chdir(d)
    with pytest.raises(RuntimeError) as ei:
        refsim.RefSolver(name, dim)
    assert text in str(ei.value), str(ei.value)
    kv, err = _dump_params(_sph_gpu(), name, d, tmp_path)
    assert kv is None and text in err, err
    with pytest.raises(P.SPHParameterError) as ej:
        P.resolve({k: v for k, v in j.items() if k != "N"}, dim)
    assert text in str(ej.value)


def test_bench_accounting_matches_survey():
    """bench.py's algorithmic FLOP constants and byte counts: the DIM = 3 DISPH values are SURVEY.md 8d's (57 Newton eval,
    78 + 57 per density / Balsara neighbour, 141 force pair, 78 / 15 gravity pair / cell; 140 / 152 B per particle), the
    other formulations follow the same counting rule; every --config resolves to the BASELINE workload it names."""
    sys.path.insert(0, U.ROOT)
    import bench
    k = bench.flop_constants(3, "disph", True)
    assert (k["newton"], k["dens"], k["bal"], k["pair"], k["pp"], k["pc"]) == (57, 78, 57, 141, 78, 15)
    b = bench.alg_bytes(3, True)
    assert (b["pre"], b["fluid"], b["gravity"]) == (140, 152, 120)
    assert bench.flop_constants(2, "gsph", True)["pair"] == 207 and bench.flop_constants(2, "disph", True)["pair"] == 123
    assert bench.flop_constants(1, "ssph", True)["bal"] == 0            # no Balsara loop in 1-D (src/pre_interaction.cpp:106)
    expect = {"c1": ("shock_tube", 1, 50), "c2": ("khi", 2, 1152), "c3": ("gresho_chan_vortex", 2, 2048), "c4": ("evrard", 3, 124),
              "c5": ("evrard", 3, 312), "c5_64m": ("evrard", 3, 496)}
    for name, (sample, dim, n) in expect.items():
        p = bench.config_params(name)
        assert (p["sample"], p["DIM"], p["N"]) == (sample, dim, n), name
    assert bench.config_params("c2")["SPHType"] == "disph" and bench.config_params("c2")["useArtificialConductivity"]
    assert bench.config_params("c3")["SPHType"] == "gsph" and bench.config_params("c3")["use2ndOrderGSPH"]
    assert bench.config_params("c5")["useGravity"] and bench.config_params("c5")["theta"] == 0.5
