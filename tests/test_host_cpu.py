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


def _slices(n, world):
    """Mirror of my_slice() in sphcode_b200/csrc/sphb_api.cu: rank r owns groups of 32 particles
    [r * slice_groups, (r + 1) * slice_groups) of the sorted order."""
    groups = -(-n // 32)
    sg = -(-groups // world)
    out = []
    for r in range(world):
        f = min(r * sg * 32, n)
        l = min(n, r * sg * 32 + sg * 32)
        out.append((f, max(0, l - f)))
    return out, sg * world * 32


@pytest.mark.parametrize("n,world", [(500, 2), (998592, 8), (33, 4), (15902832, 8), (31, 2)])
def test_slices_tile_the_particle_range(n, world):
    sl, n_pad = _slices(n, world)
    assert sum(c for _, c in sl) == n and n_pad >= n and n_pad % (32 * world) == 0
    pos = 0
    for f, c in sl:
        assert f == pos or c == 0
        assert c == 0 or f % 32 == 0
        pos += c


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
    # replicated state, sliced compute, all-gather of the slice results (what gather_d() does with NCCL)
    n = 1000
    sl, n_pad = _slices(n, world)
    mine = torch.zeros(n_pad // world, dtype=torch.float64)
    f, c = sl[rank]
    mine[:c] = torch.arange(f, f + c, dtype=torch.float64) * 2.0
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    whole = torch.cat(parts)[:n]
    # dt: min over the ranks' slice minima (all-reduce min)
    dtm = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(dtm, op=dist.ReduceOp.MIN)
    q.put((rank, bytes(uid.numpy().tobytes()), bool(torch.equal(whole, torch.arange(n, dtype=torch.float64) * 2.0)), float(dtm)))
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
    assert all(r[2] for r in res) and all(r[3] == 1.0 for r in res)


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
