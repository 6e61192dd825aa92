"""TEST INFRASTRUCTURE — ctypes access to the two CPU checkers.

* ``RefSim(flavour="tree"|"exhaustive")``: the UNMODIFIED reference translation units compiled by
  oracle/Makefile into oracle/_ref/libsphref_d{1,2,3}[_ex].so (driver: oracle/ref_driver.cpp).
* ``RefSim(flavour="port")``: the plain-C restatement oracle/sph_oracle.c (oracle/libspho.so),
  which exports the same entry points.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing here is on the product path.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def particle_dtype(dim):
    """In-memory layout of sph::SPHParticle (reference include/particle.hpp:8-33)."""
    v = (np.float64, (dim,))
    return np.dtype([
        ("pos", *v), ("vel", *v), ("vel_p", *v), ("acc", *v),
        ("mass", "f8"), ("dens", "f8"), ("pres", "f8"), ("ene", "f8"), ("ene_p", "f8"),
        ("dene", "f8"), ("sml", "f8"), ("sound", "f8"), ("balsara", "f8"), ("alpha", "f8"),
        ("gradh", "f8"), ("phi", "f8"), ("id", "i4"), ("neighbor", "i4"), ("next", "u8"),
    ], align=True)


class RefParams(C.Structure):
    """Mirror of ref_params in oracle/ref_driver.cpp (= SPHParameters, include/parameters.hpp:20-79)."""
    _fields_ = [
        ("sph_type", C.c_int), ("kernel", C.c_int),
        ("cfl_sound", C.c_double), ("cfl_force", C.c_double),
        ("av_alpha", C.c_double),
        ("use_balsara", C.c_int), ("use_tdav", C.c_int),
        ("alpha_max", C.c_double), ("alpha_min", C.c_double), ("epsilon_av", C.c_double),
        ("use_ac", C.c_int),
        ("alpha_ac", C.c_double),
        ("max_tree_level", C.c_int), ("leaf_particle_num", C.c_int),
        ("neighbor_number", C.c_int),
        ("gamma", C.c_double),
        ("iterative_sml", C.c_int),
        ("periodic", C.c_int),
        ("range_max", C.c_double * 3), ("range_min", C.c_double * 3),
        ("use_gravity", C.c_int),
        ("G", C.c_double), ("theta", C.c_double),
        ("gsph_2nd_order", C.c_int),
    ]


_SPH = {"ssph": 0, "disph": 1, "gsph": 2}
_KER = {"cubic_spline": 0, "wendland": 1}


def to_ref_params(p):
    """p: dict with the reference's JSON keys (README.md:122-152) already defaulted
    (see sphcode_b200.params.SPHParameters.as_dict)."""
    r = RefParams()
    r.sph_type = _SPH[p["SPHType"]]
    r.kernel = _KER[p["kernel"]]
    r.cfl_sound, r.cfl_force = p["cflSound"], p["cflForce"]
    r.av_alpha = p["avAlpha"]
    r.use_balsara, r.use_tdav = int(p["useBalsaraSwitch"]), int(p["useTimeDependentAV"])
    r.alpha_max, r.alpha_min, r.epsilon_av = p["alphaMax"], p["alphaMin"], p["epsilonAV"]
    r.use_ac, r.alpha_ac = int(p["useArtificialConductivity"]), p["alphaAC"]
    r.max_tree_level, r.leaf_particle_num = p["maxTreeLevel"], p["leafParticleNumber"]
    r.neighbor_number, r.gamma = p["neighborNumber"], p["gamma"]
    r.iterative_sml = int(p["iterativeSmoothingLength"])
    r.periodic = int(p["periodic"])
    for i, v in enumerate(p.get("rangeMax", [])):
        r.range_max[i] = v
    for i, v in enumerate(p.get("rangeMin", [])):
        r.range_min[i] = v
    r.use_gravity, r.G, r.theta = int(p["useGravity"]), p["G"], p["theta"]
    r.gsph_2nd_order = int(p["use2ndOrderGSPH"])
    return r


def build(target="all", verbose=False):
    """Compile the checkers (make -C oracle [all|ref|port]).  Building the checker is not using it."""
    r = subprocess.run(["make", "-C", _HERE, "-j8", target], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("oracle build failed")


def lib_path(dim, flavour):
    if flavour == "port":
        return os.path.join(_HERE, "libspho.so")
    if flavour == "gpumod":
        # the reference driver with the sph::gpu Module drop-ins in the module slots
        # (sphcode_b200/host/Makefile): integration test of the plugin boundary, needs a GPU
        return os.path.join(_HERE, "..", "sphcode_b200", "host", "_build", f"libsphgpu_d{dim}.so")
    suffix = {"tree": "", "exhaustive": "_ex"}[flavour]
    return os.path.join(_HERE, "_ref", f"libsphref_d{dim}{suffix}.so")


def available(dim, flavour):
    return os.path.exists(lib_path(dim, flavour))


_libs = {}


def _load(dim, flavour):
    key = (dim, flavour)
    if key in _libs:
        return _libs[key]
    path = lib_path(dim, flavour)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference for the _ref flavours)")
    L = C.CDLL(path)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    pre = "ref_" if flavour != "port" else "spho_"
    def f(name, res, *args):
        fn = getattr(L, pre + name)
        fn.restype, fn.argtypes = res, list(args)
        return fn
    api = dict(
        dim=f("dim", ci), sizeof_particle=f("sizeof_particle", ci),
        set_threads=f("set_threads", None, ci), max_threads=f("max_threads", ci),
        create=f("create", vp, C.POINTER(RefParams), ci, vp) if flavour != "port" else f("create", vp, C.POINTER(RefParams), ci, ci, vp),
        destroy=f("destroy", None, vp), error=f("error", C.c_char_p, vp),
        get_particles=f("get_particles", None, vp, vp), set_particles=f("set_particles", None, vp, vp),
        init_state=f("init_state", None, vp), make_tree=f("make_tree", ci, vp),
        pre=f("pre", ci, vp), fluid=f("fluid", ci, vp), gravity=f("gravity", ci, vp), timestep=f("timestep", ci, vp),
        get_dt=f("get_dt", cd, vp), set_dt=f("set_dt", None, vp, cd),
        get_h_per_v_sig=f("get_h_per_v_sig", cd, vp), set_h_per_v_sig=f("set_h_per_v_sig", None, vp, cd),
        predict=f("predict", None, vp), correct=f("correct", None, vp),
        initialize=f("initialize", ci, vp), integrate=f("integrate", ci, vp),
        energy=f("energy", None, vp, vp),
        neighbor_search_all=f("neighbor_search_all", C.c_longlong, vp, vp, ci, vp, vp, C.c_longlong),
        get_vector_array=f("get_vector_array", ci, vp, C.c_char_p, vp),
        kernel_eval=f("kernel_eval", None, vp, vp, cd, vp),
    )
    if flavour == "port":
        api["counters"] = f("counters", ci, vp, vp, ci)
    else:
        api["set_active"] = f("set_active", None, vp, ci)
        api["set_kernel"] = f("set_kernel", ci, vp)
        api["direct_gravity"] = f("direct_gravity", None, vp, ci, vp)
    _libs[key] = (L, api)
    return _libs[key]


class RefSim:
    """One reference Simulation + its four modules (src/solver.cpp:353-414)."""

    def __init__(self, params, particles, dim, flavour="tree", threads=None):
        self.dim, self.flavour = dim, flavour
        self._L, self._f = _load(dim, flavour)
        f = self._f
        if flavour != "port":
            assert f["dim"]() == dim
        self.dtype = particle_dtype(dim)
        if flavour != "port":
            assert f["sizeof_particle"]() == self.dtype.itemsize
        if threads:
            f["set_threads"](threads)
        self.threads = f["max_threads"]()
        self.n = len(particles)
        p = np.ascontiguousarray(particles, dtype=self.dtype)
        rp = to_ref_params(params)
        if flavour == "port":
            self._c = f["create"](C.byref(rp), dim, self.n, p.ctypes.data)
        else:
            self._c = f["create"](C.byref(rp), self.n, p.ctypes.data)
        self._check(0)

    def _check(self, rc):
        msg = self._f["error"](self._c)
        if rc or msg:
            raise RuntimeError(f"reference error: {msg.decode() if msg else rc}")

    def close(self):
        if getattr(self, "_c", None):
            self._f["destroy"](self._c)
            self._c = None

    __del__ = close

    @property
    def particles(self):
        out = np.empty(self.n, dtype=self.dtype)
        self._f["get_particles"](self._c, out.ctypes.data)
        return out

    @particles.setter
    def particles(self, p):
        p = np.ascontiguousarray(p, dtype=self.dtype)
        assert len(p) == self.n
        self._f["set_particles"](self._c, p.ctypes.data)

    def set_active(self, k):
        """Subsample mode: the unmodified modules compute particles 0..k-1 against ALL particles (0 = all again)."""
        self._f["set_active"](self._c, int(k))

    def set_kernel(self): self._check(self._f["set_kernel"](self._c))

    def direct_gravity(self, k):
        """Direct sum of src/gravity_force.cpp:70-84 for targets 0..k-1 over all sources: (force[k, dim], phi[k])."""
        out = np.zeros((k, self.dim + 1))
        self._f["direct_gravity"](self._c, int(k), out.ctypes.data)
        return out[:, :self.dim].copy(), out[:, self.dim].copy()

    def init_state(self): self._f["init_state"](self._c)
    def make_tree(self): self._check(self._f["make_tree"](self._c))
    def pre(self): self._check(self._f["pre"](self._c))
    def fluid(self): self._check(self._f["fluid"](self._c))
    def gravity(self): self._check(self._f["gravity"](self._c))
    def predict(self): self._f["predict"](self._c)
    def correct(self): self._f["correct"](self._c)
    def initialize(self): self._check(self._f["initialize"](self._c))
    def integrate(self):
        self._check(self._f["integrate"](self._c))
        return self.dt

    def timestep(self):
        self._check(self._f["timestep"](self._c))
        return self.dt

    @property
    def dt(self): return self._f["get_dt"](self._c)
    @dt.setter
    def dt(self, v): self._f["set_dt"](self._c, float(v))
    @property
    def h_per_v_sig(self): return self._f["get_h_per_v_sig"](self._c)
    @h_per_v_sig.setter
    def h_per_v_sig(self, v): self._f["set_h_per_v_sig"](self._c, float(v))

    COUNTER_NAMES = ("newton_evals", "newton_iters", "pre_candidates", "pre_neighbors", "force_pairs",
                     "grav_pp", "grav_pc", "grav_node_visits")

    def counters(self, reset=True):
        """Interaction counts of the reference ALGORITHM since the last reset (port flavour only): the numerators
        of the algorithmic-FLOP model, same names as sphb_counters."""
        if "counters" not in self._f:
            raise RuntimeError("interaction counters exist in the port flavour only")
        out = np.zeros(8, dtype=np.uint64)
        self._f["counters"](self._c, out.ctypes.data, int(reset))
        return dict(zip(self.COUNTER_NAMES, (int(v) for v in out)))

    def energy(self):
        out = np.zeros(3)
        self._f["energy"](self._c, out.ctypes.data)
        return out

    def neighbor_lists(self, h=None, symmetric=False, cap_total=None, exhaustive=False):
        """CSR (offsets, ids) of every particle's list; each list sorted by id (the reference
        order — by r2, unstable among ties — is not canonical)."""
        n = self.n
        cap_total = cap_total or max(64 * n, 1 << 16)
        offsets = np.zeros(n + 1, dtype=np.int64)
        ids = np.empty(cap_total, dtype=np.int32)
        hp = None
        if h is not None:
            h = np.ascontiguousarray(h, dtype=np.float64)
            hp = h.ctypes.data
        flags = int(symmetric) | (2 if (exhaustive and self.flavour == "port") else 0)   # port: bit 1 = brute force
        tot = self._f["neighbor_search_all"](self._c, hp, flags, offsets.ctypes.data, ids.ctypes.data, cap_total)
        if tot > cap_total:
            return self.neighbor_lists(h, symmetric, cap_total=int(tot), exhaustive=exhaustive)
        ids = ids[:tot]
        rows = np.repeat(np.arange(n), np.diff(offsets))
        ids = ids[np.lexsort((ids, rows))]
        return offsets, ids

    def vector_array(self, name):
        out = np.zeros((self.n, self.dim))
        self._check(self._f["get_vector_array"](self._c, name.encode(), out.ctypes.data))
        return out

    def kernel_eval(self, rij, h):
        rij = np.ascontiguousarray(rij, dtype=np.float64)
        out = np.zeros(2 + self.dim)
        self._f["kernel_eval"](self._c, rij.ctypes.data, float(h), out.ctypes.data)
        return out[0], out[1], out[2:]


# ---- the whole unmodified reference Solver (oracle/ref_solver_driver.cpp, built by `make -C oracle stock`) ---------------
class RefSolverParams(C.Structure):
    _fields_ = [
        ("t_start", C.c_double), ("t_end", C.c_double), ("t_output", C.c_double), ("t_energy", C.c_double),
        ("sph_type", C.c_int), ("kernel", C.c_int),
        ("cfl_sound", C.c_double), ("cfl_force", C.c_double), ("av_alpha", C.c_double),
        ("use_balsara", C.c_int), ("use_tdav", C.c_int),
        ("alpha_max", C.c_double), ("alpha_min", C.c_double), ("epsilon_av", C.c_double),
        ("use_ac", C.c_int), ("alpha_ac", C.c_double),
        ("max_tree_level", C.c_int), ("leaf_particle_num", C.c_int), ("neighbor_number", C.c_int),
        ("gamma", C.c_double),
        ("iterative_sml", C.c_int), ("periodic", C.c_int),
        ("range_max", C.c_double * 3), ("range_min", C.c_double * 3),
        ("use_gravity", C.c_int), ("G", C.c_double), ("theta", C.c_double),
        ("gsph_2nd_order", C.c_int), ("n_side", C.c_int), ("sample", C.c_int),
        ("output_dir", C.c_char * 512),
    ]


def solver_lib_path(dim):
    return os.path.join(_HERE, "_ref", f"libsphsolver_d{dim}.so")


def stock_binary_path(dim):
    """oracle/_ref/sph_d<dim>: the stock `./sph <sample> <threads>` of the reference (CPU baseline of BASELINE.md)."""
    return os.path.join(_HERE, "_ref", f"sph_d{dim}")


class RefSolver:
    """sph::Solver of the unmodified reference: Solver::read_parameterfile on `arg` (a sample name or a parameter
    file, resolved against the CURRENT directory like the reference does) and, on demand, its sample generator."""

    def __init__(self, arg, dim):
        path = solver_lib_path(dim)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle stock` (needs /root/reference)")
        L = C.CDLL(path)
        L.refsolver_create.restype, L.refsolver_create.argtypes = C.c_void_p, [C.c_char_p]
        L.refsolver_destroy.argtypes = [C.c_void_p]
        L.refsolver_error.restype, L.refsolver_error.argtypes = C.c_char_p, [C.c_void_p]
        L.refsolver_get_params.argtypes = [C.c_void_p, C.POINTER(RefSolverParams)]
        L.refsolver_make_ic.argtypes = [C.c_void_p]
        L.refsolver_get_particles.argtypes = [C.c_void_p, C.c_void_p]
        assert L.refsolver_dim() == dim
        self._L, self.dim = L, dim
        self._s = L.refsolver_create(str(arg).encode())
        err = L.refsolver_error(self._s)
        if err:
            msg = err.decode()
            self.close()
            raise RuntimeError(msg)

    def close(self):
        if getattr(self, "_s", None):
            self._L.refsolver_destroy(self._s)
            self._s = None

    __del__ = close

    @property
    def params(self):
        p = RefSolverParams()
        assert self._L.refsolver_get_params(self._s, C.byref(p)) == 0
        return p

    def initial_condition(self):
        n = self._L.refsolver_make_ic(self._s)
        if n < 0:
            raise RuntimeError(self._L.refsolver_error(self._s).decode())
        out = np.empty(n, dtype=particle_dtype(self.dim))
        self._L.refsolver_get_particles(self._s, out.ctypes.data)
        return out
