"""ctypes binding of libsphb.so (C ABI: include/sphb.h) — the harness tests/ and bench.py use.

There is no CPU fallback: loading fails loudly when the CUDA library is missing, and
``Context`` creation fails when no CUDA device is present.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .samples import particle_dtype

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPHB_LIB") or os.path.join(_HERE, "libsphb.so")    # SPHB_LIB: developer A/B builds
CSRC = os.path.join(_HERE, "csrc")

F_POS, F_VEL, F_VEL_P, F_ACC = 1 << 0, 1 << 1, 1 << 2, 1 << 3
F_MASS, F_DENS, F_PRES, F_ENE, F_ENE_P, F_DENE = 1 << 4, 1 << 5, 1 << 6, 1 << 7, 1 << 8, 1 << 9
F_SML, F_SOUND, F_BALSARA, F_ALPHA, F_GRADH, F_PHI = 1 << 10, 1 << 11, 1 << 12, 1 << 13, 1 << 14, 1 << 15
F_ID, F_NEIGHBOR = 1 << 16, 1 << 17
F_ALL = 0x3FFFF
T_NAMES = ("tree", "pre", "fluid", "gravity", "timestep", "predict", "correct", "exchange", "migrate", "keys", "reduce", "halo")


class SphbParams(C.Structure):
    """sphb_params (include/sphb.h) = sph::SPHParameters (reference include/parameters.hpp:20-79)."""
    _fields_ = [
        ("sph_type", C.c_int32), ("kernel", C.c_int32),
        ("cfl_sound", C.c_double), ("cfl_force", C.c_double),
        ("av_alpha", C.c_double),
        ("use_balsara_switch", C.c_int32), ("use_time_dependent_av", C.c_int32),
        ("alpha_max", C.c_double), ("alpha_min", C.c_double), ("epsilon_av", C.c_double),
        ("use_ac", C.c_int32), ("_pad0", C.c_int32),
        ("alpha_ac", C.c_double),
        ("max_tree_level", C.c_int32), ("leaf_particle_num", C.c_int32),
        ("neighbor_number", C.c_int32), ("iterative_sml", C.c_int32),
        ("gamma", C.c_double),
        ("periodic", C.c_int32), ("use_gravity", C.c_int32),
        ("range_max", C.c_double * 3), ("range_min", C.c_double * 3),
        ("G", C.c_double), ("theta", C.c_double),
        ("gsph_2nd_order", C.c_int32), ("_pad1", C.c_int32),
    ]


class SphbCounters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in (
        "n_particles", "newton_evals", "newton_iters", "pre_candidates", "pre_neighbors", "force_pairs",
        "grav_pp", "grav_pc", "grav_node_visits", "tree_nodes", "tree_leaves",
        "grav_pc_group", "grav_pp_group", "n_groups")]


_SPH = {"ssph": 0, "disph": 1, "gsph": 2}
_KER = {"cubic_spline": 0, "wendland": 1}


def to_sphb_params(p):
    """p: resolved dict of the reference's JSON keys (sphcode_b200.params.resolve)."""
    r = SphbParams()
    r.sph_type, r.kernel = _SPH[p["SPHType"]], _KER[p["kernel"]]
    r.cfl_sound, r.cfl_force = p["cflSound"], p["cflForce"]
    r.av_alpha = p["avAlpha"]
    r.use_balsara_switch, r.use_time_dependent_av = int(p["useBalsaraSwitch"]), int(p["useTimeDependentAV"])
    r.alpha_max, r.alpha_min, r.epsilon_av = p["alphaMax"], p["alphaMin"], p["epsilonAV"]
    r.use_ac, r.alpha_ac = int(p["useArtificialConductivity"]), p["alphaAC"]
    r.max_tree_level, r.leaf_particle_num = p["maxTreeLevel"], p["leafParticleNumber"]
    r.neighbor_number, r.iterative_sml = p["neighborNumber"], int(p["iterativeSmoothingLength"])
    r.gamma = p["gamma"]
    r.periodic, r.use_gravity = int(p["periodic"]), int(p["useGravity"])
    for i, v in enumerate(p.get("rangeMax", [])):
        r.range_max[i] = v
    for i, v in enumerate(p.get("rangeMin", [])):
        r.range_min[i] = v
    r.G, r.theta = p["G"], p["theta"]
    r.gsph_2nd_order = int(p["use2ndOrderGSPH"])
    return r


NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force=False, verbose=False):
    """Compile libsphb.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    hdr = os.path.join(_HERE, "..", "include", "sphb.h")
    if not force and os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(s) <= t for s in srcs + [hdr]):
            return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + [os.path.join(CSRC, "sphb_api.cu"), "-o", LIB_PATH, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(" ".join(cmd), r.stdout[-4000:], r.stderr[-4000:], sep="\n")
    if r.returncode:
        raise RuntimeError("nvcc failed building libsphb.so")
    return LIB_PATH


_lib = None

# name -> (restype, argtypes); every symbol include/sphb.h declares
_vp, _i, _d, _sz, _u32, _u64 = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_uint32, C.c_uint64
SYMBOLS = {
    "sphb_create": (_i, [C.POINTER(SphbParams), _i, _i, C.POINTER(_vp)]),
    "sphb_destroy": (None, [_vp]),
    "sphb_last_error": (C.c_char_p, [_vp]),
    "sphb_set_stream": (_i, [_vp, _vp]),
    "sphb_synchronize": (_i, [_vp]),
    "sphb_dim": (_i, [_vp]),
    "sphb_particle_num": (_i, [_vp]),
    "sphb_global_particle_num": (C.c_longlong, [_vp]),
    "sphb_first_global_index": (C.c_longlong, [_vp]),
    "sphb_halo_records": (_u64, [_vp]),
    "sphb_migrated": (_u64, [_vp]),
    "sphb_sizeof_particle": (_sz, [_i]),
    "sphb_nccl_unique_id": (_i, [_vp]),
    "sphb_set_distributed": (_i, [_vp, _i, _i, _vp]),
    "sphb_set_distributed_id": (_i, [_vp, _i, _i, _vp]),
    "sphb_upload_aos": (_i, [_vp, _vp, _i, _sz, _u32]),
    "sphb_download_aos": (_i, [_vp, _vp, _i, _sz, _u32]),
    "sphb_get_vector_array": (_i, [_vp, C.c_char_p, _vp]),
    "sphb_set_vector_array": (_i, [_vp, C.c_char_p, _vp]),
    "sphb_set_dt": (_i, [_vp, _d]),
    "sphb_get_dt": (_i, [_vp, C.POINTER(_d)]),
    "sphb_set_h_per_v_sig": (_i, [_vp, _d]),
    "sphb_get_h_per_v_sig": (_i, [_vp, C.POINTER(_d)]),
    "sphb_init_state": (_i, [_vp]),
    "sphb_make_tree": (_i, [_vp]),
    "sphb_pre_interaction": (_i, [_vp]),
    "sphb_fluid_force": (_i, [_vp]),
    "sphb_gravity_force": (_i, [_vp]),
    "sphb_gravity_direct": (_i, [_vp]),
    "sphb_gravity_direct_targets": (_i, [_vp, _i]),
    "sphb_timestep": (_i, [_vp, C.POINTER(_d)]),
    "sphb_predict": (_i, [_vp]),
    "sphb_correct": (_i, [_vp]),
    "sphb_initialize": (_i, [_vp]),
    "sphb_integrate": (_i, [_vp, C.POINTER(_d)]),
    "sphb_energy": (_i, [_vp, _vp]),
    "sphb_neighbor_lists": (_i, [_vp, _vp, _i, _vp, _vp, C.c_int64, C.POINTER(C.c_int64)]),
    "sphb_enable_counters": (_i, [_vp, _i]),
    "sphb_get_counters": (_i, [_vp, C.POINTER(SphbCounters)]),
    "sphb_enable_timers": (_i, [_vp, _i]),
    "sphb_get_timers": (_i, [_vp, _vp]),
    "sphb_launch_count": (_u64, [_vp]),
    "sphb_nonconverged": (_u64, [_vp]),
    "sphb_host_alloc": (_vp, [_sz]),
    "sphb_host_free": (None, [_vp]),
    "sphb_bench_fp64": (_i, [_i, C.POINTER(_d)]),
}


def load():
    """dlopen libsphb.so and bind every symbol of include/sphb.h; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


class SphbError(RuntimeError):
    pass


class Context:
    """One device context = Simulation + BHTree + the four Modules of the reference
    (src/solver.cpp:353-414) for one particle set."""

    def __init__(self, params, dim, device=0):
        self.L = load()
        self.dim = dim
        self.params = params
        self.dtype = particle_dtype(dim)
        assert self.L.sphb_sizeof_particle(dim) == self.dtype.itemsize
        self._c = _vp()
        sp = to_sphb_params(params)
        if self.L.sphb_create(C.byref(sp), dim, device, C.byref(self._c)):
            raise SphbError(self.L.sphb_last_error(None).decode())
        self.n = 0

    def _ck(self, rc):
        if rc:
            raise SphbError(self.L.sphb_last_error(self._c).decode())

    def close(self):
        if getattr(self, "_c", None) and self._c.value:
            self.L.sphb_destroy(self._c)
            self._c = _vp()

    __del__ = close

    # --- state
    def upload(self, particles, mask=F_ALL):
        p = np.ascontiguousarray(particles, dtype=self.dtype)
        self.n = len(p)
        self._ck(self.L.sphb_upload_aos(self._c, p.ctypes.data, len(p), self.dtype.itemsize, mask))

    def upload_raw(self, ptr, n, mask=F_ALL):
        self.n = n
        self._ck(self.L.sphb_upload_aos(self._c, ptr, n, self.dtype.itemsize, mask))

    def download(self, mask=F_ALL, out=None):
        self.n = self.local_n
        if out is None:
            out = np.zeros(self.n, dtype=self.dtype)
        self._ck(self.L.sphb_download_aos(self._c, out.ctypes.data, self.n, self.dtype.itemsize, mask))
        return out

    def download_raw(self, ptr, mask=F_ALL):
        self.n = self.local_n
        self._ck(self.L.sphb_download_aos(self._c, ptr, self.n, self.dtype.itemsize, mask))

    @property
    def particles(self):
        return self.download()

    def vector_array(self, name):
        out = np.zeros((self.n, self.dim))
        self._ck(self.L.sphb_get_vector_array(self._c, name.encode(), out.ctypes.data))
        return out

    @property
    def dt(self):
        v = _d()
        self._ck(self.L.sphb_get_dt(self._c, C.byref(v)))
        return v.value

    @dt.setter
    def dt(self, v):
        self._ck(self.L.sphb_set_dt(self._c, float(v)))

    @property
    def h_per_v_sig(self):
        v = _d()
        self._ck(self.L.sphb_get_h_per_v_sig(self._c, C.byref(v)))
        return v.value

    @h_per_v_sig.setter
    def h_per_v_sig(self, v):
        self._ck(self.L.sphb_set_h_per_v_sig(self._c, float(v)))

    # --- stages
    def init_state(self): self._ck(self.L.sphb_init_state(self._c))
    def make_tree(self): self._ck(self.L.sphb_make_tree(self._c))
    def pre(self): self._ck(self.L.sphb_pre_interaction(self._c))
    def fluid(self): self._ck(self.L.sphb_fluid_force(self._c))
    def gravity(self): self._ck(self.L.sphb_gravity_force(self._c))
    def gravity_direct(self, targets=None):
        if targets:
            self._ck(self.L.sphb_gravity_direct_targets(self._c, int(targets)))
        else:
            self._ck(self.L.sphb_gravity_direct(self._c))
    def predict(self): self._ck(self.L.sphb_predict(self._c))
    def correct(self): self._ck(self.L.sphb_correct(self._c))
    def initialize(self): self._ck(self.L.sphb_initialize(self._c))
    def synchronize(self): self._ck(self.L.sphb_synchronize(self._c))

    def timestep(self):
        v = _d()
        self._ck(self.L.sphb_timestep(self._c, C.byref(v)))
        return v.value

    def integrate(self):
        v = _d()
        self._ck(self.L.sphb_integrate(self._c, C.byref(v)))
        return v.value

    def energy(self):
        out = np.zeros(3)
        self._ck(self.L.sphb_energy(self._c, out.ctypes.data))
        return out

    # --- hooks
    def neighbor_lists(self, h=None, symmetric=False):
        n = self.n
        offsets = np.zeros(n + 1, dtype=np.int64)
        tot = C.c_int64()
        hp = None
        if h is not None:
            h = np.ascontiguousarray(h, dtype=np.float64)
            hp = h.ctypes.data
        self._ck(self.L.sphb_neighbor_lists(self._c, hp, int(symmetric), offsets.ctypes.data, None, 0, C.byref(tot)))
        ids = np.empty(max(tot.value, 1), dtype=np.int32)
        self._ck(self.L.sphb_neighbor_lists(self._c, hp, int(symmetric), offsets.ctypes.data, ids.ctypes.data, tot.value, C.byref(tot)))
        return offsets, ids[:tot.value]

    def enable_counters(self, on=True): self._ck(self.L.sphb_enable_counters(self._c, int(on)))

    def counters(self):
        c = SphbCounters()
        self._ck(self.L.sphb_get_counters(self._c, C.byref(c)))
        return {k: getattr(c, k) for k, _ in SphbCounters._fields_}

    def enable_timers(self, on=True): self._ck(self.L.sphb_enable_timers(self._c, int(on)))

    def timers(self):
        ms = (C.c_float * len(T_NAMES))()
        self._ck(self.L.sphb_get_timers(self._c, ms))
        return dict(zip(T_NAMES, list(ms)))

    @property
    def local_n(self):
        """particles this rank holds now (multi-GPU: changes with migration)"""
        return int(self.L.sphb_particle_num(self._c))

    @property
    def global_n(self): return int(self.L.sphb_global_particle_num(self._c))
    @property
    def halo_records(self): return int(self.L.sphb_halo_records(self._c))
    @property
    def migrated(self): return int(self.L.sphb_migrated(self._c))

    @property
    def launches(self): return int(self.L.sphb_launch_count(self._c))
    @property
    def nonconverged(self): return int(self.L.sphb_nonconverged(self._c))

    def set_distributed_id(self, rank, world, uid_bytes):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._ck(self.L.sphb_set_distributed_id(self._c, rank, world, buf))


def nccl_unique_id():
    L = load()
    buf = C.create_string_buffer(128)
    if L.sphb_nccl_unique_id(buf):
        raise SphbError(L.sphb_last_error(None).decode())
    return buf.raw


def fp64_peak_tflops(device=0):
    L = load()
    v = _d()
    if L.sphb_bench_fp64(device, C.byref(v)):
        raise SphbError("sphb_bench_fp64 failed")
    return v.value
