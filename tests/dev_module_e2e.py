"""Developer probe: module-mode transfers at 16 M (masked upload / downloads around the stage calls), ms per step."""
import ctypes, sys, time
import numpy as np
import parity_util  # noqa: F401
from sphcode_b200 import sample_params, make_sample, lib
F = lib
p = sample_params("evrard", N=int(sys.argv[1]) if len(sys.argv) > 1 else 312)
parts = make_sample(p); n = len(parts); rec = parts.dtype.itemsize
c = lib.Context(p, 3)
h = c.L.sphb_host_alloc(n * rec); ctypes.memmove(h, parts.ctypes.data, n * rec)
c.upload_raw(h, n); c.initialize(); c.download_raw(h)
ref = np.ctypeslib.as_array(ctypes.cast(h, ctypes.POINTER(ctypes.c_ubyte)), shape=(n * rec,)).view(parts.dtype).copy()
up = F.F_POS | F.F_VEL | F.F_ENE | F.F_SOUND
pre = F.F_SML | F.F_DENS | F.F_PRES | F.F_GRADH | F.F_BALSARA | F.F_ALPHA | F.F_NEIGHBOR
def step():
    t = [time.perf_counter()]
    c.timestep(); c.upload_raw(h, n, up); t.append(time.perf_counter())
    c.make_tree(); c.pre(); c.synchronize(); t.append(time.perf_counter())
    c.download_raw(h, pre); t.append(time.perf_counter())
    c.fluid(); c.synchronize(); t.append(time.perf_counter())
    c.download_raw(h, F.F_ACC | F.F_DENE); t.append(time.perf_counter())
    c.gravity(); c.synchronize(); t.append(time.perf_counter())
    c.download_raw(h, F.F_ACC | F.F_PHI); t.append(time.perf_counter())
    return [round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])]
step()
print("n", n, "ms: upload, tree+pre, down(pre), fluid, down(acc,dene), gravity, down(acc,phi):", step())
got = np.ctypeslib.as_array(ctypes.cast(h, ctypes.POINTER(ctypes.c_ubyte)), shape=(n * rec,)).view(parts.dtype)
full = c.particles
for f in ("sml", "dens", "acc", "phi", "neighbor", "pos", "mass"):
    assert np.array_equal(got[f], full[f]), f
print("masked transfers consistent with a full download")
