"""Developer helper (not a test): interaction counters of one Evrard step.  usage: python tests/dev_counters.py [n_side]"""
import sys
import parity_util  # noqa: F401  (sys.path)
from sphcode_b200 import sample_params, make_sample
from sphcode_b200.lib import Context

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 124
p = sample_params("evrard", N=n_side)
parts = make_sample(p)
c = Context(p, 3)
c.upload(parts)
c.initialize()
c.enable_counters(True)
c.integrate()
k = c.counters()
n = k["n_particles"]
for key, v in k.items():
    print(f"{key:18s} {v:14d}  per particle {v / n:10.3f}")
print("particles per group", n / max(k["n_groups"], 1))
c.enable_timers(True)
c.enable_counters(False)
c.integrate()
print(c.timers())
