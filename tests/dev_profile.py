"""Developer helper (not a test): a short Evrard run for ncu.  usage: python tests/dev_profile.py [n_side] [steps]"""
import sys
import parity_util  # noqa: F401  (sys.path)
from sphcode_b200 import sample_params, make_sample
from sphcode_b200.lib import Context

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 124
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = sample_params("evrard", N=n_side)
parts = make_sample(p)
c = Context(p, 3)
c.upload(parts)
c.initialize()
for _ in range(steps):
    c.integrate()
c.synchronize()
print("done", len(parts), c.launches)
