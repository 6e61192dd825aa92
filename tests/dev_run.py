"""Developer helper (not a test): run one named case on the GPU against the golden / live reference
and print the per-field errors.  usage: python tests/dev_run.py <case> [steps]"""
import sys
import numpy as np
import parity_util as U
from sphcode_b200.lib import Context

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p, parts = U.make_case(name)
c = Context(p, p["DIM"])
c.upload(parts)
c.init_state(); print("init_state ok", flush=True)
c.make_tree(); c.synchronize(); print("tree ok", flush=True)
c.pre(); c.synchronize(); print("pre ok", flush=True)
c.fluid(); c.synchronize(); print("fluid ok", flush=True)
c.gravity(); c.synchronize(); print("gravity ok", flush=True)
for s in range(steps):
    print("dt", c.integrate(), flush=True)
try:
    from oracle.refsim import RefSim
    ref = RefSim(p, parts, p["DIM"], "tree")
    ref.initialize()
    for s in range(steps):
        ref.integrate()
    print(U.field_errors(c.particles, ref.particles))
except Exception as e:  # noqa
    print("no live reference:", e)
