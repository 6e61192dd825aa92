"""Developer helper (not a test): stage timers of one step of any sample.
usage: python tests/dev_time.py <sample> <N> [key=value ...]"""
import json
import sys
import parity_util  # noqa: F401  (sys.path)
from sphcode_b200 import sample_params, make_sample
from sphcode_b200.lib import Context

sample, n_side = sys.argv[1], int(sys.argv[2])
over = {}
for kv in sys.argv[3:]:
    k, v = kv.split("=", 1)
    over[k] = json.loads(v) if v[:1] in "0123456789-[tf" else v
p = sample_params(sample, N=n_side, **over)
parts = make_sample(p)
c = Context(p, p["DIM"])
c.upload(parts)
c.initialize()
c.integrate()
c.enable_timers(True)
c.integrate()
t = c.timers()
tot = sum(t.values())
print(sample, over, "n =", len(parts), {k: round(v, 3) for k, v in t.items()}, "sum ms", round(tot, 3),
      "particle-steps/s %.3g" % (len(parts) / (tot * 1e-3)))
