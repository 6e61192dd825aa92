"""Developer helper (not a test): build an A/B variant of libsphb.so with extra nvcc flags.
usage: python tests/dev_variant.py <name> [-DMACRO=VALUE ...]  ->  _build/libsphb_<name>.so
run it with  SPHB_LIB=_build/libsphb_<name>.so python tests/dev_counters.py 312"""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sphcode_b200 import lib  # noqa: E402
os.makedirs(os.path.join(ROOT, "_build"), exist_ok=True)
out = os.path.join(ROOT, "_build", f"libsphb_{sys.argv[1]}.so")
cmd = ["nvcc"] + lib.NVCC_FLAGS + sys.argv[2:] + [os.path.join(lib.CSRC, "sphb_api.cu"), "-o", out, "-ldl"]
r = subprocess.run(cmd, capture_output=True, text=True)
print(r.stdout[-3000:], r.stderr[-3000:])
sys.exit(r.returncode)
