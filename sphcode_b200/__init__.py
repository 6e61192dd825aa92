"""sphcode_b200 — B200 (sm_100a) implementation of sphcode's per-step particle hot path.

The product is the CUDA library ``sphcode_b200/libsphb.so`` (C ABI in include/sphb.h) and the C++
host in ``sphcode_b200/host`` that keeps the reference's Module plugin surface.  This Python
package is the thin ctypes harness used by tests/ and bench.py.
"""
from .params import sample_params, resolve, SHIPPED, SAMPLES, SPHParameterError  # noqa: F401
from .samples import make_sample, particle_dtype  # noqa: F401
