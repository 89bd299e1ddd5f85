"""bisemutum-engine_b200 — B200-native wavefront path tracer behind bisemutum-engine's
`PathTracingPass` interface (reference: bisemutum/src/renderer/pass/path_tracing.{hpp,cpp}).

Layout
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/bpt/bpt.h) → libbpt.so
  host/      C++ mirror of the reference's host-side interfaces for this path → libbpt_host.so
  capi.py    ctypes view of the C ABI
  engine.py  Python mirror of Camera / LightsContext / PathTracingPass on top of the two libraries
  scenes.py  seeded procedural scene generators (synthetic input)

The product has no CPU implementation: anything that renders requires libbpt.so and a CUDA device.
"""
import os as _os

PACKAGE_DIR = _os.path.dirname(_os.path.abspath(__file__))
REPO_ROOT = _os.path.dirname(PACKAGE_DIR)
LIBBPT_PATH = _os.path.join(PACKAGE_DIR, "csrc", "libbpt.so")
LIBBPT_HOST_PATH = _os.path.join(PACKAGE_DIR, "host", "libbpt_host.so")

from . import capi, scenes  # noqa: E402,F401


def load_library():
    """Loads libbpt.so (the CUDA implementation). Raises if it has not been built."""
    return capi.Library(_os.environ.get("BPT_LIB") or LIBBPT_PATH, "bpt_", capi.BPT_ONLY_API)      # BPT_LIB: an experimental variant (tools/build_variant.sh)
