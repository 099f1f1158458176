"""libff::enter_block profile of the reference prover with the B200 drop-in (B200_REF_PROFILE=1)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle_lib import Reference, ReferenceGpu
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
which = sys.argv[2] if len(sys.argv) > 2 else "gpu"
r = ReferenceGpu() if which == "gpu" else Reference()
if which == "gpu":
    os.environ.pop("B200_REF_PROFILE", None)
    r.aurora_gf192(log_n, verify=False)
    os.environ["B200_REF_PROFILE"] = "1"
r.aurora_gf192(log_n, verify=False)
if which == "gpu":
    import ctypes as C
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    r._real.refgpu_arena_stats(C.byref(a), C.byref(b), C.byref(c))
    print(f"pinned arena: {a.value} slabs, {b.value / 2**30:.2f} GiB reserved, {c.value / 2**30:.2f} GiB free", flush=True)
