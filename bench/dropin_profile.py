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
