"""timing probe: whole Aurora prover through the drop-in, 2 warm-ups + 3 runs (PROBE_NOQUIET=1 keeps the reference's stdout)"""
import os, sys, json, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("PROBE_NUMPY", "1") == "1":
    import bench
    quiet = bench._quiet_stdout if os.environ.get("PROBE_NOQUIET") != "1" else contextlib.nullcontext
else:
    quiet = contextlib.nullcontext
from libiop_b200 import prover
with quiet():
    p = prover.DropinProver()
    t = [p.aurora_gf192(20, verify=False)[0] for _ in range(5)]
sys.stderr.write("PROBE " + json.dumps(t[2:]) + "\n")
