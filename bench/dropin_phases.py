"""Host-side phases of the Aurora prover through the drop-in that the reference's enter_block timers do not cover:
construction + registrations, produce_proof, get_transcript, destruction.  Usage: python bench/dropin_phases.py [log_n] [repeats]
(B200_DROPIN_LIB=<path of the host-simulator build> runs the same on a machine without a GPU)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200_DETERMINISTIC_RNG", "1")
from libiop_b200.prover import DropinProver
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
p = DropinProver(os.environ.get("B200_DROPIN_LIB"))
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
sys.stdout.flush(); os.dup2(devnull, 1)
try:
    rows = p.aurora_gf192_phases(log_n, reps)
finally:
    sys.stdout.flush(); os.dup2(saved, 1)
for r in rows:
    print(json.dumps({"log_n": log_n, "setup_s": round(r[0], 4), "produce_proof_s": round(r[1], 4), "get_transcript_s": round(r[2], 4),
                      "teardown_s": round(r[3], 4), "argument_dtor_s": round(r[4], 4), "total_s": round(sum(r), 4)}), flush=True)
