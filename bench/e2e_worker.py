"""One prover process of the multi-process throughput measurement (bench.py --concurrency / bench/e2e_throughput.py):
warm up, report READY, wait for GO on stdin, run R whole-prover calls, report the per-proof seconds.
Usage: python bench/e2e_worker.py <impl: dropin|cpu> <log_n> <runs> <seed>"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
impl, log_n, runs, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
out = os.fdopen(os.dup(1), "w")                      # keep a private handle on stdout, then silence the reference's printf()s
devnull = os.open(os.devnull, os.O_WRONLY); os.dup2(devnull, 1)
if impl == "dropin":
    from libiop_b200.prover import DropinProver
    p = DropinProver()
    p.aurora_gf192_repeat(log_n, 2, seed=seed)
else:
    from oracle_lib import Reference
    p = Reference()
    p.aurora_gf192(max(10, log_n - 4), seed=seed, verify=False)
out.write("READY\n"); out.flush()
sys.stdin.readline()
t0 = time.time()
secs = p.aurora_gf192_repeat(log_n, runs + 1, seed=seed)[1:]      # the first repeat absorbs the skew between the workers' instance generation
out.write(json.dumps({"prover_s": secs, "wall_s": time.time() - t0}) + "\n"); out.flush()
