"""Whole-prover throughput with P concurrent prover processes (one proof each at a time, independent instances):
drop-in provers sharing one B200, or CPU reference provers sharing the host cores.
Usage: python bench/e2e_throughput.py <dropin|cpu> <log_n> <runs per process> <P> [<P> ...]"""
import json, os, subprocess, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))


def measure(impl, log_n, runs, procs):
    ws = [subprocess.Popen([sys.executable, os.path.join(HERE, "e2e_worker.py"), impl, str(log_n), str(runs), str(0x9E3779B97F4A7C15 + i)],
                           stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for i in range(procs)]
    for w in ws:
        assert w.stdout.readline().strip() == "READY"
    t0 = time.time()
    for w in ws:
        w.stdin.write("GO\n"); w.stdin.flush()
    res = [json.loads(w.stdout.readline()) for w in ws]
    wall = time.time() - t0
    for w in ws:
        w.wait()
    mean = sum(sum(r["prover_s"]) for r in res) / (procs * runs)
    # every process proves back to back on its own instance while the others do the same: throughput = P / mean proof time
    return {"impl": impl, "log_n": log_n, "processes": procs, "proofs": procs * runs, "mean_prover_s": mean,
            "constraints_per_s": procs * (1 << log_n) / mean, "wall_s_incl_instance_generation": wall}


if __name__ == "__main__":
    impl, log_n, runs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    for P in [int(a) for a in sys.argv[4:]]:
        print(json.dumps(measure(impl, log_n, runs, P)), flush=True)
