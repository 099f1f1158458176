"""Whole provers through the drop-in at BASELINE sizes: prover seconds of several consecutive proofs, bytes over PCIe per
proof, device-resident-oracle statistics, and the transcript fingerprint checked against the CPU reference's
(tests/golden/prover_fingerprints_config.json).  Usage (under gpurun): python bench/dropin_e2e.py [aurora:20 fractal:18 ligero:18 ...]"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("B200_DETERMINISTIC_RNG", "1")
from libiop_b200.prover import DropinProver

golden = {(c["prover"], c["log_n"], c["zk"]): c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "prover_fingerprints_config.json")))}
for c in json.load(open(os.path.join(ROOT, "tests", "golden", "prover_fingerprints_extra.json"))):
    golden.setdefault((c["prover"], c["log_n"], c["zk"]), c)
specs = sys.argv[1:] or ["aurora:20", "fractal:18", "ligero:18"]
reps = int(os.environ.get("REPS", "4"))
p = DropinProver()
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
for spec in specs:
    name, log_n = spec.split(":")[0], int(spec.split(":")[1])
    zk = name.endswith("_zk")
    key = {"aurora": "aurora_gf192", "aurora_zk": "aurora_gf192", "fractal": "fractal_edwards", "ligero": "ligero_gf192", "aurora_edwards": "aurora_edwards"}[name]
    times, idx_times = [], []
    sys.stdout.flush(); os.dup2(devnull, 1)
    try:
        for r in range(reps):
            d0, l0 = p.device_stats(), p.lazy_stats()
            if key == "aurora_gf192": secs, dump, acc = p.aurora_gf192(log_n, make_zk=zk, verify=False)
            elif key == "aurora_edwards": secs, dump, acc = p.aurora_edwards(log_n, verify=False)
            elif key == "fractal_edwards": secs, dump, acc, isecs = p.fractal_edwards(log_n, verify=False); idx_times.append(round(isecs, 4))
            else: secs, dump, acc = p.ligero_gf192(log_n, make_zk=zk, verify=False)
            d1, l1 = p.device_stats(), p.lazy_stats()
            times.append(round(secs, 4))
    finally:
        sys.stdout.flush(); os.dup2(saved, 1)
    g = golden.get((key, log_n, zk))
    rec = {"prover": key, "log_n": log_n, "zk": zk, "prover_s": times, "indexer_s": idx_times or None,
           "h2d_bytes_per_proof": d1["h2d_bytes"] - d0["h2d_bytes"], "d2h_bytes_per_proof": d1["d2h_bytes"] - d0["d2h_bytes"],
           "launches_per_proof": d1["launches"] - d0["launches"],
           "host_touched_vectors": l1["faults"] - l0["faults"], "host_touched_bytes": l1["fault_bytes"] - l0["fault_bytes"],
           "query_gathers": l1["gathers"] - l0["gathers"], "transcript_bytes": len(dump), "sha256": hashlib.sha256(dump).hexdigest(),
           "equals_cpu_reference": (g["sha256"] == hashlib.sha256(dump).hexdigest()) if g else None,
           "cpu_reference_prover_s_build_container": g.get("cpu_prover_s") if g else None,
           "arena": p.arena_stats()}
    print(json.dumps(rec), flush=True)
