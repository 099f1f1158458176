"""Generates tests/golden/prover_fingerprints_config.json: sha256 of the canonical transcript bytes of the
UNMODIFIED reference provers run on the CPU (oracle/_ref/libiopref.so, `make -C oracle ref`) at the FULL sizes of
BASELINE.json's configs 2, 3, 4 and 5 (and the zero-knowledge variant of config 2's codeword size), together with the
CPU reference's own prover seconds.  The `-m gpu` tests prove the same instances through the drop-in and compare.

Run in the build container (about 10 minutes of one core, 25 GB of memory for the 2^22 case):
    python tests/golden/make_config_fingerprints.py [case ...]
"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle_lib import Reference

# (label, prover, log_n, zk)  — BASELINE.json configs[1..4]; seeds: the default seed (+ rank for config 5)
CASES = [("config4", "ligero_gf192", 18, False),
         ("config2", "aurora_gf192", 20, False),
         ("config3", "fractal_edwards", 18, False),
         ("config2_zk", "aurora_gf192", 18, True),
         ("config5_rank0", "aurora_gf192", 22, False)]
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prover_fingerprints_config.json")


def main():
    want = set(sys.argv[1:])
    r = Reference()
    out = json.load(open(PATH)) if os.path.exists(PATH) else []
    for label, which, log_n, zk in CASES:
        if want and label not in want:
            continue
        t0 = time.time()
        extra = {}
        if which == "aurora_gf192":
            secs, dump, acc = r.aurora_gf192(log_n, make_zk=zk, cap=1 << 27)
        elif which == "fractal_edwards":
            secs, dump, acc, idx = r.fractal_edwards(log_n)
            extra["indexer_s"] = round(idx, 3)
        else:
            secs, dump, acc = r.ligero_gf192(log_n, make_zk=zk)
        assert acc == 1, "the reference verifier rejected the reference's own proof"
        rec = {"config": label, "prover": which, "log_n": log_n, "zk": zk, "bytes": len(dump),
               "sha256": hashlib.sha256(dump).hexdigest(), "cpu_prover_s": round(secs, 3),
               "cpu_cores": 1, "wall_s": round(time.time() - t0, 1), **extra}
        out = [o for o in out if o["config"] != label] + [rec]
        print(rec, flush=True)
        json.dump(out, open(PATH, "w"), indent=1)
    print("wrote", PATH)


if __name__ == "__main__":
    main()
