"""Generates tests/golden/prover_fingerprints_extra.json: sha256 of the canonical transcript bytes of the
UNMODIFIED reference provers run on the CPU (oracle/_ref/libiopref.so, `make -C oracle ref`) at more sizes / variants than
hotpath_golden.npz holds.  Run in the build container: python tests/golden/make_prover_fingerprints.py"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle_lib import Reference

CASES = [("aurora_gf192", 14, False), ("aurora_gf192", 16, False), ("aurora_gf192", 12, True), ("aurora_gf192", 14, True),
         ("aurora_edwards", 14, False), ("fractal_edwards", 14, False), ("ligero_gf192", 16, False), ("ligero_gf192", 12, True)]


def main():
    r = Reference()
    out = []
    for which, log_n, zk in CASES:
        if which == "aurora_gf192":
            _, dump, acc = r.aurora_gf192(log_n, make_zk=zk)
        elif which == "aurora_edwards":
            _, dump, acc = r.aurora_edwards(log_n)
        elif which == "fractal_edwards":
            _, dump, acc, _ = r.fractal_edwards(log_n)
        else:
            _, dump, acc = r.ligero_gf192(log_n, make_zk=zk)
        assert acc == 1
        out.append({"prover": which, "log_n": log_n, "zk": zk, "bytes": len(dump), "sha256": hashlib.sha256(dump).hexdigest()})
        print(out[-1], flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prover_fingerprints_extra.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
