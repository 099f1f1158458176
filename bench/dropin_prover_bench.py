"""Whole-prover check at BASELINE sizes: the unmodified reference aurora_snark_prover on the host CPU
(oracle/_ref/libiopref.so) versus the same host code with the B200 hot path dropped in
(libiop_b200/lib/libiop_b200_dropin.so).  Prints prover seconds, transcript fingerprints (must be equal) and
verifier verdicts.  Usage (under gpurun): python bench/dropin_prover_bench.py 16 18 20"""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle_lib import Reference, ReferenceGpu

sizes = [int(a) for a in sys.argv[1:]] or [16]
cpu, gpu = Reference(), ReferenceGpu()
gpu.aurora_gf192(10, verify=False)     # warm-up: context creation, kernel load
for log_n in sizes:
    t = time.time(); gs, gd, gacc = gpu.aurora_gf192(log_n, verify=(log_n <= 18)); gwall = time.time() - t
    gs2, gd2, _ = gpu.aurora_gf192(log_n, verify=False)    # second run: per-domain plans cached
    cs, cd, _ = cpu.aurora_gf192(log_n, verify=False)
    rec = {"log_n": log_n, "cpu_prover_s": cs, "gpu_dropin_prover_s_first": gs, "gpu_dropin_prover_s_cached_plans": gs2,
           "bit_exact": cd == gd == gd2, "transcript_bytes": len(cd), "sha256": hashlib.sha256(gd).hexdigest(),
           "reference_verifier_accepts_gpu_proof": gacc,
           "cpu_constraints_per_s": (1 << log_n) / cs, "dropin_constraints_per_s": (1 << log_n) / gs2}
    print(json.dumps(rec), flush=True)
