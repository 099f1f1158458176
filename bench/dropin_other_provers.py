"""BASELINE configs 3 and 4 at full size through the drop-in: fractal_snark_indexer/prover over edwards_Fr (N = 2^18)
and ligero_snark_prover over GF(2^192) (N = 2^18), unmodified reference host code, CPU vs B200 hot path.
Usage (under gpurun): python bench/dropin_other_provers.py [fractal_log_n] [ligero_log_n]"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle_lib import Reference, ReferenceGpu

f_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
l_n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
cpu, gpu = Reference(), ReferenceGpu()
gpu.aurora_gf192(10, verify=False)
if l_n:
    gs, gd, gacc = gpu.ligero_gf192(l_n, verify=True)
    gs2, gd2, _ = gpu.ligero_gf192(l_n, verify=False)
    cs, cd, _ = cpu.ligero_gf192(l_n, verify=False)
    print(json.dumps({"prover": "ligero_gf192", "log_n": l_n, "cpu_prover_s": cs, "dropin_prover_s_first": gs, "dropin_prover_s": gs2,
                      "bit_exact": cd == gd == gd2, "transcript_bytes": len(cd), "sha256": hashlib.sha256(gd).hexdigest(),
                      "reference_verifier_accepts": gacc}), flush=True)
if f_n:
    gs, gd, gacc, gi = gpu.fractal_edwards(f_n, verify=True)
    gs2, gd2, _, gi2 = gpu.fractal_edwards(f_n, verify=False)
    cs, cd, _, ci = cpu.fractal_edwards(f_n, verify=False)
    print(json.dumps({"prover": "fractal_edwards", "log_n": f_n, "cpu_prover_s": cs, "cpu_indexer_s": ci, "dropin_prover_s_first": gs,
                      "dropin_prover_s": gs2, "dropin_indexer_s": gi2, "bit_exact": cd == gd == gd2, "transcript_bytes": len(cd),
                      "sha256": hashlib.sha256(gd).hexdigest(), "reference_verifier_accepts": gacc}), flush=True)
