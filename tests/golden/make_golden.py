"""Generates tests/golden/hotpath_golden.npz by running the UNMODIFIED reference
(oracle/_ref/libiopref.so, built by `make -C oracle ref` from /root/reference).
Only runnable in the build container; the resulting fixture is committed and is what the
GPU box tests against.  Usage: python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle_lib import Oracle, Reference, el, rand_elems, standard_basis  # noqa: E402


def main():
    r = Reference()
    o = Oracle()  # only used for random Fp sampling and domain constants
    rng = np.random.default_rng(20260925)
    g = {}
    # field ops
    g["gf_a"] = rand_elems(rng, 64); g["gf_b"] = rand_elems(rng, 64)
    g["gf_mul"] = r.gf192_mul(g["gf_a"], g["gf_b"]); g["gf_inv"] = r.gf192_inv(g["gf_a"][:8])
    g["fp_a"] = o.rand_fp(rng, 64); g["fp_b"] = o.rand_fp(rng, 64)
    g["fp_mul"] = r.fp_mul(g["fp_a"], g["fp_b"]); g["fp_add"] = r.fp_add(g["fp_a"], g["fp_b"])
    g["fp_sub"] = r.fp_sub(g["fp_a"], g["fp_b"]); g["fp_inv"] = r.fp_inv(g["fp_a"][:8])
    mod, r2, inv, s = r.fp_params()
    g["fp_mod"] = mod; g["fp_r2"] = r2; g["fp_inv_const"] = np.array([inv], dtype=np.uint64)
    g["fp_generator"] = r.fp_generator()
    for ln in (1, 4, 10, 23):
        g[f"fp_subgroup_gen_{ln}"] = r.fp_subgroup_generator(ln)
    # additive FFT / IFFT: standard basis with the Aurora shift, and a random affine subspace
    for tag, m, nc in (("std", 10, 1 << 7), ("std_full", 8, 1 << 8), ("rnd", 7, 100)):
        basis = standard_basis(m) if tag.startswith("std") else rand_elems(rng, m)
        shift = el(1 << m) if tag.startswith("std") else rand_elems(rng, 1)[0]
        c = rand_elems(rng, nc); e = rand_elems(rng, 1 << m)
        g[f"afft_{tag}_basis"] = basis; g[f"afft_{tag}_shift"] = shift
        g[f"afft_{tag}_coeffs"] = c; g[f"afft_{tag}_evals"] = r.additive_fft(c, m, basis, shift)
        g[f"aifft_{tag}_in"] = e; g[f"aifft_{tag}_out"] = r.additive_ifft(e, m, basis, shift)
    # multiplicative NTT / iNTT (coset shift = generator, and the plain subgroup)
    for tag, logn, nc, sh in (("coset", 9, 1 << 6, None), ("sub", 7, 100, 1)):
        shift = r.fp_generator() if sh is None else r.fp_from_u64(1)
        c = o.rand_fp(rng, nc); e = o.rand_fp(rng, 1 << logn)
        g[f"mfft_{tag}_shift"] = shift; g[f"mfft_{tag}_coeffs"] = c
        g[f"mfft_{tag}_evals"] = r.mult_fft(c, logn, shift)
        g[f"mifft_{tag}_in"] = e; g[f"mifft_{tag}_out"] = r.mult_ifft(e, logn, shift)
    # FRI folds (eta = 1, 2, 3), including x inside the domain
    for tag, m, eta in (("a1", 8, 1), ("a2", 8, 2), ("a3", 6, 3)):
        basis = rand_elems(rng, m); shift = rand_elems(rng, 1)[0]; x = rand_elems(rng, 1)[0]
        f = rand_elems(rng, 1 << m)
        g[f"fold_{tag}_basis"] = basis; g[f"fold_{tag}_shift"] = shift; g[f"fold_{tag}_x"] = x
        g[f"fold_{tag}_f"] = f; g[f"fold_{tag}_out"] = r.fri_fold_additive(f, m, basis, shift, eta, x)
        xin = shift ^ basis[0] ^ basis[m - 1]
        g[f"fold_{tag}_xin"] = xin; g[f"fold_{tag}_out_xin"] = r.fri_fold_additive(f, m, basis, shift, eta, xin)
    for tag, logn, eta in (("m1", 8, 1), ("m2", 8, 2), ("m3", 6, 3)):
        shift = r.fp_generator(); x = o.rand_fp(rng, 1)[0]; f = o.rand_fp(rng, 1 << logn)
        g[f"fold_{tag}_shift"] = shift; g[f"fold_{tag}_x"] = x; g[f"fold_{tag}_f"] = f
        g[f"fold_{tag}_out"] = r.fri_fold_mult(f, logn, shift, eta, x)
        gen = r.fp_subgroup_generator(logn); xin = shift.copy()
        for _ in range(5):
            xin = r.fp_mul(xin, gen)[0]
        g[f"fold_{tag}_xin"] = xin; g[f"fold_{tag}_out_xin"] = r.fri_fold_mult(f, logn, shift, eta, xin)
    # Merkle trees: (n_oracles, n_elems, coset, field_kind)
    for i, (no, n, cs, kind) in enumerate(((4, 64, 2, 1), (1, 64, 4, 1), (3, 32, 1, 1), (3, 64, 2, 2), (1, 64, 4, 2))):
        orc = [rand_elems(rng, n) if kind == 1 else o.rand_fp(rng, n) for _ in range(no)]
        g[f"mt{i}_oracles"] = np.stack(orc); g[f"mt{i}_cfg"] = np.array([no, n, cs, kind], dtype=np.uint64)
        nodes, _ = r.merkle_build(orc, cs, kind)
        g[f"mt{i}_nodes"] = np.frombuffer(nodes, dtype=np.uint8)
        r.seed(1000 + i)
        nodes, salts = r.merkle_build(orc, cs, kind, zk=True)
        g[f"mt{i}_zk_nodes"] = np.frombuffer(nodes, dtype=np.uint8)
        g[f"mt{i}_zk_salts"] = np.frombuffer(salts, dtype=np.uint8)
    # proof of work (pow.tcc:87-158): (challenge, work parameter) -> solution
    pw = []
    for i, work in enumerate((8, 12, 16, 18)):
        ch = hashlib.sha256(b"pow-challenge-%d" % i).digest()
        sol, bitlen, upper = r.pow_solve(ch, work)
        pw.append(f"{ch.hex()},{bitlen},{upper},{sol.hex()}")
    g["pow_vectors"] = np.array(pw)
    sq, pos = r.hashchain_demo()
    g["hashchain_squeeze"] = sq; g["hashchain_positions"] = pos
    # whole-prover fingerprints (sha256 of the canonical transcript dump)
    fps = []
    for log_n, zk in ((10, False), (10, True), (12, False)):
        secs, dump, acc = r.aurora_gf192(log_n, make_zk=zk)
        assert acc == 1
        fps.append(f"{log_n},{int(zk)},{len(dump)},{hashlib.sha256(dump).hexdigest()}")
    g["aurora_fingerprints"] = np.array(fps)
    # the other prover entry points (deterministic instances, see oracle/ref_shim/ref_api.cpp det_example)
    other = []
    for log_n in (10, 12):
        secs, dump, acc = r.aurora_edwards(log_n); assert acc == 1
        other.append(f"aurora_edwards,{log_n},0,{len(dump)},{hashlib.sha256(dump).hexdigest()}")
        secs, dump, acc, _ = r.fractal_edwards(log_n); assert acc == 1
        other.append(f"fractal_edwards,{log_n},0,{len(dump)},{hashlib.sha256(dump).hexdigest()}")
    for log_n, zk in ((12, False), (14, False), (10, True)):
        secs, dump, acc = r.ligero_gf192(log_n, make_zk=zk); assert acc == 1
        other.append(f"ligero_gf192,{log_n},{int(zk)},{len(dump)},{hashlib.sha256(dump).hexdigest()}")
    g["prover_fingerprints"] = np.array(other)
    fps = fps + other
    out = os.path.join(os.path.dirname(__file__), "hotpath_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes;", *fps, sep="\n")


if __name__ == "__main__":
    main()
