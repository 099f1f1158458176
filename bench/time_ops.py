"""Times the individual hot-path calls (CUDA events, device-resident).  Usage: python bench/time_ops.py [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libiop_b200 as L

dev = L.B200(0)
rng = np.random.default_rng(0)
n, m = 20, 23
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
def rnd(k):
    b = dev.alloc(k)
    for off in range(0, k, 1 << 18):
        kk = min(1 << 18, k - off)
        dev.upload(rng.integers(0, 2**64, size=(kk, 3), dtype=np.uint64), b, off)
    return b
c = rnd(1 << 21); e = dev.alloc(1 << m); o = dev.alloc(1 << 21)
def t(fn, reps=reps):
    fn(); fn(); dev.sync(); dev.timer_start()
    for _ in range(reps): fn()
    return dev.timer_stop() / reps
dom = L.Domain.standard(m)
res = {}
res["fft 2^20->2^23"] = t(lambda: dev.gf192_fft_dev(dom, c, 1 << n, e))
res["lde 2^20->2^23 fused (evals over H -> evals over L)"] = t(lambda: dev.gf192_lde_dev(L.Domain.standard(n), dom, c, e))
res["fft 2^20->2^20"] = t(lambda: dev.gf192_fft_dev(L.Domain.standard(n), c, 1 << n, o))
res["fft 32->2^23"] = t(lambda: dev.gf192_fft_dev(dom, c, 32, e))
res["ifft 2^20"] = t(lambda: dev.gf192_ifft_dev(L.Domain.standard(n), c, o))
res["ifft 2^21"] = t(lambda: dev.gf192_ifft_dev(L.Domain.standard(n + 1), c, o))
x = rng.integers(0, 2**64, size=3, dtype=np.uint64)
f = dev.alloc(1 << (m - 1))
res["fold 2^23 eta1"] = t(lambda: dev.fri_fold_dev(1, None, dom, 1, x, e, f))
d22 = L.Domain.standard(22)
res["fold 2^22 eta2"] = t(lambda: dev.fri_fold_dev(1, None, d22, 2, x, e, f))
nodes = dev.alloc_bytes(32 * (2 * (1 << (m - 1)) - 1))
e2 = [rnd(1 << m) for _ in range(3)]
res["merkle 4x2^23 c2"] = t(lambda: dev.merkle_build_dev([e] + e2, 1 << m, 2, 1, nodes))
res["merkle 1x2^23 c2"] = t(lambda: dev.merkle_build_dev([e], 1 << m, 2, 1, nodes))
res["merkle 1x2^22 c4"] = t(lambda: dev.merkle_build_dev([e], 1 << 22, 4, 1, nodes))
rows = [rnd(1 << 13) for _ in range(512)]
nodes_l = dev.alloc_bytes(32 * (2 * (1 << 13) - 1))
res["merkle ligero 512x2^13 c1"] = t(lambda: dev.merkle_build_dev(rows, 1 << 13, 1, 1, nodes_l))
# prime-field NTT (BASELINE config 3 shape: 2^20 coefficients onto the 2^23 coset, and the 2^20 inverse)
fp = L.FpParams.edwards_fr()
gen = np.array(fp.to_mont(19), dtype=np.uint64)
cd23, cd20 = L.Domain.coset(fp, 23, gen), L.Domain.coset(fp, 20, gen)
res["ntt 2^20->2^23"] = t(lambda: dev.fp_ntt_dev(fp, cd23, c, 1 << n, e))
res["intt 2^20"] = t(lambda: dev.fp_intt_dev(fp, cd20, c, o))
res["fold fp 2^23 eta1"] = t(lambda: dev.fri_fold_dev(2, fp, cd23, 1, x, e, f))
import ctypes as C
res["inv 2^23"] = t(lambda: dev._chk(dev.lib.b200_gf192_inv(dev.ctx, C.c_uint64(e2[0]), C.c_uint64(e), C.c_size_t(1 << m))))
res["inv 2^20"] = t(lambda: dev._chk(dev.lib.b200_gf192_inv(dev.ctx, C.c_uint64(o), C.c_uint64(c), C.c_size_t(1 << n))))
print(os.environ.get("B200_LIB", "default"), " | ".join(f"{k}: {v:.3f} ms" for k, v in res.items()))
