"""One invocation of ONE kernel family at the benchmark's sizes, for ncu (`--set full`, one family per run so that the
capture stays small).  Usage: python bench/profile_kernels.py <family> [reps]
families: lde lde_fused ifft ntt intt fold fold_fp merkle4 merkle1 merkle_ligero squeeze pow lincomb mul_add inv subset_sum"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libiop_b200 as L

fam = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = L.B200(0)
rng = np.random.default_rng(0)
n, m = 20, 23


def rnd(k):
    b = dev.alloc(k)
    for off in range(0, k, 1 << 18):
        kk = min(1 << 18, k - off)
        dev.upload(rng.integers(0, 2**64, size=(kk, 3), dtype=np.uint64), b, off)
    return b


x = rng.integers(0, 2**64, size=3, dtype=np.uint64)
for _ in range(reps):
    if fam == "lde":
        dev.gf192_fft_dev(L.Domain.standard(m), rnd(1 << n), 1 << n, dev.alloc(1 << m))
    elif fam == "lde_fused":
        dev.gf192_lde_dev(L.Domain.standard(n), L.Domain.standard(m), rnd(1 << n), dev.alloc(1 << m))
    elif fam == "ifft":
        dev.gf192_ifft_dev(L.Domain.standard(n), rnd(1 << n), dev.alloc(1 << n))
    elif fam in ("ntt", "intt", "fold_fp"):
        fp = L.FpParams.edwards_fr()
        gen = np.array(fp.to_mont(19), dtype=np.uint64)
        if fam == "ntt":
            dev.fp_ntt_dev(fp, L.Domain.coset(fp, m, gen), rnd(1 << n), 1 << n, dev.alloc(1 << m))
        elif fam == "intt":
            dev.fp_intt_dev(fp, L.Domain.coset(fp, n, gen), rnd(1 << n), dev.alloc(1 << n))
        else:
            dev.fri_fold_dev(2, fp, L.Domain.coset(fp, m, gen), 1, x, rnd(1 << m), dev.alloc(1 << (m - 1)))
            dev.fri_fold_dev(2, fp, L.Domain.coset(fp, m - 1, gen), 2, x, rnd(1 << (m - 1)), dev.alloc(1 << (m - 3)))
    elif fam == "fold":
        dev.fri_fold_dev(1, None, L.Domain.standard(m), 1, x, rnd(1 << m), dev.alloc(1 << (m - 1)))
        dev.fri_fold_dev(1, None, L.Domain.standard(m - 1), 2, x, rnd(1 << (m - 1)), dev.alloc(1 << (m - 3)))
    elif fam == "merkle4":
        dev.merkle_build_dev([rnd(1 << m) for _ in range(4)], 1 << m, 2, 1, dev.alloc_bytes(32 * (2 * (1 << (m - 1)) - 1)))
    elif fam == "merkle1":
        dev.merkle_build_dev([rnd(1 << m)], 1 << m, 2, 1, dev.alloc_bytes(32 * (2 * (1 << (m - 1)) - 1)))
        dev.merkle_build_dev([rnd(1 << (m - 1))], 1 << (m - 1), 4, 1, dev.alloc_bytes(32 * (2 * (1 << (m - 3)) - 1)))
    elif fam == "merkle_ligero":       # BASELINE config 4: 2^13 leaves of 512 rows x 24 B
        rows = [rnd(1 << 13) for _ in range(512)]
        dev.merkle_build_dev(rows, 1 << 13, 1, 1, dev.alloc_bytes(32 * (2 * (1 << 13) - 1)))
    elif fam == "squeeze":
        dev.squeeze_field_elements(1, None, bytes(range(40)), 1 << 18)
    elif fam == "pow":
        dev.pow_solve(bytes(range(32)), 23)
    elif fam in ("lincomb", "mul_add", "inv"):
        k = 1 << 22
        a, b, c, o = rnd(k), rnd(k), rnd(k), dev.alloc(k)
        if fam == "lincomb":
            co = rng.integers(0, 2**64, size=(3, 3), dtype=np.uint64)
            arr = (C.c_uint64 * 3)(a, b, c)
            dev._chk(dev.lib.b200_vec_lincomb(dev.ctx, C.c_int(1), None, C.c_uint64(o), C.c_size_t(k), C.c_size_t(3),
                                              co.ctypes.data_as(C.POINTER(C.c_uint64)), arr))
        elif fam == "mul_add":
            dev._chk(dev.lib.b200_vec_mul_add(dev.ctx, C.c_int(1), None, C.c_uint64(o), C.c_uint64(a), C.c_uint64(b), C.c_uint64(c),
                                              C.c_uint64(0), C.c_int(0), C.c_size_t(k)))
        else:
            dev._chk(dev.lib.b200_gf192_inv(dev.ctx, C.c_uint64(o), C.c_uint64(a), C.c_size_t(k)))
    elif fam == "subset_sum":
        t = rng.integers(0, 2**64, size=(2, m + 1, 3), dtype=np.uint64)
        o = dev.alloc(1 << m)
        dev._chk(dev.lib.b200_gf192_subset_sum_products(dev.ctx, C.c_uint64(o), C.c_uint32(m), C.c_uint32(2),
                                                        t.ctypes.data_as(C.POINTER(C.c_uint64))))
    else:
        raise SystemExit("unknown family " + fam)
dev.sync()
print("ok", fam)
