"""Short driver for ncu: one Aurora-shaped LDE (2^20 coefficients -> 2^23 evaluations), one 2^20 IFFT,
one FRI fold and one Merkle tree.  Usage (under gpurun): ncu ... python bench/profile_lde.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libiop_b200 as L

dev = L.B200(0)
rng = np.random.default_rng(0)
n, m = 20, 23
dom = L.Domain.standard(m)
c = dev.alloc(1 << n)
for off in range(0, 1 << n, 1 << 18):
    dev.upload(rng.integers(0, 2**64, size=(1 << 18, 3), dtype=np.uint64), c, off)
e = dev.alloc(1 << m)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    dev.gf192_fft_dev(dom, c, 1 << n, e)
sub = L.Domain.standard(n, shift_int=1 << m)
o = dev.alloc(1 << n)
dev.gf192_ifft_dev(sub, e, o)
f = dev.alloc(1 << (m - 1))
dev.fri_fold_dev(1, None, dom, 1, rng.integers(0, 2**64, size=3, dtype=np.uint64), e, f)
nodes = dev.alloc_bytes(32 * (2 * (1 << (m - 1)) - 1))
dev.merkle_build_dev([e], 1 << m, 2, 1, nodes)
dev.sync()
print("ok")
