"""per-launch CUDA-event timing of the tile passes of one LDE 2^20 -> 2^23 and one IFFT 2^20 (b200_profile_begin/end)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libiop_b200 as L
dev = L.B200(0); lib, ctx = dev.lib, dev.ctx
rng = np.random.default_rng(0)
c = dev.alloc(1 << 20); e = dev.alloc(1 << 23)
for off in range(0, 1 << 20, 1 << 18):
    dev.upload(rng.integers(0, 2**64, size=(1 << 18, 3), dtype=np.uint64), c, off)
dom = L.Domain.standard(23); sub = L.Domain.standard(20)
for what in ("fft", "ifft"):
    for rep in range(3):
        if rep == 2: lib.b200_profile_begin(ctx)
        if what == "fft": dev.gf192_fft_dev(dom, c, 1 << 20, e)
        else: dev.gf192_ifft_dev(sub, e, c)
    buf = C.create_string_buffer(1 << 16)
    lib.b200_profile_end(ctx, buf, C.c_size_t(len(buf)))
    print("==", what); print(buf.value.decode())
