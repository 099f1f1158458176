"""latency of small additive transforms (one tile): cluster/DSMEM kernel vs single CTA (B200_NO_CLUSTER=1)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libiop_b200 as L
dev = L.B200(0)
rng = np.random.default_rng(0)
c = dev.upload(rng.integers(0, 2**64, size=(1 << 13, 3), dtype=np.uint64)); e = dev.alloc(1 << 16)
out = []
for d, m in ((9, 11), (10, 12), (11, 13), (12, 14), (13, 15), (11, 11), (12, 12)):
    dom, sub = L.Domain.standard(m), L.Domain.standard(d)
    for name, fn in ((f"fft 2^{d}->2^{m}", lambda: dev.gf192_fft_dev(dom, c, 1 << d, e)), (f"ifft 2^{d}", lambda: dev.gf192_ifft_dev(sub, e, c))):
        fn(); fn(); dev.sync(); dev.timer_start()
        for _ in range(20): fn()
        out.append(f"{name}: {dev.timer_stop() / 20 * 1000:.0f} us")
print(("no-cluster " if os.environ.get("B200_NO_CLUSTER") else "cluster    ") + " | ".join(out))
