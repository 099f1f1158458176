"""libff::enter_block profile of the reference prover with the B200 drop-in (B200_REF_PROFILE=1); add B200_TRACE=1 for the
per-call C-ABI trace.  Usage: python bench/dropin_profile.py [log_n] [gpu|cpu]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
which = sys.argv[2] if len(sys.argv) > 2 else "gpu"
prover_name = sys.argv[3] if len(sys.argv) > 3 else "aurora_gf192"     # aurora_gf192 | aurora_edwards | fractal_edwards | ligero_gf192
if which == "gpu":
    from libiop_b200.prover import DropinProver
    r = DropinProver()
    os.environ.pop("B200_REF_PROFILE", None)
    zk = len(sys.argv) > 4 and sys.argv[4] == "zk"
    run = (lambda n, verify: r.aurora_gf192(n, make_zk=True, verify=verify)) if zk else getattr(r, prover_name)
    run(log_n, verify=False)          # warm-up: plans, pinned slabs, device blocks
    run(log_n, verify=False)
    os.environ["B200_REF_PROFILE"] = "1"
    run(log_n, verify=False)
    a = r.arena_stats(); s = r.residency_stats()
    print(f"pinned arena: {a['slabs']} slabs, {a['reserved_bytes'] / 2**30:.2f} GiB reserved, {a['free_bytes'] / 2**30:.2f} GiB free; "
          f"residency: {s['hits']} hits, {s['misses']} misses, {s['host_touched']} host-touched, {s['resident_bytes'] / 2**30:.2f} GiB resident", flush=True)
else:
    from oracle_lib import Reference
    os.environ["B200_REF_PROFILE"] = "1"
    getattr(Reference(), prover_name)(log_n, verify=False)
