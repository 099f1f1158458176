"""ctypes binding of libiop_b200/lib/libiop_b200_dropin.so: the reference's own aurora / fractal / ligero SNARK provers
(unmodified libiop host code) running with the B200 library dropped in under their hot-path headers
(libiop_b200/dropin/).  Built by `make -C libiop_b200/dropin` where the reference checkout is available."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libiop_b200_dropin.so")
DEFAULT_SEED = 0x9E3779B97F4A7C15


def available():
    return os.path.exists(LIB_PATH)


class DropinProver:
    def __init__(self, lib_path=None):
        """lib_path: another build of the same objects (tests/hostsim: the host-logic tests link them against a host simulator
        of the C ABI); the product always loads LIB_PATH."""
        lib_path = lib_path or LIB_PATH
        if not os.path.exists(lib_path):
            raise RuntimeError(f"{lib_path} is missing: `make -C libiop_b200/dropin` (needs the reference checkout at build time)")
        self.lib = C.CDLL(lib_path)
        for fn in ("b200_prover_aurora_gf192", "b200_prover_aurora_gf192_shape", "b200_prover_aurora_edwards", "b200_prover_fractal_edwards",
                   "b200_prover_ligero_gf192"):
            getattr(self.lib, fn).restype = C.c_double

    def seed(self, s):
        self.lib.b200_prover_seed(C.c_uint64(s))

    def _run(self, fn, args, cap=1 << 27, extra=()):
        buf = (C.c_uint8 * cap)()
        ln, acc = C.c_size_t(0), C.c_int(-1)
        secs = fn(*args, buf, C.c_size_t(cap), C.byref(ln), C.byref(acc), *extra)
        if ln.value > cap:
            raise RuntimeError("transcript buffer too small")
        return secs, bytes(buf[:ln.value]), acc.value

    def aurora_gf192(self, log_n, k_inputs=31, make_zk=False, rs_extra=3, fri_loc=2, seed=DEFAULT_SEED, verify=True):
        """(prover seconds, canonical transcript bytes, verifier verdict) — BASELINE configs 1/2/5"""
        return self._run(self.lib.b200_prover_aurora_gf192,
                         (C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_int(1 if make_zk else 0), C.c_uint32(rs_extra), C.c_uint32(fri_loc),
                          C.c_uint64(seed), C.c_int(1 if verify else 0)))

    def aurora_gf192_shape(self, log_n, num_variables, k_inputs=31, make_zk=False, rs_extra=3, fri_loc=2, seed=DEFAULT_SEED, verify=True):
        """2^log_n constraints over `num_variables` variables (any count >= k_inputs)"""
        return self._run(self.lib.b200_prover_aurora_gf192_shape,
                         (C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_uint64(num_variables), C.c_int(1 if make_zk else 0), C.c_uint32(rs_extra),
                          C.c_uint32(fri_loc), C.c_uint64(seed), C.c_int(1 if verify else 0)))

    def aurora_gf192_repeat(self, log_n, repeats, k_inputs=31, make_zk=False, rs_extra=3, fri_loc=2, seed=DEFAULT_SEED):
        """prover seconds of `repeats` proofs of one generated instance (instance generation excluded)"""
        secs = (C.c_double * repeats)()
        self.lib.b200_prover_aurora_gf192_repeat(C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_int(1 if make_zk else 0), C.c_uint32(rs_extra),
                                                 C.c_uint32(fri_loc), C.c_uint64(seed), C.c_uint32(repeats), secs)
        return list(secs)

    def aurora_gf192_phases(self, log_n, repeats, k_inputs=31, make_zk=False, rs_extra=3, fri_loc=2, seed=DEFAULT_SEED):
        """per proof: seconds of [construction + registrations, produce_proof, get_transcript, destruction of the protocol
        objects and oracles, destruction of the argument] (diagnostic; bench/dropin_phases.py)"""
        secs = (C.c_double * (5 * repeats))()
        self.lib.b200_prover_aurora_gf192_phases(C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_int(1 if make_zk else 0), C.c_uint32(rs_extra),
                                                 C.c_uint32(fri_loc), C.c_uint64(seed), C.c_uint32(repeats), secs)
        return [list(secs[5 * r:5 * r + 5]) for r in range(repeats)]

    def aurora_edwards(self, log_n, k_inputs=31, rs_extra=3, fri_loc=2, seed=DEFAULT_SEED, verify=True):
        return self._run(self.lib.b200_prover_aurora_edwards,
                         (C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_uint32(rs_extra), C.c_uint32(fri_loc), C.c_uint64(seed),
                          C.c_int(1 if verify else 0)))

    def fractal_edwards(self, log_n, rs_extra=3, fri_loc=2, seed=DEFAULT_SEED, verify=True, cap=1 << 27):
        """(prover seconds, transcript bytes + index roots, verdict, indexer seconds) — BASELINE config 3"""
        idx = C.c_double(0)
        secs, dump, acc = self._run(self.lib.b200_prover_fractal_edwards,
                                    (C.c_uint32(log_n), C.c_uint32(rs_extra), C.c_uint32(fri_loc), C.c_uint64(seed), C.c_int(1 if verify else 0)),
                                    cap=cap, extra=(C.byref(idx),))
        return secs, dump, acc, idx.value

    def ligero_gf192(self, log_n, ratio=0.1, rs_extra=2, make_zk=False, seed=DEFAULT_SEED, verify=True):
        """BASELINE config 4"""
        return self._run(self.lib.b200_prover_ligero_gf192,
                         (C.c_uint32(log_n), C.c_double(ratio), C.c_uint32(rs_extra), C.c_int(1 if make_zk else 0), C.c_uint64(seed),
                          C.c_int(1 if verify else 0)))

    def ligero_gf192_phases(self, log_n, repeats, ratio=0.1, rs_extra=2, seed=DEFAULT_SEED):
        """per proof: seconds of [ligero_iop construction, register_interactions, register_queries, produce_proof, get_transcript,
        destruction] (diagnostic)"""
        secs = (C.c_double * (6 * repeats))()
        self.lib.b200_prover_ligero_gf192_phases(C.c_uint32(log_n), C.c_double(ratio), C.c_uint32(rs_extra), C.c_uint64(seed),
                                                 C.c_uint32(repeats), secs)
        return [list(secs[6 * r:6 * r + 6]) for r in range(repeats)]

    def arena_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.b200_prover_arena_stats(C.byref(a), C.byref(b), C.byref(c))
        return {"slabs": a.value, "reserved_bytes": b.value, "free_bytes": c.value}

    def residency_stats(self):
        h, m, s, r = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.b200_prover_residency_stats(C.byref(h), C.byref(m), C.byref(s), C.byref(r))
        return {"hits": h.value, "misses": m.value, "host_touched": s.value, "resident_bytes": r.value}

    def lazy_stats(self):
        """device-resident oracles: host touches that forced a whole-vector download, and query-phase gathers"""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.b200_prover_lazy_stats(C.byref(a), C.byref(b), C.byref(c))
        return {"faults": a.value, "fault_bytes": b.value, "gathers": c.value}

    def device_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.b200_prover_device_stats(C.byref(a), C.byref(b), C.byref(c))
        return {"launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}
