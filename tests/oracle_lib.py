"""ctypes bindings for the parity oracle (oracle/liboracle.so) and, when it was built, the unmodified
reference (oracle/_ref/libiopref.so).  TEST INFRASTRUCTURE: imported only by tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke().

Field elements travel as numpy uint64 arrays of shape (n, 3): little-endian limbs, 24 bytes each
(gf192.hpp:98-99; Fp: Montgomery limbs, fp.hpp:42).
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


def _p(a):
    return a.ctypes.data_as(_u64p)


def _b(a):
    return a.ctypes.data_as(_u8p)


def el(lo=0, mid=0, hi=0):
    return np.array([lo, mid, hi], dtype=np.uint64)


def elems(n):
    return np.zeros((n, 3), dtype=np.uint64)


def rand_elems(rng, n):
    return rng.integers(0, 2**64, size=(n, 3), dtype=np.uint64)


def standard_basis(m):
    b = elems(m)
    for i in range(m):
        b[i, i // 64] = np.uint64(1) << np.uint64(i % 64)
    return b


class FpParams(C.Structure):
    _fields_ = [("mod", C.c_uint64 * 3), ("r2", C.c_uint64 * 3), ("one", C.c_uint64 * 3),
                ("inv", C.c_uint64), ("two_adicity", C.c_uint32), ("num_bits", C.c_uint32),
                ("generator", C.c_uint64 * 3)]


class Oracle:
    """The plain-C restatement (oracle/*.c)."""

    def __init__(self, build=True):
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if build and not os.path.exists(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(so)
        self.fp = FpParams()
        self.lib.orc_fp_params_edwards(C.byref(self.fp))
        self.lib.orc_hashchain_squeeze_position.restype = C.c_uint64

    # ---- gf192 ----
    def gf192_mul(self, a, b):
        a = np.ascontiguousarray(a).reshape(-1, 3); b = np.ascontiguousarray(b).reshape(-1, 3)
        r = elems(len(a))
        self.lib.orc_gf192_mul_vec(_p(r), _p(a), _p(b), C.c_size_t(len(a)))
        return r

    def gf192_mul_bitwise(self, a, b):
        r = elems(1)
        a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
        self.lib.orc_gf192_mul_bitwise(_p(r), _p(a), _p(b))
        return r[0]

    def gf192_inv(self, a):
        a = np.ascontiguousarray(a).reshape(-1, 3)
        r = elems(len(a))
        self.lib.orc_gf192_inv_vec(_p(r), _p(a), C.c_size_t(len(a)))
        return r

    # ---- Fp ----
    def fp_mul(self, a, b):
        a = np.ascontiguousarray(a).reshape(-1, 3); b = np.ascontiguousarray(b).reshape(-1, 3)
        r = elems(len(a))
        self.lib.orc_fp_mul_vec(C.byref(self.fp), _p(r), _p(a), _p(b), C.c_size_t(len(a)))
        return r

    def _fp2(self, fn, a, b):
        r = elems(1)
        a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
        fn(C.byref(self.fp), _p(r), _p(a), _p(b))
        return r[0]

    def fp_add(self, a, b):
        return self._fp2(self.lib.orc_fp_add, a, b)

    def fp_sub(self, a, b):
        return self._fp2(self.lib.orc_fp_sub, a, b)

    def fp_inv(self, a):
        r = elems(1); a = np.ascontiguousarray(a)
        self.lib.orc_fp_inv(C.byref(self.fp), _p(r), _p(a))
        return r[0]

    def fp_from_u64(self, v):
        r = elems(1)
        self.lib.orc_fp_from_u64(C.byref(self.fp), _p(r), C.c_uint64(v))
        return r[0]

    def fp_to_int(self, a):
        r = elems(1); a = np.ascontiguousarray(a)
        self.lib.orc_fp_to_plain(C.byref(self.fp), _p(r), _p(a))
        return int(r[0, 0]) | (int(r[0, 1]) << 64) | (int(r[0, 2]) << 128)

    def fp_generator(self):
        return np.array(list(self.fp.generator), dtype=np.uint64)

    def fp_subgroup_generator(self, logn):
        r = elems(1)
        self.lib.orc_fp_subgroup_generator(C.byref(self.fp), _p(r), C.c_uint32(logn))
        return r[0]

    def fp_modulus(self):
        return int(self.fp.mod[0]) | (int(self.fp.mod[1]) << 64) | (int(self.fp.mod[2]) << 128)

    def rand_fp(self, rng, n):
        p = self.fp_modulus()
        out = elems(n)
        for i in range(n):
            v = int.from_bytes(rng.bytes(32), "little") % p
            out[i] = [v & (2**64 - 1), (v >> 64) & (2**64 - 1), v >> 128]
        return out

    # ---- FFTs ----
    def additive_fft(self, coeffs, m, basis, shift):
        coeffs = np.ascontiguousarray(coeffs).reshape(-1, 3)
        out = elems(1 << m)
        self.lib.orc_additive_fft(_p(out), _p(coeffs), C.c_size_t(len(coeffs)), C.c_uint32(m),
                                  _p(np.ascontiguousarray(basis)), _p(np.ascontiguousarray(shift)))
        return out

    def additive_ifft(self, evals, m, basis, shift):
        evals = np.ascontiguousarray(evals).reshape(-1, 3)
        assert len(evals) == 1 << m
        out = elems(1 << m)
        self.lib.orc_additive_ifft(_p(out), _p(evals), C.c_uint32(m),
                                   _p(np.ascontiguousarray(basis)), _p(np.ascontiguousarray(shift)))
        return out

    def naive_fft_gf192(self, coeffs, m, basis, shift):
        coeffs = np.ascontiguousarray(coeffs).reshape(-1, 3)
        out = elems(1 << m)
        self.lib.orc_naive_fft_gf192(_p(out), _p(coeffs), C.c_size_t(len(coeffs)), C.c_uint32(m),
                                     _p(np.ascontiguousarray(basis)), _p(np.ascontiguousarray(shift)))
        return out

    def naive_fft_fp(self, coeffs, logn, g, shift):
        coeffs = np.ascontiguousarray(coeffs).reshape(-1, 3)
        out = elems(1 << logn)
        self.lib.orc_naive_fft_fp(C.byref(self.fp), _p(out), _p(coeffs), C.c_size_t(len(coeffs)),
                                  C.c_uint32(logn), _p(np.ascontiguousarray(g)), _p(np.ascontiguousarray(shift)))
        return out

    def mult_fft(self, coeffs, logn, g, shift):
        coeffs = np.ascontiguousarray(coeffs).reshape(-1, 3)
        out = elems(1 << logn)
        self.lib.orc_mult_fft(C.byref(self.fp), _p(out), _p(coeffs), C.c_size_t(len(coeffs)),
                              C.c_uint32(logn), _p(np.ascontiguousarray(g)), _p(np.ascontiguousarray(shift)))
        return out

    def mult_ifft(self, evals, logn, g, shift):
        evals = np.ascontiguousarray(evals).reshape(-1, 3)
        out = elems(1 << logn)
        self.lib.orc_mult_ifft(C.byref(self.fp), _p(out), _p(evals), C.c_uint32(logn),
                               _p(np.ascontiguousarray(g)), _p(np.ascontiguousarray(shift)))
        return out

    # ---- FRI ----
    def fri_fold_additive(self, f, m, basis, shift, eta, x):
        f = np.ascontiguousarray(f).reshape(-1, 3)
        out = elems(1 << (m - eta))
        self.lib.orc_fri_fold_additive(_p(out), _p(f), C.c_uint32(m), _p(np.ascontiguousarray(basis)),
                                       _p(np.ascontiguousarray(shift)), C.c_uint32(eta),
                                       _p(np.ascontiguousarray(x)))
        return out

    def fri_fold_mult(self, f, logn, g, shift, eta, x):
        f = np.ascontiguousarray(f).reshape(-1, 3)
        out = elems(1 << (logn - eta))
        self.lib.orc_fri_fold_multiplicative(C.byref(self.fp), _p(out), _p(f), C.c_uint32(logn),
                                             _p(np.ascontiguousarray(g)), _p(np.ascontiguousarray(shift)),
                                             C.c_uint32(eta), _p(np.ascontiguousarray(x)))
        return out

    # ---- hashing ----
    def blake2b(self, data, outlen=32, key=b""):
        out = (C.c_uint8 * outlen)()
        rc = self.lib.orc_blake2b(out, C.c_size_t(outlen), data, C.c_size_t(len(data)), key, C.c_size_t(len(key)))
        assert rc == 0
        return bytes(out)

    def pow_solve(self, challenge, pow_bitlen, pow_upperbound=0):
        out = (C.c_uint8 * 32)()
        self.lib.orc_pow_solve.restype = C.c_uint64
        tries = self.lib.orc_pow_solve(out, bytes(challenge), C.c_size_t(32), C.c_uint32(pow_bitlen), C.c_uint64(pow_upperbound))
        return bytes(out), tries

    def merkle_build(self, oracles, coset_size, domain_kind, salts=None, salt_len=32, digest_len=32):
        oracles = [np.ascontiguousarray(o).reshape(-1, 3) for o in oracles]
        n = len(oracles[0])
        leaves = n // coset_size
        nodes = np.zeros((2 * leaves - 1) * digest_len, dtype=np.uint8)
        arr = (_u64p * len(oracles))(*[_p(o) for o in oracles])
        sp = None
        if salts is not None:
            salts = np.ascontiguousarray(np.frombuffer(bytes(salts), dtype=np.uint8))
            sp = _b(salts)
        self.lib.orc_merkle_build(_b(nodes), arr, C.c_size_t(len(oracles)), C.c_size_t(n),
                                  C.c_size_t(coset_size), C.c_int(domain_kind), sp,
                                  C.c_size_t(salt_len), C.c_size_t(digest_len))
        return nodes.tobytes()


class Reference:
    """The unmodified reference on the CPU (oracle/_ref/libiopref.so), when present: the checker for the whole-prover
    parity tests and the CPU baseline of bench.py."""
    PREFIX = "ref_"
    LIBNAME = "libiopref.so"

    @classmethod
    def path(cls):
        return os.path.join(ORACLE_DIR, "_ref", cls.LIBNAME)

    @classmethod
    def available(cls):
        return os.path.exists(cls.path())

    def __init__(self):
        real = C.CDLL(self.path())
        prefix = self.PREFIX

        class _Renamed:
            def __getattr__(self, name):
                return getattr(real, prefix + name[4:]) if name.startswith("ref_") else getattr(real, name)
        self.lib = _Renamed()
        self._real = real
        self.lib.ref_init()
        for fn in ("ref_aurora_gf192", "ref_aurora_gf192_shape", "ref_aurora_edwards", "ref_fractal_edwards", "ref_ligero_gf192"):
            getattr(self.lib, fn).restype = C.c_double

    def seed(self, s):
        self.lib.ref_seed(C.c_uint64(s))

    def gf192_mul(self, a, b):
        a = np.ascontiguousarray(a).reshape(-1, 3); b = np.ascontiguousarray(b).reshape(-1, 3)
        r = elems(len(a)); self.lib.ref_gf192_mul(_p(r), _p(a), _p(b), C.c_size_t(len(a))); return r

    def gf192_inv(self, a):
        a = np.ascontiguousarray(a).reshape(-1, 3)
        r = elems(len(a)); self.lib.ref_gf192_inv(_p(r), _p(a), C.c_size_t(len(a))); return r

    def _fp(self, fn, a, b):
        a = np.ascontiguousarray(a).reshape(-1, 3); b = np.ascontiguousarray(b).reshape(-1, 3)
        r = elems(len(a)); fn(_p(r), _p(a), _p(b), C.c_size_t(len(a))); return r

    def fp_mul(self, a, b):
        return self._fp(self.lib.ref_fp_mul, a, b)

    def fp_add(self, a, b):
        return self._fp(self.lib.ref_fp_add, a, b)

    def fp_sub(self, a, b):
        return self._fp(self.lib.ref_fp_sub, a, b)

    def fp_inv(self, a):
        a = np.ascontiguousarray(a).reshape(-1, 3)
        r = elems(len(a)); self.lib.ref_fp_inv(_p(r), _p(a), C.c_size_t(len(a))); return r

    def fp_from_u64(self, v):
        r = elems(1); self.lib.ref_fp_from_u64(_p(r), C.c_uint64(v)); return r[0]

    def fp_generator(self):
        r = elems(1); self.lib.ref_fp_generator(_p(r)); return r[0]

    def fp_subgroup_generator(self, logn):
        r = elems(1); self.lib.ref_fp_subgroup_generator(_p(r), C.c_uint32(logn)); return r[0]

    def fp_params(self):
        mod = elems(1); r2 = elems(1); inv = C.c_uint64(); s = C.c_uint32()
        self.lib.ref_fp_params(_p(mod), _p(r2), C.byref(inv), C.byref(s))
        return mod[0], r2[0], inv.value, s.value

    def additive_fft(self, coeffs, m, basis, shift):
        coeffs = np.ascontiguousarray(coeffs).reshape(-1, 3)
        out = elems(1 << m)
        self.lib.ref_additive_fft(_p(out), _p(coeffs), C.c_size_t(len(coeffs)), C.c_uint32(m),
                                  _p(np.ascontiguousarray(basis)), _p(np.ascontiguousarray(shift)))
        return out

    def additive_ifft(self, evals, m, basis, shift):
        evals = np.ascontiguousarray(evals).reshape(-1, 3)
        out = elems(1 << m)
        self.lib.ref_additive_ifft(_p(out), _p(evals), C.c_uint32(m),
                                   _p(np.ascontiguousarray(basis)), _p(np.ascontiguousarray(shift)))
        return out

    def naive_fft_gf192(self, coeffs, m, basis, shift):
        coeffs = np.ascontiguousarray(coeffs).reshape(-1, 3)
        out = elems(1 << m)
        self.lib.ref_naive_fft_gf192(_p(out), _p(coeffs), C.c_size_t(len(coeffs)), C.c_uint32(m),
                                     _p(np.ascontiguousarray(basis)), _p(np.ascontiguousarray(shift)))
        return out

    def mult_fft(self, coeffs, logn, shift):
        coeffs = np.ascontiguousarray(coeffs).reshape(-1, 3)
        out = elems(1 << logn)
        self.lib.ref_mult_fft(_p(out), _p(coeffs), C.c_size_t(len(coeffs)), C.c_uint32(logn),
                              _p(np.ascontiguousarray(shift)))
        return out

    def mult_ifft(self, evals, logn, shift):
        evals = np.ascontiguousarray(evals).reshape(-1, 3)
        out = elems(1 << logn)
        self.lib.ref_mult_ifft(_p(out), _p(evals), C.c_uint32(logn), _p(np.ascontiguousarray(shift)))
        return out

    def fri_fold_additive(self, f, m, basis, shift, eta, x):
        f = np.ascontiguousarray(f).reshape(-1, 3)
        out = elems(1 << (m - eta))
        self.lib.ref_fri_fold_additive(_p(out), _p(f), C.c_uint32(m), _p(np.ascontiguousarray(basis)),
                                       _p(np.ascontiguousarray(shift)), C.c_uint32(eta),
                                       _p(np.ascontiguousarray(x)))
        return out

    def fri_fold_mult(self, f, logn, shift, eta, x):
        f = np.ascontiguousarray(f).reshape(-1, 3)
        out = elems(1 << (logn - eta))
        self.lib.ref_fri_fold_mult(_p(out), _p(f), C.c_uint32(logn), _p(np.ascontiguousarray(shift)),
                                   C.c_uint32(eta), _p(np.ascontiguousarray(x)))
        return out

    def merkle_build(self, oracles, coset_size, field_kind, zk=False, security=128):
        oracles = [np.ascontiguousarray(o).reshape(-1, 3) for o in oracles]
        n = len(oracles[0]); leaves = n // coset_size; dlen = (2 * security + 7) // 8
        nodes = np.zeros((2 * leaves - 1) * dlen, dtype=np.uint8)
        salts = np.zeros(leaves * dlen, dtype=np.uint8)
        arr = (_u64p * len(oracles))(*[_p(o) for o in oracles])
        self.lib.ref_merkle_build(_b(nodes), _b(salts), arr, C.c_size_t(len(oracles)), C.c_size_t(n),
                                  C.c_size_t(coset_size), C.c_int(field_kind), C.c_int(1 if zk else 0),
                                  C.c_size_t(security))
        return nodes.tobytes(), (salts.tobytes() if zk else None)

    def pow_solve(self, challenge, work_parameter, cost_per_hash=1):
        out = (C.c_uint8 * 32)(); bl = C.c_uint32(); ub = C.c_uint64()
        self.lib.ref_pow_solve(out, bytes(challenge), C.c_size_t(work_parameter), C.c_size_t(cost_per_hash), C.byref(bl), C.byref(ub))
        return bytes(out), bl.value, ub.value

    def squeeze(self, field_kind, root, index, n):
        """blake2b_FieldT_randomness_extractor<gf192 | edwards_Fr>(root, index, n) of the unmodified reference."""
        out = elems(max(n, 1))
        fn = self.lib.ref_squeeze_gf192 if field_kind == 1 else self.lib.ref_squeeze_edwards
        fn(_p(out), bytes(root), C.c_size_t(len(root)), C.c_size_t(index), C.c_size_t(n))
        return out[:n]

    def hashchain_demo(self, absorbed=b"\x00" * 32):
        sq = elems(2); pos = np.zeros(3, dtype=np.uint64)
        self.lib.ref_hashchain_demo(_p(sq), _p(pos), absorbed)
        return sq, pos

    def aurora_gf192(self, log_n, k_inputs=31, make_zk=False, rs_extra=3, fri_loc=2,
                     seed=0x9E3779B97F4A7C15, verify=True, cap=1 << 26):
        buf = np.zeros(cap, dtype=np.uint8)
        ln = C.c_size_t(); acc = C.c_int()
        secs = self.lib.ref_aurora_gf192(C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_int(int(make_zk)),
                                         C.c_uint32(rs_extra), C.c_uint32(fri_loc), C.c_uint64(seed),
                                         C.c_int(int(verify)), _b(buf), C.c_size_t(cap),
                                         C.byref(ln), C.byref(acc))
        return secs, buf[:ln.value].tobytes(), acc.value

    def aurora_gf192_shape(self, log_n, num_variables, k_inputs=31, make_zk=False, rs_extra=3, fri_loc=2,
                           seed=0x9E3779B97F4A7C15, verify=True, cap=1 << 26):
        """2^log_n constraints over `num_variables` variables (generate_r1cs_example with a free variable count)"""
        buf = np.zeros(cap, dtype=np.uint8)
        ln = C.c_size_t(); acc = C.c_int()
        secs = self.lib.ref_aurora_gf192_shape(C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_uint64(num_variables), C.c_int(int(make_zk)),
                                               C.c_uint32(rs_extra), C.c_uint32(fri_loc), C.c_uint64(seed),
                                               C.c_int(int(verify)), _b(buf), C.c_size_t(cap), C.byref(ln), C.byref(acc))
        return secs, buf[:ln.value].tobytes(), acc.value

    def aurora_gf192_repeat(self, log_n, repeats, k_inputs=31, make_zk=False, rs_extra=3, fri_loc=2, seed=0x9E3779B97F4A7C15):
        secs = (C.c_double * repeats)()
        self.lib.ref_aurora_gf192_repeat(C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_int(int(make_zk)), C.c_uint32(rs_extra),
                                         C.c_uint32(fri_loc), C.c_uint64(seed), C.c_uint32(repeats), secs)
        return list(secs)

    def _run(self, fn, args, cap=1 << 27):
        buf = np.zeros(cap, dtype=np.uint8)
        ln = C.c_size_t(); acc = C.c_int()
        extra = []
        secs = fn(*args, _b(buf), C.c_size_t(cap), C.byref(ln), C.byref(acc), *extra)
        return secs, buf[:ln.value].tobytes(), acc.value

    def aurora_edwards(self, log_n, k_inputs=31, rs_extra=3, fri_loc=2, seed=0x9E3779B97F4A7C15, verify=True):
        return self._run(self.lib.ref_aurora_edwards, (C.c_uint32(log_n), C.c_uint32(k_inputs), C.c_uint32(rs_extra),
                                                       C.c_uint32(fri_loc), C.c_uint64(seed), C.c_int(int(verify))))

    def fractal_edwards(self, log_n, rs_extra=3, fri_loc=2, seed=0x9E3779B97F4A7C15, verify=True, cap=1 << 27):
        buf = np.zeros(cap, dtype=np.uint8)
        ln = C.c_size_t(); acc = C.c_int(); idx = C.c_double()
        secs = self.lib.ref_fractal_edwards(C.c_uint32(log_n), C.c_uint32(rs_extra), C.c_uint32(fri_loc), C.c_uint64(seed),
                                            C.c_int(int(verify)), _b(buf), C.c_size_t(cap), C.byref(ln), C.byref(acc), C.byref(idx))
        return secs, buf[:ln.value].tobytes(), acc.value, idx.value

    def ligero_gf192(self, log_n, ratio=0.1, rs_extra=2, make_zk=False, seed=0x9E3779B97F4A7C15, verify=True):
        return self._run(self.lib.ref_ligero_gf192, (C.c_uint32(log_n), C.c_double(ratio), C.c_uint32(rs_extra), C.c_int(int(make_zk)),
                                                     C.c_uint64(seed), C.c_int(int(verify))))


def _dropin():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from libiop_b200 import prover
    return prover


class ReferenceGpu:
    """The product under test at whole-prover level: libiop_b200/lib/libiop_b200_dropin.so — the reference's own provers
    with the B200 library dropped in (libiop_b200/dropin/, bound by libiop_b200/prover.py).  Same call signatures as
    Reference above so the parity tests run both side by side.  Needs a CUDA device at call time."""

    @classmethod
    def available(cls):
        return _dropin().available()

    def __init__(self):
        self._p = _dropin().DropinProver()

    def __getattr__(self, name):
        return getattr(self._p, name)


def squeeze_model(root, index, n, modulus=None):
    """hashlib restatement of blake2b_FieldT_randomness_extractor (bcs/hashing/blake2b.tcc:162-256): element i =
    BLAKE2b(digest 24, key = u64 LE i, root || u64 LE index); prime fields clear the bits above the modulus' top bit
    and re-draw with key += n while the value is >= modulus."""
    import hashlib
    msg = bytes(root) + int(index).to_bytes(8, "little")
    out = np.zeros((n, 3), dtype=np.uint64)
    mask = (1 << modulus.bit_length()) - 1 if modulus else None
    for i in range(n):
        key = i
        while True:
            v = int.from_bytes(hashlib.blake2b(msg, digest_size=24, key=key.to_bytes(8, "little")).digest(), "little")
            if modulus is None:
                break
            v &= mask
            if v < modulus:
                break
            key += n
        out[i] = [v & (2**64 - 1), (v >> 64) & (2**64 - 1), v >> 128]
    return out
