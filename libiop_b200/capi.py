"""ctypes binding of include/b200_iop.h.  Field elements are numpy uint64 arrays of shape (n, 3)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb200iop.so")

_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)
_szp = C.POINTER(C.c_size_t)

DOMAIN_AFFINE, DOMAIN_MULT = 1, 2
FIELD_GF192, FIELD_FP = 1, 2


class B200Error(RuntimeError):
    pass


class FpParams(C.Structure):
    """b200_fp_params (Fp_model statics: fp.hpp:40-60)."""
    _fields_ = [("modulus", C.c_uint64 * 3), ("r2", C.c_uint64 * 3), ("one_mont", C.c_uint64 * 3),
                ("inv", C.c_uint64), ("two_adicity", C.c_uint32), ("num_bits", C.c_uint32)]

    @staticmethod
    def edwards_fr():
        """edwards_Fr (depends/libff/libff/algebra/curves/edwards/edwards_fields.cpp:24-46)."""
        p = 1552511030102430251236801561344621993261920897571225601
        R = 1 << 192
        P = FpParams()
        for i in range(3):
            P.modulus[i] = (p >> (64 * i)) & (2**64 - 1)
            P.r2[i] = ((R * R % p) >> (64 * i)) & (2**64 - 1)
            P.one_mont[i] = ((R % p) >> (64 * i)) & (2**64 - 1)
        P.inv = (-pow(p, -1, 2**64)) % 2**64
        P.two_adicity = 31
        P.num_bits = 181
        return P

    def p(self):
        return sum(int(self.modulus[i]) << (64 * i) for i in range(3))

    def to_mont(self, v):
        v = (v << 192) % self.p()
        return np.array([(v >> (64 * i)) & (2**64 - 1) for i in range(3)], dtype=np.uint64)

    def from_mont(self, limbs):
        v = sum(int(limbs[i]) << (64 * i) for i in range(3))
        return v * pow(1 << 192, -1, self.p()) % self.p()

    def subgroup_generator(self, logn, generator=19):
        """multiplicative_generator^((p-1)/2^logn) in Montgomery form (subgroup.tcc:56-59)."""
        return self.to_mont(pow(generator, (self.p() - 1) >> logn, self.p()))


class _CDomain(C.Structure):
    _fields_ = [("kind", C.c_int), ("dim", C.c_uint32), ("basis", _u64p),
                ("shift", C.c_uint64 * 3), ("generator", C.c_uint64 * 3)]


class Domain:
    """field_subset<FieldT> flattened (field_subset.hpp:25-77)."""

    def __init__(self, kind, dim, basis=None, shift=None, generator=None):
        self.kind, self.dim = kind, dim
        self.basis = None if basis is None else np.ascontiguousarray(basis, dtype=np.uint64).reshape(-1, 3)
        self.shift = np.zeros(3, dtype=np.uint64) if shift is None else np.ascontiguousarray(shift, dtype=np.uint64)
        self.generator = np.zeros(3, dtype=np.uint64) if generator is None else np.ascontiguousarray(generator, dtype=np.uint64)
        self.c = _CDomain()
        self.c.kind, self.c.dim = kind, dim
        if self.basis is not None:
            assert len(self.basis) >= dim
            self.c.basis = self.basis.ctypes.data_as(_u64p)
        for i in range(3):
            self.c.shift[i] = int(self.shift[i])
            self.c.generator[i] = int(self.generator[i])

    @staticmethod
    def affine(basis, shift):
        basis = np.ascontiguousarray(basis, dtype=np.uint64).reshape(-1, 3)
        return Domain(DOMAIN_AFFINE, len(basis), basis=basis, shift=shift)

    @staticmethod
    def standard(dim, shift_int=None):
        """affine_subspace::shifted_standard_basis (subspace.tcc:94-108,220-224): basis x^i, shift x^dim."""
        b = np.zeros((dim, 3), dtype=np.uint64)
        for i in range(dim):
            b[i, i // 64] = np.uint64(1) << np.uint64(i % 64)
        s = np.zeros(3, dtype=np.uint64)
        sv = (1 << dim) if shift_int is None else shift_int
        for i in range(3):
            s[i] = (sv >> (64 * i)) & (2**64 - 1)
        return Domain(DOMAIN_AFFINE, dim, basis=b, shift=s)

    @staticmethod
    def coset(fp, logn, shift):
        return Domain(DOMAIN_MULT, logn, shift=shift, generator=fp.subgroup_generator(logn))


def load_library():
    global LIB_PATH
    LIB_PATH = os.environ.get("B200_LIB", LIB_PATH)   # kernel-tuning experiments only
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.b200_ctx_create.restype = C.c_void_p
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_version.restype = C.c_char_p
    lib.b200_alloc.restype = C.c_uint64
    lib.b200_alloc_bytes.restype = C.c_uint64
    lib.b200_host_alloc.restype = C.c_void_p
    lib.b200_stream.restype = C.c_void_p
    return lib


class B200:
    """One context = one device + one stream (b200_ctx)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.ctx = self.lib.b200_ctx_create(C.c_int(device))
        if not self.ctx:
            raise B200Error("b200_ctx_create failed: no CUDA device visible (the hot path has no CPU fallback)")
        self.ctx = C.c_void_p(self.ctx)
        self.launches = 0

    def close(self):
        if self.ctx:
            self.lib.b200_ctx_destroy(self.ctx)
            self.ctx = None

    def _chk(self, rc):
        if rc != 0:
            raise B200Error(f"b200 error {rc}: {self.lib.b200_last_error(self.ctx).decode()}")

    # ---- memory ----
    def alloc(self, n):
        b = self.lib.b200_alloc(self.ctx, C.c_size_t(n))
        if not b:
            raise B200Error(self.lib.b200_last_error(self.ctx).decode())
        return b

    def alloc_bytes(self, n):
        b = self.lib.b200_alloc_bytes(self.ctx, C.c_size_t(n))
        if not b:
            raise B200Error(self.lib.b200_last_error(self.ctx).decode())
        return b

    def free(self, b):
        self.lib.b200_free(self.ctx, C.c_uint64(b))

    def upload(self, arr, buf=None, offset=0):
        arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 3)
        if buf is None:
            buf = self.alloc(max(len(arr), 1))
        self._chk(self.lib.b200_upload(self.ctx, C.c_uint64(buf), C.c_size_t(offset), arr.ctypes.data_as(_u64p), C.c_size_t(len(arr))))
        self.sync()  # the source array may be garbage-collected
        return buf

    def download(self, buf, n, offset=0):
        out = np.zeros((n, 3), dtype=np.uint64)
        self._chk(self.lib.b200_download(self.ctx, out.ctypes.data_as(_u64p), C.c_uint64(buf), C.c_size_t(offset), C.c_size_t(n)))
        return out

    def download_bytes(self, buf, n, offset=0):
        out = np.zeros(n, dtype=np.uint8)
        self._chk(self.lib.b200_download_bytes(self.ctx, out.ctypes.data_as(_u8p), C.c_uint64(buf), C.c_size_t(offset), C.c_size_t(n)))
        return out.tobytes()

    def upload_bytes(self, data, buf=None):
        data = np.frombuffer(bytes(data), dtype=np.uint8)
        if buf is None:
            buf = self.alloc_bytes(max(len(data), 8))
        self._chk(self.lib.b200_upload_bytes(self.ctx, C.c_uint64(buf), data.ctypes.data_as(_u8p), C.c_size_t(len(data))))
        self.sync()
        return buf

    def gather_multi(self, bufs, idx):
        """values of several device buffers at one list of positions: (len(bufs), len(idx), 3)"""
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        out = np.zeros((len(bufs), len(idx), 3), dtype=np.uint64)
        arr = (C.c_uint64 * max(len(bufs), 1))(*bufs)
        self._chk(self.lib.b200_gather_multi(self.ctx, out.ctypes.data_as(_u64p), arr, C.c_size_t(len(bufs)),
                                             idx.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(len(idx))))
        return out

    def gather(self, buf, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        out = np.zeros((len(idx), 3), dtype=np.uint64)
        self._chk(self.lib.b200_gather(self.ctx, out.ctypes.data_as(_u64p), C.c_uint64(buf), idx.ctypes.data_as(_szp), C.c_size_t(len(idx))))
        return out

    def sync(self):
        self._chk(self.lib.b200_sync(self.ctx))

    def timer_start(self):
        self._chk(self.lib.b200_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float()
        self._chk(self.lib.b200_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    # ---- device-buffer level entry points ----
    def gf192_fft_dev(self, dom, coeffs, n_coeffs, out):
        self._chk(self.lib.b200_gf192_fft(self.ctx, C.byref(dom.c), C.c_uint64(coeffs), C.c_size_t(n_coeffs), C.c_uint64(out)))

    def gf192_ifft_dev(self, dom, evals, out):
        self._chk(self.lib.b200_gf192_ifft(self.ctx, C.byref(dom.c), C.c_uint64(evals), C.c_uint64(out)))

    def gf192_lde_dev(self, dom_h, dom_l, evals, out, batch=1):
        self._chk(self.lib.b200_gf192_lde_batch(self.ctx, C.byref(dom_h.c), C.byref(dom_l.c), C.c_uint64(evals), C.c_size_t(batch), C.c_uint64(out)))

    def gf192_lde_interpolate_dev(self, dom_h, evals, coeffs_out):
        self._chk(self.lib.b200_gf192_lde_interpolate(self.ctx, C.byref(dom_h.c), C.c_uint64(evals), C.c_uint64(coeffs_out)))

    def gf192_lde_evaluate_dev(self, dom_l, coeff_dim, coeffs, out):
        self._chk(self.lib.b200_gf192_lde_evaluate(self.ctx, C.byref(dom_l.c), C.c_uint32(coeff_dim), C.c_uint64(coeffs), C.c_uint64(out)))

    def gf192_lde_halves(self, dom_h_list, evals_list, dom_l, scale=None):
        """evaluate_L( scale * sum_i interpolate_{H_i}(evals_i) ) through b200_gf192_lde_interpolate / _evaluate (the coefficient
        form in between is opaque but linear); all H_i have L's first dim(H) basis vectors.  Returns 2^dim_L evaluations."""
        d, nl = dom_h_list[0].dim, 1 << dom_l.dim
        acc = None
        for dom, ev in zip(dom_h_list, evals_list):
            a, x = self.upload(ev), self.alloc(1 << d)
            try:
                self._chk(self.lib.b200_gf192_lde_interpolate(self.ctx, C.byref(dom.c), C.c_uint64(a), C.c_uint64(x)))
                part = self.download(x, 1 << d)
            finally:
                self.free(a); self.free(x)
            acc = part if acc is None else acc ^ part
        if scale is not None:
            acc = self.gf192_mul(acc, np.tile(np.asarray(scale, dtype=np.uint64).reshape(1, 3), (1 << d, 1)))
        b, o = self.upload(acc), self.alloc(nl)
        try:
            self._chk(self.lib.b200_gf192_lde_evaluate(self.ctx, C.byref(dom_l.c), C.c_uint32(d), C.c_uint64(b), C.c_uint64(o)))
            return self.download(o, nl)
        finally:
            self.free(b); self.free(o)

    def gf192_lde(self, dom_h, dom_l, rows):
        """fused IFFT over H + FFT over L of every row: rows (batch, 2^dim_H, 3) -> (batch, 2^dim_L, 3)"""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        rows = rows.reshape(-1, 1 << dom_h.dim, 3)
        batch, nl = rows.shape[0], 1 << dom_l.dim
        a = self.upload(rows.reshape(-1, 3)) if rows.size else self.alloc(1)
        o = self.alloc(max(batch * nl, 1))
        try:
            self.gf192_lde_dev(dom_h, dom_l, a, o, batch)
            return self.download(o, batch * nl).reshape(batch, nl, 3)
        finally:
            self.free(a); self.free(o)

    def fp_ntt_dev(self, fp, dom, coeffs, n_coeffs, out):
        self._chk(self.lib.b200_fp_ntt(self.ctx, C.byref(fp), C.byref(dom.c), C.c_uint64(coeffs), C.c_size_t(n_coeffs), C.c_uint64(out)))

    def fp_intt_dev(self, fp, dom, evals, out):
        self._chk(self.lib.b200_fp_intt(self.ctx, C.byref(fp), C.byref(dom.c), C.c_uint64(evals), C.c_uint64(out)))

    def fri_fold_dev(self, field, fp, dom, eta, x, f, out):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        self._chk(self.lib.b200_fri_fold(self.ctx, C.c_int(field), C.byref(fp) if fp is not None else None, C.byref(dom.c),
                                         C.c_uint32(eta), x.ctypes.data_as(_u64p), C.c_uint64(f), C.c_uint64(out)))

    def merkle_build_dev(self, oracle_bufs, n_elems, coset, kind, nodes, salts=None, salt_len=32, digest_len=32):
        arr = (C.c_uint64 * len(oracle_bufs))(*oracle_bufs)
        sp = None
        if salts is not None:
            self._salts_keepalive = np.frombuffer(bytes(salts), dtype=np.uint8)
            sp = self._salts_keepalive.ctypes.data_as(_u8p)
        self._chk(self.lib.b200_merkle_build(self.ctx, arr, C.c_size_t(len(oracle_bufs)), C.c_size_t(n_elems), C.c_size_t(coset),
                                             C.c_int(kind), sp, C.c_size_t(salt_len), C.c_size_t(digest_len), C.c_uint64(nodes)))

    # ---- host-buffer convenience (upload, run, download): what the C++ adapter does ----
    def gf192_fft(self, dom, coeffs):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 3)
        n = 1 << dom.dim
        a = self.upload(coeffs); o = self.alloc(n)
        try:
            self.gf192_fft_dev(dom, a, len(coeffs), o)
            return self.download(o, n)
        finally:
            self.free(a); self.free(o)

    def gf192_ifft(self, dom, evals):
        evals = np.ascontiguousarray(evals, dtype=np.uint64).reshape(-1, 3)
        n = 1 << dom.dim
        assert len(evals) == n
        a = self.upload(evals); o = self.alloc(n)
        try:
            self.gf192_ifft_dev(dom, a, o)
            return self.download(o, n)
        finally:
            self.free(a); self.free(o)

    def gf192_fft_batch(self, dom, rows):
        """rows: (batch, n_coeffs, 3) -> (batch, 2^dim, 3); one b200_gf192_fft_batch call"""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        batch, n_coeffs = rows.shape[0], rows.shape[1]
        n = 1 << dom.dim
        a = self.upload(rows.reshape(-1, 3)) if batch * n_coeffs else self.alloc(1); o = self.alloc(max(batch * n, 1))
        try:
            self._chk(self.lib.b200_gf192_fft_batch(self.ctx, C.byref(dom.c), C.c_uint64(a), C.c_size_t(n_coeffs), C.c_size_t(batch), C.c_uint64(o)))
            return self.download(o, batch * n).reshape(batch, n, 3)
        finally:
            self.free(a); self.free(o)

    def gf192_ifft_batch(self, dom, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        batch, n = rows.shape[0], 1 << dom.dim
        assert rows.shape[1] == n
        a = self.upload(rows.reshape(-1, 3)) if batch else self.alloc(1); o = self.alloc(max(batch * n, 1))
        try:
            self._chk(self.lib.b200_gf192_ifft_batch(self.ctx, C.byref(dom.c), C.c_uint64(a), C.c_size_t(batch), C.c_uint64(o)))
            return self.download(o, batch * n).reshape(batch, n, 3)
        finally:
            self.free(a); self.free(o)

    def _binop(self, fn, a, b, fp=None):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 3)
        da = self.upload(a); o = self.alloc(len(a)); db = None
        try:
            args = [self.ctx] + ([C.byref(fp)] if fp is not None else []) + [C.c_uint64(o), C.c_uint64(da)]
            if b is not None:
                db = self.upload(np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 3))
                args.append(C.c_uint64(db))
            args.append(C.c_size_t(len(a)))
            self._chk(fn(*args))
            return self.download(o, len(a))
        finally:
            self.free(da); self.free(o)
            if db:
                self.free(db)

    def gf192_mul(self, a, b):
        return self._binop(self.lib.b200_gf192_mul, a, b)

    def gf192_inv(self, a):
        return self._binop(self.lib.b200_gf192_inv, a, None)

    def fp_mul(self, fp, a, b):
        return self._binop(self.lib.b200_fp_mul, a, b, fp)

    def fp_add(self, fp, a, b):
        return self._binop(self.lib.b200_fp_add, a, b, fp)

    def fp_sub(self, fp, a, b):
        return self._binop(self.lib.b200_fp_sub, a, b, fp)

    def fp_inv(self, fp, a):
        return self._binop(self.lib.b200_fp_inv, a, None, fp)

    def fp_ntt(self, fp, dom, coeffs):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 3)
        n = 1 << dom.dim
        a = self.upload(coeffs); o = self.alloc(n)
        try:
            self.fp_ntt_dev(fp, dom, a, len(coeffs), o)
            return self.download(o, n)
        finally:
            self.free(a); self.free(o)

    def fp_intt(self, fp, dom, evals):
        evals = np.ascontiguousarray(evals, dtype=np.uint64).reshape(-1, 3)
        n = 1 << dom.dim
        a = self.upload(evals); o = self.alloc(n)
        try:
            self.fp_intt_dev(fp, dom, a, o)
            return self.download(o, n)
        finally:
            self.free(a); self.free(o)

    def fri_fold(self, field, fp, dom, eta, x, f):
        f = np.ascontiguousarray(f, dtype=np.uint64).reshape(-1, 3)
        n_out = len(f) >> eta
        a = self.upload(f); o = self.alloc(n_out)
        try:
            self.fri_fold_dev(field, fp, dom, eta, x, a, o)
            return self.download(o, n_out)
        finally:
            self.free(a); self.free(o)

    def merkle_build(self, oracles, coset, kind, salts=None, salt_len=32, digest_len=32):
        oracles = [np.ascontiguousarray(o, dtype=np.uint64).reshape(-1, 3) for o in oracles]
        n = len(oracles[0]); leaves = n // coset
        bufs = [self.upload(o) for o in oracles]
        nodes = self.alloc_bytes((2 * leaves - 1) * digest_len)
        try:
            self.merkle_build_dev(bufs, n, coset, kind, nodes, salts, salt_len, digest_len)
            return self.download_bytes(nodes, (2 * leaves - 1) * digest_len)
        finally:
            for b in bufs:
                self.free(b)
            self.free(nodes)

    # ---- element-wise passes and point evaluation (N4); numpy in, numpy out ----
    def _fpref(self, fp):
        return C.byref(fp) if fp is not None else None

    def vec_lincomb(self, field, fp, coeffs, vecs):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 3)
        n = len(vecs[0]) if len(vecs) else 0
        bufs = [self.upload(v) for v in vecs]
        o = self.alloc(max(n, 1))
        try:
            arr = (C.c_uint64 * max(len(bufs), 1))(*bufs)
            self._chk(self.lib.b200_vec_lincomb(self.ctx, C.c_int(field), self._fpref(fp), C.c_uint64(o), C.c_size_t(n),
                                                C.c_size_t(len(bufs)), coeffs.ctypes.data_as(_u64p), arr))
            return self.download(o, n)
        finally:
            for b in bufs + [o]:
                self.free(b)

    def vec_mul_add(self, field, fp, a, b=None, c=None, d=None, subtract=False):
        n = len(a)
        bufs = [self.upload(v) if v is not None else 0 for v in (a, b, c, d)]
        o = self.alloc(max(n, 1))
        try:
            self._chk(self.lib.b200_vec_mul_add(self.ctx, C.c_int(field), self._fpref(fp), C.c_uint64(o), *[C.c_uint64(x) for x in bufs],
                                                C.c_int(1 if subtract else 0), C.c_size_t(n)))
            return self.download(o, n)
        finally:
            for b_ in bufs + [o]:
                if b_:
                    self.free(b_)

    def vec_affine(self, field, fp, a, mul=None, add=None):
        n = len(a)
        da = self.upload(a); o = self.alloc(max(n, 1))
        mp = np.ascontiguousarray(mul, dtype=np.uint64).ctypes.data_as(_u64p) if mul is not None else None
        ap = np.ascontiguousarray(add, dtype=np.uint64).ctypes.data_as(_u64p) if add is not None else None
        try:
            self._chk(self.lib.b200_vec_affine(self.ctx, C.c_int(field), self._fpref(fp), C.c_uint64(o), C.c_uint64(da), mp, ap, C.c_size_t(n)))
            return self.download(o, n)
        finally:
            self.free(da); self.free(o)

    def vec_mul_table(self, field, fp, a, table, shift, mask, n=None):
        n = len(a) if a is not None else n
        da = self.upload(a) if a is not None else 0
        dt = self.upload(table)
        o = self.alloc(max(n, 1))
        try:
            self._chk(self.lib.b200_vec_mul_table(self.ctx, C.c_int(field), self._fpref(fp), C.c_uint64(o), C.c_uint64(da), C.c_uint64(dt),
                                                  C.c_uint32(shift), C.c_size_t(mask), C.c_size_t(n)))
            return self.download(o, n)
        finally:
            for b_ in (da, dt, o):
                if b_:
                    self.free(b_)

    def gf192_subset_sum_products(self, dim, tables):
        """tables: (n_factors, dim+1, 3) — offset then basis, per factor."""
        tables = np.ascontiguousarray(tables, dtype=np.uint64).reshape(-1, dim + 1, 3)
        o = self.alloc(1 << dim)
        try:
            self._chk(self.lib.b200_gf192_subset_sum_products(self.ctx, C.c_uint64(o), C.c_uint32(dim), C.c_uint32(len(tables)),
                                                              tables.ctypes.data_as(_u64p)))
            return self.download(o, 1 << dim)
        finally:
            self.free(o)

    def fp_geometric(self, fp, n, a, r, add=None):
        a = np.ascontiguousarray(a, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
        addp = np.ascontiguousarray(add, dtype=np.uint64).ctypes.data_as(_u64p) if add is not None else None
        o = self.alloc(max(n, 1))
        try:
            self._chk(self.lib.b200_fp_geometric(self.ctx, C.byref(fp), C.c_uint64(o), C.c_size_t(n), a.ctypes.data_as(_u64p),
                                                 r.ctypes.data_as(_u64p), addp))
            return self.download(o, n)
        finally:
            self.free(o)

    def poly_eval(self, field, fp, coeffs, x):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 3)
        x = np.ascontiguousarray(x, dtype=np.uint64)
        out = np.zeros(3, dtype=np.uint64)
        d = self.upload(coeffs)
        try:
            self._chk(self.lib.b200_poly_eval(self.ctx, C.c_int(field), self._fpref(fp), C.c_uint64(d), C.c_size_t(len(coeffs)),
                                              x.ctypes.data_as(_u64p), out.ctypes.data_as(_u64p)))
            return out
        finally:
            self.free(d)

    def poly_divide_sparse(self, field, fp, p, deg, lead_inv, coefs, pows):
        p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 3)
        coefs = np.ascontiguousarray(coefs, dtype=np.uint64).reshape(-1, 3)
        pows = np.ascontiguousarray(pows, dtype=np.uint64)
        lead_inv = np.ascontiguousarray(lead_inv, dtype=np.uint64)
        dp = self.upload(p); dq = self.alloc(len(p) - deg); dr = self.alloc(deg)
        try:
            self._chk(self.lib.b200_poly_divide_sparse(self.ctx, C.c_int(field), self._fpref(fp), C.c_uint64(dp), C.c_size_t(len(p)),
                                                       C.c_size_t(deg), lead_inv.ctypes.data_as(_u64p), C.c_size_t(len(coefs)),
                                                       coefs.ctypes.data_as(_u64p), pows.ctypes.data_as(_szp), C.c_uint64(dq), C.c_uint64(dr)))
            return self.download(dq, len(p) - deg), self.download(dr, deg)
        finally:
            for b_ in (dp, dq, dr):
                self.free(b_)

    def spmv_csr(self, field, fp, row_ptr, col, coeff, x):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64); col = np.ascontiguousarray(col, dtype=np.uint32)
        coeff = np.ascontiguousarray(coeff, dtype=np.uint64).reshape(-1, 3)
        n = len(row_ptr) - 1
        dx = self.upload(x); dy = self.alloc(max(n, 1))
        try:
            self._chk(self.lib.b200_spmv_csr(self.ctx, C.c_int(field), self._fpref(fp), C.c_size_t(n), row_ptr.ctypes.data_as(_u64p),
                                             col.ctypes.data_as(C.POINTER(C.c_uint32)), coeff.ctypes.data_as(_u64p), C.c_uint64(dx),
                                             C.c_size_t(len(x)), C.c_uint64(dy)))
            return self.download(dy, n)
        finally:
            self.free(dx); self.free(dy)

    def gf192_scatter_mul_add(self, y, dst, src, coeff, x, scale):
        """y[dst[k]] += scale * coeff[k] * x[src[k]] (multi_lincheck's p_alpha_ABC, basic_lincheck_aux.tcc:63-84); returns the new y"""
        y = np.ascontiguousarray(y, dtype=np.uint64).reshape(-1, 3); x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 3)
        dst = np.ascontiguousarray(dst, dtype=np.uint32); src = np.ascontiguousarray(src, dtype=np.uint32)
        coeff = np.ascontiguousarray(coeff, dtype=np.uint64).reshape(-1, 3); scale = np.ascontiguousarray(scale, dtype=np.uint64)
        dy, dx = self.upload(y), self.upload(x)
        try:
            u32p = C.POINTER(C.c_uint32)
            self._chk(self.lib.b200_gf192_scatter_mul_add(self.ctx, C.c_uint64(dy), C.c_size_t(len(y)), C.c_size_t(len(dst)),
                                                          dst.ctypes.data_as(u32p), src.ctypes.data_as(u32p), coeff.ctypes.data_as(_u64p),
                                                          C.c_uint64(dx), C.c_size_t(len(x)), scale.ctypes.data_as(_u64p)))
            return self.download(dy, len(y))
        finally:
            self.free(dy); self.free(dx)

    def copy_strided(self, src, n_out, dst_stride, src_offset, src_stride, n):
        """zero-filled array of n_out elements with out[i * dst_stride] = src[src_offset + i * src_stride], i < n"""
        src = np.ascontiguousarray(src, dtype=np.uint64).reshape(-1, 3)
        ds, do = self.upload(src), self.alloc(max(n_out, 1))
        try:
            self._chk(self.lib.b200_fill_zero(self.ctx, C.c_uint64(do), C.c_size_t(0), C.c_size_t(n_out)))
            self._chk(self.lib.b200_copy_strided(self.ctx, C.c_uint64(do), C.c_size_t(dst_stride), C.c_uint64(ds), C.c_size_t(src_offset),
                                                 C.c_size_t(src_stride), C.c_size_t(n)))
            return self.download(do, n_out)
        finally:
            self.free(ds); self.free(do)

    def gather_scatter_fill(self, table, idx, map_, x_scatter, n_out, fill_value, fill_from):
        """exercises b200_gather_dev / b200_scatter_dev / b200_fill on device arrays; returns (gathered, scattered-and-filled):
        gathered[k] = table[map[idx[k]]] (map may be None); out = zeros(n_out), out[map[c]] = x_scatter[c], out[fill_from:] = fill_value"""
        table = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        dt, dg = self.upload(table), self.alloc(max(len(idx), 1))
        di = self.alloc_bytes(max(idx.nbytes, 4)); self.upload_bytes(idx.view(np.uint8), di)
        dm = 0
        if map_ is not None:
            map_ = np.ascontiguousarray(map_, dtype=np.uint32)
            dm = self.alloc_bytes(max(map_.nbytes, 4)); self.upload_bytes(map_.view(np.uint8), dm)
        x_scatter = np.ascontiguousarray(x_scatter, dtype=np.uint64).reshape(-1, 3)
        dx, do = self.upload(x_scatter), self.alloc(max(n_out, 1))
        try:
            self._chk(self.lib.b200_gather_dev(self.ctx, C.c_uint64(dg), C.c_uint64(dt), C.c_uint64(di), C.c_uint64(dm), C.c_size_t(len(idx))))
            self._chk(self.lib.b200_fill_zero(self.ctx, C.c_uint64(do), C.c_size_t(0), C.c_size_t(n_out)))
            if dm:
                self._chk(self.lib.b200_scatter_dev(self.ctx, C.c_uint64(do), C.c_uint64(dx), C.c_uint64(dm), C.c_size_t(len(x_scatter))))
            fv = np.ascontiguousarray(fill_value, dtype=np.uint64).reshape(3)
            self._chk(self.lib.b200_fill(self.ctx, C.c_uint64(do), C.c_size_t(fill_from), fv.ctypes.data_as(_u64p), C.c_size_t(n_out - fill_from)))
            return self.download(dg, len(idx)), self.download(do, n_out)
        finally:
            for b in (dt, dg, dx, do):
                self.free(b)
            self.free(di)
            if dm:
                self.free(dm)

    def csr_row_indices(self, row_ptr):
        """b200_csr_row_indices_dev: the explicit row of every entry from uploaded row pointers (u64, n_rows + 1)"""
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        n_rows, nnz = len(row_ptr) - 1, int(row_ptr[-1])
        dp = self.alloc_bytes(row_ptr.nbytes); self.upload_bytes(row_ptr.view(np.uint8), dp)
        do = self.alloc_bytes(max(4 * nnz, 4))
        try:
            self._chk(self.lib.b200_csr_row_indices_dev(self.ctx, C.c_size_t(n_rows), C.c_uint64(dp), C.c_uint64(do)))
            return np.frombuffer(self.download_bytes(do, 4 * nnz), dtype=np.uint32) if nnz else np.zeros(0, dtype=np.uint32)
        finally:
            self.free(dp); self.free(do)

    def csr_transpose(self, n_cols, col, row, coeff):
        """b200_csr_transpose_dev on uploaded entry lists; returns (col_ptr u64[n_cols+1], src u32[nnz], coeff (nnz,3))"""
        col = np.ascontiguousarray(col, dtype=np.uint32); row = np.ascontiguousarray(row, dtype=np.uint32)
        coeff = np.ascontiguousarray(coeff, dtype=np.uint64).reshape(-1, 3)
        nnz = len(col)
        dc = self.alloc_bytes(max(col.nbytes, 4)); self.upload_bytes(col.view(np.uint8), dc)
        dr = self.alloc_bytes(max(row.nbytes, 4)); self.upload_bytes(row.view(np.uint8), dr)
        dv, dvo = self.upload(coeff) if nnz else self.alloc(1), self.alloc(max(nnz, 1))
        dp, ds = self.alloc_bytes(8 * (n_cols + 1)), self.alloc_bytes(max(4 * nnz, 4))
        try:
            self._chk(self.lib.b200_csr_transpose_dev(self.ctx, C.c_size_t(n_cols), C.c_size_t(nnz), C.c_uint64(dc), C.c_uint64(dr), C.c_uint64(dv),
                                                      C.c_uint64(dp), C.c_uint64(ds), C.c_uint64(dvo)))
            col_ptr = np.frombuffer(self.download_bytes(dp, 8 * (n_cols + 1)), dtype=np.uint64)
            src = np.frombuffer(self.download_bytes(ds, 4 * nnz), dtype=np.uint32) if nnz else np.zeros(0, dtype=np.uint32)
            return col_ptr, src, self.download(dvo, nnz)
        finally:
            for b in (dc, dr, dv, dvo, dp, ds):
                self.free(b)

    def vec_rows_dot(self, field, fp, x, y, split):
        """x, y: (rows, n, 3) -> out[a] = sum_{r<split} x[r,a]*y[r,a] - sum_{r>=split} x[r,a]*y[r,a]"""
        x = np.ascontiguousarray(x, dtype=np.uint64); y = np.ascontiguousarray(y, dtype=np.uint64)
        rows, n = x.shape[0], x.shape[1]
        dx = self.upload(x.reshape(-1, 3)) if rows * n else self.alloc(1)
        dy = self.upload(y.reshape(-1, 3)) if rows * n else self.alloc(1)
        o = self.alloc(max(n, 1))
        try:
            self._chk(self.lib.b200_vec_rows_dot(self.ctx, C.c_int(field), self._fpref(fp), C.c_uint64(o), C.c_uint64(dx), C.c_uint64(dy),
                                                 C.c_size_t(n), C.c_size_t(rows), C.c_size_t(split)))
            return self.download(o, n)
        finally:
            self.free(dx); self.free(dy); self.free(o)

    def pow_solve(self, challenge, pow_bitlen, pow_upperbound=0):
        out = (C.c_uint8 * 32)()
        self._chk(self.lib.b200_pow_solve(self.ctx, bytes(challenge), C.c_size_t(32), C.c_uint32(pow_bitlen), C.c_uint64(pow_upperbound), out))
        return bytes(out)

    def squeeze_field_elements(self, field, fp, root_plus_index, n):
        """blake2b_FieldT_randomness_extractor (bcs/hashing/blake2b.tcc:231-256): n elements from root || u64 index."""
        msg = bytes(root_plus_index)
        out = self.alloc(max(n, 1))
        try:
            self._chk(self.lib.b200_squeeze_field_elements(self.ctx, C.c_int(field), self._fpref(fp), msg, C.c_size_t(len(msg)),
                                                           C.c_size_t(n), C.c_uint64(out)))
            return self.download(out, n)
        finally:
            self.free(out)

    def blake2b_batch(self, msgs, msg_len, digest_len=32):
        n = len(msgs) // msg_len
        i = self.upload_bytes(msgs); o = self.alloc_bytes(n * digest_len)
        try:
            self._chk(self.lib.b200_blake2b_batch(self.ctx, C.c_uint64(i), C.c_size_t(msg_len), C.c_size_t(n), C.c_size_t(digest_len), C.c_uint64(o)))
            return self.download_bytes(o, n * digest_len)
        finally:
            self.free(i); self.free(o)
