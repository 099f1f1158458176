"""GPU parity tests for the element-wise passes between the LDE / Merkle / FRI calls and
polynomial evaluation at a point (N4), through the C ABI.  The checker restates the reference's scalar loops with the
oracle's field arithmetic (oracle/gf192.c, oracle/fp.c); bit-exact."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import el, elems, rand_elems, standard_basis

pytestmark = pytest.mark.gpu

GF, FP = 1, 2


@pytest.fixture(scope="module")
def dev():
    import libiop_b200 as L
    d = L.B200(0)
    yield d
    d.close()


@pytest.fixture(scope="module")
def fp():
    import libiop_b200 as L
    return L.FpParams.edwards_fr()


def bcast(e, n):
    return np.broadcast_to(np.asarray(e, dtype=np.uint64).reshape(1, 3), (n, 3)).copy()


class GfOps:
    def __init__(self, o): self.o = o
    def mul(self, a, b): return self.o.gf192_mul(a, b)
    def add(self, a, b): return a ^ b
    def sub(self, a, b): return a ^ b
    def rand(self, rng, n): return rand_elems(rng, n)


def _to_int(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 3).astype(object)
    return a[:, 0] + (a[:, 1] << 64) + (a[:, 2] << 128)


def _from_int(v):
    m = (1 << 64) - 1
    return np.stack([(v & m), ((v >> 64) & m), (v >> 128)], axis=1).astype(np.uint64)


class FpOps:
    """vector add/sub on the Montgomery representatives (Fp_model::operator+=/-=, fp.tcc:323-532: plain modular add/sub)"""
    def __init__(self, o): self.o, self.p = o, o.fp_modulus()
    def mul(self, a, b): return self.o.fp_mul(a, b)
    def add(self, a, b): return _from_int((_to_int(a) + _to_int(b)) % self.p)
    def sub(self, a, b): return _from_int((_to_int(a) - _to_int(b)) % self.p)
    def rand(self, rng, n): return self.o.rand_fp(rng, n)


def ops_for(field, oracle):
    return GfOps(oracle) if field == GF else FpOps(oracle)


@pytest.mark.parametrize("field", [GF, FP])
@pytest.mark.parametrize("n,k", [(1, 1), (1000, 3), (4097, 4), (70000, 13), (333, 0)])
def test_vec_lincomb(dev, oracle, fp, field, n, k):
    """multi_lincheck f_combined_Mz (basic_lincheck_aux.tcc:117-123) / combined_LDT maximal oracles (ldt_reducer_aux.tcc:52-66)."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(n * 31 + k)
    vecs = [F.rand(rng, n) for _ in range(k)]
    coeffs = F.rand(rng, max(k, 1))[:k]
    want = np.zeros((n, 3), dtype=np.uint64)
    for j in range(k):
        want = F.add(want, F.mul(bcast(coeffs[j], n), vecs[j]))
    fpref = C.byref(fp) if field == FP else None
    if k == 0:                                # empty combination = zero vector; needs the explicit length of the C call
        o = dev.upload(F.rand(rng, n))
        dev._chk(dev.lib.b200_vec_lincomb(dev.ctx, C.c_int(field), fpref, C.c_uint64(o), C.c_size_t(n), C.c_size_t(0), None, None))
        got = dev.download(o, n)
        dev.free(o)
    else:
        got = dev.vec_lincomb(field, fp if field == FP else None, coeffs, vecs)
    assert (got == want).all()


@pytest.mark.parametrize("field", [GF, FP])
@pytest.mark.parametrize("form", ["ab", "ab+c", "ab-cd", "a-cd", "a+c"])
def test_vec_mul_add(dev, oracle, fp, field, form):
    """fz = fw*Z + f1v (r1cs_rs_iop.tcc:216-221); lincheck combined*p1 - fz*p2 (basic_lincheck_aux.tcc:131-137);
    Az*Bz - Cz (rowcheck.tcc:61-82)."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(7)
    n = 50001
    a, b, c, d = (F.rand(rng, n) for _ in range(4))
    fpp = fp if field == FP else None
    if form == "ab":
        got, want = dev.vec_mul_add(field, fpp, a, b), F.mul(a, b)
    elif form == "ab+c":
        got, want = dev.vec_mul_add(field, fpp, a, b, c), F.add(F.mul(a, b), c)
    elif form == "ab-cd":
        got, want = dev.vec_mul_add(field, fpp, a, b, c, d, subtract=True), F.sub(F.mul(a, b), F.mul(c, d))
    elif form == "a-cd":
        got, want = dev.vec_mul_add(field, fpp, a, None, c, d, subtract=True), F.sub(a, F.mul(c, d))
    else:
        got, want = dev.vec_mul_add(field, fpp, a, None, c), F.add(a, c)
    assert (got == want).all()


@pytest.mark.parametrize("field", [GF, FP])
@pytest.mark.parametrize("shift,mask", [(10, 7), (0, 7), (3, 0xFFFF)])
def test_vec_mul_table(dev, oracle, fp, field, shift, mask):
    """rowcheck's Z_H^-1 per coset of H in L (rowcheck.tcc:36-84): additive index = pos / |H|, multiplicative = pos mod #cosets."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(11)
    n = 1 << 13
    a = F.rand(rng, n)
    table = F.rand(rng, min(mask, (n - 1) >> shift) + 1)
    idx = (np.arange(n) >> shift) & mask
    fpp = fp if field == FP else None
    assert (dev.vec_mul_table(field, fpp, a, table, shift, mask) == F.mul(a, table[idx])).all()
    assert (dev.vec_mul_table(field, fpp, None, table, shift, mask, n=n) == table[idx]).all()


@pytest.mark.parametrize("dim,k", [(1, 1), (8, 1), (13, 1), (13, 3), (10, 20)])
def test_gf192_subset_sum_products(dev, oracle, dim, k):
    """all_subset_sums (utils.tcc:8-30) and subspace_element_powers (exponentiation.tcc:22-54)."""
    rng = np.random.default_rng(dim * 100 + k)
    tables = rand_elems(rng, k * (dim + 1)).reshape(k, dim + 1, 3)
    n = 1 << dim
    idx = np.arange(n)
    want = None
    for f in range(k):
        v = bcast(tables[f, 0], n)
        for j in range(dim):
            v[(idx >> j) & 1 == 1] ^= tables[f, 1 + j]
        want = v if want is None else oracle.gf192_mul(want, v)
    assert (dev.gf192_subset_sum_products(dim, tables) == want).all()


def test_subspace_element_powers_matches_repeated_multiplication(dev, oracle):
    """x^e over an affine subspace via one factor per set bit of e == e-fold product of the elements."""
    dim, e = 9, 11
    basis, shift = standard_basis(dim), el(1 << dim)
    n = 1 << dim
    elems_tab = np.concatenate([shift.reshape(1, 3), basis]).reshape(1, dim + 1, 3)
    xs = dev.gf192_subset_sum_products(dim, elems_tab)
    want = xs.copy()
    for _ in range(e - 1):
        want = oracle.gf192_mul(want, xs)

    def frob(t, times):
        for _ in range(times):
            t = oracle.gf192_mul(t, t)
        return t
    tabs = np.stack([frob(elems_tab[0].copy(), b) for b in range(4) if (e >> b) & 1])
    assert (dev.gf192_subset_sum_products(dim, tabs) == want).all()


@pytest.mark.parametrize("n", [1, 15, 16, 17, 5000, 1 << 15])
def test_fp_geometric(dev, oracle, fp, n):
    """coset_element_powers (exponentiation.tcc:56-72); vanishing polynomial over a coset (vanishing_polynomial.tcc:115-165)."""
    rng = np.random.default_rng(n)
    a, r, c = oracle.rand_fp(rng, 3)
    want = np.zeros((n, 3), dtype=np.uint64)
    cur = a.reshape(1, 3)
    for i in range(n):                       # sequential like the reference
        want[i] = cur[0]
        cur = oracle.fp_mul(cur, r.reshape(1, 3))
    assert (dev.fp_geometric(fp, n, a, r) == want).all()
    assert (dev.fp_geometric(fp, n, a, r, add=c) == FpOps(oracle).add(want, bcast(c, n))).all()


@pytest.mark.parametrize("field", [GF, FP])
@pytest.mark.parametrize("n", [1, 7, 8, 2047, 2048, 2049, 100000, (1 << 22) + 5])
def test_poly_eval(dev, oracle, fp, field, n):
    """polynomial::evaluation_at_point (polynomial.tcc:98-109): Horner."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(n)
    coeffs = F.rand(rng, n)
    x = F.rand(rng, 1)
    got = dev.poly_eval(field, fp if field == FP else None, coeffs, x[0])
    # checker: blocked Horner with the oracle's vector multiply (log-depth, same value as the scalar recurrence)
    vals, y = coeffs, x
    while len(vals) > 1:
        if len(vals) % 2:
            vals = np.concatenate([vals, np.zeros((1, 3), dtype=np.uint64)])
        vals = F.add(vals[0::2], F.mul(vals[1::2], np.repeat(y, len(vals) // 2, axis=0)))
        y = F.mul(y, y)
    assert (got == vals[0]).all()
    if n <= 2049:                            # and the literal scalar recurrence at small sizes
        acc = np.zeros((1, 3), dtype=np.uint64)
        for i in range(n - 1, -1, -1):
            acc = F.add(F.mul(acc, x), coeffs[i:i + 1])
        assert (got == acc[0]).all()


def _inv(F, field, oracle, x):
    return oracle.gf192_inv(x.reshape(1, 3))[0] if field == GF else oracle.fp_inv(x)


@pytest.mark.parametrize("field", [GF, FP])
@pytest.mark.parametrize("log_deg,n_p,linearized", [(3, 20, True), (4, 17, True), (5, 64, True), (6, 200, True), (4, 100, False), (5, 33, False)])
def test_poly_divide_sparse(dev, oracle, fp, field, log_deg, n_p, linearized):
    """polynomial_over_linearized_polynomial (linearized_polynomial.tcc:238-291) and
    polynomial_over_multiplicative_vanishing_polynomial (vanishing_polynomial.tcc:312-358): the checker is the reference's
    sequential coefficient-clearing loop; P = Q*Z + R is verified by evaluating both sides at a random point as well."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(log_deg * 1000 + n_p)
    deg = 1 << log_deg
    P = F.rand(rng, n_p)
    if linearized:                         # Z = lead*y^deg + sum_j Z[j] y^(2^(j-1)) + Z[0]
        pows = [0] + [1 << k for k in range(log_deg)]
    else:                                  # Z = y^deg + Z[0]
        pows = [0]
    coefs = F.rand(rng, len(pows))
    lead = F.rand(rng, 1)[0]
    lead_inv = _inv(F, field, oracle, lead)
    # sequential restatement
    quo = P[deg:].copy(); rem = P[:deg].copy()
    for i in range(len(quo) - 1, -1, -1):
        tw = F.mul(quo[i:i + 1], lead_inv.reshape(1, 3))
        quo[i] = tw[0]
        for c, pw in zip(coefs, pows):
            t = F.mul(tw, c.reshape(1, 3))
            if i + pw < deg:
                rem[i + pw] = F.sub(rem[i + pw:i + pw + 1], t)[0]
            else:
                quo[i + pw - deg] = F.sub(quo[i + pw - deg:i + pw - deg + 1], t)[0]
    q, r = dev.poly_divide_sparse(field, fp if field == FP else None, P, deg, lead_inv, coefs, pows)
    assert (q == quo).all() and (r == rem).all()


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1001, 65537])
def test_gf192_inv_ragged_and_zeros(dev, oracle, n):
    """b200_gf192_inv shares one inversion per run of 8 elements (Montgomery's trick, algebra/utils.tcc:38-103): ragged
    tails and zero elements (inverse taken as zero, like the reference's has_zeroes path) must not disturb neighbours."""
    rng = np.random.default_rng(n)
    a = rand_elems(rng, n)
    zeros = rng.integers(0, n, size=max(1, n // 5))
    a[zeros] = 0
    if n > 20:
        a[8:16] = 0                      # a whole run of zeros
    got = dev.gf192_inv(a)
    nz = (a != 0).any(axis=1)
    want = np.zeros_like(a)
    if nz.any():
        want[nz] = oracle.gf192_inv(a[nz])
    assert (got == want).all()
    one = np.zeros((int(nz.sum()), 3), dtype=np.uint64); one[:, 0] = 1
    assert (oracle.gf192_mul(got[nz], a[nz]) == one).all()


@pytest.mark.parametrize("field", [GF, FP])
def test_vec_affine(dev, oracle, fp, field):
    """out = a*mul + add with scalars: f - claimed_eval (boundary_constraint.tcc:56-60), k * v (utils.tcc:38-69)."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(3)
    n = 30011
    a = F.rand(rng, n)
    m, c = F.rand(rng, 2)
    fpp = fp if field == FP else None
    assert (dev.vec_affine(field, fpp, a, m, c) == F.add(F.mul(a, bcast(m, n)), bcast(c, n))).all()
    assert (dev.vec_affine(field, fpp, a, m, None) == F.mul(a, bcast(m, n))).all()
    assert (dev.vec_affine(field, fpp, a, None, c) == F.add(a, bcast(c, n))).all()
    assert (dev.vec_affine(field, fpp, a, None, None) == a).all()


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000, 40001])
def test_fp_inv_ragged_and_zeros(dev, oracle, fp, n):
    """b200_fp_inv shares one Fermat inversion per run of 16 elements (Montgomery's trick, algebra/utils.tcc:38-103)."""
    rng = np.random.default_rng(n + 5)
    a = oracle.rand_fp(rng, n)
    a[rng.integers(0, n, size=max(1, n // 7))] = 0
    if n > 40:
        a[16:32] = 0
    got = dev.fp_inv(fp, a)
    nz = (a != 0).any(axis=1)
    assert (got[~nz] == 0).all()
    one = bcast(np.array(list(fp.one_mont), dtype=np.uint64), int(nz.sum()))
    if nz.any():
        assert (oracle.fp_mul(got[nz], a[nz]) == one).all()
        k = min(50, int(nz.sum()))
        idx = np.nonzero(nz)[0][:k]
        for i in idx:
            assert (got[i] == oracle.fp_inv(a[i])).all()


@pytest.mark.parametrize("field", [GF, FP])
def test_spmv_csr(dev, oracle, fp, field):
    """create_Az_Bz_Cz_from_variable_assignment (relations/r1cs.tcc:236-266): rows with 0..4 terms, repeated columns."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(99)
    n_rows, n_x = 5000, 777
    lens = rng.integers(0, 5, size=n_rows)
    lens[:3] = [0, 4, 0]
    row_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    nnz = int(row_ptr[-1])
    col = rng.integers(0, n_x, size=nnz).astype(np.uint32)
    coeff = F.rand(rng, nnz)
    x = F.rand(rng, n_x)
    prod = F.mul(coeff, x[col])
    want = np.zeros((n_rows, 3), dtype=np.uint64)
    for j in range(4):                                     # j-th term of every row that has one
        rows = np.nonzero(lens > j)[0]
        idx = (row_ptr[rows] + j).astype(np.int64)
        want[rows] = F.add(want[rows], prod[idx])
    got = dev.spmv_csr(field, fp if field == FP else None, row_ptr, col, coeff, x)
    assert (got == want).all()
    with pytest.raises(Exception):
        bad = col.copy(); bad[5] = n_x
        dev.spmv_csr(field, fp if field == FP else None, row_ptr, bad, coeff, x)


@pytest.mark.parametrize("field", [GF, FP])
@pytest.mark.parametrize("rows,split,n", [(1, 1, 5), (7, 3, 1000), (6, 0, 129), (5, 5, 128), (0, 0, 9)])
def test_vec_rows_dot(dev, oracle, fp, field, rows, split, n):
    """sum over the first `split` rows of x_r * y_r minus the sum over the others: the response evaluations of Ligero's
    interleaved lincheck (interleaved_lincheck_ot.tcc:283-330), restated with the oracle's field arithmetic."""
    F = ops_for(field, oracle)
    rng = np.random.default_rng(31 * rows + n)
    x = F.rand(rng, rows * n).reshape(rows, n, 3); y = F.rand(rng, rows * n).reshape(rows, n, 3)
    want = np.zeros((n, 3), dtype=np.uint64)
    for r in range(rows):
        prod = F.mul(x[r], y[r])
        want = F.add(want, prod) if r < split else F.sub(want, prod)
    got = dev.vec_rows_dot(field, fp if field == FP else None, x, y, split)
    assert (got == want).all()
    with pytest.raises(Exception):
        dev.vec_rows_dot(field, fp if field == FP else None, x, y, rows + 1)


# ------------------------------------------------------------------ N3: transposed sparse product of multi_lincheck (p_alpha_ABC)
@pytest.mark.parametrize("n_y,n_x,nnz", [(1 << 10, 1 << 10, 3000), (1 << 12, 1 << 11, 1), (64, 64, 0), (1 << 14, 1 << 14, 50000)])
def test_gf192_scatter_mul_add(dev, oracle, n_y, n_x, nnz):
    """b200_gf192_scatter_mul_add: y[dst[k]] += scale * coeff[k] * x[src[k]] — the loop of multi_lincheck_virtual_oracle::set_challenge
    (protocols/encoded/lincheck/basic_lincheck_aux.tcc:63-84) over the flattened entries of one matrix; heavy collisions on a
    few destinations included (atomic XOR accumulation)."""
    rng = np.random.default_rng(n_y + nnz)
    y, x = rand_elems(rng, n_y), rand_elems(rng, n_x)
    dst = rng.integers(0, n_y, size=nnz, dtype=np.uint32); src = rng.integers(0, n_x, size=nnz, dtype=np.uint32)
    if nnz > 100:
        dst[:nnz // 4] = dst[0]                       # a quarter of the entries hit one destination
    coeff, scale = rand_elems(rng, nnz), rand_elems(rng, 1)[0]
    want = y.copy()
    if nnz:
        prod = oracle.gf192_mul(oracle.gf192_mul(coeff, np.repeat(scale.reshape(1, 3), nnz, 0)), x[src])
        for k in range(nnz):
            want[dst[k]] ^= prod[k]
    assert (dev.gf192_scatter_mul_add(y, dst, src, coeff, x, scale) == want).all()


def test_gf192_scatter_rejects_out_of_range_entries(dev):
    import libiop_b200 as L
    rng = np.random.default_rng(3)
    y, x = rand_elems(rng, 16), rand_elems(rng, 16)
    with pytest.raises(L.B200Error):
        dev.gf192_scatter_mul_add(y, [16], [0], rand_elems(rng, 1), x, rand_elems(rng, 1)[0])
    with pytest.raises(L.B200Error):
        dev.gf192_scatter_mul_add(y, [0], [99], rand_elems(rng, 1), x, rand_elems(rng, 1)[0])


@pytest.mark.parametrize("n_src,n_out,ds,so,ss,n", [(1 << 12, 1 << 9, 1, 0, 8, 1 << 9), (1 << 9, 1 << 12, 8, 0, 1, 1 << 9), (100, 40, 2, 3, 5, 19), (8, 8, 1, 0, 1, 0)])
def test_copy_strided(dev, n_src, n_out, ds, so, ss, n):
    """b200_copy_strided: sub-coset gather (fft.tcc:446-455) and placement of a sub-coset inside the larger one (subgroup.tcc:150-162)"""
    rng = np.random.default_rng(n_src + n)
    src = rand_elems(rng, n_src)
    want = elems(n_out)
    for i in range(n):
        want[i * ds] = src[so + i * ss]
    assert (dev.copy_strided(src, n_out, ds, so, ss, n) == want).all()


@pytest.mark.parametrize("n_table,n_idx,with_map", [(1 << 10, 1 << 12, True), (1 << 10, 1 << 12, False), (37, 5, True), (16, 0, True)])
def test_gather_scatter_fill_on_device(dev, n_table, n_idx, with_map):
    """b200_gather_dev / b200_scatter_dev / b200_fill: the index-driven copies behind matrix_indexer::compute_oracles_over_K
    (fractal_indexer.tcc:52-123) and compute_p_alpha_M's placement (lincheck/common.tcc:22-27); pure data movement, bit-exact"""
    rng = np.random.default_rng(n_table * 7 + n_idx)
    table = rand_elems(rng, n_table)
    idx = rng.integers(0, n_table, size=n_idx, dtype=np.uint32)
    perm = rng.permutation(n_table).astype(np.uint32)
    x = rand_elems(rng, n_table)
    n_out = n_table + 9
    fill = rand_elems(rng, 1)[0]
    got_g, got_s = dev.gather_scatter_fill(table, idx, perm if with_map else None, x, n_out, fill, n_table + 2)
    want_g = table[perm[idx]] if with_map else table[idx]
    assert got_g.shape == (n_idx, 3) and (got_g == want_g).all()
    want_s = elems(n_out)
    if with_map:
        want_s[perm] = x
    want_s[n_table + 2:] = fill
    assert (got_s == want_s).all()


@pytest.mark.parametrize("n_bufs,n,k", [(1, 100, 7), (37, 1 << 12, 211), (1100, 256, 64), (3, 50, 0)])
def test_gather_multi(dev, n_bufs, n, k):
    """b200_gather_multi: every oracle of a Merkle tree at every queried position with one launch (bcs_prover.tcc:197-207)"""
    rng = np.random.default_rng(n_bufs * 1000 + k)
    data = [rand_elems(rng, n) for _ in range(n_bufs)]
    bufs = [dev.upload(d) for d in data]
    idx = rng.integers(0, n, size=k)
    try:
        got = dev.gather_multi(bufs, idx)
    finally:
        for b in bufs:
            dev.free(b)
    assert got.shape == (n_bufs, k, 3)
    for s in range(n_bufs):
        assert (got[s] == data[s][idx]).all()


@pytest.mark.parametrize("n_rows,max_terms", [(1 << 12, 1), (5000, 4), (100, 0), (1, 7)])
def test_csr_row_indices_on_device(dev, n_rows, max_terms):
    """b200_csr_row_indices_dev: row_out[k] = i for k in [row_ptr[i], row_ptr[i+1]), rows with no entries included"""
    rng = np.random.default_rng(n_rows + max_terms)
    counts = rng.integers(0, max_terms + 1, size=n_rows)
    row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    got = dev.csr_row_indices(row_ptr)
    assert (got == np.repeat(np.arange(n_rows, dtype=np.uint32), counts)).all()


@pytest.mark.parametrize("n_rows,n_cols,nnz", [(1 << 12, 1 << 12, 3 << 12), (100, 37, 1000), (5000, 70000, 20000), (8, 8, 0)])
def test_csr_transpose_on_device(dev, n_rows, n_cols, nnz):
    """b200_csr_transpose_dev: the entries of an (row, col, coeff) list regrouped by column (counting sort with atomic cursors).
    The order inside a column is unspecified, so each column is compared as a multiset."""
    rng = np.random.default_rng(n_rows + n_cols + nnz)
    row = np.sort(rng.integers(0, n_rows, size=nnz, dtype=np.uint32))
    col = rng.integers(0, n_cols, size=nnz, dtype=np.uint32)
    coeff = rand_elems(rng, nnz)
    col_ptr, src, out = dev.csr_transpose(n_cols, col, row, coeff)
    counts = np.bincount(col, minlength=n_cols)
    assert (col_ptr == np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)).all()
    order = np.argsort(col, kind="stable")
    want = np.concatenate([row[order].reshape(-1, 1).astype(np.uint64), coeff[order]], axis=1)
    got = np.concatenate([src.reshape(-1, 1).astype(np.uint64), out], axis=1)
    for c in np.unique(col):
        lo, hi = int(col_ptr[c]), int(col_ptr[c + 1])
        assert sorted(map(tuple, got[lo:hi])) == sorted(map(tuple, want[lo:hi]))


@pytest.mark.parametrize("dim", [1, 5, 12])
def test_powers_by_subset_sum_products(dev, oracle, dim):
    """alpha^i for i < 2^dim as a product of dim index-bit-affine factors (libiop_b200/host/b200_elementwise.hpp powers_of):
    factor j is 1 when bit j of i is clear and alpha^(2^j) when it is set."""
    rng = np.random.default_rng(dim)
    alpha = rand_elems(rng, 1)[0]
    tables = np.zeros((dim, dim + 1, 3), dtype=np.uint64)
    sq = alpha.copy()
    for j in range(dim):
        tables[j, 0] = el(1)
        tables[j, 1 + j] = sq ^ el(1)
        sq = oracle.gf192_mul(sq.reshape(1, 3), sq.reshape(1, 3))[0]
    got = dev.gf192_subset_sum_products(dim, tables)
    want = elems(1 << dim); want[0] = el(1)
    for i in range(1, 1 << dim):
        want[i] = oracle.gf192_mul(want[i - 1].reshape(1, 3), alpha.reshape(1, 3))[0]
    assert (got == want).all()
