"""GPU parity tests (run on the B200 box): the CUDA hot path, called through the C ABI
(include/b200_iop.h via libiop_b200/capi.py), against the oracle and the reference-generated
golden fixtures.  Bit-exact: everything is integer / carry-less arithmetic."""
import hashlib

import numpy as np
import pytest

from oracle_lib import el, elems, rand_elems, standard_basis

pytestmark = pytest.mark.gpu

AFF, MUL = 1, 2


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


def dom_affine(basis, shift):
    import libiop_b200 as L
    return L.Domain.affine(basis, shift)


def dom_coset(fp, logn, shift):
    import libiop_b200 as L
    return L.Domain.coset(fp, logn, shift)


# ------------------------------------------------------------------ field arithmetic (T0)
def test_gf192_field_ops(dev, oracle, golden):
    g = golden
    assert (dev.gf192_mul(g["gf_a"], g["gf_b"]) == g["gf_mul"]).all()
    assert (dev.gf192_inv(g["gf_a"][:8]) == g["gf_inv"]).all()
    rng = np.random.default_rng(1)
    a, b = rand_elems(rng, 1000000), rand_elems(rng, 1000000)     # 10^6 random pairs
    a[:4] = 0xFFFFFFFFFFFFFFFF; b[:2] = 0xFFFFFFFFFFFFFFFF
    assert (dev.gf192_mul(a, b) == oracle.gf192_mul(a, b)).all()
    assert (dev.gf192_inv(a[:2000]) == oracle.gf192_inv(a[:2000])).all()


def test_fp_field_ops(dev, oracle, golden, fp):
    g = golden
    assert (dev.fp_mul(fp, g["fp_a"], g["fp_b"]) == g["fp_mul"]).all()
    assert (dev.fp_add(fp, g["fp_a"], g["fp_b"]) == g["fp_add"]).all()
    assert (dev.fp_sub(fp, g["fp_a"], g["fp_b"]) == g["fp_sub"]).all()
    assert (dev.fp_inv(fp, g["fp_a"][:8]) == g["fp_inv"]).all()
    rng = np.random.default_rng(2)
    a, b = oracle.rand_fp(rng, 20000), oracle.rand_fp(rng, 20000)
    assert (dev.fp_mul(fp, a, b) == oracle.fp_mul(a, b)).all()
    assert list(fp.modulus) == list(oracle.fp.mod) and list(fp.r2) == list(oracle.fp.r2)
    assert fp.inv == oracle.fp.inv and list(fp.one_mont) == list(oracle.fp.one)


# ------------------------------------------------------------------ additive FFT (T1, T2)
@pytest.mark.parametrize("tag,m", [("std", 10), ("std_full", 8), ("rnd", 7)])
def test_additive_fft_golden(dev, golden, tag, m):
    g = golden
    d = dom_affine(g[f"afft_{tag}_basis"], g[f"afft_{tag}_shift"])
    assert (dev.gf192_fft(d, g[f"afft_{tag}_coeffs"]) == g[f"afft_{tag}_evals"]).all()
    assert (dev.gf192_ifft(d, g[f"aifft_{tag}_in"]) == g[f"aifft_{tag}_out"]).all()


@pytest.mark.parametrize("m", list(range(1, 12)))
def test_additive_fft_vs_naive(dev, oracle, m):
    """test_fft.cpp:27-52: FFT == naive Horner evaluation, random affine subspaces, m <= 11."""
    rng = np.random.default_rng(100 + m)
    basis, shift = rand_elems(rng, m), rand_elems(rng, 1)[0]
    c = rand_elems(rng, 1 << m)
    want = oracle.naive_fft_gf192(c, m, basis, shift) if m <= 9 else oracle.additive_fft(c, m, basis, shift)
    d = dom_affine(basis, shift)
    assert (dev.gf192_fft(d, c) == want).all()
    assert (dev.gf192_ifft(d, want) == c).all()


@pytest.mark.parametrize("m,nc", [(12, 4096), (13, 1000), (14, 1 << 11), (16, 1 << 13), (16, 1 << 16), (17, 3), (15, 1), (18, (1 << 15) + 1)])
def test_additive_fft_vs_oracle(dev, oracle, m, nc):
    """Aurora-shaped domains (standard basis, shift x^m) incl. ragged / tiny coefficient counts."""
    rng = np.random.default_rng(200 + m + nc)
    basis, shift = standard_basis(m), el(1 << m)
    c = rand_elems(rng, nc)
    assert (dev.gf192_fft(dom_affine(basis, shift), c) == oracle.additive_fft(c, m, basis, shift)).all()


@pytest.mark.parametrize("m", [12, 13, 15, 17])
def test_additive_ifft_vs_oracle(dev, oracle, m):
    rng = np.random.default_rng(300 + m)
    basis, shift = (standard_basis(m), el(1 << m)) if m % 2 else (rand_elems(rng, m), rand_elems(rng, 1)[0])
    e = rand_elems(rng, 1 << m)
    assert (dev.gf192_ifft(dom_affine(basis, shift), e) == oracle.additive_ifft(e, m, basis, shift)).all()


def test_additive_small_transforms_without_the_cluster_kernel():
    """Small transforms normally run on additive_cluster_pass (tile spread over a thread-block cluster); B200_NO_CLUSTER=1
    keeps them on additive_tile_pass (one CTA per tile).  Both must agree with the oracle."""
    import subprocess, sys, os
    code = r"""
import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import libiop_b200 as L
from oracle_lib import Oracle, el, rand_elems, standard_basis
o, dev = Oracle(), L.B200(0)
rng = np.random.default_rng(5)
for m, nc in ((9, 512), (11, 2048), (13, 2048), (13, 5000), (15, 1 << 12)):
    basis, shift = standard_basis(m), el(1 << m)
    c = rand_elems(rng, nc)
    ev = dev.gf192_fft(L.Domain.standard(m), c)
    assert (ev == o.additive_fft(c, m, basis, shift)).all(), (m, nc)
    d = max(1, int(np.ceil(np.log2(nc))))
    assert (dev.gf192_ifft(L.Domain.standard(m), ev)[:nc] == c).all(), (m, nc)
print('OK')
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env in ({"B200_NO_CLUSTER": "1"}, {}):
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


@pytest.mark.parametrize("m,n_coeffs,batch", [(13, 2048, 5), (13, 2048, 1), (12, 1000, 3), (15, 1 << 13, 2), (11, 2048, 37),
                                              (8, 64, 300), (13, 1, 4), (14, 1 << 14, 3), (16, 1 << 13, 3)])
def test_additive_fft_batch_equals_the_oracle_row_by_row(dev, oracle, m, n_coeffs, batch):
    """b200_gf192_fft_batch: the rows of Ligero's interleaved codes (ligero.tcc:293-331) in one call; every row must be
    the reference's additive_FFT of that row (non-power-of-two coefficient counts are zero-padded per row)."""
    import libiop_b200 as L
    rng = np.random.default_rng(1000 * m + batch)
    basis, shift = standard_basis(m), el(1 << m)
    rows = rand_elems(rng, batch * n_coeffs).reshape(batch, n_coeffs, 3)
    got = dev.gf192_fft_batch(L.Domain.standard(m), rows)
    assert got.shape == (batch, 1 << m, 3)
    for r in range(batch):
        assert (got[r] == oracle.additive_fft(rows[r], m, basis, shift)).all(), r


@pytest.mark.parametrize("m,batch", [(11, 7), (11, 1), (13, 3), (5, 300), (14, 2), (1, 5), (0, 4)])
def test_additive_ifft_batch_equals_the_oracle_row_by_row(dev, oracle, m, batch):
    import libiop_b200 as L
    rng = np.random.default_rng(77 * m + batch)
    basis, shift = standard_basis(m), el(1 << m)
    rows = rand_elems(rng, batch << m).reshape(batch, 1 << m, 3)
    got = dev.gf192_ifft_batch(L.Domain.standard(m), rows)
    for r in range(batch):
        assert (got[r] == oracle.additive_ifft(rows[r], m, basis, shift)).all(), r


def test_additive_batch_lde_at_ligero_size_round_trips(dev):
    """512 rows 2^11 -> 2^13 (BASELINE config 4): IFFT over the systematic domain, FFT over the codeword domain, and the
    first 2^11 coefficients of the IFFT over the codeword domain are the row's coefficients again."""
    import libiop_b200 as L
    rng = np.random.default_rng(4)
    rows = rand_elems(rng, 512 << 11).reshape(512, 1 << 11, 3)
    coeffs = dev.gf192_ifft_batch(L.Domain.standard(11), rows)
    code = dev.gf192_fft_batch(L.Domain.standard(13), coeffs)
    back = dev.gf192_ifft_batch(L.Domain.standard(13), code)
    assert (back[:, :1 << 11] == coeffs).all() and not back[:, 1 << 11:].any()
    for r in (0, 255, 511):
        assert (code[r] == dev.gf192_fft(L.Domain.standard(13), coeffs[r])).all()


@pytest.mark.parametrize("d,m,batch,shift_h", [(10, 13, 1, None), (12, 15, 1, 5), (13, 16, 1, None), (14, 16, 2, None), (11, 13, 7, 3), (5, 8, 40, None),
                                               (0, 4, 3, None), (9, 9, 1, None), (16, 19, 1, None), (3, 14, 2, 1)])
def test_fused_lde_equals_ifft_then_fft(dev, oracle, d, m, batch, shift_h):
    """b200_gf192_lde: evaluations over H -> evaluations over L without the monomial coefficients (the reference interpolates and
    re-evaluates: r1cs_rs_iop.tcc:449-479).  Bit-exact against the oracle's additive_IFFT over H followed by additive_FFT over L,
    Aurora-shaped domains (standard basis, shifts x^d / x^m) and other shifts of H, batches, tiny and tile-crossing sizes."""
    import libiop_b200 as L
    rng = np.random.default_rng(31 * d + m + batch)
    basis_h, basis_l = standard_basis(d), standard_basis(m)
    sh = el(1 << d) if shift_h is None else rand_elems(rng, 1)[0]
    sl = el(1 << m)
    dom_h = L.Domain.affine(basis_h, sh) if d else L.Domain.affine(elems(0), sh)
    dom_l = L.Domain.affine(basis_l, sl)
    rows = rand_elems(rng, batch << d).reshape(batch, 1 << d, 3)
    got = dev.gf192_lde(dom_h, dom_l, rows)
    for r in range(batch):
        coeffs = oracle.additive_ifft(rows[r], d, basis_h, sh) if d else rows[r]
        assert (got[r] == oracle.additive_fft(coeffs, m, basis_l, sl)).all(), r


@pytest.mark.parametrize("d,m", [(6, 9), (10, 13), (13, 15)])
def test_lde_halves_and_the_quotient_by_the_vanishing_polynomial(dev, oracle, d, m):
    """b200_gf192_lde_interpolate / _evaluate: (1) the two halves compose to the fused extension, i.e. to the oracle's
    additive_IFFT over H then additive_FFT over L; (2) the form in between is linear and shift-independent, which gives the
    sumcheck quotient h = floor(f / Z_H) (sumcheck.tcc:344-388) from f's values on two cosets of H:
        h|L = evaluate_L( (interpolate_{C_0}(f|C_0) + interpolate_{C_1}(f|C_1)) / (Z_H(b) - Z_H(0)) ),  C_1 = C_0 + b.
    Checked against the oracle: f = g + Z_H * h for random g (deg < |H|), h (deg < |H|), Z_H expanded by multiplication."""
    import libiop_b200 as L
    rng = np.random.default_rng(17 * d + m)
    basis_l = standard_basis(m)
    sl = el(1 << m)
    dom_l = L.Domain.affine(basis_l, sl)
    nh = 1 << d
    # (1) composition
    sh = rand_elems(rng, 1)[0]
    ev = rand_elems(rng, nh)
    got = dev.gf192_lde_halves([L.Domain.affine(basis_l[:d], sh)], [ev], dom_l)
    assert (got == oracle.additive_fft(oracle.additive_ifft(ev, d, basis_l[:d], sh), m, basis_l, sl)).all()
    # (2) quotient: Z_H = prod_{a in H} (x - a), H = span(basis[0..d)) unshifted, as monomial coefficients by repeated products
    g, h = rand_elems(rng, nh), rand_elems(rng, nh)
    dom2 = (d + 1, basis_l[:d + 1], sl)                                   # the first 2|H| positions of L
    h_on_2h = oracle.additive_fft(h, *dom2)
    g_on_2h = oracle.additive_fft(g, *dom2)
    pts = oracle.additive_fft(np.array([[0, 0, 0], [1, 0, 0]], dtype=np.uint64), *dom2)          # the polynomial x: the points themselves
    H_elems = oracle.additive_fft(np.array([[0, 0, 0], [1, 0, 0]], dtype=np.uint64), d, basis_l[:d], el(0))
    Z = np.tile(el(1).reshape(1, 3), (2 * nh, 1))
    for a in H_elems:
        Z = oracle.gf192_mul(Z, pts ^ a.reshape(1, 3))
    f_on_2h = g_on_2h ^ oracle.gf192_mul(Z, h_on_2h)
    # w = Z_H(b) - Z_H(0) with b = basis[d]; Z_H(0) = 0 for the unshifted H
    w = np.tile(el(1).reshape(1, 3), (1, 1))
    for a in H_elems:
        w = oracle.gf192_mul(w, (basis_l[d] ^ a).reshape(1, 3))
    w_inv = dev.gf192_inv(w)[0]
    c0 = L.Domain.affine(basis_l[:d], sl)
    c1 = L.Domain.affine(basis_l[:d], sl ^ basis_l[d])
    got_h = dev.gf192_lde_halves([c0, c1], [f_on_2h[:nh], f_on_2h[nh:]], dom_l, scale=w_inv)
    assert (got_h == oracle.additive_fft(h, m, basis_l, sl)).all()


def test_fused_lde_rejects_unrelated_bases(dev):
    import libiop_b200 as L
    rng = np.random.default_rng(9)
    dom_h = L.Domain.affine(rand_elems(rng, 6), el(7))
    with pytest.raises(L.B200Error):
        dev.gf192_lde(dom_h, L.Domain.standard(9), rand_elems(rng, 64))


def test_fused_lde_at_config_size_matches_the_two_transforms(dev):
    """2^20 -> 2^23 (BASELINE config 2): the fused extension equals IFFT over H then FFT over L computed by the device's own two
    transforms (which the oracle pins at smaller sizes)."""
    import libiop_b200 as L
    rng = np.random.default_rng(20)
    ev = rng.integers(0, 2**64, size=(1 << 20, 3), dtype=np.uint64)
    dom_h, dom_l = L.Domain.standard(20), L.Domain.standard(23)
    fused = dev.gf192_lde(dom_h, dom_l, ev)[0]
    two = dev.gf192_fft(dom_l, dev.gf192_ifft(dom_h, ev))
    assert hashlib.sha256(fused.tobytes()).hexdigest() == hashlib.sha256(two.tobytes()).hexdigest()


def test_additive_fft_empty_and_errors(dev):
    import libiop_b200 as L
    d = L.Domain.standard(6)
    out = dev.gf192_fft(d, elems(0))
    assert not out.any()
    with pytest.raises(L.B200Error):
        dev.gf192_fft(d, elems(65))                       # more coefficients than domain elements
    with pytest.raises(L.B200Error):
        dev.fp_ntt(L.FpParams.edwards_fr(), d, elems(4))  # wrong domain kind (fft.tcc:333)


def test_additive_lde_roundtrip_large(dev):
    """BASELINE config-2 shape at full size: 2^20 coefficients -> 2^23 evaluations (rate 1/8);
    checked through a size-independent property: IFFT over the first 2^20-point sub-domain of L
    (IFFT_of_known_degree semantics, fft.tcc:458-475) returns the coefficients, and linearity."""
    import libiop_b200 as L
    rng = np.random.default_rng(7)
    m, dlog = 23, 20
    n, nd = 1 << m, 1 << dlog
    dom = L.Domain.standard(m)
    c1, c2 = rand_elems(rng, nd), rand_elems(rng, nd)
    b1, b2, b3 = dev.upload(c1), dev.upload(c2), dev.upload(c1 ^ c2)
    e1, e2, e3 = dev.alloc(n), dev.alloc(n), dev.alloc(n)
    dev.gf192_fft_dev(dom, b1, nd, e1); dev.gf192_fft_dev(dom, b2, nd, e2); dev.gf192_fft_dev(dom, b3, nd, e3)
    # linearity on a strided sample of the 2^23 outputs
    idx = rng.integers(0, n, size=4096).astype(np.uint64)
    assert (dev.gather(e1, idx) ^ dev.gather(e2, idx) == dev.gather(e3, idx)).all()
    # interpolate back from the first 2^20 evaluations (sub-domain = first 20 basis vectors, same shift)
    sub = L.Domain.standard(dlog, shift_int=1 << m)
    back = dev.alloc(nd)
    dev.gf192_ifft_dev(sub, e1, back)
    assert (dev.download(back, nd) == c1).all()
    for b in (b1, b2, b3, e1, e2, e3, back):
        dev.free(b)


# ------------------------------------------------------------------ multiplicative NTT
@pytest.mark.parametrize("tag,logn", [("coset", 9), ("sub", 7)])
def test_ntt_golden(dev, golden, fp, tag, logn):
    g = golden
    d = dom_coset(fp, logn, g[f"mfft_{tag}_shift"])
    assert (dev.fp_ntt(fp, d, g[f"mfft_{tag}_coeffs"]) == g[f"mfft_{tag}_evals"]).all()
    assert (dev.fp_intt(fp, d, g[f"mifft_{tag}_in"]) == g[f"mifft_{tag}_out"]).all()


@pytest.mark.parametrize("logn,nc,coset", [(1, 2, True), (3, 8, False), (6, 64, True), (10, 1 << 10, True), (13, 1 << 10, True),
                                           (14, 1000, False), (16, 1 << 16, True), (17, 1 << 14, True), (15, 1, True), (12, 3, False)])
def test_ntt_vs_oracle(dev, oracle, fp, logn, nc, coset):
    rng = np.random.default_rng(400 + logn + nc)
    shift = oracle.fp_generator() if coset else oracle.fp_from_u64(1)
    gen = oracle.fp_subgroup_generator(logn)
    assert (fp.subgroup_generator(logn) == gen).all()
    c = oracle.rand_fp(rng, nc)
    d = dom_coset(fp, logn, shift)
    want = oracle.naive_fft_fp(c, logn, gen, shift) if logn <= 6 else oracle.mult_fft(c, logn, gen, shift)
    assert (dev.fp_ntt(fp, d, c) == want).all()
    full = oracle.rand_fp(rng, 1 << logn)
    assert (dev.fp_intt(fp, d, full) == oracle.mult_ifft(full, logn, gen, shift)).all()


def test_ntt_roundtrip_large(dev, oracle, fp):
    """Fractal config shape: 2^20 coefficients on the 2^23 coset; iNTT over the order-2^20 sub-coset
    (every 8th evaluation, fft.tcc:446-455) returns the coefficients."""
    import libiop_b200 as L
    rng = np.random.default_rng(8)
    logn, dlog = 23, 20
    shift = oracle.fp_generator()
    c = oracle.rand_fp(rng, 1 << 10)
    c = np.tile(c, (1 << (dlog - 10), 1))
    c[:, 0] ^= np.arange(len(c), dtype=np.uint64) & np.uint64(0xFFFF)   # distinct, still < p
    dom = dom_coset(fp, logn, shift)
    b = dev.upload(c); e = dev.alloc(1 << logn)
    dev.fp_ntt_dev(fp, dom, b, len(c), e)
    idx = (np.arange(1 << dlog, dtype=np.uint64) * np.uint64(8))
    sub_evals = dev.gather(e, idx)
    sub = dom_coset(fp, dlog, shift)
    assert (dev.fp_intt(fp, sub, sub_evals) == c).all()
    dev.free(b); dev.free(e)


# ------------------------------------------------------------------ FRI fold (T3)
@pytest.mark.parametrize("tag,m,eta", [("a1", 8, 1), ("a2", 8, 2), ("a3", 6, 3)])
def test_fold_additive_golden(dev, golden, tag, m, eta):
    g = golden
    d = dom_affine(g[f"fold_{tag}_basis"], g[f"fold_{tag}_shift"])
    assert (dev.fri_fold(1, None, d, eta, g[f"fold_{tag}_x"], g[f"fold_{tag}_f"]) == g[f"fold_{tag}_out"]).all()
    assert (dev.fri_fold(1, None, d, eta, g[f"fold_{tag}_xin"], g[f"fold_{tag}_f"]) == g[f"fold_{tag}_out_xin"]).all()


@pytest.mark.parametrize("tag,logn,eta", [("m1", 8, 1), ("m2", 8, 2), ("m3", 6, 3)])
def test_fold_mult_golden(dev, golden, fp, tag, logn, eta):
    g = golden
    d = dom_coset(fp, logn, g[f"fold_{tag}_shift"])
    assert (dev.fri_fold(2, fp, d, eta, g[f"fold_{tag}_x"], g[f"fold_{tag}_f"]) == g[f"fold_{tag}_out"]).all()
    assert (dev.fri_fold(2, fp, d, eta, g[f"fold_{tag}_xin"], g[f"fold_{tag}_f"]) == g[f"fold_{tag}_out_xin"]).all()


@pytest.mark.parametrize("m,eta", [(14, 1), (14, 2), (13, 3), (12, 4), (16, 2)])
def test_fold_additive_vs_oracle(dev, oracle, m, eta):
    rng = np.random.default_rng(500 + m + eta)
    basis, shift, x = rand_elems(rng, m), rand_elems(rng, 1)[0], rand_elems(rng, 1)[0]
    f = rand_elems(rng, 1 << m)
    assert (dev.fri_fold(1, None, dom_affine(basis, shift), eta, x, f) == oracle.fri_fold_additive(f, m, basis, shift, eta, x)).all()


@pytest.mark.parametrize("logn,eta", [(14, 1), (14, 2), (13, 3), (12, 4), (16, 2)])
def test_fold_mult_vs_oracle(dev, oracle, fp, logn, eta):
    rng = np.random.default_rng(600 + logn + eta)
    shift, x = oracle.fp_generator(), oracle.rand_fp(rng, 1)[0]
    gen = oracle.fp_subgroup_generator(logn)
    f = oracle.rand_fp(rng, 1 << logn)
    assert (dev.fri_fold(2, fp, dom_coset(fp, logn, shift), eta, x, f) == oracle.fri_fold_mult(f, logn, gen, shift, eta, x)).all()


def test_fold_of_low_degree_is_evaluation(dev, oracle):
    """test_fri_aux.cpp:16-86 at BASELINE round size: folding the LDE of a degree < 2^eta polynomial gives P(x)."""
    import libiop_b200 as L
    rng = np.random.default_rng(9)
    m, eta = 22, 2
    dom = L.Domain.standard(m)
    c = rand_elems(rng, 1 << eta)
    x = rand_elems(rng, 1)[0]
    b = dev.upload(c); e = dev.alloc(1 << m); o = dev.alloc(1 << (m - eta))
    dev.gf192_fft_dev(dom, b, len(c), e)
    dev.fri_fold_dev(1, None, dom, eta, x, e, o)
    px = oracle.naive_fft_gf192(c, 0, elems(0), x)[0]
    got = dev.download(o, 1 << (m - eta))
    assert (got == px).all()
    for t in (b, e, o):
        dev.free(t)


# ------------------------------------------------------------------ Merkle (T4)
def test_blake2b_batch(dev):
    rng = np.random.default_rng(10)
    for msg_len in (0, 8, 24, 48, 64, 96, 128, 136, 192, 256, 264, 1000 * 8):
        n = 37
        data = rng.bytes(msg_len * n) if msg_len else b""
        if msg_len == 0:
            continue
        got = dev.blake2b_batch(data, msg_len)
        want = b"".join(hashlib.blake2b(data[i * msg_len:(i + 1) * msg_len], digest_size=32).digest() for i in range(n))
        assert got == want, msg_len


@pytest.mark.parametrize("i", range(5))
def test_merkle_golden(dev, golden, i):
    g = golden
    no, n, cs, kind = (int(v) for v in g[f"mt{i}_cfg"])
    orc = list(g[f"mt{i}_oracles"])
    assert dev.merkle_build(orc, cs, kind) == g[f"mt{i}_nodes"].tobytes()
    assert dev.merkle_build(orc, cs, kind, salts=g[f"mt{i}_zk_salts"].tobytes()) == g[f"mt{i}_zk_nodes"].tobytes()


@pytest.mark.parametrize("no,n,cs,kind", [(4, 1 << 12, 2, 1), (1, 1 << 12, 2, 1), (1, 1 << 13, 4, 1), (70, 1 << 9, 1, 1),
                                          (3, 1 << 12, 2, 2), (1, 1 << 12, 4, 2), (2, 4, 2, 1), (1, 1 << 15, 1, 2)])
def test_merkle_vs_oracle(dev, oracle, no, n, cs, kind):
    rng = np.random.default_rng(700 + no + n + cs + kind)
    orc = [rand_elems(rng, n) if kind == 1 else oracle.rand_fp(rng, n) for _ in range(no)]
    assert dev.merkle_build(orc, cs, kind) == oracle.merkle_build(orc, cs, kind)
    salts = rng.bytes((n // cs) * 32)
    assert dev.merkle_build(orc, cs, kind, salts=salts) == oracle.merkle_build(orc, cs, kind, salts=salts)


@pytest.mark.parametrize("no,n,cs,kind", [(130, 1 << 10, 1, 1), (43, 8, 1, 1), (20, 1 << 8, 4, 1), (44, 1 << 11, 1, 2), (12, 1 << 9, 4, 2),
                                          (513, 1 << 9, 1, 1)])
def test_merkle_wide_leaves_warp_cooperative(dev, oracle, no, n, cs, kind):
    """Leaves of >= 1 KB (Ligero's interleaved rows, merkle_tree.tcc:123-145) are hashed by merkle_leaf_quad_kernel: four lanes per
    BLAKE2b state, message blocks staged in shared memory.  Same trees as the one-thread-per-hash kernels (B200_NO_QUAD_LEAF is
    compared in test_merkle_wide_leaves_both_kernels_agree) and as the oracle: ragged last blocks, zk salts, both leaf layouts,
    more than 64 oracles (device pointer table)."""
    rng = np.random.default_rng(900 + no + n + cs + kind)
    orc = [rand_elems(rng, n) if kind == 1 else oracle.rand_fp(rng, n) for _ in range(no)]
    assert dev.merkle_build(orc, cs, kind) == oracle.merkle_build(orc, cs, kind)
    salts = rng.bytes((n // cs) * 32)
    assert dev.merkle_build(orc, cs, kind, salts=salts) == oracle.merkle_build(orc, cs, kind, salts=salts)


def test_merkle_wide_leaves_both_kernels_agree():
    """Ligero's tree at full size (512 rows x 2^13, BASELINE config 4) with and without the warp-cooperative leaf kernel."""
    import subprocess, sys, os
    code = r"""
import sys, hashlib
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import libiop_b200 as L
dev = L.B200(0)
rng = np.random.default_rng(11)
rows = [dev.upload(rng.integers(0, 2**64, size=(1 << 13, 3), dtype=np.uint64)) for _ in range(512)]
nodes = dev.alloc_bytes(32 * (2 * (1 << 13) - 1))
dev.merkle_build_dev(rows, 1 << 13, 1, 1, nodes)
print('TREE', hashlib.sha256(dev.download_bytes(nodes, 32 * (2 * (1 << 13) - 1))).hexdigest())
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    got = []
    for env in ({"B200_NO_QUAD_LEAF": "1"}, {}):
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "TREE" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
        got.append(r.stdout.split("TREE")[1].split()[0])
    assert got[0] == got[1]


def test_merkle_errors(dev):
    import libiop_b200 as L
    with pytest.raises(L.B200Error):
        dev.merkle_build([elems(6)], 2, 1)        # 3 leaves: not a power of two (merkle_tree.tcc:27-31)
    with pytest.raises(L.B200Error):
        dev.merkle_build([elems(2)], 2, 1)        # 1 leaf


def test_merkle_large_consistency(dev):
    """BASELINE tree0 shape (4 oracles x 2^22, coset 2 -> 2^21 leaves here to keep the download small):
    root recomputed on the host from the device's leaf layer with hashlib."""
    rng = np.random.default_rng(11)
    n, cs, no = 1 << 18, 2, 4
    orc = [rand_elems(rng, n) for _ in range(no)]
    nodes = dev.merkle_build(orc, cs, 1)
    leaves = n // cs
    layer = [nodes[(leaves - 1 + i) * 32:(leaves + i) * 32] for i in range(leaves)]
    # spot-check leaves, then rebuild all inner nodes
    for i in (0, 1, leaves // 2, leaves - 1):
        msg = b"".join(orc[k][i * cs + j].tobytes() for k in range(no) for j in range(cs))
        assert layer[i] == hashlib.blake2b(msg, digest_size=32).digest()
    while len(layer) > 1:
        layer = [hashlib.blake2b(layer[2 * i] + layer[2 * i + 1], digest_size=32).digest() for i in range(len(layer) // 2)]
    assert layer[0] == nodes[:32]


# ------------------------------------------------------------------ C++ adapter (libiop_b200/host)
def test_cpp_adapter_reference_style_tests():
    """tests/cpp/test_host_api.cpp mirrors the reference's test_fft / test_fri_aux / test_merkle_tree gtests."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_host_api.bin")
    assert os.path.exists(exe), "build() must have produced tests/cpp/test_host_api.bin"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "0 failed" in out.stdout


def test_lazy_vectors_on_the_device():
    """tests/hostsim/lazy_vector_test.cpp built against the real library (tests/cpp/lazy_vector_test.bin): device-resident host
    vectors — ownership-keyed registration, materialisation on any host access through the SIGSEGV handler (a real
    cudaMemcpy inside it), clones, query gathers."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "lazy_vector_test.bin")
    assert os.path.exists(exe), "build() must have produced tests/cpp/lazy_vector_test.bin"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "0 failed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


# ------------------------------------------------------------------ whole-prover drop-in parity (T5, T6)
@pytest.mark.parametrize("log_n,zk", [(10, False), (10, True), (12, False), (14, False)])
def test_reference_prover_with_b200_dropin_is_bit_exact(golden, log_n, zk):
    """The UNMODIFIED reference aurora_snark_prover, compiled with the B200 adapter shadowing its hot-path
    headers (libiop_b200/dropin -> libiop_b200/lib/libiop_b200_dropin.so), must emit the same
    transcript bytes as the CPU reference (canonical dump of SURVEY.md §8c), and the reference verifier
    must accept it.  Fingerprints of the CPU run are in the golden fixture (N=2^10, 2^12); where the CPU
    reference library travelled to this box it is also run side by side."""
    import hashlib
    from oracle_lib import Reference, ReferenceGpu
    if not ReferenceGpu.available():
        pytest.skip("libiop_b200/lib/libiop_b200_dropin.so not built (needs /root/reference at build time)")
    g = ReferenceGpu()
    secs, dump, accepted = g.aurora_gf192(log_n, make_zk=zk, verify=True)
    assert accepted == 1, "reference verifier rejected the GPU-built proof"
    fp = hashlib.sha256(dump).hexdigest()
    known = {}
    for line in golden["aurora_fingerprints"]:
        n, z, ln, h = str(line).split(",")
        known[(int(n), bool(int(z)))] = (int(ln), h)
    if (log_n, zk) in known:
        assert (len(dump), fp) == known[(log_n, zk)]
    if Reference.available():
        r = Reference()
        _, cpu_dump, cpu_ok = r.aurora_gf192(log_n, make_zk=zk, verify=False)
        assert cpu_dump == dump


@pytest.mark.parametrize("which,log_n,zk", [("aurora_edwards", 10, False), ("aurora_edwards", 12, False),
                                            ("fractal_edwards", 10, False), ("fractal_edwards", 12, False),
                                            ("ligero_gf192", 12, False), ("ligero_gf192", 14, False), ("ligero_gf192", 10, True)])
def test_other_reference_provers_with_b200_dropin_are_bit_exact(golden, which, log_n, zk):
    """fractal_snark_indexer/prover over edwards_Fr (multiplicative NTT path), ligero_snark_prover over GF(2^192)
    (hundreds of small LDEs + one wide-leaf Merkle tree) and Aurora over edwards_Fr, all UNMODIFIED reference host
    code with the B200 hot path dropped in: same transcript bytes as the CPU reference (fingerprints generated by
    tests/golden/make_golden.py from oracle/_ref/libiopref.so), and the reference verifiers accept."""
    import hashlib
    from oracle_lib import ReferenceGpu
    if not ReferenceGpu.available():
        pytest.skip("libiop_b200/lib/libiop_b200_dropin.so not built (needs /root/reference at build time)")
    g = ReferenceGpu()
    if which == "aurora_edwards":
        _, dump, acc = g.aurora_edwards(log_n)
    elif which == "fractal_edwards":
        _, dump, acc, _ = g.fractal_edwards(log_n)
    else:
        _, dump, acc = g.ligero_gf192(log_n, make_zk=zk)
    assert acc == 1, "reference verifier rejected the GPU-built proof"
    want = {}
    for line in golden["prover_fingerprints"]:
        w, n, z, ln, h = str(line).split(",")
        want[(w, int(n), bool(int(z)))] = (int(ln), h)
    assert (len(dump), hashlib.sha256(dump).hexdigest()) == want[(which, log_n, zk)]


def _extra_fingerprints():
    import json, os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prover_fingerprints_extra.json")
    return json.load(open(path))


@pytest.mark.parametrize("case", _extra_fingerprints(), ids=lambda c: f"{c['prover']}-2^{c['log_n']}{'-zk' if c['zk'] else ''}")
def test_dropin_provers_match_cpu_reference_fingerprints(case):
    """More sizes and the zero-knowledge variants: the transcript bytes of the drop-in provers must hash to the fingerprints
    of the CPU reference runs recorded by tests/golden/make_prover_fingerprints.py, and the reference verifier must accept."""
    import hashlib
    from oracle_lib import ReferenceGpu
    if not ReferenceGpu.available():
        pytest.skip("libiop_b200/lib/libiop_b200_dropin.so not built (needs /root/reference at build time)")
    g = ReferenceGpu()
    which, log_n, zk = case["prover"], case["log_n"], case["zk"]
    if which == "aurora_gf192":
        _, dump, acc = g.aurora_gf192(log_n, make_zk=zk)
    elif which == "aurora_edwards":
        _, dump, acc = g.aurora_edwards(log_n)
    elif which == "fractal_edwards":
        _, dump, acc, _ = g.fractal_edwards(log_n)
    else:
        _, dump, acc = g.ligero_gf192(log_n, make_zk=zk)
    assert acc == 1, "reference verifier rejected the GPU-built proof"
    assert (len(dump), hashlib.sha256(dump).hexdigest()) == (case["bytes"], case["sha256"])


def _config_fingerprints():
    import json, os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prover_fingerprints_config.json")
    return json.load(open(path)) if os.path.exists(path) else []


@pytest.mark.gpu
@pytest.mark.parametrize("case", _config_fingerprints(), ids=lambda c: f"{c['config']}-{c['prover']}-2^{c['log_n']}{'-zk' if c['zk'] else ''}")
def test_dropin_provers_at_baseline_config_sizes(case):
    """BASELINE.json configs 2-5 at FULL size (Aurora GF(2^192) N=2^20, Fractal edwards N=2^18, Ligero GF(2^192) N=2^18, Aurora
    N=2^22 = one rank of config 5, and the zero-knowledge Aurora with the same 2^23-point codeword as config 2): the transcript
    of the drop-in prover must hash to the fingerprint of the UNMODIFIED CPU reference run recorded (with its CPU seconds) by
    tests/golden/make_config_fingerprints.py, and the reference verifier must accept the GPU-built proof."""
    import hashlib
    from oracle_lib import ReferenceGpu
    if not ReferenceGpu.available():
        pytest.skip("libiop_b200/lib/libiop_b200_dropin.so not built (needs /root/reference at build time)")
    g = ReferenceGpu()
    which, log_n, zk = case["prover"], case["log_n"], case["zk"]
    if which == "aurora_gf192":
        _, dump, acc = g.aurora_gf192(log_n, make_zk=zk)
    elif which == "fractal_edwards":
        _, dump, acc, _ = g.fractal_edwards(log_n)
    else:
        _, dump, acc = g.ligero_gf192(log_n, make_zk=zk)
    assert acc == 1, "reference verifier rejected the GPU-built proof"
    assert (len(dump), hashlib.sha256(dump).hexdigest()) == (case["bytes"], case["sha256"])


def test_device_spmv_route_of_the_dropin_is_bit_exact():
    """create_Az_Bz_Cz_from_variable_assignment goes to b200_spmv_csr only from three terms per row (measured crossover);
    B200_SPMV_MIN_TERMS_PER_ROW=0 forces that route for the one-term rows of generate_r1cs_example — same transcripts."""
    import subprocess, sys, os
    from oracle_lib import ReferenceGpu
    if not ReferenceGpu.available():
        pytest.skip("libiop_b200/lib/libiop_b200_dropin.so not built (needs /root/reference at build time)")
    code = r"""
import hashlib, sys
sys.path.insert(0, 'tests')
from oracle_lib import ReferenceGpu
g = ReferenceGpu()
for which, n in (('aurora_gf192', 14), ('aurora_edwards', 14)):
    _, d, a = getattr(g, which)(n)
    print('RES', which, n, len(d), hashlib.sha256(d).hexdigest(), a)
print('LAUNCHED', g.device_stats()['launches'])
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, B200_SPMV_MIN_TERMS_PER_ROW="0", B200_TRACE="1"),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "spmv_csr" in r.stderr, "the device SpMV route was not taken"
    want = {(c["prover"], c["log_n"]): (c["bytes"], c["sha256"]) for c in _extra_fingerprints() if not c["zk"]}
    seen = 0
    for line in r.stdout.splitlines():
        if line.startswith("RES "):
            _, w, n, ln, h, acc = line.split()
            assert int(acc) == 1 and (int(ln), h) == want[(w, int(n))], line
            seen += 1
    assert seen == 2


def test_standalone_cpp_program_on_the_reference_api():
    """libiop_b200/dropin/examples/aurora_prove.cpp: a plain C++ executable calling aurora_snark_prover / aurora_snark_verifier with
    the library dropped in (replacement operator new in the executable itself): runs, verifies, and is deterministic."""
    import subprocess, os, re
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libiop_b200", "lib", "aurora_prove")
    if not os.path.exists(exe):
        pytest.skip("libiop_b200/lib/aurora_prove not built (needs /root/reference at build time)")
    r = subprocess.run([exe, "14", "0", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    sums = re.findall(r"\(accept\), transcript (\d+) bytes, fnv1a ([0-9a-f]{16})", r.stderr)
    assert len(sums) == 2 and sums[0] == sums[1], r.stderr[-1000:]
    assert int(sums[0][0]) == 149800            # same length as the CPU reference's transcript at N = 2^14 (prover_fingerprints_extra.json)


@pytest.mark.parametrize("lazy", ["1", "0"])
def test_device_resident_oracles_do_not_change_the_transcripts(golden, lazy):
    """Oracles produced by the adapter stay on the device behind lazy host vectors (libiop_b200/host/b200_residency.hpp); with
    B200_LAZY=0 every result is downloaded instead.  Both ways all provers must emit the CPU reference's transcripts; with lazy
    vectors on, the device copies must actually be used (hits) and the provers must not read any of them on the host (a
    host read would be served correctly by the fault handler, but as a whole-vector download)."""
    import subprocess, sys, os
    from oracle_lib import ReferenceGpu
    if not ReferenceGpu.available():
        pytest.skip("libiop_b200/lib/libiop_b200_dropin.so not built (needs /root/reference at build time)")
    code = r"""
import hashlib, sys, ctypes as C
sys.path.insert(0, 'tests')
from oracle_lib import ReferenceGpu
g = ReferenceGpu()
out = []
_, d, a = g.aurora_gf192(12, make_zk=False, verify=False); out.append(('aurora', 12, 0, len(d), hashlib.sha256(d).hexdigest()))
_, d, a = g.aurora_gf192(10, make_zk=True, verify=False); out.append(('aurora', 10, 1, len(d), hashlib.sha256(d).hexdigest()))
_, d, a = g.aurora_edwards(12, verify=False); out.append(('aurora_edwards', 12, 0, len(d), hashlib.sha256(d).hexdigest()))
_, d, a, _ = g.fractal_edwards(12, verify=False); out.append(('fractal_edwards', 12, 0, len(d), hashlib.sha256(d).hexdigest()))
_, d, a = g.ligero_gf192(14, make_zk=False, verify=False); out.append(('ligero_gf192', 14, 0, len(d), hashlib.sha256(d).hexdigest()))
st, lz = g.residency_stats(), g.lazy_stats()
for o in out: print('RES', *o)
print('STATS', st['hits'], lz['faults'], lz['gathers'])
"""
    env = dict(os.environ, B200_LAZY=lazy)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    want = {}
    for line in golden["aurora_fingerprints"]:
        n, z, ln, h = str(line).split(",")
        want[("aurora", int(n), int(z))] = (int(ln), h)
    for line in golden["prover_fingerprints"]:
        w, n, z, ln, h = str(line).split(",")
        want[(w, int(n), int(z))] = (int(ln), h)
    seen = 0
    for line in r.stdout.splitlines():
        if line.startswith("RES "):
            _, w, n, z, ln, h = line.split()
            assert (int(ln), h) == want[(w, int(n), int(z))], line
            seen += 1
        if line.startswith("STATS "):
            hits, faults, gathers = (int(x) for x in line.split()[1:4])
            if lazy == "1":
                assert hits > 50 and gathers > 10 and faults == 0, line
            else:
                assert hits == 0 and faults == 0, line
    assert seen == 5


# ------------------------------------------------------------------ proof of work
def test_pow_solve(dev, oracle, golden):
    """b200_pow_solve returns the FIRST accepted candidate in the reference's order (pow.tcc:87-102)."""
    for line in golden["pow_vectors"]:
        ch, bl, ub, sol = str(line).split(",")
        assert dev.pow_solve(bytes.fromhex(ch), int(bl), int(ub)).hex() == sol
    rng = np.random.default_rng(12)
    for bitlen in (0, 1, 5, 10, 14, 20, 23):
        ch = rng.bytes(32)
        want, tries = oracle.pow_solve(ch, bitlen, 0)
        got = dev.pow_solve(ch, bitlen, 0)
        assert got == want, (bitlen, tries)
        h = hashlib.blake2b(ch + got, digest_size=32).digest()
        assert int.from_bytes(h[24:], "little") & ((1 << bitlen) - 1) == 0


# ------------------------------------------------------------------ verifier random messages of the BCS hash chain
@pytest.mark.parametrize("field", [1, 2])
@pytest.mark.parametrize("n", [1, 255, 256, 257, 5000])
def test_squeeze_field_elements(dev, oracle, fp, field, n):
    """b200_squeeze_field_elements == blake2b_FieldT_randomness_extractor (bcs/hashing/blake2b.tcc:231-256): keyed
    BLAKE2b per element; for edwards_Fr about half of the draws are rejected and re-drawn with key += n."""
    from oracle_lib import squeeze_model, Reference
    root = np.random.default_rng(100 + n).bytes(32)
    index = 3 + n
    modulus = oracle.fp_modulus() if field == 2 else None
    got = dev.squeeze_field_elements(field, fp if field == 2 else None, root + index.to_bytes(8, "little"), n)
    assert (got == squeeze_model(root, index, n, modulus)).all()
    if Reference.available():
        assert (got == Reference().squeeze(field, root, index, n)).all()


def test_squeeze_field_elements_at_ligero_size(dev, oracle, fp):
    """One Ligero lincheck challenge at BASELINE config 4 is 2^18 elements (interleaved_lincheck_ot.tcc:79)."""
    from oracle_lib import squeeze_model
    root, index, n = bytes(range(32)), 7, 1 << 18
    got = dev.squeeze_field_elements(1, None, root + index.to_bytes(8, "little"), n)
    assert (got == squeeze_model(root, index, n)).all()
    got = dev.squeeze_field_elements(2, fp, root + index.to_bytes(8, "little"), n)
    p = oracle.fp_modulus()
    v = got.astype(object); v = v[:, 0] + (v[:, 1] << 64) + (v[:, 2] << 128)
    assert (v < p).all()
    pick = np.arange(0, n, 997)
    want = squeeze_model_at(root, index, n, pick, p)
    assert (got[pick] == want).all()


def squeeze_model_at(root, index, n, positions, modulus):
    """squeeze_model restricted to some positions (the draws are independent per position)"""
    msg = root + int(index).to_bytes(8, "little")
    mask = (1 << modulus.bit_length()) - 1
    out = np.zeros((len(positions), 3), dtype=np.uint64)
    for r, i in enumerate(positions):
        key = int(i)
        while True:
            v = int.from_bytes(hashlib.blake2b(msg, digest_size=24, key=key.to_bytes(8, "little")).digest(), "little") & mask
            if v < modulus:
                break
            key += n
        out[r] = [v & (2**64 - 1), (v >> 64) & (2**64 - 1), v >> 128]
    return out


def test_squeeze_rejects_bad_arguments(dev, fp):
    with pytest.raises(Exception):
        dev.squeeze_field_elements(1, None, b"\x00" * 36, 4)          # not a multiple of 8
    with pytest.raises(Exception):
        dev.squeeze_field_elements(1, None, b"\x00" * 136, 4)         # longer than one block
    assert len(dev.squeeze_field_elements(1, None, b"\x00" * 40, 0)) == 0
