"""CPU tests: the oracle (oracle/*.c) against the reference's own KATs, SURVEY.md Appendix C
(reference-generated), Python's hashlib, and tests/golden/hotpath_golden.npz (reference output)."""
import hashlib
import struct

import numpy as np
import pytest

from oracle_lib import el, elems, rand_elems, standard_basis

# depends/libff/libff/algebra/fields/tests/test_binary_fields.cpp:125-134 (a_hi,a_mid,a_lo,b_hi,b_mid,b_lo -> hi,mid,lo)
GF192_KATS = [
    (0x0cdc30af062ba589, 0x8418b817f5caf84c, 0x3590e0c94d76b16a, 0x4193e34b8241b8b6, 0x8d73a6d58a9833ad, 0xf7903a9c58e4223c, 0xedb2d77042f07f3e, 0xc223197d04425354, 0x601e9baa99e6b2fd),
    (0x56aa86373c498fdd, 0x81f59dadf3516810, 0x657e286088c96142, 0x0fe9a9f10ea6b3be, 0x614becf0f6981970, 0x12356e70b09e27a6, 0xcf0642476758587a, 0x2ed578d8f170b814, 0x0b86792ebec24165),
    (0x5d1bce4867374275, 0xf764a67447957a65, 0x897bb93561e04d72, 0xf507be43450e596c, 0x7625671a07a1b127, 0xb6f0e7fe834a305f, 0x2d43614daf895767, 0xa0d43a4a59167a19, 0x1d01f4493cf9f6bf),
    (0x23fad70584b9ff24, 0x00a7fced16ffa59b, 0x7009d6ea6cbc3723, 0x5756b56a1d208f91, 0xac2c97ebcf121998, 0xb4dab96b1454919d, 0x0b31e0f4e429e2d8, 0x91b61a18bdb18fd4, 0x35ce63ac64f58838),
    (0xc1c42adfa999b078, 0x633c9cefc089eb74, 0x0fd73239d93bd077, 0xfae308c406eb08fe, 0x49bfcd0bd4d96b01, 0xf55ab02f9dae69b8, 0xecd19c6c59d10ede, 0x5c081d84ee58552d, 0x598aefe615b5a49a),
    (0xf2e6b1294f40ac62, 0x512288f4b06917d3, 0xd2c36d17d828d9b4, 0x206901073d267fa5, 0x9a2789bfb11c03c4, 0x1cf009d26ee1f80f, 0x68fd98a8545f9688, 0x1e6f3c0f65f3b25d, 0x75fb9afdb21f1278),
    (0xc78b2478a98afb85, 0x579c25544b3ba640, 0x7766723a1141eddc, 0x1d5e8949c38296e0, 0x224aa1e6e025b316, 0x84eb9e2187501666, 0xd33304755a3790f5, 0xd9f5f3be81d6054d, 0x742cb900f1d5bcf9),
    (0x75bee645ed32bf73, 0xfaaf7393e729adf5, 0xe4e8ae96b691f6d2, 0x94659c82eed44ed6, 0x0d6f1491ffab6313, 0x034858f4f1f7b14b, 0x998f56abe7e4348f, 0xcafb5a359d76233d, 0x930bfcefe10d3166),
    (0x9ded3331f7e20e08, 0xae8f8592118650df, 0x4105e1bce7fda1a5, 0x22677e786437ded8, 0xeab6bb9efffe16f2, 0xae333a6c2e525a7e, 0x981afe97f038b8e7, 0x2803fd759a41f4e0, 0xf80ac10929fc7e3f),
    (0x13ee2aa2bc56bb9f, 0x4b28b4a78b34aed6, 0xc9a973387db34f3c, 0x5980393a96b7e262, 0xbee979d0d3e73491, 0x724efc0872e555d7, 0xce5179b296413b74, 0x90b969c3def4555f, 0x2de0b46a441b4738),
]


def hexel(s):
    """48 hex digits limb2||limb1||limb0 (gf192::print, gf192.cpp:324-327) -> limbs."""
    v = int(s, 16)
    return el(v & (2**64 - 1), (v >> 64) & (2**64 - 1), v >> 128)


def test_gf192_kats(oracle):
    for ah, am, al, bh, bm, bl, ch, cm, cl in GF192_KATS:
        a, b, c = el(al, am, ah), el(bl, bm, bh), el(cl, cm, ch)
        assert (oracle.gf192_mul(a, b)[0] == c).all()
        assert (oracle.gf192_mul_bitwise(a, b) == c).all()
    rng = np.random.default_rng(5)
    a = rand_elems(rng, 16)
    one = el(1)
    assert (oracle.gf192_mul(a, oracle.gf192_inv(a)) == one).all()


def test_field_golden(oracle, golden):
    g = golden
    assert (oracle.gf192_mul(g["gf_a"], g["gf_b"]) == g["gf_mul"]).all()
    assert (oracle.gf192_inv(g["gf_a"][:8]) == g["gf_inv"]).all()
    assert (oracle.fp_mul(g["fp_a"], g["fp_b"]) == g["fp_mul"]).all()
    for i in range(64):
        assert (oracle.fp_add(g["fp_a"][i], g["fp_b"][i]) == g["fp_add"][i]).all()
        assert (oracle.fp_sub(g["fp_a"][i], g["fp_b"][i]) == g["fp_sub"][i]).all()
    for i in range(8):
        assert (oracle.fp_inv(g["fp_a"][i]) == g["fp_inv"][i]).all()
    assert list(oracle.fp.mod) == list(g["fp_mod"])
    assert list(oracle.fp.r2) == list(g["fp_r2"])
    assert oracle.fp.inv == int(g["fp_inv_const"][0]) == 0xdde553277fffffff  # edwards_fields.cpp:34
    assert oracle.fp.two_adicity == 31 and oracle.fp.num_bits == 181        # edwards_fields.cpp:40-42
    assert (oracle.fp_generator() == g["fp_generator"]).all()
    for ln in (1, 4, 10, 23):
        assert (oracle.fp_subgroup_generator(ln) == g[f"fp_subgroup_gen_{ln}"]).all()


@pytest.mark.parametrize("tag,m", [("std", 10), ("std_full", 8), ("rnd", 7)])
def test_additive_fft_golden(oracle, golden, tag, m):
    g = golden
    basis, shift = g[f"afft_{tag}_basis"], g[f"afft_{tag}_shift"]
    assert (oracle.additive_fft(g[f"afft_{tag}_coeffs"], m, basis, shift) == g[f"afft_{tag}_evals"]).all()
    assert (oracle.additive_ifft(g[f"aifft_{tag}_in"], m, basis, shift) == g[f"aifft_{tag}_out"]).all()
    # IFFT o FFT = id (test_fft.cpp:123-139)
    back = oracle.additive_ifft(g[f"afft_{tag}_evals"], m, basis, shift)
    nc = len(g[f"afft_{tag}_coeffs"])
    assert (back[:nc] == g[f"afft_{tag}_coeffs"]).all() and not back[nc:].any()


def test_additive_fft_vs_naive(oracle):
    """test_fft.cpp:27-52: additive FFT == naive_FFT for m <= 11 (here m <= 7 for speed)."""
    rng = np.random.default_rng(11)
    for m in range(1, 8):
        basis, shift = rand_elems(rng, m), rand_elems(rng, 1)[0]
        c = rand_elems(rng, 1 << m)
        assert (oracle.additive_fft(c, m, basis, shift) == oracle.naive_fft_gf192(c, m, basis, shift)).all()


@pytest.mark.parametrize("tag,logn", [("coset", 9), ("sub", 7)])
def test_mult_fft_golden(oracle, golden, tag, logn):
    g = golden
    shift = g[f"mfft_{tag}_shift"]
    gen = oracle.fp_subgroup_generator(logn)
    assert (oracle.mult_fft(g[f"mfft_{tag}_coeffs"], logn, gen, shift) == g[f"mfft_{tag}_evals"]).all()
    assert (oracle.mult_ifft(g[f"mifft_{tag}_in"], logn, gen, shift) == g[f"mifft_{tag}_out"]).all()


def test_mult_fft_vs_naive(oracle):
    """test_fft.cpp:54-121."""
    rng = np.random.default_rng(12)
    for logn in range(1, 7):
        gen = oracle.fp_subgroup_generator(logn)
        for shift in (oracle.fp_generator(), oracle.fp_from_u64(1)):
            c = oracle.rand_fp(rng, 1 << logn)
            assert (oracle.mult_fft(c, logn, gen, shift) == oracle.naive_fft_fp(c, logn, gen, shift)).all()


@pytest.mark.parametrize("tag,m,eta", [("a1", 8, 1), ("a2", 8, 2), ("a3", 6, 3)])
def test_fold_additive_golden(oracle, golden, tag, m, eta):
    g = golden
    args = (g[f"fold_{tag}_f"], m, g[f"fold_{tag}_basis"], g[f"fold_{tag}_shift"], eta)
    assert (oracle.fri_fold_additive(*args, g[f"fold_{tag}_x"]) == g[f"fold_{tag}_out"]).all()
    assert (oracle.fri_fold_additive(*args, g[f"fold_{tag}_xin"]) == g[f"fold_{tag}_out_xin"]).all()


@pytest.mark.parametrize("tag,logn,eta", [("m1", 8, 1), ("m2", 8, 2), ("m3", 6, 3)])
def test_fold_mult_golden(oracle, golden, tag, logn, eta):
    g = golden
    gen = oracle.fp_subgroup_generator(logn)
    args = (g[f"fold_{tag}_f"], logn, gen, g[f"fold_{tag}_shift"], eta)
    assert (oracle.fri_fold_mult(*args, g[f"fold_{tag}_x"]) == g[f"fold_{tag}_out"]).all()
    assert (oracle.fri_fold_mult(*args, g[f"fold_{tag}_xin"]) == g[f"fold_{tag}_out_xin"]).all()


def test_fold_is_polynomial_evaluation(oracle):
    """test_fri_aux.cpp:16-86: folding the evaluations of a degree < coset polynomial gives P(x)."""
    rng = np.random.default_rng(13)
    m, eta = 6, 2
    basis, shift, x = rand_elems(rng, m), rand_elems(rng, 1)[0], rand_elems(rng, 1)[0]
    c = rand_elems(rng, 1 << eta)
    f = oracle.naive_fft_gf192(c, m, basis, shift)
    px = oracle.naive_fft_gf192(c, 0, elems(0), x)[0]
    out = oracle.fri_fold_additive(f, m, basis, shift, eta, x)
    assert (out == px).all()


def test_blake2b_vs_hashlib(oracle):
    rng = np.random.default_rng(14)
    for n in (0, 1, 24, 48, 64, 96, 127, 128, 129, 192, 255, 256, 257, 1000, 12288):
        d = rng.bytes(n)
        assert oracle.blake2b(d) == hashlib.blake2b(d, digest_size=32).digest()
        assert oracle.blake2b(d, 24, b"\x01\x00\x00\x00\x00\x00\x00\x00") == \
            hashlib.blake2b(d, digest_size=24, key=b"\x01\x00\x00\x00\x00\x00\x00\x00").digest()
        assert oracle.blake2b(d, 8, b"abcdefgh") == hashlib.blake2b(d, digest_size=8, key=b"abcdefgh").digest()


@pytest.mark.parametrize("i", range(5))
def test_merkle_golden(oracle, golden, i):
    g = golden
    no, n, cs, kind = (int(v) for v in g[f"mt{i}_cfg"])
    orc = list(g[f"mt{i}_oracles"])
    assert oracle.merkle_build(orc, cs, kind) == g[f"mt{i}_nodes"].tobytes()
    assert oracle.merkle_build(orc, cs, kind, salts=g[f"mt{i}_zk_salts"].tobytes()) == g[f"mt{i}_zk_nodes"].tobytes()


# ---- SURVEY.md Appendix C (reference output, cross-checked there against an independent model) ----
def test_appendix_c1_c2(oracle):
    c = elems(4)
    for i in range(4):
        c[i] = el(3 * i + 1, 0xfedcba9876543210 ^ i, (0x0123456789abcdef * (i + 1)) % 2**64)
    basis, shift = standard_basis(3), el(8)
    ev = oracle.additive_fft(c, 3, basis, shift)
    want = ["d34822f50da2199a0662f8aaf99d0018000000000000115e", "d1c7a6ea3bed8eedb97530eca86426500000000000001300",
            "2922a797e8041d174386ac0d9e5b770a0000000000001454", "30835ecc09a93beafa5afb1afa5afd280000000000001636",
            "04153c1cc3ab67801512c91caea9773c0000000000001f2b", "5c7fb79905ab55cbbcbc42bc4343b8080000000000001dfd",
            "270c53589918ef5d66aa673266aa637e0000000000001bc1", "6448a59988fa6c9cc9cf73c3beb80020000000000000192b"]
    assert (ev == np.stack([hexel(s) for s in want])).all()
    x = el(0x3333, 0x2222, 0x1111)
    fold = oracle.fri_fold_additive(ev, 3, basis, shift, 1, x)
    want2 = ["6baca615d9510f36d8bf08658fb71152fa57bc0442dc0ef2", "eb678a56d02b21d1e1e4434cb6fa87ec9ced1672240628c6",
             "69164f9fee4de9644f64b293181a66d7adc841311400dbc3", "e9dd63dce737c783763ff9ba2157f069cb72eb4772dafdf7"]
    assert (fold == np.stack([hexel(s) for s in want2])).all()


def test_appendix_c3_merkle_root(oracle):
    a, b = elems(8), elems(8)
    for i in range(8):
        a[i] = el(100 + i, 7 * i + 3, i + 1)
        b[i] = el(200 + i, 9 * i + 1, 50 + i)
    nodes = oracle.merkle_build([a, b], 2, 1)
    assert nodes[:32].hex() == "63727104797a17bf8c14d619f4afabd2b7e90c50c99afcec8d990c8a5bd83685"
    leaf0 = b"".join(struct.pack("<3Q", *map(int, e)) for e in (a[0], a[1], b[0], b[1]))
    assert nodes[3 * 32:4 * 32] == hashlib.blake2b(leaf0, digest_size=32).digest()


def test_appendix_c4_hashchain(oracle, golden):
    import ctypes as C
    from oracle_lib import _p
    hc = (C.c_uint8 * 128)()  # orc_hashchain is 64 + 8 + 8 bytes
    oracle.lib.orc_hashchain_init(hc, C.c_size_t(128))
    oracle.lib.orc_hashchain_absorb(hc)
    sq = elems(2)
    oracle.lib.orc_hashchain_squeeze_gf192(hc, _p(sq), C.c_size_t(2))
    assert (sq[0] == hexel("c20190344550b8f8d5409013363fd62f5e8628e516918ce3")).all()
    pos = [oracle.lib.orc_hashchain_squeeze_position(hc, C.c_uint64(1 << 20)) for _ in range(3)]
    assert pos == [26513, 930642, 629782]
    assert (sq == golden["hashchain_squeeze"]).all() and pos == [int(v) for v in golden["hashchain_positions"]]


def test_appendix_c5_c6(oracle):
    c = elems(4)
    for i in range(4):
        c[i] = oracle.fp_from_u64(1000003 * (i + 1) + 7)
    gen, shift = oracle.fp_subgroup_generator(3), oracle.fp_generator()
    ev = oracle.mult_fft(c, 3, gen, shift)
    want = ["000cb81dfbbd09eafd0b9dc5c2750df831f4eb5b76aeaa91", "0002b32ecb6ad8cd9e9c66a87dcf009a12e55a6c76ef538b",
            "000d6455481e604bf69b2c3b32fdd5aa089f97b2c27b2b19", "00005a65d2938deb6d0d8c94bccda7401e2ce8953f30d007",
            "0006614127694a1d8941632a8b26e4da777e27f89d8688b5", "0002cbf1210343f53b2e6ad4fea75ca397bb36174882028a",
            "0005129359caed0d048d1228c0ff1b82ab394e23e5543f34", "000f8142dec3685c75b9fcae4f26ccb576992cea3d627776"]
    assert (ev == np.stack([hexel(s) for s in want])).all()
    back = oracle.mult_ifft(ev, 3, gen, shift)
    assert (back[:4] == c).all() and not back[4:].any()
    fold = oracle.fri_fold_mult(ev, 3, gen, shift, 1, oracle.fp_from_u64(123456789))
    want2 = ["000d8399c40d8cb81fbee8620e54edc473591fb3c6ad0005", "000f790f3cbc5a763be70c8a09a8bc6a8eea9571da337fb1",
             "000f90c7970df7f2fde120e7fa0c4b3cc176a7f680586950", "000d9b521e5f2a34e1b8fcbffeb87c96a5e532386cd1e9a4"]
    assert (fold == np.stack([hexel(s) for s in want2])).all()


def test_oracle_equals_reference_when_built(oracle, reference):
    """Differential test against the unmodified reference (only where oracle/_ref exists)."""
    rng = np.random.default_rng(99)
    a, b = rand_elems(rng, 500), rand_elems(rng, 500)
    assert (oracle.gf192_mul(a, b) == reference.gf192_mul(a, b)).all()
    m = 9
    basis, shift = rand_elems(rng, m), rand_elems(rng, 1)[0]
    c = rand_elems(rng, 300)
    assert (oracle.additive_fft(c, m, basis, shift) == reference.additive_fft(c, m, basis, shift)).all()


def test_pow_golden(oracle, golden):
    """pow::solve_pow (pow.tcc:87-158): reference-generated (challenge, bit length) -> first accepted candidate."""
    for line in golden["pow_vectors"]:
        ch, bl, ub, sol = str(line).split(",")
        got, tries = oracle.pow_solve(bytes.fromhex(ch), int(bl), int(ub))
        assert got.hex() == sol
    ch = hashlib.sha256(b"zero-bits").digest()
    assert oracle.pow_solve(ch, 0, 0)[0] == ch          # the challenge itself is the first candidate


@pytest.mark.parametrize("field", [1, 2])
def test_squeeze_model_matches_the_reference(reference, oracle, field):
    """The hashlib model the GPU squeeze test checks against == the unmodified reference's
    blake2b_FieldT_randomness_extractor (bcs/hashing/blake2b.tcc:231-256), rejection sampling included."""
    from oracle_lib import squeeze_model
    modulus = oracle.fp_modulus() if field == 2 else None
    for n, index in ((1, 1), (5, 2), (300, 9)):
        root = bytes((7 * n + k) & 255 for k in range(32))
        assert (reference.squeeze(field, root, index, n) == squeeze_model(root, index, n, modulus)).all()
