"""CPU-side checks of the drop-in layer (no GPU needed): the libsodium stand-in used by the drop-in build hashes like
RFC 7693 / Python's hashlib, its randombytes stream is the one the parity harness expects, and the drop-in library —
when it has been built (needs the reference checkout) — loads and exports the whole-prover entry points."""
import ctypes as C
import hashlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "libiop_b200", "dropin", "shims", "sodium_shim.c")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libsodium_shim.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, SHIM])
    return C.CDLL(so)


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1000, 4096 + 7])
def test_shim_blake2b_matches_hashlib(shim, n):
    """crypto_generichash_blake2b stand-in (libiop/bcs/hashing/blake2b.cpp:15,37,62 call sites): unkeyed 32-byte digests
    (leaf / two-to-one hashes), keyed short digests (hash chain), block-boundary lengths."""
    data = bytes((i * 131 + 7) & 0xFF for i in range(n))
    for outlen, key in ((32, b""), (64, b""), (24, bytes(range(8))), (16, bytes(range(64)))):
        out = (C.c_uint8 * outlen)()
        rc = shim.crypto_generichash_blake2b(out, C.c_size_t(outlen), data, C.c_ulonglong(n), key if key else None, C.c_size_t(len(key)))
        assert rc == 0
        assert bytes(out) == hashlib.blake2b(data, digest_size=outlen, key=key).digest()


def test_shim_randombytes_stream_is_the_seeded_xorshift(shim):
    """randombytes_buf stand-in: xorshift64 (13, 7, 17), byte = state >> 32, default seed 0x9E3779B97F4A7C15 —
    the same stream oracle/ref_shim/sodium_impl.c gives the CPU reference, so both provers draw identical instances."""
    def stream(seed, n):
        s, out = seed, bytearray()
        for _ in range(n):
            s ^= (s << 13) & (2**64 - 1); s ^= s >> 7; s ^= (s << 17) & (2**64 - 1)
            out.append((s >> 32) & 0xFF)
        return bytes(out)
    buf = (C.c_uint8 * 100)()
    shim.b200_shim_randombytes_seed(C.c_uint64(0))
    shim.randombytes_buf(buf, C.c_size_t(100))
    assert bytes(buf) == stream(0x9E3779B97F4A7C15, 100)
    shim.b200_shim_randombytes_seed(C.c_uint64(12345))
    shim.randombytes_buf(buf, C.c_size_t(37))
    assert bytes(buf)[:37] == stream(12345, 37)


def test_shim_chacha20_core_matches_rfc8439(shim):
    """the 8-block vector core of the default randombytes_buf against RFC 8439: the block function test vector (2.3.2: key
    00..1f, nonce 00:00:00:09:00:00:00:4a:00:00:00:00, counter 1) and the counter-mode continuation (lanes 1..7 = counters 2..8,
    checked against a scalar restatement of the block function)."""
    def block(key, counter, nonce):
        import struct
        def rotl(v, n): return ((v << n) | (v >> (32 - n))) & 0xffffffff
        st = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + list(struct.unpack("<8I", key)) + [counter] + list(struct.unpack("<3I", nonce))
        x = st[:]
        def qr(a, b, c, d):
            x[a] = (x[a] + x[b]) & 0xffffffff; x[d] = rotl(x[d] ^ x[a], 16)
            x[c] = (x[c] + x[d]) & 0xffffffff; x[b] = rotl(x[b] ^ x[c], 12)
            x[a] = (x[a] + x[b]) & 0xffffffff; x[d] = rotl(x[d] ^ x[a], 8)
            x[c] = (x[c] + x[d]) & 0xffffffff; x[b] = rotl(x[b] ^ x[c], 7)
        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        return struct.pack("<16I", *[(a + b) & 0xffffffff for a, b in zip(x, st)])
    key = bytes(range(32))
    nonce = bytes.fromhex("000000090000004a00000000")
    want0 = bytes.fromhex("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                          "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")
    assert block(key, 1, nonce) == want0                    # the restatement itself against the RFC vector
    out = C.create_string_buffer(512)
    shim.b200_shim_chacha20_keystream(key, C.c_uint32(1), nonce, out, C.c_size_t(8))
    assert out.raw[:64] == want0
    for lane in range(1, 8):
        assert out.raw[64 * lane:64 * lane + 64] == block(key, 1 + lane, nonce), lane


def test_shim_randombytes_is_unpredictable_unless_the_parity_switch_is_thrown(shim):
    """the default randombytes_buf must not be the seeded stream — zk masks and Merkle salts come from it.
    Default = ChaCha20 keyed from getrandom(2); the xorshift stream only after b200_shim_randombytes_deterministic/seed."""
    a, b = (C.c_uint8 * 64)(), (C.c_uint8 * 64)()
    fresh = C.CDLL(shim._name)                       # same mapping, state as loaded: nothing has switched the mode yet in a new process
    fresh.b200_shim_randombytes_secure()
    assert fresh.b200_shim_randombytes_is_deterministic() == 0
    fresh.randombytes_buf(a, C.c_size_t(64)); fresh.randombytes_buf(b, C.c_size_t(64))
    assert bytes(a) != bytes(b) and len(set(bytes(a))) > 16
    fresh.b200_shim_randombytes_deterministic(C.c_uint64(7))
    assert fresh.b200_shim_randombytes_is_deterministic() == 1
    fresh.randombytes_buf(a, C.c_size_t(64))
    fresh.b200_shim_randombytes_deterministic(C.c_uint64(7))
    fresh.randombytes_buf(b, C.c_size_t(64))
    assert bytes(a) == bytes(b)
    fresh.b200_shim_randombytes_secure()
    fresh.randombytes_buf(b, C.c_size_t(64))
    assert bytes(a) != bytes(b)
    # a process that never calls the switch starts in the secure mode
    code = ("import ctypes as C; l = C.CDLL(%r); print(l.b200_shim_randombytes_is_deterministic())" % shim._name)
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert out.stdout.strip() == "0", out.stderr


def test_dropin_library_exports_prover_entry_points():
    from libiop_b200 import prover
    if not prover.available():
        pytest.skip("libiop_b200/lib/libiop_b200_dropin.so not built (needs /root/reference at build time)")
    lib = C.CDLL(prover.LIB_PATH)
    for name in ("b200_prover_aurora_gf192", "b200_prover_aurora_edwards", "b200_prover_fractal_edwards", "b200_prover_ligero_gf192",
                 "b200_prover_seed", "b200_prover_arena_stats", "b200_prover_residency_stats", "b200_prover_device_stats"):
        assert hasattr(lib, name), name
    # nothing of the test oracle is linked into the product
    syms = subprocess.run(["nm", "-D", prover.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_every_shadow_file_includes_the_reference_file_it_shadows():
    """libiop_b200/dropin/libiop/<path>.tcc must pull in the unmodified reference <path>.tcc (no copied bodies)."""
    base = os.path.join(ROOT, "libiop_b200", "dropin", "libiop")
    n = 0
    for d, _, files in os.walk(base):
        for f in files:
            if f.endswith((".tcc", ".hpp")):
                rel = os.path.relpath(os.path.join(d, f), os.path.dirname(base))
                text = open(os.path.join(d, f)).read()
                assert f"#include B200_REF_FILE({rel})" in text, rel
                n += 1
    assert n >= 18


def test_dropin_host_logic_against_the_reference_on_cpu():
    """libiop_b200/dropin/examples/dropin_host_test.cpp (built by `make -C libiop_b200/dropin` where the reference checkout is):
    the flat zk-salt sampler consumes randombytes exactly like merkle_tree::sample_leaf_randomness, and the pruned
    authentication-path node list equals what the reference's get_set_membership_proof emits — no device involved."""
    exe = os.path.join(ROOT, "libiop_b200", "lib", "dropin_host_test")
    if not os.path.exists(exe):
        pytest.skip("libiop_b200/lib/dropin_host_test not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "0 failed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
