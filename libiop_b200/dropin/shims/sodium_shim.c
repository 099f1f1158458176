/*
 * sodium_shim.c — stand-ins for the two libsodium symbols libiop links against (INSTALL.md:10,21;
 * libiop/bcs/hashing/blake2b.cpp:15,37,62, blake2b.tcc:56,146,172,200; libff random_element via randombytes_buf),
 * for images without libsodium.  A deployment with libsodium installed links that instead and drops this file.
 *   crypto_generichash_blake2b : RFC 7693 BLAKE2b, sequential mode, optional key, 1..64 byte digests
 *   randombytes_buf            : ChaCha20 keystream keyed from getrandom(2) by default; the seeded xorshift64 stream of the
 *                                parity harness only after b200_shim_randombytes_deterministic() (see below).
 * Host-side control-plane hashing only (hash chain, query positions, verifier): bulk hashing runs on the device.
 */
#define _GNU_SOURCE
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static const uint64_t b2_iv[8] = {
    0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };

static inline uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

/* message word schedule: sigma[r][i], generated from the RFC's first ten permutations (rounds 10, 11 repeat 0, 1) */
static const uint8_t b2_sigma[10][16] = {
    {  0,  1,  2,  3,  4,  5,  6,  7,  8,  9, 10, 11, 12, 13, 14, 15 }, { 14, 10,  4,  8,  9, 15, 13,  6,  1, 12,  0,  2, 11,  7,  5,  3 },
    { 11,  8, 12,  0,  5,  2, 15, 13, 10, 14,  3,  6,  7,  1,  9,  4 }, {  7,  9,  3,  1, 13, 12, 11, 14,  2,  6,  5, 10,  4,  0, 15,  8 },
    {  9,  0,  5,  7,  2,  4, 10, 15, 14,  1, 11, 12,  6,  8,  3, 13 }, {  2, 12,  6, 10,  0, 11,  8,  3,  4, 13,  7,  5, 15, 14,  1,  9 },
    { 12,  5,  1, 15, 14, 13,  4, 10,  0,  7,  6,  3,  9,  2,  8, 11 }, { 13, 11,  7, 14, 12,  1,  3,  9,  5,  0, 15,  4,  8,  6,  2, 10 },
    {  6, 15, 14,  9, 11,  3,  0,  8, 12,  2, 13,  7,  1,  4, 10,  5 }, { 10,  2,  8,  4,  7,  6,  1,  5, 15, 11,  9, 14,  3, 12, 13,  0 } };

typedef struct { uint64_t h[8]; uint64_t t; uint8_t block[128]; size_t fill; } b2_state;

static void b2_compress(b2_state *s, int last)
{
    uint64_t m[16], v[16];
    memcpy(m, s->block, 128);                      /* little-endian host (x86-64) */
    for (int i = 0; i < 8; ++i) { v[i] = s->h[i]; v[8 + i] = b2_iv[i]; }
    v[12] ^= s->t;                                 /* message lengths below 2^64 bytes: high counter word is zero */
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 12; ++r) {
        const uint8_t *sg = b2_sigma[r % 10];
#define B2_MIX(a, b, c, d, x, y)                     \
        v[a] += v[b] + m[x]; v[d] = rotr64(v[d] ^ v[a], 32); \
        v[c] += v[d];        v[b] = rotr64(v[b] ^ v[c], 24); \
        v[a] += v[b] + m[y]; v[d] = rotr64(v[d] ^ v[a], 16); \
        v[c] += v[d];        v[b] = rotr64(v[b] ^ v[c], 63);
        B2_MIX(0, 4,  8, 12, sg[0],  sg[1])  B2_MIX(1, 5,  9, 13, sg[2],  sg[3])
        B2_MIX(2, 6, 10, 14, sg[4],  sg[5])  B2_MIX(3, 7, 11, 15, sg[6],  sg[7])
        B2_MIX(0, 5, 10, 15, sg[8],  sg[9])  B2_MIX(1, 6, 11, 12, sg[10], sg[11])
        B2_MIX(2, 7,  8, 13, sg[12], sg[13]) B2_MIX(3, 4,  9, 14, sg[14], sg[15])
#undef B2_MIX
    }
    for (int i = 0; i < 8; ++i) s->h[i] ^= v[i] ^ v[8 + i];
}

static void b2_absorb(b2_state *s, const uint8_t *in, size_t n)
{
    while (n) {
        if (s->fill == 128) { s->t += 128; b2_compress(s, 0); s->fill = 0; }   /* a full block is only final if no byte follows */
        const size_t take = (128 - s->fill < n) ? 128 - s->fill : n;
        memcpy(s->block + s->fill, in, take);
        s->fill += take; in += take; n -= take;
    }
}

int crypto_generichash_blake2b(unsigned char *out, size_t outlen, const unsigned char *in, unsigned long long inlen,
                               const unsigned char *key, size_t keylen)
{
    if (outlen == 0 || outlen > 64 || keylen > 64) return -1;
    b2_state s;
    memset(&s, 0, sizeof s);
    memcpy(s.h, b2_iv, sizeof s.h);
    s.h[0] ^= 0x01010000ULL ^ ((uint64_t)keylen << 8) ^ (uint64_t)outlen;   /* fanout = depth = 1 */
    if (keylen) { uint8_t kb[128] = {0}; memcpy(kb, key, keylen); b2_absorb(&s, kb, 128); }
    b2_absorb(&s, in, (size_t)inlen);
    s.t += s.fill;
    memset(s.block + s.fill, 0, 128 - s.fill);
    b2_compress(&s, 1);
    memcpy(out, s.h, outlen);
    return 0;
}

/* ---- randombytes_buf ---------------------------------------------------------------------------------------------
 * All of the reference's prover randomness comes through this symbol: the zero-knowledge masking polynomials
 * (gf192::random_element), Ligero's blinding rows and the Merkle leaf salts (merkle_tree.tcc:37-72).
 *   default        : a ChaCha20 keystream keyed from getrandom(2) (fast key erasure, re-keyed after fork) — unpredictable.
 *   deterministic  : the seeded xorshift64 stream of the parity harness (same stream as oracle/ref_shim/sodium_impl.c, so
 *                    the CPU reference and the drop-in draw identical instances and salts).  NOT zero-knowledge: only
 *                    b200_shim_randombytes_deterministic() switches it on, and the prover harness (prover_api.cpp) leaves
 *                    it on across a make_zk proof only when B200_DETERMINISTIC_RNG=1 is set (tests/conftest.py). */
#include <errno.h>
#include <stdlib.h>
#include <sys/random.h>
#include <sys/types.h>
#include <unistd.h>

static int shim_deterministic = 0;
static uint64_t shim_rng_state = 0x9E3779B97F4A7C15ULL;

void b200_shim_randombytes_deterministic(uint64_t seed)
{
    shim_deterministic = 1;
    shim_rng_state = seed ? seed : 0x9E3779B97F4A7C15ULL;
}
void b200_shim_randombytes_secure(void) { shim_deterministic = 0; }
int b200_shim_randombytes_is_deterministic(void) { return shim_deterministic; }
/* historical name used by the harnesses: same as ..._deterministic */
void b200_shim_randombytes_seed(uint64_t seed) { b200_shim_randombytes_deterministic(seed); }

/* ChaCha20 (RFC 8439 2.3), eight consecutive blocks at a time: the lane loops below are plain C that the compiler turns into
   8 x 32-bit vector code (the avx2 clone is picked at run time; the generic one is the same source). */
#include <pthread.h>
#define CC_LANES 8
typedef uint32_t cc_v8 __attribute__((vector_size(32)));        /* one state word of eight blocks (GCC vector extension) */
#define CC_ROTL(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define CC_QR(a, b, c, d) x[a] += x[b]; x[d] ^= x[a]; x[d] = CC_ROTL(x[d], 16); x[c] += x[d]; x[b] ^= x[c]; x[b] = CC_ROTL(x[b], 12); \
                          x[a] += x[b]; x[d] ^= x[a]; x[d] = CC_ROTL(x[d], 8);  x[c] += x[d]; x[b] ^= x[c]; x[b] = CC_ROTL(x[b], 7);
#define CC_BLOCKS8_BODY                                                                     \
    cc_v8 x[16], s[16];                                                                     \
    for (int i = 0; i < 16; ++i) s[i] = (cc_v8){ in[i], in[i], in[i], in[i], in[i], in[i], in[i], in[i] }; \
    s[12] += (cc_v8){ 0, 1, 2, 3, 4, 5, 6, 7 };                                             \
    for (int i = 0; i < 16; ++i) x[i] = s[i];                                               \
    for (int r = 0; r < 10; ++r) {                                                          \
        CC_QR(0, 4, 8, 12) CC_QR(1, 5, 9, 13) CC_QR(2, 6, 10, 14) CC_QR(3, 7, 11, 15)       \
        CC_QR(0, 5, 10, 15) CC_QR(1, 6, 11, 12) CC_QR(2, 7, 8, 13) CC_QR(3, 4, 9, 14)       \
    }                                                                                       \
    for (int i = 0; i < 16; ++i) {                                                          \
        x[i] += s[i];                                                                       \
        for (int l = 0; l < CC_LANES; ++l) { const uint32_t v = x[i][l]; memcpy(out + 64 * l + 4 * i, &v, 4); }   /* little-endian host */ \
    }
static void chacha20_blocks8_generic(const uint32_t in[16], uint8_t *out) { CC_BLOCKS8_BODY }
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static void chacha20_blocks8_avx2(const uint32_t in[16], uint8_t *out) { CC_BLOCKS8_BODY }
#endif
static void chacha20_blocks8(const uint32_t in[16], uint8_t *out)
{
#if defined(__x86_64__) && defined(__GNUC__)
    static int has_avx2 = -1;
    if (has_avx2 < 0) has_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
    if (has_avx2) { chacha20_blocks8_avx2(in, out); return; }
#endif
    chacha20_blocks8_generic(in, out);
}

/* Fast-key-erasure generator over a keystream pool: a refill produces CC_POOL bytes under the current key, the first 32 of
   which become the next key (so earlier output cannot be recomputed from a later state); requests are served from the rest
   and every byte handed out is wiped from the pool.  Small requests — one field element is 24 bytes, a zero-knowledge proof
   draws ~3 * 10^6 of them — cost a memcpy instead of two block functions each; large ones (Merkle salts: 2^20-byte batches)
   run at the vector code's rate.  The state is re-keyed from getrandom(2) in a forked child (pthread_atfork) and guarded by a
   mutex (libsodium's randombytes_buf is thread-safe too). */
#define CC_POOL 4096
static uint32_t cc_state[16];
static uint8_t cc_pool[CC_POOL];
static size_t cc_pool_pos = CC_POOL;
static int cc_keyed = 0;
static pthread_mutex_t cc_mu = PTHREAD_MUTEX_INITIALIZER;

static void cc_forget(void) { cc_keyed = 0; cc_pool_pos = CC_POOL; pthread_mutex_init(&cc_mu, NULL); }
static void secure_key(void)
{
    static int atfork_set = 0;
    uint8_t key[44];
    size_t got = 0;
    while (got < sizeof key) {
        const ssize_t r = getrandom(key + got, sizeof key - got, 0);
        if (r < 0) { if (errno == EINTR) continue; abort(); }   /* no entropy source: never fall back to something guessable */
        got += (size_t)r;
    }
    cc_state[0] = 0x61707865u; cc_state[1] = 0x3320646eu; cc_state[2] = 0x79622d32u; cc_state[3] = 0x6b206574u;
    memcpy(&cc_state[4], key, 32);
    cc_state[12] = 0; memcpy(&cc_state[13], key + 32, 12);
    memset(key, 0, sizeof key);
    cc_pool_pos = CC_POOL;
    if (!atfork_set) { pthread_atfork(NULL, NULL, cc_forget); atfork_set = 1; }
    cc_keyed = 1;
}
static void cc_refill(void)
{
    cc_state[12] = 0;
    for (size_t off = 0; off < CC_POOL; off += 64 * CC_LANES) { chacha20_blocks8(cc_state, cc_pool + off); cc_state[12] += CC_LANES; }
    memcpy(&cc_state[4], cc_pool, 32);
    memset(cc_pool, 0, 32);
    cc_pool_pos = 32;
}
static void secure_bytes(unsigned char *p, size_t size)
{
    pthread_mutex_lock(&cc_mu);
    if (!cc_keyed) secure_key();
    while (size) {
        if (cc_pool_pos == CC_POOL) cc_refill();
        const size_t avail = CC_POOL - cc_pool_pos, k = size < avail ? size : avail;
        memcpy(p, cc_pool + cc_pool_pos, k);
        memset(cc_pool + cc_pool_pos, 0, k);
        cc_pool_pos += k; p += k; size -= k;
    }
    pthread_mutex_unlock(&cc_mu);
}

/* keystream of ChaCha20 for (key, counter, nonce) through the 8-block core, for known-answer tests (RFC 8439 2.3.2 / 2.4.2):
   n_blocks <= 8 blocks of 64 bytes starting at `counter` */
void b200_shim_chacha20_keystream(const uint8_t key[32], uint32_t counter, const uint8_t nonce[12], uint8_t *out, size_t n_blocks)
{
    uint32_t st[16] = { 0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u };
    uint8_t all[64 * CC_LANES];
    memcpy(&st[4], key, 32); st[12] = counter; memcpy(&st[13], nonce, 12);
    chacha20_blocks8(st, all);
    memcpy(out, all, 64 * (n_blocks < CC_LANES ? n_blocks : CC_LANES));
}

void randombytes_buf(void *buf, size_t size)
{
    unsigned char *p = (unsigned char *)buf;
    if (!shim_deterministic) { secure_bytes(p, size); return; }
    for (size_t i = 0; i < size; ++i) {
        shim_rng_state ^= shim_rng_state << 13;
        shim_rng_state ^= shim_rng_state >> 7;
        shim_rng_state ^= shim_rng_state << 17;
        p[i] = (unsigned char)(shim_rng_state >> 32);
    }
}
