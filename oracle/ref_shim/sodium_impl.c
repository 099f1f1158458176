/*
 * oracle/ref_shim/sodium_impl.c — the two libsodium symbols the reference links against
 * (SURVEY.md §0.3), provided for the oracle/_ref build (TEST INFRASTRUCTURE).
 *   crypto_generichash_blake2b -> oracle/blake2b.c (RFC 7693)
 *   randombytes_buf            -> deterministic xorshift64 stream, byte = state >> 32
 *                                 (default seed 0x9E3779B97F4A7C15), reseedable.
 */
#include <stddef.h>
#include <stdint.h>
#include "../oracle.h"

int crypto_generichash_blake2b(unsigned char *out, size_t outlen,
                               const unsigned char *in, unsigned long long inlen,
                               const unsigned char *key, size_t keylen)
{
    return orc_blake2b(out, outlen, in, (size_t)inlen, key, keylen);
}

static uint64_t g_state = 0x9E3779B97F4A7C15ull;

void b200_shim_randombytes_seed(uint64_t seed) { g_state = seed ? seed : 0x9E3779B97F4A7C15ull; }

void randombytes_buf(void *buf, size_t size)
{
    unsigned char *p = (unsigned char *)buf;
    for (size_t i = 0; i < size; ++i) {
        g_state ^= g_state << 13;
        g_state ^= g_state >> 7;
        g_state ^= g_state << 17;
        p[i] = (unsigned char)(g_state >> 32);
    }
}
