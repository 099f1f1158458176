/*
 * oracle/gf192.c — GF(2^192) arithmetic, CPU restatement (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Field: GF(2)[x]/(x^192 + x^7 + x^2 + x + 1), 3 x u64 little-endian limbs
 * (depends/libff/libff/algebra/fields/binary/gf192.hpp:19-20,98-99).
 * Pinned by tests/test_oracle.py against the ten libff multiplication KATs
 * (depends/libff/libff/algebra/fields/tests/test_binary_fields.cpp:124-135).
 */
#include "oracle.h"
#include <string.h>

#if defined(__PCLMUL__)
#include <immintrin.h>
#endif

/* Bit-serial shift-and-add multiplication: restates the portable (non-USE_ASM) branch of
   gf192::operator*= (gf192.cpp:143-177). Used as the ground truth for the fast path below. */
void orc_gf192_mul_bitwise(f192 *r, const f192 *a, const f192 *b)
{
    uint64_t sh[3] = { a->w[0], a->w[1], a->w[2] };
    uint64_t acc[3] = { 0, 0, 0 };
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 64; ++j) {
            if ((b->w[i] >> j) & 1) { acc[0] ^= sh[0]; acc[1] ^= sh[1]; acc[2] ^= sh[2]; }
            const uint64_t carry = sh[2] >> 63;
            sh[2] = (sh[2] << 1) | (sh[1] >> 63);
            sh[1] = (sh[1] << 1) | (sh[0] >> 63);
            sh[0] = (sh[0] << 1) ^ (carry ? 0x87ull : 0ull);
        }
    }
    r->w[0] = acc[0]; r->w[1] = acc[1]; r->w[2] = acc[2];
}

/* 64x64 -> 128 carry-less product */
static inline void clmul64(uint64_t a, uint64_t b, uint64_t *lo, uint64_t *hi)
{
#if defined(__PCLMUL__)
    const __m128i p = _mm_clmulepi64_si128(_mm_set_epi64x(0, (long long)a),
                                           _mm_set_epi64x(0, (long long)b), 0x00);
    *lo = (uint64_t)_mm_cvtsi128_si64(p);
    *hi = (uint64_t)_mm_extract_epi64(p, 1);
#else
    uint64_t l = 0, h = 0;
    for (int i = 0; i < 64; ++i) {
        if ((b >> i) & 1) {
            l ^= a << i;
            if (i) h ^= a >> (64 - i);
        }
    }
    *lo = l; *hi = h;
#endif
}

/* Schoolbook 3x3 limb product followed by reduction modulo x^192 + x^7 + x^2 + x + 1.
   Same function as gf192::operator*= (gf192.cpp:72-184); the product in GF(2^192) is unique,
   so the choice of multiplication schedule does not affect the bytes. */
void orc_gf192_mul(f192 *r, const f192 *a, const f192 *b)
{
    uint64_t c[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            uint64_t lo, hi;
            clmul64(a->w[i], b->w[j], &lo, &hi);
            c[i + j] ^= lo;
            c[i + j + 1] ^= hi;
        }
    /* fold limbs 5,4,3 down: x^192 = x^7 + x^2 + x + 1 */
    for (int i = 5; i >= 3; --i) {
        const uint64_t h = c[i];
        c[i - 3] ^= h ^ (h << 1) ^ (h << 2) ^ (h << 7);
        c[i - 2] ^= (h >> 63) ^ (h >> 62) ^ (h >> 57);
    }
    r->w[0] = c[0]; r->w[1] = c[1]; r->w[2] = c[2];
}

/* a^(2^192-2) with the reference's addition chain (gf192.cpp:245-288). */
void orc_gf192_inv(f192 *r, const f192 *el)
{
    f192 a = *el, result = {{0,0,0}}, prev = {{0,0,0}}, b;
    for (int i = 0; i <= 6; ++i) {
        b = a;
        for (uint64_t j = 0; j < (1ull << i); ++j) orc_gf192_mul(&b, &b, &b);
        orc_gf192_mul(&a, &a, &b);
        prev = result;
        if (i == 0) result = b; else orc_gf192_mul(&result, &result, &b);
    }
    for (int i = 0; i < 64; ++i) orc_gf192_mul(&result, &result, &result);
    orc_gf192_mul(&prev, &prev, &prev);
    f192 sq; orc_gf192_mul(&sq, el, el);
    orc_gf192_mul(&result, &result, &prev);
    orc_gf192_mul(r, &result, &sq);
}

void orc_gf192_mul_vec(f192 *r, const f192 *a, const f192 *b, size_t n)
{
    for (size_t i = 0; i < n; ++i) orc_gf192_mul(&r[i], &a[i], &b[i]);
}

void orc_gf192_inv_vec(f192 *r, const f192 *a, size_t n)
{
    for (size_t i = 0; i < n; ++i) orc_gf192_inv(&r[i], &a[i]);
}
