/*
 * oracle/fp.c — 3-limb Montgomery prime-field arithmetic, CPU restatement
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Restates libff::Fp_model<3, modulus> (depends/libff/libff/algebra/fields/prime_base/fp.tcc):
 *   mul_reduce  fp.tcc:25-194   -> a*b*R^-1 mod p, result in [0,p), R = 2^192
 *   operator+=  fp.tcc:323-437  -> (a+b) mod p
 *   operator-=  fp.tcc:440-532  -> (a-b) mod p
 *   invert      fp.tcc:670-722  -> a^-1 (unique; computed here as a^(p-2))
 * Parameters for edwards_Fr: depends/libff/libff/algebra/curves/edwards/edwards_fields.cpp:24-46.
 */
#include "oracle.h"
#include <string.h>

typedef unsigned __int128 u128;

static int ge3(const uint64_t a[3], const uint64_t b[3])
{
    for (int i = 2; i >= 0; --i) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}

static uint64_t sub3(uint64_t r[3], const uint64_t a[3], const uint64_t b[3])
{
    uint64_t borrow = 0;
    for (int i = 0; i < 3; ++i) {
        const u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

static uint64_t add3(uint64_t r[3], const uint64_t a[3], const uint64_t b[3])
{
    uint64_t carry = 0;
    for (int i = 0; i < 3; ++i) {
        const u128 s = (u128)a[i] + b[i] + carry;
        r[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    return carry;
}

/* CIOS Montgomery multiplication. */
void orc_fp_mul(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b)
{
    uint64_t t[5] = { 0, 0, 0, 0, 0 };
    for (int i = 0; i < 3; ++i) {
        uint64_t carry = 0;
        for (int j = 0; j < 3; ++j) {
            const u128 s = (u128)a->w[j] * b->w[i] + t[j] + carry;
            t[j] = (uint64_t)s; carry = (uint64_t)(s >> 64);
        }
        u128 s = (u128)t[3] + carry;
        t[3] = (uint64_t)s; t[4] = (uint64_t)(s >> 64);

        const uint64_t k = t[0] * P->inv;
        s = (u128)k * P->mod[0] + t[0];
        carry = (uint64_t)(s >> 64);
        for (int j = 1; j < 3; ++j) {
            s = (u128)k * P->mod[j] + t[j] + carry;
            t[j - 1] = (uint64_t)s; carry = (uint64_t)(s >> 64);
        }
        s = (u128)t[3] + carry;
        t[2] = (uint64_t)s;
        t[3] = t[4] + (uint64_t)(s >> 64);
        t[4] = 0;
    }
    if (t[3] || ge3(t, P->mod)) sub3(t, t, P->mod);
    r->w[0] = t[0]; r->w[1] = t[1]; r->w[2] = t[2];
}

void orc_fp_add(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b)
{
    uint64_t t[3];
    const uint64_t c = add3(t, a->w, b->w);
    if (c || ge3(t, P->mod)) sub3(t, t, P->mod);
    memcpy(r->w, t, sizeof t);
}

void orc_fp_sub(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b)
{
    uint64_t t[3];
    if (sub3(t, a->w, b->w)) add3(t, t, P->mod);
    memcpy(r->w, t, sizeof t);
}

void orc_fp_pow_u64(const orc_fp_params *P, f192 *r, const f192 *a, uint64_t e)
{
    f192 acc, base = *a;
    memcpy(acc.w, P->one, sizeof acc.w);
    while (e) {
        if (e & 1) orc_fp_mul(P, &acc, &acc, &base);
        orc_fp_mul(P, &base, &base, &base);
        e >>= 1;
    }
    *r = acc;
}

static void fp_pow_big(const orc_fp_params *P, f192 *r, const f192 *a, const uint64_t e[3])
{
    f192 acc, base = *a;
    memcpy(acc.w, P->one, sizeof acc.w);
    for (int i = 0; i < 192; ++i) {
        if ((e[i / 64] >> (i % 64)) & 1) orc_fp_mul(P, &acc, &acc, &base);
        orc_fp_mul(P, &base, &base, &base);
    }
    *r = acc;
}

void orc_fp_inv(const orc_fp_params *P, f192 *r, const f192 *a)
{
    uint64_t e[3], two[3] = { 2, 0, 0 };
    sub3(e, P->mod, two);
    fp_pow_big(P, r, a, e);
}

void orc_fp_from_u64(const orc_fp_params *P, f192 *r, uint64_t v)
{
    f192 t = {{ v, 0, 0 }}, r2;
    memcpy(r2.w, P->r2, sizeof r2.w);
    orc_fp_mul(P, r, &t, &r2);
}

void orc_fp_to_plain(const orc_fp_params *P, f192 *r, const f192 *a)
{
    f192 one = {{ 1, 0, 0 }};
    orc_fp_mul(P, r, a, &one);
}

int orc_fp_params_init(orc_fp_params *P, const uint64_t mod[3], uint64_t generator_small)
{
    memset(P, 0, sizeof *P);
    memcpy(P->mod, mod, sizeof P->mod);
    if (!(mod[0] & 1)) return -1;
    /* inv = -p^{-1} mod 2^64 by Newton iteration */
    uint64_t x = 1;
    for (int i = 0; i < 6; ++i) x *= 2 - mod[0] * x;
    P->inv = (uint64_t)0 - x;
    /* R mod p and R^2 mod p by repeated doubling of 1 */
    uint64_t t[3] = { 1, 0, 0 };
    for (int i = 0; i < 384; ++i) {
        const uint64_t c = add3(t, t, t);
        if (c || ge3(t, mod)) sub3(t, t, mod);
        if (i == 191) memcpy(P->one, t, sizeof t);
    }
    memcpy(P->r2, t, sizeof t);
    /* bit length */
    int nb = 192;
    while (nb > 0 && !((mod[(nb - 1) / 64] >> ((nb - 1) % 64)) & 1)) --nb;
    P->num_bits = (uint32_t)nb;
    /* 2-adicity of p-1 */
    uint64_t pm1[3], one[3] = { 1, 0, 0 };
    sub3(pm1, mod, one);
    uint32_t s = 0;
    while (s < 191 && !((pm1[s / 64] >> (s % 64)) & 1)) ++s;
    P->two_adicity = s;
    f192 g; orc_fp_from_u64(P, &g, generator_small);
    memcpy(P->generator, g.w, sizeof g.w);
    return 0;
}

/* edwards_Fr modulus r = 1552511030102430251236801561344621993261920897571225601,
   multiplicative generator 19 (edwards_fields.cpp:32-46). */
void orc_fp_params_edwards(orc_fp_params *P)
{
    /* decimal -> limbs by schoolbook multiply-add */
    const char *dec = "1552511030102430251236801561344621993261920897571225601";
    uint64_t m[3] = { 0, 0, 0 };
    for (const char *c = dec; *c; ++c) {
        u128 carry = (u128)(*c - '0');
        for (int i = 0; i < 3; ++i) {
            const u128 v = (u128)m[i] * 10 + carry;
            m[i] = (uint64_t)v; carry = v >> 64;
        }
    }
    orc_fp_params_init(P, m, 19);
}

void orc_fp_subgroup_generator(const orc_fp_params *P, f192 *g, uint32_t logn)
{
    /* g = generator^((p-1)/2^logn)  (subgroup.tcc:56-59) */
    uint64_t e[3], one[3] = { 1, 0, 0 };
    sub3(e, P->mod, one);
    for (uint32_t i = 0; i < logn; ++i) {
        e[0] = (e[0] >> 1) | (e[1] << 63);
        e[1] = (e[1] >> 1) | (e[2] << 63);
        e[2] >>= 1;
    }
    f192 gen; memcpy(gen.w, P->generator, sizeof gen.w);
    fp_pow_big(P, g, &gen, e);
}

void orc_fp_mul_vec(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b, size_t n)
{
    for (size_t i = 0; i < n; ++i) orc_fp_mul(P, &r[i], &a[i], &b[i]);
}
