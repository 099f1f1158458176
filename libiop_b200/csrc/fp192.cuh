/*
 * fp192.cuh — 3x64-bit (6x32-bit limb) Montgomery prime-field arithmetic for sm_100a.
 *
 * Same values and encoding as libff::Fp_model<3, modulus> (depends/libff/libff/algebra/fields/
 * prime_base/fp.tcc: mul_reduce 25-194, += 323-437, -= 440-532, invert 670-722): elements are kept
 * in Montgomery form a*R mod p with R = 2^192 (fp.hpp:42), fully reduced to [0, p).
 * The modulus is a run-time parameter (b200_fp_params); edwards_Fr is the benchmark instance
 * (depends/libff/libff/algebra/curves/edwards/edwards_fields.cpp:24-46).
 *
 * __host__ __device__ so the arithmetic can be unit-tested on the CPU build box.
 */
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace b200 {

struct fp192 { uint32_t w[6]; };

struct fp_dev_params {
    uint32_t mod[6];
    uint32_t one[6];  /* R mod p */
    uint32_t r2[6];   /* R^2 mod p */
    uint32_t inv32;   /* -p^-1 mod 2^32 */
};

#define B200_FHD __host__ __device__ __forceinline__

B200_FHD fp192 fp_zero() { fp192 r; for (int i = 0; i < 6; ++i) r.w[i] = 0; return r; }
B200_FHD bool fp_eq(const fp192 &a, const fp192 &b)
{
    uint32_t x = 0;
    for (int i = 0; i < 6; ++i) x |= a.w[i] ^ b.w[i];
    return x == 0;
}
B200_FHD bool fp_is_zero(const fp192 &a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3] | a.w[4] | a.w[5]) == 0; }

/* r = a - b, returns borrow */
B200_FHD uint32_t sub6(uint32_t r[6], const uint32_t a[6], const uint32_t b[6])
{
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const uint64_t d = (uint64_t)a[i] - b[i] - borrow;
        r[i] = (uint32_t)d;
        borrow = (d >> 32) & 1u;
    }
    return (uint32_t)borrow;
}
B200_FHD uint32_t add6(uint32_t r[6], const uint32_t a[6], const uint32_t b[6])
{
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const uint64_t s = (uint64_t)a[i] + b[i] + carry;
        r[i] = (uint32_t)s;
        carry = s >> 32;
    }
    return (uint32_t)carry;
}
B200_FHD bool geq6(const uint32_t a[6], const uint32_t b[6])
{
    uint32_t t[6];
    return sub6(t, a, b) == 0;
}

B200_FHD fp192 fp_add(const fp192 &a, const fp192 &b, const fp_dev_params &P)
{
    fp192 s, t;
    const uint32_t carry = add6(s.w, a.w, b.w);
    const uint32_t borrow = sub6(t.w, s.w, P.mod);
    return (carry || !borrow) ? t : s;
}
B200_FHD fp192 fp_sub(const fp192 &a, const fp192 &b, const fp_dev_params &P)
{
    fp192 d, t;
    const uint32_t borrow = sub6(d.w, a.w, b.w);
    add6(t.w, d.w, P.mod);
    return borrow ? t : d;
}
B200_FHD fp192 fp_neg(const fp192 &a, const fp_dev_params &P)
{
    if (fp_is_zero(a)) return a;
    fp192 r; sub6(r.w, P.mod, a.w); return r;
}

/* CIOS Montgomery product a*b*R^-1 mod p, 32-bit limbs */
B200_FHD fp192 fp_mul(const fp192 &a, const fp192 &b, const fp_dev_params &P)
{
    uint32_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const uint64_t s = (uint64_t)a.w[j] * b.w[i] + t[j] + carry;
            t[j] = (uint32_t)s; carry = s >> 32;
        }
        uint64_t s = (uint64_t)t[6] + carry;
        t[6] = (uint32_t)s; t[7] = (uint32_t)(s >> 32);

        const uint32_t m = t[0] * P.inv32;
        s = (uint64_t)m * P.mod[0] + t[0];
        carry = s >> 32;
#pragma unroll
        for (int j = 1; j < 6; ++j) {
            s = (uint64_t)m * P.mod[j] + t[j] + carry;
            t[j - 1] = (uint32_t)s; carry = s >> 32;
        }
        s = (uint64_t)t[6] + carry;
        t[5] = (uint32_t)s;
        t[6] = t[7] + (uint32_t)(s >> 32);
    }
    fp192 r, u;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.w[i] = t[i];
    const uint32_t borrow = sub6(u.w, r.w, P.mod);
    return (t[6] || !borrow) ? u : r;
}

B200_FHD fp192 fp_one(const fp_dev_params &P) { fp192 r; for (int i = 0; i < 6; ++i) r.w[i] = P.one[i]; return r; }

B200_FHD fp192 fp_pow_u64(fp192 base, uint64_t e, const fp_dev_params &P)
{
    fp192 acc = fp_one(P);
    while (e) {
        if (e & 1) acc = fp_mul(acc, base, P);
        base = fp_mul(base, base, P);
        e >>= 1;
    }
    return acc;
}

/* a^(p-2) */
B200_FHD fp192 fp_inv(const fp192 &a, const fp_dev_params &P)
{
    uint32_t e[6], two[6] = {2, 0, 0, 0, 0, 0};
    sub6(e, P.mod, two);
    fp192 acc = fp_one(P), base = a;
    for (int i = 0; i < 192; ++i) {
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fp_mul(acc, base, P);
        base = fp_mul(base, base, P);
    }
    return acc;
}

B200_FHD fp192 fp_from_u64(uint64_t v, const fp_dev_params &P)
{
    fp192 t = fp_zero(), r2;
    t.w[0] = (uint32_t)v; t.w[1] = (uint32_t)(v >> 32);
    for (int i = 0; i < 6; ++i) r2.w[i] = P.r2[i];
    return fp_mul(t, r2, P);
}

inline fp_dev_params make_fp_dev_params(const uint64_t mod[3], const uint64_t r2[3], const uint64_t one[3], uint64_t inv)
{
    fp_dev_params d;
    for (int i = 0; i < 3; ++i) {
        d.mod[2 * i] = (uint32_t)mod[i]; d.mod[2 * i + 1] = (uint32_t)(mod[i] >> 32);
        d.r2[2 * i] = (uint32_t)r2[i]; d.r2[2 * i + 1] = (uint32_t)(r2[i] >> 32);
        d.one[2 * i] = (uint32_t)one[i]; d.one[2 * i + 1] = (uint32_t)(one[i] >> 32);
    }
    d.inv32 = (uint32_t)inv;
    return d;
}

} // namespace b200
