/*
 * gf192.cuh — GF(2^192) = GF(2)[x]/(x^192 + x^7 + x^2 + x + 1) arithmetic for sm_100a.
 *
 * Same field and element encoding as libff::gf192 (depends/libff/libff/algebra/fields/binary/
 * gf192.hpp:19-20,98-99; mul gf192.cpp:72-184; inverse gf192.cpp:245-288).  NVIDIA GPUs have no
 * carry-less multiplier, so the 32x32 carry-less product is emulated on the integer-multiply pipe
 * ("multiplication with holes"): an operand limb is split into S residue classes (bit positions
 * congruent mod S); the ordinary integer product of two such pieces has the wanted GF(2) product
 * bits at positions congruent to (i+j) mod S, and the integer carries land only in the holes as long
 * as no position accumulates 2^S or more partial products.
 *
 *   S = 4 : 16 IMAD.WIDE per 32x32, products XOR-combined           (clmul32_s4)
 *   S = 5 : 25 IMAD.WIDE per 32x32, products summed inside the IMAD accumulator (no XORs):
 *           a class collects at most 7+6+6+6+6 = 31 < 32 partial products per position, except
 *           class 1 which could reach 32 and is therefore split into two chains  (clmul32_s5)
 * The 192x192 product is 3-way Karatsuba over 64-bit digits, each 64x64 a 2-way Karatsuba over
 * 32-bit limbs (18 clmul32), followed by reduction with x^192 = x^7 + x^2 + x + 1.
 *
 * Everything is __host__ __device__ so the arithmetic can be unit-tested on the CPU build box
 * (tests/cpp/field_host_test.cu) — the product path itself only ever runs it on the device.
 */
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#ifndef B200_GF_SPACING
#define B200_GF_SPACING 4 /* measured on B200: LDE 2^20->2^23 6.54 ms (S=4) vs 8.06 ms (S=5) */
#endif

namespace b200 {

struct gf192 {
    uint32_t w[6]; /* little-endian 32-bit limbs (same bytes as the reference's 3 x u64) */
};

#define B200_HD __host__ __device__ __forceinline__

B200_HD gf192 gf_zero() { gf192 r; for (int i = 0; i < 6; ++i) r.w[i] = 0; return r; }
B200_HD gf192 gf_one() { gf192 r = gf_zero(); r.w[0] = 1; return r; }
B200_HD gf192 gf_add(const gf192 &a, const gf192 &b)
{
    gf192 r;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.w[i] = a.w[i] ^ b.w[i];
    return r;
}
B200_HD void gf_addto(gf192 &a, const gf192 &b)
{
#pragma unroll
    for (int i = 0; i < 6; ++i) a.w[i] ^= b.w[i];
}
B200_HD bool gf_is_zero(const gf192 &a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3] | a.w[4] | a.w[5]) == 0; }
B200_HD bool gf_eq(const gf192 &a, const gf192 &b)
{
    return ((a.w[0] ^ b.w[0]) | (a.w[1] ^ b.w[1]) | (a.w[2] ^ b.w[2]) |
            (a.w[3] ^ b.w[3]) | (a.w[4] ^ b.w[4]) | (a.w[5] ^ b.w[5])) == 0;
}

/* 32x32 -> 64 multiply(-add): one IMAD.WIDE.U32.  On the device the PTX instruction is named explicitly: written as
   (uint64_t)a * b, nvcc 12.9 lowers through the generic 64x64 pattern and leaves an add of a zero uniform register on
   the high word of every product (96 wasted VIADDs per field multiplication in the SASS). */
B200_HD uint64_t mulw(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__) && !defined(B200_NO_PTX_MULW)
    uint64_t r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
#else
    return (uint64_t)a * (uint64_t)b;
#endif
}
B200_HD uint64_t madw(uint32_t a, uint32_t b, uint64_t c)
{
#if defined(__CUDA_ARCH__) && !defined(B200_NO_PTX_MULW)
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#else
    return (uint64_t)a * (uint64_t)b + c;
#endif
}

/* ---- spacing 4 -------------------------------------------------------------------------------- */
struct pieces4 { uint32_t p[4]; };
B200_HD pieces4 split4(uint32_t x)
{
    pieces4 r;
    r.p[0] = x & 0x11111111u; r.p[1] = x & 0x22222222u; r.p[2] = x & 0x44444444u; r.p[3] = x & 0x88888888u;
    return r;
}
B200_HD pieces4 split4x(uint32_t x, uint32_t y) /* pieces of x ^ y: one LOP3 each */
{
    pieces4 r;
    r.p[0] = (x ^ y) & 0x11111111u; r.p[1] = (x ^ y) & 0x22222222u;
    r.p[2] = (x ^ y) & 0x44444444u; r.p[3] = (x ^ y) & 0x88888888u;
    return r;
}
B200_HD uint64_t clmul32_s4(const pieces4 &a, const pieces4 &b)
{
    /* class k = (i + j) mod 4.  Two products of one class may share the IMAD accumulator when their
       position-count maxima do not coincide (i+j differs by 4): the sum then stays <= 15. */
    uint64_t c0 = madw(a.p[1], b.p[3], mulw(a.p[0], b.p[0])) ^ mulw(a.p[2], b.p[2]) ^ mulw(a.p[3], b.p[1]);
    uint64_t c1 = madw(a.p[2], b.p[3], mulw(a.p[0], b.p[1])) ^ madw(a.p[3], b.p[2], mulw(a.p[1], b.p[0]));
    uint64_t c2 = madw(a.p[3], b.p[3], mulw(a.p[0], b.p[2])) ^ mulw(a.p[1], b.p[1]) ^ mulw(a.p[2], b.p[0]);
    uint64_t c3 = mulw(a.p[0], b.p[3]) ^ mulw(a.p[1], b.p[2]) ^ mulw(a.p[2], b.p[1]) ^ mulw(a.p[3], b.p[0]);
    return (c0 & 0x1111111111111111ull) | (c1 & 0x2222222222222222ull) |
           (c2 & 0x4444444444444444ull) | (c3 & 0x8888888888888888ull);
}

/* ---- spacing 5 -------------------------------------------------------------------------------- */
struct pieces5 { uint32_t p[5]; };
#define B200_M5_0 0x42108421u /* bits 0,5,10,15,20,25,30 */
#define B200_M5_1 0x84210842u /* bits 1,6,...,31         */
#define B200_M5_2 0x08421084u /* bits 2,7,...,27         */
#define B200_M5_3 0x10842108u /* bits 3,8,...,28         */
#define B200_M5_4 0x21084210u /* bits 4,9,...,29         */
B200_HD pieces5 split5(uint32_t x)
{
    pieces5 r;
    r.p[0] = x & B200_M5_0; r.p[1] = x & B200_M5_1; r.p[2] = x & B200_M5_2; r.p[3] = x & B200_M5_3; r.p[4] = x & B200_M5_4;
    return r;
}
B200_HD pieces5 split5x(uint32_t x, uint32_t y)
{
    pieces5 r;
    r.p[0] = (x ^ y) & B200_M5_0; r.p[1] = (x ^ y) & B200_M5_1; r.p[2] = (x ^ y) & B200_M5_2;
    r.p[3] = (x ^ y) & B200_M5_3; r.p[4] = (x ^ y) & B200_M5_4;
    return r;
}
B200_HD uint64_t clmul32_s5(const pieces5 &a, const pieces5 &b)
{
    /* class k = (i + j) mod 5; five products per class chained through the IMAD accumulator.
       Worst-case partial products per position: classes 0,2: 31, classes 3,4: 30, class 1: 32 (split). */
    uint64_t c0 = madw(a.p[4], b.p[1], madw(a.p[3], b.p[2], madw(a.p[2], b.p[3], madw(a.p[1], b.p[4], mulw(a.p[0], b.p[0])))));
    uint64_t c1 = madw(a.p[4], b.p[2], madw(a.p[3], b.p[3], mulw(a.p[2], b.p[4]))) ^ madw(a.p[1], b.p[0], mulw(a.p[0], b.p[1]));
    uint64_t c2 = madw(a.p[4], b.p[3], madw(a.p[3], b.p[4], madw(a.p[2], b.p[0], madw(a.p[1], b.p[1], mulw(a.p[0], b.p[2])))));
    uint64_t c3 = madw(a.p[4], b.p[4], madw(a.p[3], b.p[0], madw(a.p[2], b.p[1], madw(a.p[1], b.p[2], mulw(a.p[0], b.p[3])))));
    uint64_t c4 = madw(a.p[4], b.p[0], madw(a.p[3], b.p[1], madw(a.p[2], b.p[2], madw(a.p[1], b.p[3], mulw(a.p[0], b.p[4])))));
    return (c0 & 0x1084210842108421ull) | (c1 & 0x2108421084210842ull) | (c2 & 0x4210842108421084ull) |
           (c3 & 0x8421084210842108ull) | (c4 & 0x0842108421084210ull);
}

#if B200_GF_SPACING == 5
typedef pieces5 pieces;
B200_HD pieces split(uint32_t x) { return split5(x); }
B200_HD pieces splitx(uint32_t x, uint32_t y) { return split5x(x, y); }
B200_HD uint64_t clmul32(const pieces &a, const pieces &b) { return clmul32_s5(a, b); }
#else
typedef pieces4 pieces;
B200_HD pieces split(uint32_t x) { return split4(x); }
B200_HD pieces splitx(uint32_t x, uint32_t y) { return split4x(x, y); }
B200_HD uint64_t clmul32(const pieces &a, const pieces &b) { return clmul32_s4(a, b); }
#endif

/* A 64-bit digit prepared for multiplication: pieces of its low limb, high limb and their XOR. */
struct digit64 { pieces lo, hi, mid; };
B200_HD digit64 prep_digit(uint32_t lo, uint32_t hi)
{
    digit64 d; d.lo = split(lo); d.hi = split(hi); d.mid = splitx(lo, hi); return d;
}

/* 64x64 -> 128 carry-less, 2-way Karatsuba; result as 4 x u32 (little endian) */
B200_HD void clmul64(const digit64 &a, const digit64 &b, uint32_t r[4])
{
    const uint64_t z0 = clmul32(a.lo, b.lo);
    const uint64_t z2 = clmul32(a.hi, b.hi);
    const uint64_t zm = clmul32(a.mid, b.mid) ^ z0 ^ z2;
    const uint64_t lo = z0 ^ (zm << 32);
    const uint64_t hi = z2 ^ (zm >> 32);
    r[0] = (uint32_t)lo; r[1] = (uint32_t)(lo >> 32); r[2] = (uint32_t)hi; r[3] = (uint32_t)(hi >> 32);
}

/* Operand prepared once and reused for several multiplications (e.g. a twiddle). */
struct gf192_prepared { digit64 d0, d1, d2, d01, d02, d12; };
B200_HD gf192_prepared gf_prepare(const gf192 &a)
{
    gf192_prepared p;
    p.d0 = prep_digit(a.w[0], a.w[1]);
    p.d1 = prep_digit(a.w[2], a.w[3]);
    p.d2 = prep_digit(a.w[4], a.w[5]);
    p.d01 = prep_digit(a.w[0] ^ a.w[2], a.w[1] ^ a.w[3]);
    p.d02 = prep_digit(a.w[0] ^ a.w[4], a.w[1] ^ a.w[5]);
    p.d12 = prep_digit(a.w[2] ^ a.w[4], a.w[3] ^ a.w[5]);
    return p;
}

/* reduce a 384-bit carry-less product modulo x^192 + x^7 + x^2 + x + 1 */
B200_HD gf192 gf_reduce(const uint32_t c[12])
{
    /* T = H * (x^7 + x^2 + x + 1), H = c[6..11] (191 bits) -> 198 bits in t[0..6] */
    uint32_t t[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const uint32_t h = (i < 6) ? c[6 + i] : 0u;
        const uint32_t l = (i > 0) ? c[5 + i] : 0u;
        t[i] = h ^ ((h << 1) | (l >> 31)) ^ ((h << 2) | (l >> 30)) ^ ((h << 7) | (l >> 25));
    }
    /* second fold: t[6] holds bits 192..197 of T (at most 7 bits) */
    const uint32_t h2 = t[6];
    gf192 r;
    r.w[0] = c[0] ^ t[0] ^ h2 ^ (h2 << 1) ^ (h2 << 2) ^ (h2 << 7);
#pragma unroll
    for (int i = 1; i < 6; ++i) r.w[i] = c[i] ^ t[i];
    return r;
}

B200_HD gf192 gf_mul_prepared(const gf192_prepared &a, const gf192_prepared &b)
{
    uint32_t p0[4], p1[4], p2[4], p01[4], p02[4], p12[4];
    clmul64(a.d0, b.d0, p0);
    clmul64(a.d1, b.d1, p1);
    clmul64(a.d2, b.d2, p2);
    clmul64(a.d01, b.d01, p01);
    clmul64(a.d02, b.d02, p02);
    clmul64(a.d12, b.d12, p12);
    /* c0 = p0, c1 = p01+p0+p1, c2 = p02+p0+p2+p1, c3 = p12+p1+p2, c4 = p2; c_i at 64*i bits */
    uint32_t c[12];
    uint32_t c1[4], c2[4], c3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c1[i] = p01[i] ^ p0[i] ^ p1[i];
        c2[i] = p02[i] ^ p0[i] ^ p2[i] ^ p1[i];
        c3[i] = p12[i] ^ p1[i] ^ p2[i];
    }
    c[0] = p0[0]; c[1] = p0[1];
    c[2] = p0[2] ^ c1[0]; c[3] = p0[3] ^ c1[1];
    c[4] = c1[2] ^ c2[0]; c[5] = c1[3] ^ c2[1];
    c[6] = c2[2] ^ c3[0]; c[7] = c2[3] ^ c3[1];
    c[8] = c3[2] ^ p2[0]; c[9] = c3[3] ^ p2[1];
    c[10] = p2[2]; c[11] = p2[3];
    return gf_reduce(c);
}

B200_HD gf192 gf_mul(const gf192 &a, const gf192 &b) { return gf_mul_prepared(gf_prepare(a), gf_prepare(b)); }
B200_HD gf192 gf_mul(const gf192 &a, const gf192_prepared &b) { return gf_mul_prepared(gf_prepare(a), b); }

/* Squaring is GF(2)-linear: spread the bits (insert zeros) and reduce. */
B200_HD uint32_t spread16(uint32_t x)
{
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
B200_HD gf192 gf_sqr(const gf192 &a)
{
    uint32_t c[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) { c[2 * i] = spread16(a.w[i]); c[2 * i + 1] = spread16(a.w[i] >> 16); }
    return gf_reduce(c);
}

/* a^(2^192 - 2), Itoh–Tsujii chain (same value as gf192::inverse, gf192.cpp:245-288; the inverse is unique) */
B200_HD gf192 gf_inv(const gf192 &el)
{
    /* t_k = el^(2^k - 1): t_{2k} = t_k^(2^k) * t_k ; t_{k+1} = t_k^2 * el */
    gf192 t1 = el;
    gf192 t2 = gf_mul(gf_sqr(t1), t1);
    gf192 t3 = gf_mul(gf_sqr(t2), el);
    gf192 x = t3;
    for (int i = 0; i < 3; ++i) x = gf_sqr(x);
    gf192 t6 = gf_mul(x, t3);
    x = t6;
    for (int i = 0; i < 6; ++i) x = gf_sqr(x);
    gf192 t12 = gf_mul(x, t6);
    x = t12;
    for (int i = 0; i < 12; ++i) x = gf_sqr(x);
    gf192 t24 = gf_mul(x, t12);
    x = t24;
    for (int i = 0; i < 24; ++i) x = gf_sqr(x);
    gf192 t48 = gf_mul(x, t24);
    x = t48;
    for (int i = 0; i < 48; ++i) x = gf_sqr(x);
    gf192 t96 = gf_mul(x, t48);
    x = t96;
    for (int i = 0; i < 48; ++i) x = gf_sqr(x);
    gf192 t144 = gf_mul(x, t48);
    x = t144;
    for (int i = 0; i < 24; ++i) x = gf_sqr(x);
    gf192 t168 = gf_mul(x, t24);
    x = t168;
    for (int i = 0; i < 12; ++i) x = gf_sqr(x);
    gf192 t180 = gf_mul(x, t12);
    x = t180;
    for (int i = 0; i < 6; ++i) x = gf_sqr(x);
    gf192 t186 = gf_mul(x, t6);
    x = t186;
    for (int i = 0; i < 3; ++i) x = gf_sqr(x);
    gf192 t189 = gf_mul(x, t3);
    x = gf_sqr(gf_sqr(t189));
    gf192 t191 = gf_mul(x, t2);
    return gf_sqr(t191); /* el^(2^192 - 2) */
}

} // namespace b200
