/*
 * oracle/fft.c — additive (Gao–Mateer) FFT over GF(2^192) and radix-2 NTT over Fp,
 * CPU restatement (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * additive_FFT          libiop/algebra/fft.tcc:39-124
 * additive_IFFT         libiop/algebra/fft.tcc:126-204
 * naive_FFT             libiop/algebra/fft.tcc:14-37
 * multiplicative_FFT_degree_aware   libiop/algebra/fft.tcc:236-317
 * multiplicative_IFFT_internal      libiop/algebra/fft.tcc:343-361 ->
 *      libfqfft basic_radix2_domain::iFFT/icosetFFT (basic_radix2_domain.tcc:49-74),
 *      _basic_serial_radix2_FFT (basic_radix2_domain_aux.tcc:45-79),
 *      _multiply_by_coset (basic_radix2_domain_aux.tcc:172-180)
 * all_subset_sums       libiop/algebra/utils.tcc:8-30
 * bitreverse_vector     libiop/algebra/utils.tcc:144-159
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static inline void xadd(f192 *r, const f192 *a)
{
    r->w[0] ^= a->w[0]; r->w[1] ^= a->w[1]; r->w[2] ^= a->w[2];
}

static size_t bitrev(size_t x, uint32_t bits)
{
    size_t r = 0;
    for (uint32_t i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

static void bitreverse_vec(f192 *a, uint32_t logn)
{
    const size_t n = (size_t)1 << logn;
    for (size_t k = 0; k < n; ++k) {
        const size_t rk = bitrev(k, logn);
        if (k < rk) { f192 t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
}

/* utils.tcc:8-30 */
static void all_subset_sums(f192 *out, const f192 *basis, size_t m, const f192 *shift)
{
    out[0] = *shift;
    size_t len = 1;
    for (size_t i = 0; i < m; ++i) {
        for (size_t j = 0; j < len; ++j) {
            out[len + j] = out[j];
            xadd(&out[len + j], &basis[i]);
        }
        len <<= 1;
    }
}

/* fft.tcc:39-124 */
void orc_additive_fft(f192 *S, const f192 *coeffs, size_t n_coeffs,
                      uint32_t m, const f192 *basis, const f192 *shift)
{
    const size_t n = (size_t)1 << m;
    memset(S, 0, n * sizeof(f192));
    memcpy(S, coeffs, (n_coeffs < n ? n_coeffs : n) * sizeof(f192));
    if (m == 0) return;

    f192 *recursed_betas = (f192 *)calloc((size_t)(m + 1) * m / 2 + 1, sizeof(f192));
    f192 *recursed_shifts = (f192 *)calloc(m, sizeof(f192));
    f192 *betas2 = (f192 *)malloc(m * sizeof(f192));
    memcpy(betas2, basis, m * sizeof(f192));
    f192 shift2 = *shift;
    size_t ptr = 0;

    for (uint32_t j = 0; j < m; ++j) {
        const f192 beta = betas2[m - 1 - j];
        f192 betai = {{ 1, 0, 0 }};
        /* twist by beta (fft.tcc:61-70) */
        for (size_t ofs = 0; ofs < n; ofs += ((size_t)1 << j)) {
            for (size_t p = 0; p < ((size_t)1 << j); ++p)
                orc_gf192_mul(&S[ofs + p], &S[ofs + p], &betai);
            orc_gf192_mul(&betai, &betai, &beta);
        }
        /* radix conversion (fft.tcc:72-83) */
        for (size_t stride = n / 4; stride >= ((size_t)1 << j) && stride > 0; stride >>= 1) {
            for (size_t ofs = 0; ofs < n; ofs += stride * 4)
                for (size_t i = 0; i < stride; ++i) {
                    xadd(&S[ofs + 2 * stride + i], &S[ofs + 3 * stride + i]);
                    xadd(&S[ofs + 1 * stride + i], &S[ofs + 2 * stride + i]);
                }
        }
        /* deltas for the reverse process (fft.tcc:85-96) */
        f192 betainv; orc_gf192_inv(&betainv, &beta);
        for (size_t i = 0; i + 1 + j < m; ++i) {
            f192 nb, sq;
            orc_gf192_mul(&nb, &betas2[i], &betainv);
            recursed_betas[ptr++] = nb;
            orc_gf192_mul(&sq, &nb, &nb);
            xadd(&sq, &nb);
            betas2[i] = sq;
        }
        f192 ns, sq;
        orc_gf192_mul(&ns, &shift2, &betainv);
        recursed_shifts[j] = ns;
        orc_gf192_mul(&sq, &ns, &ns);
        xadd(&sq, &ns);
        shift2 = sq;
    }

    bitreverse_vec(S, m);

    /* unwind the recursion (fft.tcc:101-120) */
    f192 *sums = (f192 *)malloc((n / 2 ? n / 2 : 1) * sizeof(f192));
    for (uint32_t j = 0; j < m; ++j) {
        ptr -= j;
        all_subset_sums(sums, &recursed_betas[ptr], j, &recursed_shifts[m - 1 - j]);
        const size_t stride = (size_t)1 << j;
        for (size_t ofs = 0; ofs < n; ofs += 2 * stride)
            for (size_t i = 0; i < stride; ++i) {
                f192 t;
                orc_gf192_mul(&t, &S[ofs + stride + i], &sums[i]);
                xadd(&S[ofs + i], &t);
                xadd(&S[ofs + stride + i], &S[ofs + i]);
            }
    }
    free(sums); free(betas2); free(recursed_shifts); free(recursed_betas);
}

/* fft.tcc:126-204 */
void orc_additive_ifft(f192 *S, const f192 *evals,
                       uint32_t m, const f192 *basis, const f192 *shift)
{
    const size_t n = (size_t)1 << m;
    memcpy(S, evals, n * sizeof(f192));
    if (m == 0) return;

    f192 *recursed_twists = (f192 *)calloc(m, sizeof(f192));
    f192 *betas2 = (f192 *)malloc(m * sizeof(f192));
    f192 *newbetas = (f192 *)malloc(m * sizeof(f192));
    f192 *sums = (f192 *)malloc((n / 2 ? n / 2 : 1) * sizeof(f192));
    memcpy(betas2, basis, m * sizeof(f192));
    f192 shift2 = *shift;

    for (uint32_t j = 0; j < m; ++j) {
        const f192 beta = betas2[m - 1 - j];
        f192 betainv; orc_gf192_inv(&betainv, &beta);
        recursed_twists[j] = betainv;
        for (size_t i = 0; i + 1 + j < m; ++i) {
            f192 nb, sq;
            orc_gf192_mul(&nb, &betas2[i], &betainv);
            newbetas[i] = nb;
            orc_gf192_mul(&sq, &nb, &nb);
            xadd(&sq, &nb);
            betas2[i] = sq;
        }
        f192 ns, sq;
        orc_gf192_mul(&ns, &shift2, &betainv);
        orc_gf192_mul(&sq, &ns, &ns);
        xadd(&sq, &ns);
        shift2 = sq;

        all_subset_sums(sums, newbetas, m - 1 - j, &ns);
        const size_t half = (size_t)1 << (m - 1 - j);
        for (size_t ofs = 0; ofs < n; ofs += 2 * half)
            for (size_t p = 0; p < half; ++p) {
                f192 t;
                xadd(&S[ofs + half + p], &S[ofs + p]);
                orc_gf192_mul(&t, &S[ofs + half + p], &sums[p]);
                xadd(&S[ofs + p], &t);
            }
    }

    bitreverse_vec(S, m);

    for (uint32_t j = 0; j < m; ++j) {
        size_t N = (size_t)4 << (m - 1 - j);
        /* radix combinations (fft.tcc:175-188) */
        while (N <= n) {
            const size_t quarter = N / 4;
            for (size_t ofs = 0; ofs < n; ofs += N)
                for (size_t i = 0; i < quarter; ++i) {
                    xadd(&S[ofs + 1 * quarter + i], &S[ofs + 2 * quarter + i]);
                    xadd(&S[ofs + 2 * quarter + i], &S[ofs + 3 * quarter + i]);
                }
            N *= 2;
        }
        /* twist by beta^-1 (fft.tcc:190-200) */
        const f192 betainv = recursed_twists[m - 1 - j];
        f192 betainvi = {{ 1, 0, 0 }};
        const size_t blk = (size_t)1 << (m - 1 - j);
        for (size_t ofs = 0; ofs < n; ofs += blk) {
            for (size_t p = 0; p < blk; ++p)
                orc_gf192_mul(&S[ofs + p], &S[ofs + p], &betainvi);
            orc_gf192_mul(&betainvi, &betainvi, &betainv);
        }
    }
    free(sums); free(newbetas); free(betas2); free(recursed_twists);
}

/* fft.tcc:14-37 with the domain enumeration of subspace.tcc:54-71,161-170 */
void orc_naive_fft_gf192(f192 *out, const f192 *coeffs, size_t n_coeffs,
                         uint32_t m, const f192 *basis, const f192 *shift)
{
    const size_t n = (size_t)1 << m;
    for (size_t idx = 0; idx < n; ++idx) {
        f192 p = *shift;
        for (uint32_t k = 0; k < m; ++k)
            if ((idx >> k) & 1) xadd(&p, &basis[k]);
        f192 v = {{ 0, 0, 0 }};
        for (size_t i = n_coeffs; i--; ) {
            orc_gf192_mul(&v, &v, &p);
            xadd(&v, &coeffs[i]);
        }
        out[idx] = v;
    }
}

/* fft.tcc:14-37 with the domain enumeration of subgroup.tcc:280-304 (shift * g^idx) */
void orc_naive_fft_fp(const orc_fp_params *P, f192 *out, const f192 *coeffs, size_t n_coeffs,
                      uint32_t logn, const f192 *g, const f192 *shift)
{
    const size_t n = (size_t)1 << logn;
    f192 p = *shift;
    for (size_t idx = 0; idx < n; ++idx) {
        f192 v = {{ 0, 0, 0 }};
        for (size_t i = n_coeffs; i--; ) {
            orc_fp_mul(P, &v, &v, &p);
            orc_fp_add(P, &v, &v, &coeffs[i]);
        }
        out[idx] = v;
        orc_fp_mul(P, &p, &p, g);
    }
}

static int is_one(const orc_fp_params *P, const f192 *a)
{
    return a->w[0] == P->one[0] && a->w[1] == P->one[1] && a->w[2] == P->one[2];
}

/* Cooley–Tukey DIT over `a` (bit-reversed input), twiddles = powers of omega.
   _basic_serial_radix2_FFT, basic_radix2_domain_aux.tcc:45-79 */
static void radix2_dit(const orc_fp_params *P, f192 *a, uint32_t logn, const f192 *omega)
{
    const size_t n = (size_t)1 << logn;
    bitreverse_vec(a, logn);
    size_t m = 1;
    for (uint32_t s = 1; s <= logn; ++s) {
        f192 w_m; orc_fp_pow_u64(P, &w_m, omega, n / (2 * m));
        for (size_t k = 0; k < n; k += 2 * m) {
            f192 w; memcpy(w.w, P->one, sizeof w.w);
            for (size_t j = 0; j < m; ++j) {
                f192 t;
                orc_fp_mul(P, &t, &w, &a[k + j + m]);
                orc_fp_sub(P, &a[k + j + m], &a[k + j], &t);
                orc_fp_add(P, &a[k + j], &a[k + j], &t);
                orc_fp_mul(P, &w, &w, &w_m);
            }
        }
        m *= 2;
    }
}

/* fft.tcc:236-317.  The degree-aware level skipping (fft.tcc:252-289) is an optimisation that
   leaves the output unchanged: the skipped levels combine with zero operands. */
void orc_mult_fft(const orc_fp_params *P, f192 *a, const f192 *coeffs, size_t n_coeffs,
                  uint32_t logn, const f192 *g, const f192 *shift)
{
    const size_t n = (size_t)1 << logn;
    memset(a, 0, n * sizeof(f192));
    memcpy(a, coeffs, n_coeffs * sizeof(f192));
    if (!is_one(P, shift)) {                 /* _multiply_by_coset, aux.tcc:172-180 */
        f192 u = *shift;
        for (size_t i = 1; i < n_coeffs; ++i) {
            orc_fp_mul(P, &a[i], &a[i], &u);
            orc_fp_mul(P, &u, &u, shift);
        }
    }
    radix2_dit(P, a, logn, g);
}

/* fft.tcc:343-361 */
void orc_mult_ifft(const orc_fp_params *P, f192 *a, const f192 *evals,
                   uint32_t logn, const f192 *g, const f192 *shift)
{
    const size_t n = (size_t)1 << logn;
    memcpy(a, evals, n * sizeof(f192));
    f192 ginv; orc_fp_inv(P, &ginv, g);
    radix2_dit(P, a, logn, &ginv);                     /* basic_radix2_domain.tcc:53 */
    f192 nn, sconst;
    orc_fp_from_u64(P, &nn, (uint64_t)n);
    orc_fp_inv(P, &sconst, &nn);                       /* basic_radix2_domain.tcc:55-59 */
    for (size_t i = 0; i < n; ++i) orc_fp_mul(P, &a[i], &a[i], &sconst);
    if (!is_one(P, shift)) {                           /* icosetFFT, basic_radix2_domain.tcc:70-74 */
        f192 sinv; orc_fp_inv(P, &sinv, shift);
        f192 u = sinv;
        for (size_t i = 1; i < n; ++i) {
            orc_fp_mul(P, &a[i], &a[i], &u);
            orc_fp_mul(P, &u, &u, &sinv);
        }
    }
}
