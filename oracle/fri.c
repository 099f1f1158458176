/*
 * oracle/fri.c — FRI commit-phase coset folding, CPU restatement
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * additive_evaluate_next_f_i_over_entire_domain        libiop/protocols/ldt/fri/fri_aux.tcc:37-103
 * multiplicative_evaluate_next_f_i_over_entire_domain  libiop/protocols/ldt/fri/fri_aux.tcc:107-249
 * vanishing_polynomial_from_subspace                   libiop/algebra/polynomials/vanishing_polynomial.tcc:374-395
 * linearized_polynomial::evaluation_at_point           libiop/algebra/polynomials/linearized_polynomial.tcc:30-48
 * batch_inverse_and_mul                                libiop/algebra/utils.tcc:38-103
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static inline void xadd(f192 *r, const f192 *a)
{
    r->w[0] ^= a->w[0]; r->w[1] ^= a->w[1]; r->w[2] ^= a->w[2];
}
static inline int is_zero(const f192 *a) { return !(a->w[0] | a->w[1] | a->w[2]); }
static inline int eq(const f192 *a, const f192 *b)
{
    return a->w[0] == b->w[0] && a->w[1] == b->w[1] && a->w[2] == b->w[2];
}

/* linearized polynomial: c[0] + c[1] X + c[2] X^2 + c[3] X^4 + ... (linearized_polynomial.tcc:30-48) */
static void linpoly_eval(f192 *r, const f192 *c, size_t len, const f192 *x)
{
    f192 res = c[0], xp = *x;
    for (size_t i = 1; i < len; ++i) {
        f192 t; orc_gf192_mul(&t, &c[i], &xp);
        xadd(&res, &t);
        orc_gf192_mul(&xp, &xp, &xp);
    }
    *r = res;
}

/* vanishing polynomial of span(basis[0..k)) (vanishing_polynomial.tcc:374-395, shift = 0) */
static size_t subspace_vp(f192 *c /* k+2 */, const f192 *basis, size_t k)
{
    size_t len = 2;
    memset(c, 0, (k + 2) * sizeof(f192));
    c[1].w[0] = 1;
    for (size_t b = 0; b < k; ++b) {
        f192 pc; linpoly_eval(&pc, c, len, &basis[b]);
        /* poly = poly^2 + poly * pc ; squaring shifts coefficients up by one and squares them */
        f192 sq[66], sc[66];
        memset(sq, 0, sizeof sq);
        for (size_t i = len; i > 0; --i) orc_gf192_mul(&sq[i], &c[i - 1], &c[i - 1]);
        /* linearized_polynomial::square keeps the constant term's square at index 1 only if the
           constant is non-zero; here the constant term is always zero (unshifted subspace). */
        sq[0] = (f192){{0,0,0}};
        for (size_t i = 0; i < len; ++i) orc_gf192_mul(&sc[i], &c[i], &pc);
        sc[len] = (f192){{0,0,0}};
        len += 1;
        for (size_t i = 0; i < len; ++i) { c[i] = sq[i]; xadd(&c[i], &sc[i]); }
    }
    return len;
}

/* utils.tcc:38-69 (gf192 instance) */
static void batch_inverse_and_mul_gf192(f192 *R, const f192 *vec, size_t n, const f192 *k)
{
    f192 c = vec[0];
    R[0] = c;
    for (size_t i = 1; i < n; ++i) { orc_gf192_mul(&c, &c, &vec[i]); R[i] = c; }
    f192 c_inv; orc_gf192_inv(&c_inv, &c);
    orc_gf192_mul(&c_inv, &c_inv, k);
    for (size_t i = n - 1; i > 0; --i) {
        orc_gf192_mul(&R[i], &R[i - 1], &c_inv);
        orc_gf192_mul(&c_inv, &c_inv, &vec[i]);
    }
    R[0] = c_inv;
}

/* fri_aux.tcc:37-103 */
void orc_fri_fold_additive(f192 *out, const f192 *f, uint32_t m, const f192 *basis,
                           const f192 *shift, uint32_t eta, const f192 *x)
{
    const size_t n = (size_t)1 << m, coset = (size_t)1 << eta, num_cosets = n / coset;
    f192 vp[66];
    const size_t vplen = subspace_vp(vp, basis, eta);
    f192 vp_x; linpoly_eval(&vp_x, vp, vplen, x);
    f192 inv_lin; orc_gf192_inv(&inv_lin, &vp[1]);

    f192 *elems = (f192 *)malloc(coset * sizeof(f192));
    f192 *diffs = (f192 *)malloc(coset * sizeof(f192));
    f192 *lag = (f192 *)malloc(coset * sizeof(f192));
    for (size_t j = 0; j < num_cosets; ++j) {
        /* elements of coset j: positions j*coset + k (subspace.tcc:87-91, 161-170) */
        for (size_t k = 0; k < coset; ++k) {
            const size_t idx = j * coset + k;
            f192 p = *shift;
            for (uint32_t b = 0; b < m; ++b) if ((idx >> b) & 1) xadd(&p, &basis[b]);
            elems[k] = p;
        }
        f192 vp_shift; linpoly_eval(&vp_shift, vp, vplen, &elems[0]);
        f192 shifted_vp_x = vp_x; xadd(&shifted_vp_x, &vp_shift);
        const int x_in_domain = is_zero(&shifted_vp_x);
        f192 interp = {{0,0,0}};
        int done = 0;
        for (size_t k = 0; k < coset; ++k) {
            if (x_in_domain && eq(x, &elems[k])) { interp = f[j * coset + k]; done = 1; break; }
            diffs[k] = *x; xadd(&diffs[k], &elems[k]);
        }
        if (!x_in_domain) {
            f192 kk; orc_gf192_mul(&kk, &inv_lin, &shifted_vp_x);
            batch_inverse_and_mul_gf192(lag, diffs, coset, &kk);
            for (size_t k = 0; k < coset; ++k) {
                f192 t; orc_gf192_mul(&t, &f[j * coset + k], &lag[k]);
                xadd(&interp, &t);
            }
        }
        (void)done;
        out[j] = interp;
    }
    free(lag); free(diffs); free(elems);
}

/* fri_aux.tcc:107-249.  Replicates the reference statement by statement, including the fact that
   when x lies in coset j the `continue` at fri_aux.tcc:205 skips the h update of lines 213-216. */
void orc_fri_fold_multiplicative(const orc_fp_params *P, f192 *out, const f192 *f, uint32_t logn,
                                 const f192 *g_dom, const f192 *shift, uint32_t eta, const f192 *x)
{
    const size_t n = (size_t)1 << logn, coset = (size_t)1 << eta, num_cosets = n / coset;
    f192 one; memcpy(one.w, P->one, sizeof one.w);
    const f192 zero = {{0,0,0}};

    const f192 h_inc = *g_dom;
    f192 t, h_inc_to_coset_inv_plus_one;
    orc_fp_pow_u64(P, &t, &h_inc, coset);
    orc_fp_inv(P, &t, &t);
    orc_fp_mul(P, &h_inc_to_coset_inv_plus_one, &t, &h_inc);
    /* generator of the order-`coset` subgroup (field_subset(coset_size, one)) */
    f192 g; orc_fp_subgroup_generator(P, &g, eta);
    f192 g_inv; orc_fp_inv(P, &g_inv, &g);
    f192 x_to_order; orc_fp_pow_u64(P, &x_to_order, x, coset);

    f192 *shifted_x = (f192 *)malloc(coset * sizeof(f192));
    shifted_x[0] = *x;
    for (size_t i = 1; i < coset; ++i) orc_fp_mul(P, &shifted_x[i], &shifted_x[i - 1], &g_inv);

    f192 cur_h = *shift;
    f192 first; orc_fp_pow_u64(P, &t, &cur_h, coset); orc_fp_inv(P, &t, &t);
    orc_fp_mul(P, &first, &t, &cur_h);
    f192 cur_const_plus_h; orc_fp_mul(P, &cur_const_plus_h, &x_to_order, &first);

    f192 *to_invert = (f192 *)malloc(n * sizeof(f192));
    f192 *coset_const = (f192 *)malloc(num_cosets * sizeof(f192));
    f192 all_const, cs; orc_fp_from_u64(P, &cs, (uint64_t)coset); orc_fp_inv(P, &all_const, &cs);
    int x_ever = 0; size_t x_coset_index = 0, x_index_in_domain = 0;

    for (size_t j = 0; j < num_cosets; ++j) {
        f192 cc; orc_fp_sub(P, &cc, &cur_const_plus_h, &cur_h);
        coset_const[j] = cc;
        if (eq(&cc, &zero)) {
            x_ever = 1; x_coset_index = j;
            f192 cur = cur_h;
            for (size_t k = 0; k < coset; ++k) {
                if (eq(&cur, x)) x_index_in_domain = k * num_cosets + j;
                orc_fp_mul(P, &cur, &cur, &g);
                to_invert[j * coset + k] = one;
            }
            continue;
        }
        for (size_t k = 0; k < coset; ++k)
            orc_fp_sub(P, &to_invert[j * coset + k], &shifted_x[k], &cur_h);
        orc_fp_mul(P, &cur_h, &cur_h, &h_inc);
        orc_fp_mul(P, &cur_const_plus_h, &cur_const_plus_h, &h_inc_to_coset_inv_plus_one);
    }
    /* batch_inverse_and_mul (utils.tcc:38-69) */
    f192 *R = (f192 *)malloc(n * sizeof(f192));
    {
        f192 c = to_invert[0]; R[0] = c;
        for (size_t i = 1; i < n; ++i) { orc_fp_mul(P, &c, &c, &to_invert[i]); R[i] = c; }
        f192 c_inv; orc_fp_inv(P, &c_inv, &c); orc_fp_mul(P, &c_inv, &c_inv, &all_const);
        for (size_t i = n - 1; i > 0; --i) {
            orc_fp_mul(P, &R[i], &R[i - 1], &c_inv);
            orc_fp_mul(P, &c_inv, &c_inv, &to_invert[i]);
        }
        R[0] = c_inv;
    }
    for (size_t j = 0; j < num_cosets; ++j) {
        f192 interp = zero;
        for (size_t k = 0; k < coset; ++k) {
            orc_fp_mul(P, &t, &f[k * num_cosets + j], &R[j * coset + k]);
            orc_fp_add(P, &interp, &interp, &t);
        }
        orc_fp_mul(P, &interp, &interp, &coset_const[j]);
        out[j] = interp;
    }
    if (x_ever) out[x_coset_index] = f[x_index_in_domain];
    free(R); free(coset_const); free(to_invert); free(shifted_x);
}
