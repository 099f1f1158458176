/*
 * fri.cu — FRI commit-phase coset folding, sm_100a.
 *
 * Replaces evaluate_next_f_i_over_entire_domain (libiop/protocols/ldt/fri/fri_aux.tcc:6-34):
 *   additive       fri_aux.tcc:37-103   cosets = 2^eta consecutive positions      (subspace.tcc:74-91)
 *   multiplicative fri_aux.tcc:107-249  cosets = positions j + k*|L|/2^eta         (subgroup.tcc:176-197)
 * Output j is the value at x of the unique polynomial of degree < 2^eta interpolating f on coset j.
 *
 * The reference spends one field inversion per coset (additive, ~208 multiplications) or a global
 * batch inversion (multiplicative).  Neither is needed: with a_k = x - v_k,
 *     P(x) = sum_k f_k * prod_{k' != k} a_k' / prod_{k' != k} (v_k - v_k')
 * and the denominators do not depend on the coset (additive: the product of the non-zero elements of
 * span(basis[0..eta)), i.e. the linear coefficient of its vanishing polynomial, fri_aux.tcc:60-61;
 * multiplicative: 2^eta * h^(2^eta-1) * g^-k).  One thread folds one coset with 4*2^eta + O(1)
 * multiplications and no inversion; when x lies in a coset the formula returns f(x) by itself.
 */
#include <algorithm>
#include <cstring>

#include "ctx.h"
#include "fp192.cuh"
#include "gf192.cuh"

namespace b200 {

#define B200_FOLD_MAX_ETA 5

struct AddFoldParams {
    const gf192 *f;
    gf192 *out;
    gf192 basis[32];     /* basis[b] for index bit b (b >= eta used for the coset offset) */
    gf192 w[1 << B200_FOLD_MAX_ETA]; /* subset sums of basis[0..eta) */
    gf192 x_plus_shift;  /* x - shift */
    gf192 denom_inv;     /* (prod of non-zero elements of span(basis[0..eta)))^-1 */
    uint32_t m, eta;
    size_t n_out;
};

__device__ __noinline__ gf192 gf_mul_ni(const gf192 &a, const gf192 &b) { return gf_mul(a, b); }

/* generic localization parameter (eta >= 3): looped, arrays in local memory, shared multiplier body */
__global__ void __launch_bounds__(128) fold_additive_generic(const AddFoldParams P)
{
    const int S = 1 << P.eta;
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= P.n_out) return;
    gf192 a0 = P.x_plus_shift;
    for (uint32_t b = P.eta; b < P.m; ++b)
        if ((j >> (b - P.eta)) & 1u) gf_addto(a0, P.basis[b]);
    gf192 a[1 << B200_FOLD_MAX_ETA], pre[1 << B200_FOLD_MAX_ETA];
    const uint64_t *src = reinterpret_cast<const uint64_t *>(P.f + j * S);
#pragma unroll 1
    for (int k = 0; k < S; ++k) {
        a[k] = gf_add(a0, P.w[k]);
        pre[k] = k ? gf_mul_ni(pre[k - 1], a[k]) : a[0];
    }
    gf192 acc = gf_zero(), suf = gf_one();
#pragma unroll 1
    for (int k = S - 1; k >= 0; --k) {       /* N_k = pre[k-1] * suf(k+1..) */
        gf192 fv;
        const uint64_t q0 = __ldg(src + 3 * k), q1 = __ldg(src + 3 * k + 1), q2 = __ldg(src + 3 * k + 2);
        fv.w[0] = (uint32_t)q0; fv.w[1] = (uint32_t)(q0 >> 32);
        fv.w[2] = (uint32_t)q1; fv.w[3] = (uint32_t)(q1 >> 32);
        fv.w[4] = (uint32_t)q2; fv.w[5] = (uint32_t)(q2 >> 32);
        const gf192 nk = k ? gf_mul_ni(pre[k - 1], suf) : suf;
        gf_addto(acc, gf_mul_ni(fv, nk));
        suf = gf_mul_ni(suf, a[k]);
    }
    acc = gf_mul_ni(acc, P.denom_inv);
    uint64_t *dst = reinterpret_cast<uint64_t *>(P.out + j);
    dst[0] = (uint64_t)acc.w[0] | ((uint64_t)acc.w[1] << 32);
    dst[1] = (uint64_t)acc.w[2] | ((uint64_t)acc.w[3] << 32);
    dst[2] = (uint64_t)acc.w[4] | ((uint64_t)acc.w[5] << 32);
}

/* eta = 1, 2: every quantity that depends on the coset is an F2-linear function of the coset offset
   (z = x - v_0, z*c, (z^2 + b0 z)*c are additive in z), so it is rebuilt from per-basis-vector tables
   with XORs only.
     eta = 1:  P(x) = f0 + (f0 + f1) * t,                       t = z / b0                    (1 multiplication)
     eta = 2:  with q(Z) = Z^2 + b0 Z, s = q(z)/q(b1), t = z/b0, c2 = b1/b0:
               P(x) = A + t * B,  A = f0 + (f0 + f2 + c2 (f2+f3)) s,  B = (f0+f1) + (f0+f1+f2+f3) s   (4 multiplications) */
struct AddFoldLinParams {
    const gf192 *f;
    gf192 *out;
    gf192 t_tab[32];   /* t(basis[b]) for index bit b >= eta */
    gf192 s_tab[32];   /* s(basis[b]) (eta = 2) */
    gf192 t0, s0;      /* t(x - shift), s(x - shift) */
    gf192 c2;          /* b1 / b0 (eta = 2) */
    uint32_t m, eta;
    size_t n_out;
};

__device__ __forceinline__ gf192 ld_elem(const gf192 *p)
{
    const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
    const uint64_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    gf192 r;
    r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
    r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
    r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
    return r;
}
__device__ __forceinline__ void st_elem(gf192 *p, const gf192 &v)
{
    uint64_t *dst = reinterpret_cast<uint64_t *>(p);
    dst[0] = (uint64_t)v.w[0] | ((uint64_t)v.w[1] << 32);
    dst[1] = (uint64_t)v.w[2] | ((uint64_t)v.w[3] << 32);
    dst[2] = (uint64_t)v.w[4] | ((uint64_t)v.w[5] << 32);
}

__global__ void __launch_bounds__(256) fold_additive_eta1(const AddFoldLinParams P)
{
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= P.n_out) return;
    gf192 t = P.t0;
    for (uint32_t b = 1; b < P.m; ++b)
        if ((j >> (b - 1)) & 1u) gf_addto(t, P.t_tab[b]);
    const gf192 f0 = ld_elem(P.f + 2 * j), f1 = ld_elem(P.f + 2 * j + 1);
    st_elem(P.out + j, gf_add(f0, gf_mul(gf_add(f0, f1), t)));
}

__global__ void __launch_bounds__(256) fold_additive_eta2(const AddFoldLinParams P)
{
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= P.n_out) return;
    gf192 t = P.t0, s = P.s0;
    for (uint32_t b = 2; b < P.m; ++b)
        if ((j >> (b - 2)) & 1u) { gf_addto(t, P.t_tab[b]); gf_addto(s, P.s_tab[b]); }
    const gf192 f0 = ld_elem(P.f + 4 * j), f1 = ld_elem(P.f + 4 * j + 1);
    const gf192 f2 = ld_elem(P.f + 4 * j + 2), f3 = ld_elem(P.f + 4 * j + 3);
    const gf192 f01 = gf_add(f0, f1), f23 = gf_add(f2, f3);
    const gf192 a1 = gf_add(f2, gf_mul(f23, P.c2));                 /* A(q(b1)) */
    const gf192 A = gf_add(f0, gf_mul(gf_add(f0, a1), s));
    const gf192 B = gf_add(f01, gf_mul(gf_add(f01, f23), s));       /* b0 * B(q(z)) */
    st_elem(P.out + j, gf_add(A, gf_mul(t, B)));
}

static int fold_additive(b200_ctx *c, const b200_domain *L, uint32_t eta, const uint64_t x[3], b200_buf f, b200_buf out)
{
    const uint32_t m = L->dim;
    if (eta == 0 || eta > m || eta > B200_FOLD_MAX_ETA)
        return set_err(c, B200_ERR_INVALID, "unsupported localization parameter %u for a 2^%u domain", eta, m);
    if (m > 31) return set_err(c, B200_ERR_INVALID, "domain too large");
    AddFoldParams P; memset(&P, 0, sizeof P);
    P.f = (const gf192 *)f; P.out = (gf192 *)out; P.m = m; P.eta = eta; P.n_out = (size_t)1 << (m - eta);
    memcpy(P.basis, L->basis, (size_t)m * sizeof(gf192));
    gf192 xs, sh; memcpy(&xs, x, 24); memcpy(&sh, L->shift, 24);
    P.x_plus_shift = gf_add(xs, sh);
    const uint32_t S = 1u << eta;
    P.w[0] = gf_zero();
    for (uint32_t v = 1; v < S; ++v) {
        const uint32_t lowbit = v & (~v + 1u);
        P.w[v] = gf_add(P.w[v ^ lowbit], P.basis[31 - __builtin_clz(lowbit)]);
    }
    gf192 prod = gf_one();
    for (uint32_t v = 1; v < S; ++v) {
        if (gf_is_zero(P.w[v])) return set_err(c, B200_ERR_INVALID, "coset basis is linearly dependent");
        prod = gf_mul(prod, P.w[v]);
    }
    P.denom_inv = gf_inv(prod);
    if (eta <= 2) {
        AddFoldLinParams Q; memset(&Q, 0, sizeof Q);
        Q.f = P.f; Q.out = P.out; Q.m = m; Q.eta = eta; Q.n_out = P.n_out;
        const gf192 b0 = P.basis[0];
        const gf192 c0 = gf_inv(b0);
        gf192 c1 = gf_zero();
        if (eta == 2) {
            const gf192 b1 = P.basis[1];
            c1 = gf_inv(gf_add(gf_sqr(b1), gf_mul(b0, b1)));        /* 1 / q(b1) */
            Q.c2 = gf_mul(b1, c0);
        }
        auto tfun = [&](const gf192 &z) { return gf_mul(z, c0); };
        auto sfun = [&](const gf192 &z) { return gf_mul(gf_add(gf_sqr(z), gf_mul(b0, z)), c1); };
        Q.t0 = tfun(P.x_plus_shift); Q.s0 = sfun(P.x_plus_shift);
        for (uint32_t b = eta; b < m; ++b) { Q.t_tab[b] = tfun(P.basis[b]); Q.s_tab[b] = sfun(P.basis[b]); }
        const unsigned blocks = (unsigned)((Q.n_out + 255) / 256);
        if (eta == 1) fold_additive_eta1<<<blocks, 256, 0, c->stream>>>(Q);
        else fold_additive_eta2<<<blocks, 256, 0, c->stream>>>(Q);
    } else {
        const unsigned blocks = (unsigned)((P.n_out + 127) / 128);
        fold_additive_generic<<<blocks, 128, 0, c->stream>>>(P);
    }
    B200_LAUNCH_CHECK(c);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
struct MulFoldParams {
    const fp192 *f;
    fp192 *out;
    fp_dev_params fp;
    fp192 x;
    fp192 gpow[1 << B200_FOLD_MAX_ETA];  /* g^k, g = generator of the order-2^eta subgroup */
    fp192 h_lo_base, h_hi_base;          /* h_j = shift * G^j = shift * Glo^(j & 1023) * Ghi^(j >> 10) */
    fp192 shift;
    fp192 d_shift, d_lo_base, d_hi_base; /* h_j^-(S-1) / S, same decomposition */
    const fp192 *tab;                    /* device tables: [0,1024) G^lo, [1024, 2048) D^lo, then G^(1024 hi), D^(1024 hi) */
    uint32_t hi_entries;
    uint32_t logn, eta;
    size_t n_out;
    /* reference quirk when x lies in the domain (fri_aux.tcc:190-206): every coset from the one containing x on
       evaluates to zero and the last output is overwritten with f at (k*, last coset) */
    int *xflag;                          /* device: [0] = j* (or -1), [1] = k* */
};

__device__ __forceinline__ fp192 ldg_fp2(const fp192 *p)
{
    const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
    const uint64_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    fp192 r;
    r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
    r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
    r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
    return r;
}

__device__ __noinline__ fp192 fp_mul_ni(const fp192 &a, const fp192 &b, const fp_dev_params &P) { return fp_mul(a, b, P); }

__global__ void __launch_bounds__(128) fold_multiplicative_generic(const MulFoldParams P)
{
    const int S = 1 << P.eta;
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= P.n_out) return;
    const fp_dev_params &fp = P.fp;
    const fp192 *t = P.tab;
    fp192 h = fp_mul_ni(P.shift, ldg_fp2(t + (j & 1023)), fp);
    fp192 dj = fp_mul_ni(P.d_shift, ldg_fp2(t + 1024 + (j & 1023)), fp);
    if (j >> 10) {
        h = fp_mul_ni(h, ldg_fp2(t + 2048 + (j >> 10)), fp);
        dj = fp_mul_ni(dj, ldg_fp2(t + 2048 + P.hi_entries + (j >> 10)), fp);
    }
    fp192 a[1 << B200_FOLD_MAX_ETA], pre[1 << B200_FOLD_MAX_ETA];
    bool hit = false; int khit = 0;
#pragma unroll 1
    for (int k = 0; k < S; ++k) {
        const fp192 vk = k ? fp_mul_ni(h, P.gpow[k], fp) : h;
        a[k] = fp_sub(P.x, vk, fp);
        if (fp_is_zero(a[k])) { hit = true; khit = k; }
        pre[k] = k ? fp_mul_ni(pre[k - 1], a[k], fp) : a[0];
    }
    if (hit && P.xflag) { P.xflag[0] = (int)j; P.xflag[1] = khit; }
    fp192 acc = fp_zero(), suf = fp_one(fp);
#pragma unroll 1
    for (int k = S - 1; k >= 0; --k) {
        fp192 fv = ldg_fp2(P.f + (size_t)k * P.n_out + j);
        if (k) fv = fp_mul_ni(fv, P.gpow[k], fp);
        const fp192 nk = k ? fp_mul_ni(pre[k - 1], suf, fp) : suf;
        acc = fp_add(acc, fp_mul_ni(fv, nk, fp), fp);
        suf = fp_mul_ni(suf, a[k], fp);
    }
    acc = fp_mul_ni(acc, dj, fp);
    uint64_t *dst = reinterpret_cast<uint64_t *>(P.out + j);
    dst[0] = (uint64_t)acc.w[0] | ((uint64_t)acc.w[1] << 32);
    dst[1] = (uint64_t)acc.w[2] | ((uint64_t)acc.w[3] << 32);
    dst[2] = (uint64_t)acc.w[4] | ((uint64_t)acc.w[5] << 32);
}

template <int ETA>
__global__ void __launch_bounds__(128) fold_multiplicative_kernel(const MulFoldParams P)
{
    constexpr int S = 1 << ETA;
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= P.n_out) return;
    const fp_dev_params &fp = P.fp;
    /* h = shift * G^j and dj = h^-(S-1) / S from two-level power tables */
    const fp192 *t = P.tab;
    fp192 h = fp_mul(P.shift, ldg_fp2(t + (j & 1023)), fp);
    fp192 dj = fp_mul(P.d_shift, ldg_fp2(t + 1024 + (j & 1023)), fp);
    if (j >> 10) {
        h = fp_mul(h, ldg_fp2(t + 2048 + (j >> 10)), fp);
        dj = fp_mul(dj, ldg_fp2(t + 2048 + P.hi_entries + (j >> 10)), fp);
    }
    fp192 a[S], fv[S];
    bool hit = false; int khit = 0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
        const fp192 vk = (k == 0) ? h : fp_mul(h, P.gpow[k], fp);
        a[k] = fp_sub(P.x, vk, fp);
        if (fp_is_zero(a[k])) { hit = true; khit = k; }
        fv[k] = ldg_fp2(P.f + (size_t)k * P.n_out + j);
        if (k) fv[k] = fp_mul(fv[k], P.gpow[k], fp);       /* f_k * g^k */
    }
    if (hit && P.xflag) { P.xflag[0] = (int)j; P.xflag[1] = khit; }
    fp192 acc = fp_zero();
    if (S == 2) {
        acc = fp_add(fp_mul(fv[0], a[1], fp), fp_mul(fv[1], a[0], fp), fp);
    } else {
        fp192 pre[S], suf[S];
        pre[0] = a[0];
#pragma unroll
        for (int k = 1; k < S; ++k) pre[k] = fp_mul(pre[k - 1], a[k], fp);
        suf[S - 1] = a[S - 1];
#pragma unroll
        for (int k = S - 2; k >= 0; --k) suf[k] = fp_mul(suf[k + 1], a[k], fp);
#pragma unroll
        for (int k = 0; k < S; ++k) {
            fp192 nk;
            if (k == 0) nk = suf[1];
            else if (k == S - 1) nk = pre[S - 2];
            else nk = fp_mul(pre[k - 1], suf[k + 1], fp);
            acc = fp_add(acc, fp_mul(fv[k], nk, fp), fp);
        }
    }
    acc = fp_mul(acc, dj, fp);
    uint64_t *dst = reinterpret_cast<uint64_t *>(P.out + j);
    dst[0] = (uint64_t)acc.w[0] | ((uint64_t)acc.w[1] << 32);
    dst[1] = (uint64_t)acc.w[2] | ((uint64_t)acc.w[3] << 32);
    dst[2] = (uint64_t)acc.w[4] | ((uint64_t)acc.w[5] << 32);
}

/* applies the reference's behaviour for x inside the domain (see MulFoldParams::xflag) */
__global__ void fold_multiplicative_fixup(fp192 *out, const fp192 *f, size_t n_out, const int *xflag)
{
    const int jstar = xflag[0], kstar = xflag[1];
    if (jstar < 0) return;
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n_out; j += (size_t)gridDim.x * blockDim.x) {
        if (j < (size_t)jstar) continue;
        out[j] = (j == n_out - 1) ? f[(size_t)kstar * n_out + (n_out - 1)] : fp_zero();
    }
}

__global__ void fp_pow_table_step(fp192 *tab, fp192 factor, uint32_t len, fp_dev_params P)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < len) tab[len + e] = fp_mul(tab[e], factor, P);
}

static int build_pow(b200_ctx *c, fp192 *tab, fp192 base, uint32_t log_entries, const fp_dev_params &fp)
{
    const fp192 one = fp_one(fp);
    B200_CUDA(c, cudaMemcpyAsync(tab, &one, sizeof one, cudaMemcpyHostToDevice, c->stream));
    for (uint32_t b = 0; b < log_entries; ++b) {
        const uint32_t len = 1u << b;
        fp_pow_table_step<<<(len + 255) / 256, 256, 0, c->stream>>>(tab, base, len, fp);
        base = fp_mul(base, base, fp);
    }
    B200_LAUNCH_CHECK(c);
    return 0;
}

static fp192 fp_from_limbs(const uint64_t v[3])
{
    fp192 r;
    for (int i = 0; i < 3; ++i) { r.w[2 * i] = (uint32_t)v[i]; r.w[2 * i + 1] = (uint32_t)(v[i] >> 32); }
    return r;
}

static int fold_multiplicative(b200_ctx *c, const b200_fp_params *Pm, const b200_domain *L, uint32_t eta,
                               const uint64_t x[3], b200_buf f, b200_buf out)
{
    const uint32_t logn = L->dim;
    if (!Pm) return set_err(c, B200_ERR_INVALID, "field parameters required for the multiplicative fold");
    if (eta == 0 || eta > logn || eta > B200_FOLD_MAX_ETA)
        return set_err(c, B200_ERR_INVALID, "unsupported localization parameter %u for a 2^%u domain", eta, logn);
    if (logn > 31) return set_err(c, B200_ERR_INVALID, "domain too large");
    const fp_dev_params fp = make_fp_dev_params(Pm->modulus, Pm->r2, Pm->one_mont, Pm->inv);
    MulFoldParams P; memset(&P, 0, sizeof P);
    P.f = (const fp192 *)f; P.out = (fp192 *)out; P.fp = fp; P.logn = logn; P.eta = eta;
    P.n_out = (size_t)1 << (logn - eta);
    const uint32_t S = 1u << eta;
    const fp192 G = fp_from_limbs(L->generator), shift = fp_from_limbs(L->shift);
    P.x = fp_from_limbs(x); P.shift = shift;
    /* g = G^(n/S): generator of the order-S subgroup (field_subset(coset_size, one), fri_aux.tcc:158-160) */
    const fp192 g = fp_pow_u64(G, (uint64_t)P.n_out, fp);
    P.gpow[0] = fp_one(fp);
    for (uint32_t k = 1; k < S; ++k) P.gpow[k] = fp_mul(P.gpow[k - 1], g, fp);
    /* D = G^-(S-1); d_shift = shift^-(S-1) / S */
    const fp192 D = fp_inv(fp_pow_u64(G, S - 1, fp), fp);
    P.d_shift = fp_inv(fp_mul(fp_pow_u64(shift, S - 1, fp), fp_from_u64(S, fp), fp), fp);
    const uint32_t out_bits = logn - eta;
    const uint32_t hi_bits = out_bits > 10 ? out_bits - 10 : 0;
    P.hi_entries = 1u << hi_bits;
    const size_t tab_elems = 2048 + 2 * (size_t)P.hi_entries;
    fp192 *tab = (fp192 *)scratch_get(c, 2, tab_elems * sizeof(fp192) + 64);
    if (!tab) return set_err(c, B200_ERR_NOMEM, "cannot allocate fold tables");
    int rc;
    if ((rc = build_pow(c, tab, G, 10, fp))) return rc;
    if ((rc = build_pow(c, tab + 1024, D, 10, fp))) return rc;
    if ((rc = build_pow(c, tab + 2048, fp_pow_u64(G, 1024, fp), hi_bits, fp))) return rc;
    if ((rc = build_pow(c, tab + 2048 + P.hi_entries, fp_pow_u64(D, 1024, fp), hi_bits, fp))) return rc;
    P.tab = tab;
    /* is x in the domain?  x^n == shift^n  (subgroup.tcc:306-310) */
    const bool x_in_domain = fp_eq(fp_pow_u64(P.x, (uint64_t)1 << logn, fp), fp_pow_u64(shift, (uint64_t)1 << logn, fp));
    int *xflag = nullptr;
    if (x_in_domain) {
        xflag = (int *)(tab + tab_elems);
        const int init[2] = {-1, 0};
        B200_CUDA(c, cudaMemcpyAsync(xflag, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    }
    P.xflag = xflag;
    const unsigned blocks = (unsigned)((P.n_out + 127) / 128);
    switch (eta) {
    case 1: fold_multiplicative_kernel<1><<<blocks, 128, 0, c->stream>>>(P); break;
    case 2: fold_multiplicative_kernel<2><<<blocks, 128, 0, c->stream>>>(P); break;
    default: fold_multiplicative_generic<<<blocks, 128, 0, c->stream>>>(P); break;
    }
    B200_LAUNCH_CHECK(c);
    if (x_in_domain) {
        fold_multiplicative_fixup<<<(unsigned)std::min<size_t>((P.n_out + 255) / 256, 1184), 256, 0, c->stream>>>(
            P.out, P.f, P.n_out, xflag);
        B200_LAUNCH_CHECK(c);
    }
    return 0;
}

int fri_fold(b200_ctx *c, int field, const b200_fp_params *fp, const b200_domain *L, uint32_t eta,
             const uint64_t x[3], b200_buf f, b200_buf out)
{
    if (!L || !x) return set_err(c, B200_ERR_INVALID, "null domain or challenge");
    if (field == B200_FIELD_GF192 && L->kind == B200_DOMAIN_AFFINE_SUBSPACE) return fold_additive(c, L, eta, x, f, out);
    if (field == B200_FIELD_FP && L->kind == B200_DOMAIN_MULTIPLICATIVE_COSET) return fold_multiplicative(c, fp, L, eta, x, f, out);
    return set_err(c, B200_ERR_INVALID, "f_i_domain is of unsupported domain type"); /* fri_aux.tcc:33 */
}

} // namespace b200
