/*
 * b200_hostsim.cpp — HOST SIMULATOR of the C ABI in include/b200_iop.h.  TEST INFRASTRUCTURE ONLY.
 *
 * Purpose: the drop-in layer (libiop_b200/dropin/, libiop_b200/host/) is several thousand lines of HOST logic — which
 * vector goes to which call in which layout, residency and lifetime of device copies, lazily materialised host vectors,
 * the shadows of the reference's IOP / protocol code.  None of that needs a GPU to be wrong.  This file implements every
 * entry point of the C ABI on the CPU ("device memory" = malloc) on top of the parity oracle (the C files under oracle/), so that
 * tests/test_dropin_hostsim.py can run the reference's whole provers through the SAME drop-in objects in the build
 * container (no GPU there) and compare their transcripts with the CPU reference's fingerprints.
 *
 * It is NOT a fallback: it is built only by tests/hostsim/Makefile into tests/hostsim/_build/, it links oracle/liboracle.so
 * (which product code must never do), nothing under libiop_b200/ or include/ refers to it, and the product library
 * libiop_b200/lib/libb200iop.so has no code path to it (b200_ctx_create there fails without a CUDA device).
 * Speed is irrelevant here; clarity is: every function is the plain statement of the header's contract.
 */
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_iop.h"
#include "../../oracle/oracle.h"

struct b200_ctx {
    char err[512];
    unsigned long long launches = 0, h2d = 0, d2h = 0;
    std::chrono::steady_clock::time_point t0;
};

namespace {

int fail(b200_ctx *c, int code, const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(c->err, sizeof c->err, fmt, ap); va_end(ap);
    return code;
}
f192 *E(b200_buf b) { return (f192 *)(uintptr_t)b; }
f192 el(const uint64_t *p) { f192 r; std::memcpy(&r, p, 24); return r; }
f192 zero() { f192 r; std::memset(&r, 0, 24); return r; }
bool is_zero(const f192 &a) { return (a.w[0] | a.w[1] | a.w[2]) == 0; }

/* field policy: GF(2^192) or the 3-limb Montgomery prime field */
struct Field {
    int kind;
    orc_fp_params P;
    f192 add(const f192 &a, const f192 &b) const
    {
        f192 r;
        if (kind == B200_FIELD_GF192) { for (int i = 0; i < 3; ++i) r.w[i] = a.w[i] ^ b.w[i]; }
        else orc_fp_add(&P, &r, &a, &b);
        return r;
    }
    f192 sub(const f192 &a, const f192 &b) const
    {
        if (kind == B200_FIELD_GF192) return add(a, b);
        f192 r; orc_fp_sub(&P, &r, &a, &b); return r;
    }
    f192 mul(const f192 &a, const f192 &b) const
    {
        f192 r;
        if (kind == B200_FIELD_GF192) orc_gf192_mul(&r, &a, &b); else orc_fp_mul(&P, &r, &a, &b);
        return r;
    }
    f192 inv(const f192 &a) const
    {
        f192 r;
        if (is_zero(a)) return zero();
        if (kind == B200_FIELD_GF192) orc_gf192_inv(&r, &a); else orc_fp_inv(&P, &r, &a);
        return r;
    }
};
orc_fp_params to_orc(const b200_fp_params *fp)
{
    orc_fp_params P;
    std::memset(&P, 0, sizeof P);
    std::memcpy(P.mod, fp->modulus, 24); std::memcpy(P.r2, fp->r2, 24); std::memcpy(P.one, fp->one_mont, 24);
    P.inv = fp->inv; P.two_adicity = fp->two_adicity; P.num_bits = fp->num_bits;
    /* b200_fp_params carries no multiplicative generator (the kernels derive the order-2^eta subgroup from the domain's
       generator); the oracle's fold wants one: the tests' prime field is edwards_Fr, whose constants the oracle knows */
    orc_fp_params ed;
    orc_fp_params_edwards(&ed);
    if (std::memcmp(ed.mod, P.mod, 24) == 0) std::memcpy(P.generator, ed.generator, 24);
    return P;
}
bool make_field(b200_ctx *c, int field, const b200_fp_params *fp, Field *f)
{
    f->kind = field;
    if (field == B200_FIELD_GF192) return true;
    if (field == B200_FIELD_FP && fp) { f->P = to_orc(fp); return true; }
    fail(c, B200_ERR_INVALID, "unknown field id %d (or missing prime-field parameters)", field);
    return false;
}

} // namespace

extern "C" {

const char *b200_version(void) { return "b200 host simulator (tests only)"; }
b200_ctx *b200_ctx_create(int) { b200_ctx *c = new b200_ctx(); c->err[0] = 0; return c; }
void b200_ctx_destroy(b200_ctx *c) { delete c; }
const char *b200_last_error(b200_ctx *c) { return c ? c->err : "no context"; }
int b200_sync(b200_ctx *) { return 0; }
void *b200_stream(b200_ctx *) { return nullptr; }
int b200_device_info(b200_ctx *, int *sm, size_t *mem, int *maj, int *min)
{
    if (sm) *sm = 148;
    if (mem) *mem = (size_t)180 << 30;
    if (maj) *maj = 10;
    if (min) *min = 0;
    return 0;
}
b200_buf b200_alloc(b200_ctx *c, size_t n) { return b200_alloc_bytes(c, n * 24); }
void b200_free(b200_ctx *, b200_buf b) { std::free((void *)(uintptr_t)b); }
int b200_trim(b200_ctx *) { return 0; }
b200_bytes b200_alloc_bytes(b200_ctx *, size_t n) { return (b200_bytes)(uintptr_t)std::malloc(n ? n : 8); }
void b200_free_bytes(b200_ctx *, b200_bytes b) { std::free((void *)(uintptr_t)b); }
void *b200_host_alloc(b200_ctx *, size_t n) { void *p = nullptr; return posix_memalign(&p, 4096, n ? n : 8) == 0 ? p : nullptr; }
void b200_host_free(b200_ctx *, void *p) { std::free(p); }

int b200_upload(b200_ctx *c, b200_buf dst, size_t off, const uint64_t *src, size_t n)
{ c->h2d += n * 24; std::memcpy(E(dst) + off, src, n * 24); return 0; }
int b200_download(b200_ctx *c, uint64_t *dst, b200_buf src, size_t off, size_t n)
{ c->d2h += n * 24; std::memcpy(dst, E(src) + off, n * 24); return 0; }
int b200_upload_bytes(b200_ctx *c, b200_bytes dst, const uint8_t *src, size_t n)
{ c->h2d += n; std::memcpy((void *)(uintptr_t)dst, src, n); return 0; }
int b200_download_bytes(b200_ctx *c, uint8_t *dst, b200_bytes src, size_t off, size_t n)
{ c->d2h += n; std::memcpy(dst, (const uint8_t *)(uintptr_t)src + off, n); return 0; }
int b200_copy(b200_ctx *, b200_buf dst, size_t doff, b200_buf src, size_t soff, size_t n)
{ std::memmove(E(dst) + doff, E(src) + soff, n * 24); return 0; }
int b200_wait_uploads(b200_ctx *) { return 0; }
int b200_copy_strided(b200_ctx *, b200_buf dst, size_t dstride, b200_buf src, size_t soff, size_t sstride, size_t n)
{ for (size_t i = 0; i < n; ++i) E(dst)[i * dstride] = E(src)[soff + i * sstride]; return 0; }
int b200_fill_zero(b200_ctx *, b200_buf dst, size_t off, size_t n) { std::memset(E(dst) + off, 0, n * 24); return 0; }
int b200_gather(b200_ctx *c, uint64_t *dst, b200_buf src, const size_t *idx, size_t k)
{
    c->d2h += k * 24;
    for (size_t i = 0; i < k; ++i) std::memcpy(dst + 3 * i, E(src) + idx[i], 24);
    return 0;
}
int b200_gather_dev(b200_ctx *, b200_buf out, b200_buf table, b200_bytes idx, b200_bytes map, size_t n)
{
    const uint32_t *ix = reinterpret_cast<const uint32_t *>(idx), *mp = reinterpret_cast<const uint32_t *>(map);
    for (size_t k = 0; k < n; ++k) E(out)[k] = E(table)[mp ? mp[ix[k]] : ix[k]];
    return 0;
}
int b200_scatter_dev(b200_ctx *, b200_buf out, b200_buf x, b200_bytes map, size_t n)
{
    const uint32_t *mp = reinterpret_cast<const uint32_t *>(map);
    if (!mp) return B200_ERR_INVALID;
    for (size_t c = 0; c < n; ++c) E(out)[mp[c]] = E(x)[c];
    return 0;
}
int b200_fill(b200_ctx *, b200_buf dst, size_t off, const uint64_t value[3], size_t n)
{ for (size_t i = 0; i < n; ++i) std::memcpy(E(dst) + off + i, value, 24); return 0; }
int b200_csr_row_indices_dev(b200_ctx *, size_t n_rows, b200_bytes row_ptr, b200_bytes row_out)
{
    const uint64_t *rp = reinterpret_cast<const uint64_t *>(row_ptr);
    uint32_t *ro = reinterpret_cast<uint32_t *>(row_out);
    for (size_t i = 0; i < n_rows; ++i) for (uint64_t k = rp[i]; k < rp[i + 1]; ++k) ro[k] = (uint32_t)i;
    return 0;
}
int b200_csr_transpose_dev(b200_ctx *, size_t n_cols, size_t nnz, b200_bytes col, b200_bytes row, b200_buf coeff,
                           b200_bytes col_ptr_out, b200_bytes src_out, b200_buf coeff_out)
{
    const uint32_t *cl = reinterpret_cast<const uint32_t *>(col), *rw = reinterpret_cast<const uint32_t *>(row);
    uint64_t *cp = reinterpret_cast<uint64_t *>(col_ptr_out);
    uint32_t *src = reinterpret_cast<uint32_t *>(src_out);
    std::vector<uint64_t> cursor(n_cols + 1, 0);
    for (size_t k = 0; k < nnz; ++k) ++cursor[cl[k] + 1];
    for (size_t c = 0; c < n_cols; ++c) cursor[c + 1] += cursor[c];
    for (size_t c = 0; c <= n_cols; ++c) cp[c] = cursor[c];
    for (size_t k = 0; k < nnz; ++k) { const uint64_t at = cursor[cl[k]]++; src[at] = rw[k]; E(coeff_out)[at] = E(coeff)[k]; }
    return 0;
}
int b200_gather_multi(b200_ctx *c, uint64_t *dst, const b200_buf *srcs, size_t n_srcs, const size_t *idx, size_t k)
{
    c->d2h += n_srcs * k * 24;
    for (size_t s = 0; s < n_srcs; ++s)
        for (size_t j = 0; j < k; ++j) std::memcpy(dst + 3 * (s * k + j), E(srcs[s]) + idx[j], 24);
    return 0;
}
int b200_timer_start(b200_ctx *c) { c->t0 = std::chrono::steady_clock::now(); return 0; }
int b200_timer_stop(b200_ctx *c, float *ms)
{ *ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - c->t0).count(); return 0; }

/* ---- additive FFT ---- */
static int check_affine(b200_ctx *c, const b200_domain *L, const char *what)
{
    if (!L || L->kind != B200_DOMAIN_AFFINE_SUBSPACE) return fail(c, B200_ERR_INVALID, "%s needs an affine subspace domain", what);
    return 0;
}
int b200_gf192_fft_batch(b200_ctx *c, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, size_t batch, b200_buf out)
{
    if (check_affine(c, L, "additive FFT")) return B200_ERR_INVALID;
    const size_t n = (size_t)1 << L->dim;
    if (n_coeffs > n) return fail(c, B200_ERR_INVALID, "more coefficients (%zu) than domain elements (%zu)", n_coeffs, n);
    const f192 shift = el(L->shift);
    for (size_t r = 0; r < batch; ++r) {
        ++c->launches;
        if (n_coeffs == 0) { std::memset(E(out) + r * n, 0, n * 24); continue; }
        orc_additive_fft(E(out) + r * n, E(coeffs) + r * n_coeffs, n_coeffs, L->dim, (const f192 *)L->basis, &shift);
    }
    return 0;
}
int b200_gf192_fft(b200_ctx *c, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf out)
{ return b200_gf192_fft_batch(c, L, coeffs, n_coeffs, 1, out); }
int b200_gf192_ifft_batch(b200_ctx *c, const b200_domain *H, b200_buf evals, size_t batch, b200_buf out)
{
    if (check_affine(c, H, "additive IFFT")) return B200_ERR_INVALID;
    const size_t n = (size_t)1 << H->dim;
    const f192 shift = el(H->shift);
    for (size_t r = 0; r < batch; ++r) {
        ++c->launches;
        std::vector<f192> tmp(n);            /* evals and out may alias */
        orc_additive_ifft(tmp.data(), E(evals) + r * n, H->dim, (const f192 *)H->basis, &shift);
        std::memcpy(E(out) + r * n, tmp.data(), n * 24);
    }
    return 0;
}
int b200_gf192_ifft(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf out) { return b200_gf192_ifft_batch(c, H, evals, 1, out); }
int b200_gf192_lde_batch(b200_ctx *c, const b200_domain *H, const b200_domain *L, b200_buf evals, size_t batch, b200_buf out)
{
    if (check_affine(c, H, "fused extension") || check_affine(c, L, "fused extension")) return B200_ERR_INVALID;
    if (H->dim > L->dim || (H->dim && std::memcmp(H->basis, L->basis, (size_t)H->dim * 24) != 0))
        return fail(c, B200_ERR_INVALID, "the codeword domain does not extend the basis of the interpolation domain");
    const size_t nh = (size_t)1 << H->dim, nl = (size_t)1 << L->dim;
    const f192 sh = el(H->shift), sl = el(L->shift);
    for (size_t r = 0; r < batch; ++r) {              /* the two transforms the kernel fuses */
        ++c->launches;
        std::vector<f192> coeffs(nh);
        orc_additive_ifft(coeffs.data(), E(evals) + r * nh, H->dim, (const f192 *)H->basis, &sh);
        orc_additive_fft(E(out) + r * nl, coeffs.data(), nh, L->dim, (const f192 *)L->basis, &sl);
    }
    return 0;
}
int b200_gf192_lde(b200_ctx *c, const b200_domain *H, const b200_domain *L, b200_buf evals, b200_buf out)
{ return b200_gf192_lde_batch(c, H, L, evals, 1, out); }
/* The halves of the extension: the device's intermediate coefficient form is opaque but LINEAR in the polynomial; the
   simulator uses the monomial coefficients for it (the one property callers rely on holds for any linear form). */
int b200_gf192_lde_interpolate(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf out)
{
    if (evals == out) return fail(c, B200_ERR_INVALID, "the interpolation half does not run in place");
    return b200_gf192_ifft_batch(c, H, evals, 1, out);
}
int b200_gf192_lde_evaluate(b200_ctx *c, const b200_domain *L, uint32_t coeff_dim, b200_buf coeffs, b200_buf out)
{
    if (check_affine(c, L, "fused extension")) return B200_ERR_INVALID;
    if (coeff_dim > L->dim) return fail(c, B200_ERR_INVALID, "extension from dimension %u to %u is not possible", coeff_dim, L->dim);
    if (coeffs == out) return fail(c, B200_ERR_INVALID, "the evaluation half does not run in place");
    return b200_gf192_fft_batch(c, L, coeffs, (size_t)1 << coeff_dim, 1, out);
}
int b200_gf192_mul(b200_ctx *c, b200_buf out, b200_buf a, b200_buf b, size_t n) { ++c->launches; orc_gf192_mul_vec(E(out), E(a), E(b), n); return 0; }
int b200_gf192_inv(b200_ctx *c, b200_buf out, b200_buf a, size_t n)
{
    ++c->launches;
    Field f; f.kind = B200_FIELD_GF192;
    for (size_t i = 0; i < n; ++i) E(out)[i] = f.inv(E(a)[i]);
    return 0;
}

/* ---- NTT ---- */
static int check_coset(b200_ctx *c, const b200_fp_params *fp, const b200_domain *L, const char *what)
{
    if (!fp) return fail(c, B200_ERR_INVALID, "%s: missing prime-field parameters", what);
    if (!L || L->kind != B200_DOMAIN_MULTIPLICATIVE_COSET) return fail(c, B200_ERR_INVALID, "%s needs a multiplicative coset domain", what);
    return 0;
}
int b200_fp_ntt(b200_ctx *c, const b200_fp_params *fp, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf out)
{
    if (check_coset(c, fp, L, "NTT")) return B200_ERR_INVALID;
    const size_t n = (size_t)1 << L->dim;
    if (n_coeffs > n) return fail(c, B200_ERR_INVALID, "more coefficients (%zu) than domain elements (%zu)", n_coeffs, n);
    ++c->launches;
    if (n_coeffs == 0) { std::memset(E(out), 0, n * 24); return 0; }
    const orc_fp_params P = to_orc(fp);
    const f192 g = el(L->generator), s = el(L->shift);
    std::vector<f192> tmp(n);
    orc_mult_fft(&P, tmp.data(), E(coeffs), n_coeffs, L->dim, &g, &s);
    std::memcpy(E(out), tmp.data(), n * 24);
    return 0;
}
int b200_fp_intt(b200_ctx *c, const b200_fp_params *fp, const b200_domain *H, b200_buf evals, b200_buf out)
{
    if (check_coset(c, fp, H, "inverse NTT")) return B200_ERR_INVALID;
    const size_t n = (size_t)1 << H->dim;
    ++c->launches;
    const orc_fp_params P = to_orc(fp);
    const f192 g = el(H->generator), s = el(H->shift);
    std::vector<f192> tmp(n);
    orc_mult_ifft(&P, tmp.data(), E(evals), H->dim, &g, &s);
    std::memcpy(E(out), tmp.data(), n * 24);
    return 0;
}
static int fp_binop(b200_ctx *c, const b200_fp_params *fp, b200_buf out, b200_buf a, b200_buf b, size_t n, int op)
{
    if (!fp) return fail(c, B200_ERR_INVALID, "missing prime-field parameters");
    Field f; f.kind = B200_FIELD_FP; f.P = to_orc(fp);
    ++c->launches;
    for (size_t i = 0; i < n; ++i)
        E(out)[i] = op == 0 ? f.mul(E(a)[i], E(b)[i]) : op == 1 ? f.add(E(a)[i], E(b)[i]) : op == 2 ? f.sub(E(a)[i], E(b)[i]) : f.inv(E(a)[i]);
    return 0;
}
int b200_fp_mul(b200_ctx *c, const b200_fp_params *fp, b200_buf o, b200_buf a, b200_buf b, size_t n) { return fp_binop(c, fp, o, a, b, n, 0); }
int b200_fp_add(b200_ctx *c, const b200_fp_params *fp, b200_buf o, b200_buf a, b200_buf b, size_t n) { return fp_binop(c, fp, o, a, b, n, 1); }
int b200_fp_sub(b200_ctx *c, const b200_fp_params *fp, b200_buf o, b200_buf a, b200_buf b, size_t n) { return fp_binop(c, fp, o, a, b, n, 2); }
int b200_fp_inv(b200_ctx *c, const b200_fp_params *fp, b200_buf o, b200_buf a, size_t n) { return fp_binop(c, fp, o, a, a, n, 3); }

/* ---- FRI fold ---- */
int b200_fri_fold(b200_ctx *c, int field, const b200_fp_params *fp, const b200_domain *L, uint32_t eta, const uint64_t x_i[3],
                  b200_buf f_i, b200_buf out)
{
    if (!L) return fail(c, B200_ERR_INVALID, "fold: no domain");
    if (eta > L->dim) return fail(c, B200_ERR_INVALID, "fold: localization beyond the domain dimension");
    const f192 x = el(x_i), s = el(L->shift);
    ++c->launches;
    const size_t n_out = (size_t)1 << (L->dim - eta);
    std::vector<f192> tmp(n_out);
    if (field == B200_FIELD_GF192) {
        if (check_affine(c, L, "additive fold")) return B200_ERR_INVALID;
        orc_fri_fold_additive(tmp.data(), E(f_i), L->dim, (const f192 *)L->basis, &s, eta, &x);
    } else if (field == B200_FIELD_FP) {
        if (check_coset(c, fp, L, "multiplicative fold")) return B200_ERR_INVALID;
        const orc_fp_params P = to_orc(fp);
        const f192 g = el(L->generator);
        orc_fri_fold_multiplicative(&P, tmp.data(), E(f_i), L->dim, &g, &s, eta, &x);
    } else return fail(c, B200_ERR_INVALID, "unknown field id %d", field);
    std::memcpy(E(out), tmp.data(), n_out * 24);
    return 0;
}

/* ---- Merkle ---- */
int b200_merkle_build(b200_ctx *c, const b200_buf *oracles, size_t n_oracles, size_t n_elems, size_t coset_size, int domain_kind,
                      const uint8_t *salts, size_t salt_len, size_t digest_len, b200_bytes nodes_out)
{
    if (!n_oracles || !coset_size || n_elems % coset_size) return fail(c, B200_ERR_LOGIC, "merkle: bad leaf geometry");
    const size_t leaves = n_elems / coset_size;
    if (leaves < 2 || (leaves & (leaves - 1))) return fail(c, B200_ERR_INVALID, "Merkle tree size must be a power of two, and at least 2.");
    std::vector<const f192 *> ptr(n_oracles);
    for (size_t i = 0; i < n_oracles; ++i) ptr[i] = E(oracles[i]);
    ++c->launches;
    if (salts) c->h2d += leaves * salt_len;
    orc_merkle_build((uint8_t *)(uintptr_t)nodes_out, ptr.data(), n_oracles, n_elems, coset_size, domain_kind, salts, salt_len, digest_len);
    return 0;
}
int b200_merkle_root(b200_ctx *c, b200_bytes nodes, uint8_t *root, size_t digest_len)
{ c->d2h += digest_len; std::memcpy(root, (const void *)(uintptr_t)nodes, digest_len); return 0; }
int b200_merkle_nodes(b200_ctx *c, b200_bytes nodes, const size_t *idx, size_t k, size_t digest_len, uint8_t *out)
{
    c->d2h += k * digest_len;
    for (size_t i = 0; i < k; ++i) std::memcpy(out + i * digest_len, (const uint8_t *)(uintptr_t)nodes + idx[i] * digest_len, digest_len);
    return 0;
}
int b200_blake2b_batch(b200_ctx *c, b200_bytes in, size_t msg_len, size_t n_msgs, size_t digest_len, b200_bytes out)
{
    ++c->launches;
    for (size_t i = 0; i < n_msgs; ++i)
        orc_blake2b((uint8_t *)(uintptr_t)out + i * digest_len, digest_len, (const uint8_t *)(uintptr_t)in + i * msg_len, msg_len, nullptr, 0);
    return 0;
}

/* ---- element-wise passes ---- */
int b200_vec_lincomb(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, size_t n, size_t k, const uint64_t *coeffs, const b200_buf *vecs)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    ++c->launches;
    for (size_t i = 0; i < n; ++i) {
        f192 acc = zero();
        for (size_t j = 0; j < k; ++j) acc = f.add(acc, f.mul(el(coeffs + 3 * j), E(vecs[j])[i]));
        E(out)[i] = acc;
    }
    return 0;
}
int b200_vec_mul_add(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, b200_buf b, b200_buf cc, b200_buf d,
                     int subtract, size_t n)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    if (!n) return 0;
    if (!a) return fail(c, B200_ERR_INVALID, "b200_vec_mul_add: operand a is required");
    ++c->launches;
    for (size_t i = 0; i < n; ++i) {
        f192 r = E(a)[i];
        if (b) r = f.mul(r, E(b)[i]);
        if (cc) { f192 s = E(cc)[i]; if (d) s = f.mul(s, E(d)[i]); r = subtract ? f.sub(r, s) : f.add(r, s); }
        E(out)[i] = r;
    }
    return 0;
}
int b200_vec_rows_dot(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf x, b200_buf y, size_t n, size_t rows, size_t split)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    if (split > rows) return fail(c, B200_ERR_INVALID, "b200_vec_rows_dot: split beyond the number of rows");
    ++c->launches;
    for (size_t a = 0; a < n; ++a) {
        f192 plus = zero(), minus = zero();
        for (size_t r = 0; r < split; ++r) plus = f.add(plus, f.mul(E(x)[r * n + a], E(y)[r * n + a]));
        for (size_t r = split; r < rows; ++r) minus = f.add(minus, f.mul(E(x)[r * n + a], E(y)[r * n + a]));
        E(out)[a] = f.sub(plus, minus);
    }
    return 0;
}
int b200_vec_affine(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, const uint64_t *mul, const uint64_t *add, size_t n)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    ++c->launches;
    for (size_t i = 0; i < n; ++i) {
        f192 r = E(a)[i];
        if (mul) r = f.mul(r, el(mul));
        if (add) r = f.add(r, el(add));
        E(out)[i] = r;
    }
    return 0;
}
int b200_vec_find_zero(b200_ctx *c, b200_buf a, size_t n, size_t *index_out)
{
    ++c->launches;
    size_t i = 0;
    while (i < n && !is_zero(E(a)[i])) ++i;
    *index_out = i;
    return 0;
}
int b200_vec_mul_table(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, b200_buf table, uint32_t shift, size_t mask, size_t n)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    ++c->launches;
    for (size_t i = 0; i < n; ++i) {
        const f192 t = E(table)[(i >> shift) & mask];
        E(out)[i] = a ? f.mul(E(a)[i], t) : t;
    }
    return 0;
}
int b200_gf192_subset_sum_products(b200_ctx *c, b200_buf out, uint32_t dim, uint32_t n_factors, const uint64_t *tables)
{
    if (dim > 40 || n_factors == 0 || n_factors > 64) return fail(c, B200_ERR_INVALID, "subset sums: dim <= 40 and 1..64 factors");
    Field f; f.kind = B200_FIELD_GF192;
    ++c->launches;
    for (size_t i = 0; i < ((size_t)1 << dim); ++i) {
        f192 acc = zero();
        for (uint32_t k = 0; k < n_factors; ++k) {
            const uint64_t *tb = tables + (size_t)k * (dim + 1) * 3;
            f192 v = el(tb);
            for (uint32_t j = 0; j < dim; ++j) if ((i >> j) & 1u) v = f.add(v, el(tb + 3 * (1 + j)));
            acc = k ? f.mul(acc, v) : v;
        }
        E(out)[i] = acc;
    }
    return 0;
}
int b200_fp_geometric(b200_ctx *c, const b200_fp_params *fp, b200_buf out, size_t n, const uint64_t a[3], const uint64_t r[3], const uint64_t add[3])
{
    if (!fp) return fail(c, B200_ERR_INVALID, "missing prime-field parameters");
    Field f; f.kind = B200_FIELD_FP; f.P = to_orc(fp);
    ++c->launches;
    f192 cur = el(a);
    const f192 rr = el(r);
    for (size_t i = 0; i < n; ++i) { E(out)[i] = add ? f.add(cur, el(add)) : cur; cur = f.mul(cur, rr); }
    return 0;
}
int b200_poly_eval(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf coeffs, size_t n, const uint64_t x[3], uint64_t out[3])
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    ++c->launches;
    f192 acc = zero();
    const f192 y = el(x);
    for (size_t i = n; i-- > 0;) acc = f.add(f.mul(acc, y), E(coeffs)[i]);
    std::memcpy(out, &acc, 24);
    return 0;
}
int b200_spmv_csr(b200_ctx *c, int field, const b200_fp_params *fp, size_t n_rows, const uint64_t *row_ptr, const uint32_t *col,
                  const uint64_t *coeff, b200_buf x, size_t n_x, b200_buf y_out)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    ++c->launches;
    for (size_t i = 0; i < n_rows; ++i) {
        f192 acc = zero();
        for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
            if (col[k] >= n_x) return fail(c, B200_ERR_INVALID, "spmv: column index %u outside the vector (%zu)", col[k], n_x);
            acc = f.add(acc, f.mul(el(coeff + 3 * k), E(x)[col[k]]));
        }
        E(y_out)[i] = acc;
    }
    return 0;
}
int b200_spmv_csr_dev(b200_ctx *c, int field, const b200_fp_params *fp, size_t n_rows, b200_bytes row_ptr, b200_bytes col, b200_buf coeff,
                      b200_buf x, b200_buf y_out)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    ++c->launches;
    const uint64_t *rp = (const uint64_t *)(uintptr_t)row_ptr;
    const uint32_t *cl = (const uint32_t *)(uintptr_t)col;
    for (size_t i = 0; i < n_rows; ++i) {
        f192 acc = zero();
        for (uint64_t k = rp[i]; k < rp[i + 1]; ++k) acc = f.add(acc, f.mul(E(coeff)[k], E(x)[cl[k]]));
        E(y_out)[i] = acc;
    }
    return 0;
}
int b200_gf192_scatter_mul_add_dev(b200_ctx *c, b200_buf y, size_t nnz, b200_bytes dst, b200_bytes src, b200_buf coeff, b200_buf x,
                                   const uint64_t scale[3])
{
    Field f; f.kind = B200_FIELD_GF192;
    ++c->launches;
    const uint32_t *d = (const uint32_t *)(uintptr_t)dst, *sr = (const uint32_t *)(uintptr_t)src;
    for (size_t k = 0; k < nnz; ++k) E(y)[d[k]] = f.add(E(y)[d[k]], f.mul(f.mul(E(coeff)[k], el(scale)), E(x)[sr[k]]));
    return 0;
}
int b200_gf192_scatter_mul_add(b200_ctx *c, b200_buf y, size_t n_y, size_t nnz, const uint32_t *dst, const uint32_t *src,
                               const uint64_t *coeff, b200_buf x, size_t n_x, const uint64_t scale[3])
{
    Field f; f.kind = B200_FIELD_GF192;
    ++c->launches;
    c->h2d += nnz * 32;
    for (size_t k = 0; k < nnz; ++k) {
        if (dst[k] >= n_y || src[k] >= n_x) return fail(c, B200_ERR_INVALID, "scatter: entry %zu outside the vectors", k);
        E(y)[dst[k]] = f.add(E(y)[dst[k]], f.mul(f.mul(el(coeff + 3 * k), el(scale)), E(x)[src[k]]));
    }
    return 0;
}
int b200_poly_divide_sparse(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf p, size_t n_p, size_t deg, const uint64_t lead_inv[3],
                            size_t n_terms, const uint64_t *coefs, const size_t *pows, b200_buf q_out, b200_buf r_out)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    if (n_p <= deg) return fail(c, B200_ERR_INVALID, "polynomial division: dividend degree below divisor degree");
    for (size_t j = 0; j < n_terms; ++j)
        if (pows[j] >= deg) return fail(c, B200_ERR_INVALID, "polynomial division: term degree not below the leading degree");
    ++c->launches;
    /* schoolbook from the top on a working copy (the reference's order) */
    std::vector<f192> w(E(p), E(p) + n_p);
    const size_t qn = n_p - deg;
    const f192 li = el(lead_inv);
    for (size_t i = qn; i-- > 0;) {
        const f192 q = f.mul(w[deg + i], li);
        E(q_out)[i] = q;
        for (size_t j = 0; j < n_terms; ++j) w[i + pows[j]] = f.sub(w[i + pows[j]], f.mul(el(coefs + 3 * j), q));
    }
    if (r_out) std::memcpy(E(r_out), w.data(), deg * 24);
    return 0;
}

/* ---- proof of work / hash chain ---- */
int b200_pow_solve(b200_ctx *c, const uint8_t *challenge, size_t digest_len, uint32_t pow_bitlen, uint64_t pow_upperbound, uint8_t *pow_out)
{
    ++c->launches;
    orc_pow_solve(pow_out, challenge, digest_len, pow_bitlen, pow_upperbound);
    return 0;
}
int b200_squeeze_field_elements(b200_ctx *c, int field, const b200_fp_params *fp, const uint8_t *msg, size_t msg_len, size_t n, b200_buf out)
{
    Field f; if (!make_field(c, field, fp, &f)) return B200_ERR_INVALID;
    if (msg_len < 8 || msg_len > 128 || (msg_len & 7)) return fail(c, B200_ERR_INVALID, "squeeze: message length must be 8..128 and a multiple of 8");
    ++c->launches;
    for (size_t i = 0; i < n; ++i) {
        uint64_t key = i;
        for (;;) {
            f192 v;
            orc_blake2b((uint8_t *)&v, 24, msg, msg_len, (const uint8_t *)&key, 8);
            if (field == B200_FIELD_GF192) { E(out)[i] = v; break; }
            /* blake2b.tcc:185-229: clear the bits above the modulus' top bit, re-draw with key += n while >= modulus */
            const uint32_t bits = fp->num_bits;
            for (uint32_t b = bits; b < 192; ++b) v.w[b / 64] &= ~(1ull << (b % 64));
            bool less = false;
            for (int k = 2; k >= 0; --k) { if (v.w[k] != fp->modulus[k]) { less = v.w[k] < fp->modulus[k]; break; } }
            if (less) { E(out)[i] = v; break; }
            key += n;
        }
    }
    return 0;
}

unsigned long long b200_launch_count(b200_ctx *c) { return c->launches; }
void b200_transfer_bytes(b200_ctx *c, unsigned long long *h2d, unsigned long long *d2h) { if (h2d) *h2d = c->h2d; if (d2h) *d2h = c->d2h; }
int b200_profile_begin(b200_ctx *) { return 0; }
int b200_profile_end(b200_ctx *, char *buf, size_t cap) { if (cap) buf[0] = 0; return 0; }

} /* extern "C" */
