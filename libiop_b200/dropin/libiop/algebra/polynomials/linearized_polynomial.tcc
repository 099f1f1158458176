/* Shadow of libiop/algebra/polynomials/linearized_polynomial.tcc for the drop-in build.
   polynomial_over_linearized_polynomial (linearized_polynomial.tcc:238-291) — the division of the sumcheck polynomial by
   Z_H that yields h (sumcheck.tcc:363-366) — clears one coefficient at a time on the host (|P|-|H| steps of log|H|+1
   multiplications each, loop-carried).  b200_poly_divide_sparse runs the same recurrence block-parallel on the device.
   Small inputs keep the reference loop. */
#include <cstdio>
#include <cstdlib>
#include "b200_dropin_common.hpp"
#include "libiop/algebra/fft.hpp"

#define polynomial_over_linearized_polynomial ref_cpu_polynomial_over_linearized_polynomial
#include B200_REF_FILE(libiop/algebra/polynomials/linearized_polynomial.tcc)
#undef polynomial_over_linearized_polynomial

namespace libiop {
namespace b200_glue {

/* ---- division by a LOW-degree linearized polynomial under a long dividend (f_w / Z_{1,v}: r1cs_rs_iop.tcc:554-568, degree 2^20
   over degree 32).  The recurrence of b200_poly_divide_sparse advances only deg_Z / 2 coefficients per step there; instead:
     1. the roots of Z form an affine subspace S of dimension log2(deg Z): its direction space is the kernel of the F2-linear
        part of Z and its offset solves L(y) = Z(0) — 192 x 192 Gaussian elimination over GF(2) on the host;
     2. R = P mod Z is the interpolation over S of P's values there (deg R < |S|): |S| device evaluations of P;
     3. Q = (P - R) / Z exactly, so on any domain D' disjoint from S the values of Q are (P - R)(x) / Z(x): one FFT, one
        element-wise inversion and product, one IFFT over a 2^ceil(log2 |P|)-point D'.
   Quotient and remainder of a polynomial division are unique, so these are the reference's coefficients.  Returns false
   (nothing computed) whenever a step does not apply; the caller then runs the reference loop. */
inline bool gf2_bit(const uint64_t w[3], unsigned i) { return (w[i >> 6] >> (i & 63)) & 1u; }
template<typename FieldT>
bool roots_of_linearized(const linearized_polynomial<FieldT> &Z, std::vector<FieldT> &basis, FieldT &offset)
{
    static_assert(sizeof(FieldT) == 24, "GF(2^192) elements are 3 x u64");
    struct row { uint64_t img[3], pre[3]; };
    const auto linear_part = [&Z](const FieldT &y) { return Z.evaluation_at_point(y) - Z.constant_coefficient(); };
    std::vector<row> pivots;                         /* reduced rows with distinct leading image bits */
    std::vector<int> lead;
    const auto reduce = [&](row &r) {
        for (std::size_t p = 0; p < pivots.size(); ++p)
            if (gf2_bit(r.img, (unsigned)lead[p])) for (int k = 0; k < 3; ++k) { r.img[k] ^= pivots[p].img[k]; r.pre[k] ^= pivots[p].pre[k]; }
    };
    const auto top_bit = [](const uint64_t w[3]) { for (int i = 191; i >= 0; --i) if (gf2_bit(w, (unsigned)i)) return i; return -1; };
    basis.clear();
    for (unsigned i = 0; i < 192; ++i) {
        row r; std::memset(&r, 0, sizeof r);
        r.pre[i >> 6] = uint64_t(1) << (i & 63);
        FieldT e; std::memcpy(&e, r.pre, 24);
        const FieldT image = linear_part(e);
        std::memcpy(r.img, &image, 24);
        reduce(r);
        const int t = top_bit(r.img);
        if (t < 0) { FieldT k; std::memcpy(&k, r.pre, 24); basis.push_back(k); continue; }     /* a kernel vector */
        for (std::size_t p = 0; p < pivots.size(); ++p)      /* keep the pivot set fully reduced */
            if (gf2_bit(pivots[p].img, (unsigned)t)) for (int k = 0; k < 3; ++k) { pivots[p].img[k] ^= r.img[k]; pivots[p].pre[k] ^= r.pre[k]; }
        pivots.push_back(r); lead.push_back(t);
    }
    row target; std::memset(&target, 0, sizeof target);
    const FieldT c = Z.constant_coefficient();
    std::memcpy(target.img, &c, 24);
    reduce(target);
    if (top_bit(target.img) >= 0) return false;          /* Z(0) is not in the image: no roots in the field */
    std::memcpy(&offset, target.pre, 24);                 /* L(offset) = Z(0), i.e. Z(offset) = 0 in characteristic 2 */
    return (std::size_t(1) << basis.size()) == Z.degree();
}

template<typename FieldT>
bool divide_by_low_degree_linearized(const polynomial<FieldT> &P, const linearized_polynomial<FieldT> &Z,
                                     std::pair<polynomial<FieldT>, polynomial<FieldT>> &result)
{
    using namespace libiop_b200;
    if (field_traits<FieldT>::kind != B200_FIELD_GF192) return false;
    const std::size_t n_p = P.coefficients().size(), deg = Z.degree(), n_q = n_p - deg;
    std::vector<FieldT> root_basis;
    FieldT root_offset;
    if (deg < 2 || deg > 4096 || !roots_of_linearized<FieldT>(Z, root_basis, root_offset)) return false;
    const field_subset<FieldT> S(affine_subspace<FieldT>(root_basis, root_offset));
    /* 2. remainder: interpolate P over S */
    const device_view d_P(P.coefficients());
    std::vector<FieldT> at_roots(deg);
    for (std::size_t i = 0; i < deg; ++i) at_roots[i] = poly_eval<FieldT>(d_P.buf(), n_p, S.element_by_index(i));
    std::vector<FieldT> R = ref_cpu_IFFT_over_field_subset<FieldT>(at_roots, S);          /* |S| <= 4096: host */
    R.resize(deg, FieldT::zero());
    /* 3. quotient through evaluations over D' = x^k' + <1, x, ..., x^(k-1)>, k = ceil(log2 |P|), Z(x) != 0 on D' */
    std::size_t k = 0;
    while ((std::size_t(1) << k) < n_p) ++k;
    const std::size_t n = std::size_t(1) << k;
    std::vector<FieldT> std_basis;
    for (std::size_t i = 0; i < k; ++i) { uint64_t w[3] = {0, 0, 0}; w[i >> 6] = uint64_t(1) << (i & 63); FieldT b; std::memcpy(&b, w, 24); std_basis.push_back(b); }
    device_buffer d_Z(n), d_work(n), d_eval(n);
    bool found = false;
    field_subset<FieldT> D;
    for (std::size_t s = k; s < k + 6 && s < 192 && !found; ++s) {
        uint64_t w[3] = {0, 0, 0}; w[s >> 6] = uint64_t(1) << (s & 63);
        FieldT shift; std::memcpy(&shift, w, 24);
        D = field_subset<FieldT>(affine_subspace<FieldT>(std_basis, shift));
        subspace_function_product<FieldT>(d_Z, k, { affine_function_table<FieldT>(D, [&Z](const FieldT &x) { return Z.evaluation_at_point(x); }, FieldT::one()) });
        std::size_t zero_at = 0;
        throw_status(context(), b200_vec_find_zero(context(), d_Z.get(), n, &zero_at));
        found = zero_at == n;
    }
    if (!found) return false;
    /* P - R on the device: only the low deg coefficients change (subtraction = addition in characteristic 2) */
    throw_status(context(), b200_fill_zero(context(), d_work.get(), 0, n));
    throw_status(context(), b200_copy(context(), d_work.get(), 0, d_P.get(), 0, n_p));
    device_buffer d_R(deg);
    d_R.upload(R);
    throw_status(context(), b200_vec_mul_add(context(), B200_FIELD_GF192, nullptr, d_work.get(), d_work.get(), 0, d_R.get(), 0, 1, deg));
    FFT_over_field_subset_device<FieldT>(d_work, n, D, d_eval);
    vec_inverse<FieldT>(d_Z, d_Z, n);
    vec_mul_add<FieldT>(d_eval, d_eval, &d_Z, nullptr, nullptr, false, n);
    IFFT_over_field_subset_device<FieldT>(d_eval, D, d_work);
    result = std::make_pair(polynomial<FieldT>(d_work.template download<FieldT>(n_q)), polynomial<FieldT>(std::move(R)));
    if (std::getenv("B200_DIVIDE_TRACE")) std::fprintf(stderr, "[b200] division by a degree-%zu linearized polynomial through evaluations: |P| = %zu, domain 2^%zu\n", deg, n_p, k);
    return true;
}

template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> divide_linearized(const polynomial<FieldT> &P, const linearized_polynomial<FieldT> &Z, std::true_type)
{
    const std::size_t deg_Z = Z.degree();
    /* the device recurrence advances deg_Z/2 quotient coefficients per launch: a low-degree divisor under a long dividend
       (Z_{1,v} under f_w, r1cs_rs_iop.tcc:554-568) goes through evaluations instead (divide_by_low_degree_linearized) */
    if (Z.num_terms() >= 2 && P.degree() >= deg_Z && P.coefficients().size() - deg_Z >= 8192 &&
        (P.coefficients().size() - deg_Z) / (deg_Z - deg_Z / 2) > 64) {
        std::pair<polynomial<FieldT>, polynomial<FieldT>> qr;
        if (divide_by_low_degree_linearized<FieldT>(P, Z, qr)) return qr;
    }
    if (Z.num_terms() < 2 || P.degree() < deg_Z || P.coefficients().size() - deg_Z < 4096 ||
        (P.coefficients().size() - deg_Z) / (deg_Z - deg_Z / 2) > 64)
        return ref_cpu_polynomial_over_linearized_polynomial<FieldT>(P, Z);
    std::vector<FieldT> coefs;
    std::vector<std::size_t> pows;
    std::size_t p_pow = deg_Z / 2;                       /* Z[j] multiplies y^(2^(j-1)) for j >= 1, Z[0] is the constant */
    for (std::size_t j = Z.num_terms() - 1; j--; ) { coefs.push_back(Z[j]); pows.push_back(p_pow); p_pow /= 2; }
    auto qr = libiop_b200::poly_divide_sparse<FieldT>(P.coefficients(), deg_Z, Z[Z.num_terms() - 1], coefs, pows);
    return std::make_pair(polynomial<FieldT>(std::move(qr.first)), polynomial<FieldT>(std::move(qr.second)));
}
template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> divide_linearized(const polynomial<FieldT> &P, const linearized_polynomial<FieldT> &Z, std::false_type)
{ return ref_cpu_polynomial_over_linearized_polynomial<FieldT>(P, Z); }
}

template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> polynomial_over_linearized_polynomial(const polynomial<FieldT> &P,
                                                                                        const linearized_polynomial<FieldT> &Z)
{ return b200_glue::divide_linearized<FieldT>(P, Z, libiop_b200::b200_enabled<FieldT>()); }
} // namespace libiop
