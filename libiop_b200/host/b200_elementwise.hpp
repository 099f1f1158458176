/*
 * b200_elementwise.hpp — C++ face of the element-wise device passes (include/b200_iop.h, "element-wise passes"):
 * the building blocks of the reference's virtual-oracle evaluations over device buffers, plus the
 * domain-derived vectors they need (elements of L, a vanishing polynomial over L, powers of the elements of L),
 * produced on the device from O(dim) host values instead of being materialised by host loops
 * (algebra/utils.tcc:8-30, algebra/exponentiation.tcc:3-88, algebra/polynomials/vanishing_polynomial.tcc:98-165).
 *
 * Generic over the field type (field_traits<FieldT>) and over anything with the reference's field_subset /
 * polynomial method names.  Host arithmetic on the handful of table entries uses FieldT's own operators.
 */
#ifndef LIBIOP_B200_ELEMENTWISE_HPP_
#define LIBIOP_B200_ELEMENTWISE_HPP_

#include <cstring>
#include <list>
#include <mutex>
#include <stdexcept>
#include <utility>
#include <vector>

#include "b200_device.hpp"
#include "b200_fields.hpp"

namespace libiop_b200 {

template<typename FieldT> inline const uint64_t *limbs(const FieldT &x) { return reinterpret_cast<const uint64_t *>(&x); }

template<typename FieldT> inline FieldT field_pow(FieldT base, std::size_t e)
{
    FieldT r = FieldT::one();
    while (e) { if (e & 1) r *= base; base *= base; e >>= 1; }
    return r;
}

/* out[i] = sum_k coeffs[k] * vecs[k][i] */
template<typename FieldT>
void vec_lincomb(device_buffer &out, std::size_t n, const std::vector<FieldT> &coeffs, const std::vector<const device_buffer *> &vecs)
{
    static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
    std::vector<b200_buf> h;
    for (auto *v : vecs) h.push_back(v->get());
    throw_status(context(), b200_vec_lincomb(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), out.get(), n, h.size(),
                                             reinterpret_cast<const uint64_t *>(coeffs.data()), h.data()));
}

/* out = a*b (+|-) c*d with absent operands (nullptr) read as "one" (b, d) or "no second term" (c) */
template<typename FieldT>
void vec_mul_add(device_buffer &out, const device_buffer &a, const device_buffer *b, const device_buffer *c, const device_buffer *d,
                 bool subtract, std::size_t n)
{
    throw_status(context(), b200_vec_mul_add(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), out.get(), a.get(),
                                             b ? b->get() : 0, c ? c->get() : 0, d ? d->get() : 0, subtract ? 1 : 0, n));
}

/* out = a * mul + add with scalar mul / add (nullptr = absent) */
template<typename FieldT>
void vec_affine(device_buffer &out, const device_buffer &a, const FieldT *mul, const FieldT *add, std::size_t n)
{
    throw_status(context(), b200_vec_affine(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), out.get(), a.get(),
                                            mul ? limbs(*mul) : nullptr, add ? limbs(*add) : nullptr, n));
}

/* out[i] = a[i]^-1 (zero stays zero, as in the reference's batch_inverse with has_zeroes, algebra/utils.tcc:73-103) */
template<typename FieldT>
void vec_inverse(device_buffer &out, const device_buffer &a, std::size_t n)
{
    if (field_traits<FieldT>::kind == B200_FIELD_GF192) throw_status(context(), b200_gf192_inv(context(), out.get(), a.get(), n));
    else throw_status(context(), b200_fp_inv(context(), field_traits<FieldT>::fp(), out.get(), a.get(), n));
}

/* out[i] = a[i] * table[(i >> shift) & mask] */
template<typename FieldT>
void vec_mul_table(device_buffer &out, const device_buffer *a, const std::vector<FieldT> &table, unsigned shift, std::size_t mask, std::size_t n)
{
    device_buffer t(table.size());
    t.upload(table);
    throw_status(context(), b200_vec_mul_table(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), out.get(),
                                               a ? a->get() : 0, t.get(), shift, mask, n));
    throw_status(context(), b200_sync(context()));   /* t is released on return */
}

/* One F2-affine function over an affine subspace, given by its values at the shift and at shift + basis[j]:
   table = { f(shift), f(shift + basis[0]) - f(shift), ... }.  `scale` multiplies the whole function. */
template<typename FieldT, typename DomainT, typename Fn>
std::vector<FieldT> affine_function_table(const DomainT &domain, Fn f, const FieldT &scale)
{
    const FieldT at_shift = f(domain.shift());
    std::vector<FieldT> t;
    t.reserve(domain.dimension() + 1);
    t.push_back(at_shift * scale);
    for (const FieldT &b : domain.basis()) t.push_back((f(domain.shift() + b) - at_shift) * scale);
    return t;
}

/* out[i] = prod_k (affine function k)(element i of the subspace): b200_gf192_subset_sum_products */
template<typename FieldT>
void subspace_function_product(device_buffer &out, std::size_t dim, const std::vector<std::vector<FieldT>> &tables)
{
    std::vector<FieldT> flat;
    for (auto &t : tables) flat.insert(flat.end(), t.begin(), t.end());
    throw_status(context(), b200_gf192_subset_sum_products(context(), out.get(), (uint32_t)dim, (uint32_t)tables.size(),
                                                           reinterpret_cast<const uint64_t *>(flat.data())));
}

/* out[i] = a * r^i + add over n positions (prime field) */
template<typename FieldT>
void geometric_sequence(device_buffer &out, std::size_t n, const FieldT &a, const FieldT &r, const FieldT &add)
{
    throw_status(context(), b200_fp_geometric(context(), field_traits<FieldT>::fp(), out.get(), n, limbs(a), limbs(r), limbs(add)));
}

/* out[i] = base^i, i < n (n a power of two for GF(2^192)): the alpha powers of multi_lincheck's set_challenge
   (lincheck/basic_lincheck_aux.tcc:38-46).  Binary field: base^i = prod_j (bit j of i ? base^(2^j) : 1), i.e. log2(n) factors
   that are affine in the index bits, generated by b200_gf192_subset_sum_products; prime field: a geometric sequence. */
template<typename FieldT>
void powers_of(device_buffer &out, std::size_t n, const FieldT &base)
{
    if (field_traits<FieldT>::kind != B200_FIELD_GF192) { geometric_sequence<FieldT>(out, n, FieldT::one(), base, FieldT::zero()); return; }
    std::size_t dim = 0;
    while ((std::size_t(1) << dim) < n) ++dim;
    if ((std::size_t(1) << dim) != n) throw std::invalid_argument("powers_of: length must be a power of two");
    if (dim == 0) { const std::vector<FieldT> one(1, FieldT::one()); out.upload(one); return; }
    std::vector<std::vector<FieldT>> tables(dim, std::vector<FieldT>(dim + 1, FieldT::zero()));
    FieldT sq = base;
    for (std::size_t j = 0; j < dim; ++j) {
        tables[j][0] = FieldT::one();
        tables[j][1 + j] = sq - FieldT::one();          /* factor j is 1 when bit j is clear, base^(2^j) when it is set */
        sq *= sq;
    }
    subspace_function_product<FieldT>(out, dim, tables);
}

/* GF(2^192): y[dst[k]] += scale * coeff[k] * x[src[k]] over a list of matrix entries (b200_gf192_scatter_mul_add) */
template<typename FieldT>
void scatter_mul_add(device_buffer &y, std::size_t n_y, const std::vector<uint32_t> &dst, const std::vector<uint32_t> &src,
                     const std::vector<FieldT> &coeff, const device_buffer &x, std::size_t n_x, const FieldT &scale)
{
    static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
    throw_status(context(), b200_gf192_scatter_mul_add(context(), y.get(), n_y, dst.size(), dst.data(), src.data(),
                                                       reinterpret_cast<const uint64_t *>(coeff.data()), x.get(), n_x, limbs(scale)));
}

/* scale * x^exponent for every element x of the domain, in domain order
   (subset_element_powers, algebra/exponentiation.tcc:74-86, times a constant) */
template<typename FieldT, typename DomainT>
void domain_element_powers(device_buffer &out, const DomainT &domain, std::size_t exponent, const FieldT &scale)
{
    if ((int)domain.type() == B200_DOMAIN_AFFINE_SUBSPACE) {
        std::vector<std::vector<FieldT>> tables;
        for (std::size_t bit = 0; bit < 8 * sizeof(exponent); ++bit) {
            if (!((exponent >> bit) & 1)) continue;
            /* x -> x^(2^bit) is F2-linear */
            tables.push_back(affine_function_table<FieldT>(domain,
                [bit](FieldT x) { for (std::size_t s = 0; s < bit; ++s) x *= x; return x; },
                tables.empty() ? scale : FieldT::one()));
        }
        if (tables.empty())   /* exponent 0: the constant `scale` */
            tables.push_back(affine_function_table<FieldT>(domain, [](FieldT) { return FieldT::one(); }, scale));
        subspace_function_product<FieldT>(out, domain.dimension(), tables);
    } else {
        geometric_sequence<FieldT>(out, domain.num_elements(), scale * field_pow(domain.shift(), exponent),
                                   field_pow(domain.generator(), exponent), FieldT::zero());
    }
}

/* sign * x + offset for every element x of the domain, sign = +1 or -1 (all_elements / all_subset_sums with a shifted
   offset, e.g. x - eval_point in boundary_constraint.tcc:30-50 or x_0 - y in lagrange_polynomial.tcc:72-85) */
template<typename FieldT, typename DomainT>
void domain_elements_affine(device_buffer &out, const DomainT &domain, bool negate, const FieldT &offset)
{
    const FieldT sign = negate ? -FieldT::one() : FieldT::one();
    if ((int)domain.type() == B200_DOMAIN_AFFINE_SUBSPACE) {
        std::vector<FieldT> t = affine_function_table<FieldT>(domain, [](const FieldT &x) { return x; }, sign);
        t[0] += offset;
        subspace_function_product<FieldT>(out, domain.dimension(), { t });
    } else {
        geometric_sequence<FieldT>(out, domain.num_elements(), sign * domain.shift(), domain.generator(), offset);
    }
}

/* x^-1 for every element x of an affine subspace, in domain order (zero stays zero): a table that depends on the domain
   alone, like the transforms' twiddle tables — sumcheck's g oracle needs it over the codeword domain in every proof
   (sumcheck.tcc:82-99: x^(|H|-1) = x^|H| * x^-1; 4.2 ms of inversions at |L| = 2^23).  The last two domains' tables stay on
   the device (|L| * 24 bytes each); the returned buffer is valid until two other domains have been asked for. */
template<typename FieldT, typename DomainT>
const device_buffer &subspace_element_inverses(const DomainT &domain)
{
    static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
    struct entry { std::vector<FieldT> key; device_buffer inv; };
    static std::list<entry> *cache = new std::list<entry>();           /* never destroyed: device buffers outlive no context */
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    std::vector<FieldT> key(1, domain.shift());
    key.insert(key.end(), domain.basis().begin(), domain.basis().end());
    for (auto it = cache->begin(); it != cache->end(); ++it)
        if (it->key == key) { cache->splice(cache->begin(), *cache, it); return cache->front().inv; }
    const std::size_t n = domain.num_elements();
    device_buffer inv(n);
    domain_elements_affine<FieldT>(inv, domain, false, FieldT::zero());
    vec_inverse<FieldT>(inv, inv, n);
    while (cache->size() >= 2) { throw_status(context(), b200_sync(context())); cache->pop_back(); }
    cache->push_front(entry{ std::move(key), std::move(inv) });
    return cache->front().inv;
}

/* Z(x) for every element x of the domain, Z a vanishing polynomial of a subset H of the same kind
   (vanishing_polynomial::evaluations_over_field_subset, vanishing_polynomial.tcc:98-165):
   additive: Z is F2-affine; multiplicative: Z = X^|H| - const. */
template<typename FieldT, typename DomainT, typename VanishingT>
void vanishing_polynomial_over_domain(device_buffer &out, const VanishingT &Z, const DomainT &domain)
{
    if ((int)domain.type() == B200_DOMAIN_AFFINE_SUBSPACE) {
        subspace_function_product<FieldT>(out, domain.dimension(),
            { affine_function_table<FieldT>(domain, [&Z](const FieldT &x) { return Z.evaluation_at_point(x); }, FieldT::one()) });
    } else {
        const std::size_t deg = Z.degree();
        const FieldT a = field_pow(domain.shift(), deg);
        geometric_sequence<FieldT>(out, domain.num_elements(), a, field_pow(domain.generator(), deg), Z.evaluation_at_point(domain.shift()) - a);
    }
}

/* polynomial::evaluation_at_point on the device (coefficients resident in `coeffs`) */
template<typename FieldT>
FieldT poly_eval(const device_buffer &coeffs, std::size_t n, const FieldT &x)
{
    FieldT r;
    throw_status(context(), b200_poly_eval(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), coeffs.get(), n, limbs(x),
                                           reinterpret_cast<uint64_t *>(&r)));
    return r;
}

/* P = Q*Z + R for a sparse divisor Z = lead*y^deg + sum_j coefs[j]*y^pows[j] (b200_poly_divide_sparse), device to device:
   n_p > deg coefficients in d_p; d_q receives n_p - deg, d_r (may be null) deg coefficients */
template<typename FieldT>
void poly_divide_sparse_device(const device_buffer &d_p, std::size_t n_p, std::size_t deg, const FieldT &lead,
                               const std::vector<FieldT> &coefs, const std::vector<std::size_t> &pows, device_buffer &d_q, device_buffer *d_r)
{
    const FieldT lead_inv = lead.inverse();
    throw_status(context(), b200_poly_divide_sparse(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), d_p.get(), n_p, deg,
                                                    limbs(lead_inv), coefs.size(), reinterpret_cast<const uint64_t *>(coefs.data()),
                                                    pows.data(), d_q.get(), d_r ? d_r->get() : 0));
}
/* host vectors: returns {Q, R} with |Q| = |P| - deg and |R| = deg.  Requires |P| > deg. */
template<typename FieldT>
std::pair<std::vector<FieldT>, std::vector<FieldT>> poly_divide_sparse(const std::vector<FieldT> &P, std::size_t deg, const FieldT &lead,
                                                                       const std::vector<FieldT> &coefs, const std::vector<std::size_t> &pows)
{
    device_buffer d_p(P.size()), d_q(P.size() - deg), d_r(deg);
    d_p.upload_async(P);
    poly_divide_sparse_device<FieldT>(d_p, P.size(), deg, lead, coefs, pows, d_q, &d_r);
    std::vector<FieldT> q = d_q.template download<FieldT>(P.size() - deg);
    std::vector<FieldT> r = d_r.template download<FieldT>(deg);
    return std::make_pair(std::move(q), std::move(r));
}

/*
 * Small cache of device copies of immutable host vectors, keyed by (address, length) and validated by a sample of
 * the contents (first, last and 30 strided elements) so that a recycled allocation is never mistaken for the vector
 * it replaced.  Used for polynomial coefficient vectors that are evaluated at many points
 * (bcs_prover::get_transcript -> multi_lincheck_virtual_oracle::evaluation_at_point, basic_lincheck_aux.tcc:160-161).
 */
template<typename FieldT>
class device_vector_cache {
    struct entry { const FieldT *ptr; std::size_t n; std::vector<FieldT> sample; device_buffer buf; };
    std::list<entry> entries_;
    std::size_t max_entries_;
    static std::vector<FieldT> sample_of(const FieldT *p, std::size_t n)
    {
        std::vector<FieldT> s;
        const std::size_t k = n < 32 ? n : 32;
        for (std::size_t i = 0; i < k; ++i) s.push_back(p[(k > 1) ? (i * (n - 1)) / (k - 1) : 0]);
        return s;
    }
public:
    explicit device_vector_cache(std::size_t max_entries) : max_entries_(max_entries) {}
    const device_buffer &get(const std::vector<FieldT> &v)
    {
        const std::vector<FieldT> s = sample_of(v.data(), v.size());
        for (auto it = entries_.begin(); it != entries_.end(); ++it) {
            if (it->ptr == v.data() && it->n == v.size() && std::memcmp(it->sample.data(), s.data(), s.size() * sizeof(FieldT)) == 0) {
                entries_.splice(entries_.begin(), entries_, it);
                return entries_.front().buf;
            }
        }
        if (entries_.size() >= max_entries_) entries_.pop_back();
        entries_.emplace_front(entry{v.data(), v.size(), s, device_buffer(v.size())});
        entries_.front().buf.upload(v);
        return entries_.front().buf;
    }
    void clear() { entries_.clear(); }
};

/* verifier random message of the BCS hash chain (blake2b_FieldT_randomness_extractor, bcs/hashing/blake2b.tcc:231-256):
   n elements drawn by keyed BLAKE2b from `root || squeeze index` */
template<typename FieldT>
std::vector<FieldT> squeeze_field_elements(const unsigned char *root_plus_index, std::size_t len, std::size_t n)
{
    device_buffer out(n);
    throw_status(context(), b200_squeeze_field_elements(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(),
                                                        root_plus_index, len, n, out.get()));
    return out.download<FieldT>(n);
}

} // namespace libiop_b200
#endif
