/*
 * b200_hotpath.hpp — the reference's hot-path function templates, same names / argument order /
 * by-value-by-reference conventions / element orderings / exceptions, forwarding to the C ABI:
 *   FFT_over_field_subset, IFFT_over_field_subset, IFFT_of_known_degree_over_field_subset
 *                                                  libiop/algebra/fft.hpp:62-88 (fft.tcc:407-475)
 *   evaluate_next_f_i_over_entire_domain          libiop/protocols/ldt/fri/fri_aux.hpp:23-28 (fri_aux.tcc:6-34)
 * The templates are generic over the field type and the domain type: anything with the reference's
 * field_subset method names (type(), dimension(), num_elements(), basis(), shift(), generator(),
 * get_subset_of_order()) works — libiop_b200::field_subset or libiop::field_subset itself — as long as
 * field_traits<FieldT> is specialised (b200_fields.hpp; INTEGRATION.md for libff types).
 */
#ifndef LIBIOP_B200_HOTPATH_HPP_
#define LIBIOP_B200_HOTPATH_HPP_

#include <cstring>
#include <memory>
#include <vector>

#include "b200_device.hpp"
#include "b200_fields.hpp"
#include "b200_residency.hpp"

namespace libiop_b200 {

/* flatten a field_subset-like object into the C ABI's b200_domain; `basis_store` keeps the limbs alive */
template<typename FieldT, typename DomainT>
b200_domain make_domain(const DomainT &domain, std::vector<uint64_t> &basis_store)
{
    b200_domain d;
    std::memset(&d, 0, sizeof d);
    d.dim = (uint32_t)domain.dimension();
    const FieldT shift = domain.shift();
    std::memcpy(d.shift, &shift, 24);
    if ((int)domain.type() == B200_DOMAIN_AFFINE_SUBSPACE) {
        d.kind = B200_DOMAIN_AFFINE_SUBSPACE;
        const std::vector<FieldT> &b = domain.basis();
        basis_store.resize(3 * b.size() + 3);
        if (!b.empty()) std::memcpy(basis_store.data(), b.data(), 24 * b.size());
        d.basis = basis_store.data();
    } else if ((int)domain.type() == B200_DOMAIN_MULTIPLICATIVE_COSET) {
        d.kind = B200_DOMAIN_MULTIPLICATIVE_COSET;
        const FieldT g = domain.generator();
        std::memcpy(d.generator, &g, 24);
    } else {
        throw std::invalid_argument("field_subset type unsupported.");
    }
    return d;
}

/* fft.hpp:62-68.  Device-resident variant first (what a device-aware prover calls), host vectors below. */
template<typename FieldT, typename DomainT>
void FFT_over_field_subset_device(const device_buffer &coeffs, std::size_t n_coeffs, const DomainT &domain, device_buffer &evals_out)
{
    std::vector<uint64_t> store;
    const b200_domain d = make_domain<FieldT>(domain, store);
    int rc;
    if (field_traits<FieldT>::kind == B200_FIELD_GF192) rc = b200_gf192_fft(context(), &d, coeffs.get(), n_coeffs, evals_out.get());
    else rc = b200_fp_ntt(context(), field_traits<FieldT>::fp(), &d, coeffs.get(), n_coeffs, evals_out.get());
    throw_status(context(), rc);
}

template<typename FieldT, typename DomainT>
void IFFT_over_field_subset_device(const device_buffer &evals, const DomainT &domain, device_buffer &coeffs_out)
{
    std::vector<uint64_t> store;
    const b200_domain d = make_domain<FieldT>(domain, store);
    int rc;
    if (field_traits<FieldT>::kind == B200_FIELD_GF192) rc = b200_gf192_ifft(context(), &d, evals.get(), coeffs_out.get());
    else rc = b200_fp_intt(context(), field_traits<FieldT>::fp(), &d, evals.get(), coeffs_out.get());
    throw_status(context(), rc);
}

/* Fused low-degree extension (b200_gf192_lde): evaluations over H -> evaluations over L, equal to
   FFT_over_field_subset(IFFT_over_field_subset(evals, H), L).  Returns false — nothing done — when the pair of domains does
   not qualify (prime field, or L does not extend H's basis) and the caller should run the two transforms. */
template<typename FieldT, typename DomainT>
bool LDE_over_field_subsets_device(const device_buffer &evals, const DomainT &H, const DomainT &L, device_buffer &evals_out)
{
    if (field_traits<FieldT>::kind != B200_FIELD_GF192) return false;
    if ((int)H.type() != B200_DOMAIN_AFFINE_SUBSPACE || (int)L.type() != B200_DOMAIN_AFFINE_SUBSPACE) return false;
    if (H.dimension() > L.dimension()) return false;
    const std::vector<FieldT> &bh = H.basis(), &bl = L.basis();
    for (std::size_t k = 0; k < bh.size(); ++k)
        if (!(bh[k] == bl[k])) return false;
    std::vector<uint64_t> store_h, store_l;
    const b200_domain dh = make_domain<FieldT>(H, store_h), dl = make_domain<FieldT>(L, store_l);
    throw_status(context(), b200_gf192_lde(context(), &dh, &dl, evals.get(), evals_out.get()));
    return true;
}
/* host vectors in, a (device-resident) codeword vector out */
template<typename FieldT, typename DomainT>
std::vector<FieldT> LDE_over_field_subsets(const std::vector<FieldT> &evals, const DomainT &H, const DomainT &L)
{
    if (evals.size() != H.num_elements()) throw std::invalid_argument("LDE_over_field_subsets: evaluation vector does not match the domain");
    const device_view in(evals);
    device_buffer out(L.num_elements());
    if (!LDE_over_field_subsets_device<FieldT>(in.buf(), H, L, out)) {
        device_buffer coeffs(H.num_elements());
        IFFT_over_field_subset_device<FieldT>(in.buf(), H, coeffs);
        FFT_over_field_subset_device<FieldT>(coeffs, H.num_elements(), L, out);
    }
    return make_lazy_vector<FieldT>(std::move(out), L.num_elements());
}

/* The reference's signatures take the vectors BY VALUE (fft.hpp:62-88); the implementations below take references so that a
   caller that already holds the vector (the drop-in dispatchers, which received the reference's by-value copy) does not
   pay a second copy of up to 192 MiB.  The by-value overloads with the reference's names forward to them. */
template<typename FieldT, typename DomainT>
std::vector<FieldT> FFT_over_field_subset_cref(const std::vector<FieldT> &coeffs, const DomainT &domain)
{
    if (coeffs.size() > domain.num_elements())
        throw std::invalid_argument("FFT_over_field_subset: more coefficients than domain elements");
    const device_view in(coeffs);
    device_buffer out(domain.num_elements());
    FFT_over_field_subset_device<FieldT>(in.buf(), coeffs.size(), domain, out);
    return make_lazy_vector<FieldT>(std::move(out), domain.num_elements());   /* the codeword stays on the device for its consumers */
}
template<typename FieldT, typename DomainT>
std::vector<FieldT> FFT_over_field_subset(const std::vector<FieldT> coeffs, DomainT domain)
{ return FFT_over_field_subset_cref<FieldT>(coeffs, domain); }

/* fft.hpp:70-76; `n_evals` leading evaluations of `evals` are transformed (the whole vector by default) */
template<typename FieldT, typename DomainT>
std::vector<FieldT> IFFT_over_field_subset_cref(const std::vector<FieldT> &evals, const DomainT &domain, std::size_t n_evals = ~std::size_t(0))
{
    if (n_evals == ~std::size_t(0)) n_evals = evals.size();
    if (n_evals != domain.num_elements() || n_evals > evals.size())
        throw std::invalid_argument("IFFT_over_field_subset: evaluation vector does not match the domain");
    if (n_evals == 1) return std::vector<FieldT>(evals.begin(), evals.begin() + 1);           /* fft.tcc:397-401 */
    const device_view in(evals.data(), n_evals);              /* a prefix of a device-resident codeword is used in place */
    device_buffer out(n_evals);
    IFFT_over_field_subset_device<FieldT>(in.buf(), domain, out);
    return out.template download<FieldT>(n_evals);             /* coefficients are consumed by host polynomial code */
}
template<typename FieldT, typename DomainT>
std::vector<FieldT> IFFT_over_field_subset(const std::vector<FieldT> evals, DomainT domain)
{ return IFFT_over_field_subset_cref<FieldT>(evals, domain); }

/* fft.hpp:78-88 (fft.tcc:435-475): additive -> first 2^k evaluations over the first k basis vectors (a prefix of the
   caller's vector: used in place when it is device-resident, uploaded otherwise), multiplicative -> every (n/2^k)-th
   evaluation over the order-2^k sub-coset.  The coefficients (2^k of them) stay on the device. */
template<typename FieldT, typename DomainT>
device_buffer IFFT_of_known_degree_over_field_subset_to_device(const std::vector<FieldT> &evals, std::size_t degree, const DomainT &domain)
{
    std::size_t closest_power_of_two = 1;
    while (closest_power_of_two < degree) closest_power_of_two <<= 1;
    if (closest_power_of_two > evals.size() || evals.size() != domain.num_elements())
        throw std::invalid_argument("IFFT_of_known_degree_over_field_subset: evaluation vector does not match the domain");
    const DomainT minimal = domain.get_subset_of_order(closest_power_of_two);
    device_buffer out(closest_power_of_two);
    if ((int)domain.type() == B200_DOMAIN_AFFINE_SUBSPACE) {
        const device_view in(evals.data(), closest_power_of_two);
        if (closest_power_of_two == 1) throw_status(context(), b200_copy(context(), out.get(), 0, in.get(), 0, 1));
        else IFFT_over_field_subset_device<FieldT>(in.buf(), minimal, out);
        return out;
    }
    const std::size_t freq = domain.num_elements() / closest_power_of_two;
    device_buffer sub(closest_power_of_two);
    if (const b200_buf resident = residency::instance().find(evals.data(), evals.size()))
        throw_status(context(), b200_copy_strided(context(), sub.get(), 1, resident, 0, freq, closest_power_of_two));
    else {
        std::vector<FieldT> picked;
        picked.reserve(closest_power_of_two);
        for (std::size_t i = 0; i < domain.num_elements(); i += freq) picked.emplace_back(evals[i]);
        sub.upload(picked);
    }
    if (closest_power_of_two == 1) return sub;
    IFFT_over_field_subset_device<FieldT>(sub, minimal, out);
    return out;
}
template<typename FieldT, typename DomainT>
std::vector<FieldT> IFFT_of_known_degree_over_field_subset_cref(const std::vector<FieldT> &evals, std::size_t degree, const DomainT &domain)
{
    std::size_t closest_power_of_two = 1;
    while (closest_power_of_two < degree) closest_power_of_two <<= 1;
    if ((int)domain.type() == B200_DOMAIN_AFFINE_SUBSPACE)
        return IFFT_over_field_subset_cref<FieldT>(evals, domain.get_subset_of_order(closest_power_of_two), closest_power_of_two);
    const device_buffer coeffs = IFFT_of_known_degree_over_field_subset_to_device<FieldT>(evals, degree, domain);
    return coeffs.template download<FieldT>(closest_power_of_two);
}
template<typename FieldT, typename DomainT>
std::vector<FieldT> IFFT_of_known_degree_over_field_subset(const std::vector<FieldT> evals, std::size_t degree, DomainT domain)
{ return IFFT_of_known_degree_over_field_subset_cref<FieldT>(evals, degree, domain); }

/* Interleaved Reed-Solomon encoding (Ligero): `n_rows` rows, row r holding the |H| evaluations at rows + r * |H|, are
   interpolated over the systematic domain H and evaluated over the codeword domain L — the reference does this one row at
   a time with IFFT_over_field_subset + FFT_over_field_subset (ligero.tcc:293-331, interleaved_lincheck_ot.tcc:283-330);
   here all rows go through b200_gf192_ifft_batch / b200_gf192_fft_batch.  Returns n_rows * |L| elements, row-major.
   `rows` must stay alive and unmodified until the next synchronising call on the context. */
template<typename FieldT, typename DomainT>
device_buffer encode_rows_device(const FieldT *rows, std::size_t n_rows, const DomainT &H, const DomainT &L)
{
    static_assert(field_traits<FieldT>::kind == B200_FIELD_GF192, "batched row encoding is implemented for additive domains");
    const std::size_t nH = H.num_elements(), nL = L.num_elements();
    device_buffer coeffs(n_rows * nH), out(n_rows * nL);
    if (n_rows == 0) return out;
    const device_view in(rows, n_rows * nH);
    std::vector<uint64_t> store_h, store_l;
    const b200_domain dh = make_domain<FieldT>(H, store_h), dl = make_domain<FieldT>(L, store_l);
    throw_status(context(), b200_gf192_ifft_batch(context(), &dh, in.get(), n_rows, coeffs.get()));
    throw_status(context(), b200_gf192_fft_batch(context(), &dl, coeffs.get(), nH, n_rows, out.get()));
    return out;
}

/* the same, handed back as one host vector per row (what oracle<FieldT> wants); every row keeps a device copy in the
   residency cache for the Merkle tree and the linchecks that read it next */
template<typename FieldT, typename DomainT>
std::vector<std::vector<FieldT>> encode_rows(const FieldT *rows, std::size_t n_rows, const DomainT &H, const DomainT &L)
{
    const std::size_t nL = L.num_elements();
    const device_buffer all = encode_rows_device<FieldT>(rows, n_rows, H, L);
    std::vector<std::vector<FieldT>> result;
    result.reserve(n_rows);
    for (std::size_t r = 0; r < n_rows; ++r) {
        device_buffer row(nL);
        throw_status(context(), b200_copy(context(), row.get(), 0, all.get(), r * nL, nL));
        result.emplace_back(make_lazy_vector<FieldT>(std::move(row), nL));
    }
    return result;
}

/* fri_aux.hpp:23-28 */
template<typename FieldT, typename DomainT>
void evaluate_next_f_i_over_entire_domain_device(const device_buffer &f_i, const DomainT &f_i_domain, std::size_t coset_size,
                                                 const FieldT x_i, device_buffer &f_next_out)
{
    std::vector<uint64_t> store;
    const b200_domain d = make_domain<FieldT>(f_i_domain, store);
    uint32_t eta = 0;
    while ((std::size_t(1) << eta) < coset_size) ++eta;
    uint64_t x[3];
    std::memcpy(x, &x_i, 24);
    throw_status(context(), b200_fri_fold(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), &d, eta, x,
                                          f_i.get(), f_next_out.get()));
}

template<typename FieldT, typename DomainT>
std::shared_ptr<std::vector<FieldT>> evaluate_next_f_i_over_entire_domain(
    const std::shared_ptr<std::vector<FieldT>> &f_i_evals, const DomainT &f_i_domain, const std::size_t coset_size, const FieldT x_i)
{
    const std::size_t n = f_i_evals->size();
    const device_view in(*f_i_evals);
    device_buffer out(n / coset_size);
    evaluate_next_f_i_over_entire_domain_device<FieldT>(in.buf(), f_i_domain, coset_size, x_i, out);
    return std::make_shared<std::vector<FieldT>>(make_lazy_vector<FieldT>(std::move(out), n / coset_size));
}

} // namespace libiop_b200
#endif
