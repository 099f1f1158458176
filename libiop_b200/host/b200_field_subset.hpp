/*
 * b200_field_subset.hpp — domain objects of the adapter, mirroring the reference's plugin surface:
 *   field_subset<FieldT>            libiop/algebra/field_subset/field_subset.hpp:25-77
 *   affine_subspace<FieldT>         libiop/algebra/field_subset/subspace.hpp:17-75, subspace.tcc:54-108,161-170,220-224
 *   multiplicative_coset<FieldT>    libiop/algebra/field_subset/subgroup.hpp:21-105, subgroup.tcc:56-59,176-197,280-315
 * Same method names, orderings and index math; only what the hot path needs is kept.
 */
#ifndef LIBIOP_B200_FIELD_SUBSET_HPP_
#define LIBIOP_B200_FIELD_SUBSET_HPP_

#include <cstddef>
#include <stdexcept>
#include <vector>

#include "b200_fields.hpp"

namespace libiop_b200 {

enum field_subset_type { affine_subspace_type = 1, multiplicative_coset_type = 2 }; /* field_subset.hpp:18-21 */

inline std::size_t log2_ceil(std::size_t n) { std::size_t r = 0; while ((std::size_t(1) << r) < n) ++r; return r; }

template<typename FieldT>
class affine_subspace {
    std::vector<FieldT> basis_;
    FieldT shift_;
public:
    affine_subspace() = default;
    affine_subspace(const std::vector<FieldT> &basis, const FieldT &shift = FieldT(0)) : basis_(basis), shift_(shift) {}
    /* subspace.tcc:94-108,220-224: basis x^i (the integer 1<<i), arbitrary shift */
    static affine_subspace shifted_standard_basis(std::size_t dimension, const FieldT &shift)
    {
        std::vector<FieldT> b;
        for (std::size_t i = 0; i < dimension; ++i) b.emplace_back(FieldT(uint64_t(1) << i));
        return affine_subspace(b, shift);
    }
    std::size_t dimension() const { return basis_.size(); }
    std::size_t num_elements() const { return std::size_t(1) << basis_.size(); }
    const std::vector<FieldT> &basis() const { return basis_; }
    const FieldT shift() const { return shift_; }
    FieldT element_by_index(std::size_t index) const /* subspace.tcc:54-71,167-170 */
    {
        if (index >= num_elements()) throw std::invalid_argument("element index out of bounds");
        FieldT r = shift_;
        for (std::size_t i = 0; i < basis_.size(); ++i) if (index & (std::size_t(1) << i)) r += basis_[i];
        return r;
    }
};

template<typename FieldT>
class multiplicative_coset {
    std::size_t order_ = 0;
    FieldT g_, shift_;
public:
    multiplicative_coset() = default;
    multiplicative_coset(std::size_t order, const FieldT &shift) : order_(order), shift_(shift)
    {
        if (order == 0 || (order & (order - 1))) throw std::invalid_argument("The order of the subgroup must be a power of two.");
        g_ = FieldT::subgroup_generator((uint32_t)log2_ceil(order)); /* subgroup.tcc:56-59 */
    }
    std::size_t dimension() const { return log2_ceil(order_); }
    std::size_t num_elements() const { return order_; }
    FieldT generator() const { return g_; }
    FieldT shift() const { return shift_; }
    FieldT element_by_index(std::size_t index) const { return shift_ * g_.pow(index); } /* subgroup.tcc:296-304 */
};

template<typename FieldT>
class field_subset {
    field_subset_type type_;
    affine_subspace<FieldT> subspace_;
    multiplicative_coset<FieldT> coset_;
    template<typename F, int K> struct maker;
    template<typename F> struct maker<F, B200_FIELD_GF192> {
        static void make(field_subset &s, std::size_t n, bool has_shift, const F &shift)
        {
            /* field_subset.tcc: additive fields get the standard basis; the one-argument form is unshifted */
            s.type_ = affine_subspace_type;
            s.subspace_ = affine_subspace<F>::shifted_standard_basis(log2_ceil(n), has_shift ? shift : F(0));
        }
    };
    template<typename F> struct maker<F, B200_FIELD_FP> {
        static void make(field_subset &s, std::size_t n, bool has_shift, const F &shift)
        {
            s.type_ = multiplicative_coset_type;
            s.coset_ = multiplicative_coset<F>(n, has_shift ? shift : F::one());
        }
    };
public:
    field_subset() : type_(affine_subspace_type) {}
    explicit field_subset(std::size_t num_elements) { maker<FieldT, field_traits<FieldT>::kind>::make(*this, num_elements, false, FieldT()); }
    field_subset(std::size_t num_elements, const FieldT coset_shift) { maker<FieldT, field_traits<FieldT>::kind>::make(*this, num_elements, true, coset_shift); }
    field_subset(const affine_subspace<FieldT> subspace) : type_(affine_subspace_type), subspace_(subspace) {}
    field_subset(const multiplicative_coset<FieldT> coset) : type_(multiplicative_coset_type), coset_(coset) {}

    field_subset_type type() const { return type_; }
    std::size_t dimension() const { return type_ == affine_subspace_type ? subspace_.dimension() : coset_.dimension(); }
    std::size_t num_elements() const { return std::size_t(1) << dimension(); }
    const affine_subspace<FieldT> &subspace() const { return subspace_; }
    const multiplicative_coset<FieldT> &coset() const { return coset_; }
    const std::vector<FieldT> &basis() const { return subspace_.basis(); }
    const FieldT shift() const { return type_ == affine_subspace_type ? subspace_.shift() : coset_.shift(); }
    FieldT generator() const { return coset_.generator(); }
    FieldT element_by_index(std::size_t i) const { return type_ == affine_subspace_type ? subspace_.element_by_index(i) : coset_.element_by_index(i); }

    /* coset index math: subspace.tcc:74-91, subgroup.tcc:176-197 */
    std::size_t coset_index(std::size_t position, std::size_t coset_size) const
    {
        return type_ == affine_subspace_type ? position / coset_size : position % (num_elements() / coset_size);
    }
    std::size_t intra_coset_index(std::size_t position, std::size_t coset_size) const
    {
        return type_ == affine_subspace_type ? position % coset_size : position / (num_elements() / coset_size);
    }
    std::size_t position_by_coset_indices(std::size_t coset_index, std::size_t intra, std::size_t coset_size) const
    {
        return type_ == affine_subspace_type ? coset_index * coset_size + intra : coset_index + intra * (num_elements() / coset_size);
    }
    /* field_subset.tcc:218-237: first log2(order) basis vectors, same shift / every (n/order)-th element */
    field_subset get_subset_of_order(std::size_t order) const
    {
        if (type_ == affine_subspace_type) {
            std::vector<FieldT> b = subspace_.basis();
            b.resize(log2_ceil(order));
            return field_subset(affine_subspace<FieldT>(b, subspace_.shift()));
        }
        return field_subset(order, coset_.shift());
    }
};

} // namespace libiop_b200
#endif
