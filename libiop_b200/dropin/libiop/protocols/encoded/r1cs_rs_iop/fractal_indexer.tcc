/* Shadow of libiop/protocols/encoded/r1cs_rs_iop/fractal_indexer.tcc for the drop-in build (Fractal, BASELINE config 3).
   matrix_indexer::compute_oracles_over_K (fractal_indexer.tcc:52-123) walks the non-zero entries of the matrix and, per
   entry, exponentiates the domain generator twice (element_by_index on a coset whose element cache is empty,
   subgroup.tcc:295-303), evaluates the formal derivative of Z_H (another exponentiation, vanishing_polynomial.tcc:55-74)
   and inverts it — about 2^18 x 3 modular exponentiations and inversions per matrix, run again by the prover
   (holographic_lincheck.tcc:452-462).  The values only depend on the column index, so the specialisation below
   tabulates them once per domain: the elements of H from the domain itself, D Z_H over H on the device (|H| x^(|H|-1) is a
   geometric sequence on a coset; a constant on a subspace) followed by one element-wise inversion, and then fills the
   row / col / val / row*col vectors with table look-ups and one multiplication each.  Same values, same order. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/r1cs_rs_iop/fractal_indexer.tcc)

namespace libiop {
namespace b200_glue {
/* 1 / (D Z_H)(x) for every x of H, in domain order */
template<typename FieldT>
std::vector<FieldT> inverse_vanishing_derivative_over(const field_subset<FieldT> &H)
{
    using namespace libiop_b200;
    const std::size_t n = H.num_elements();
    const vanishing_polynomial<FieldT> Z(H);
    if (H.type() == affine_subspace_type || n < residency::min_elems)      /* constant (additive) or tiny: host */
    {
        std::vector<FieldT> d;
        d.reserve(n);
        const std::vector<FieldT> elems = H.all_elements();
        for (std::size_t i = 0; i < n; ++i) d.emplace_back(Z.formal_derivative_at_point(elems[i]));
        return batch_inverse<FieldT>(d);
    }
    device_buffer d(n);
    domain_element_powers<FieldT>(d, H, n - 1, FieldT(n));                  /* |H| * x^(|H|-1), vanishing_polynomial.tcc:57-62 */
    vec_inverse<FieldT>(d, d, n);
    return d.template download<FieldT>(n);
}
}

#define B200_MATRIX_INDEXER_SPECIALISE(FIELD)                                                                              \
template<> inline std::vector<std::vector<FIELD>> matrix_indexer<FIELD>::compute_oracles_over_K()                          \
{                                                                                                                          \
    this->bivariate_lagrange_poly_ = bivariate_lagrange_polynomial<FIELD>(this->matrix_domain_);                           \
    const std::size_t K = this->index_domain_.num_elements();                                                              \
    const std::vector<FIELD> H = this->matrix_domain_.all_elements();                         /* H[i] = i-th element of H */ \
    const std::vector<FIELD> inv_dZ = b200_glue::inverse_vanishing_derivative_over<FIELD>(this->matrix_domain_);           \
    /* one entry per non-zero M[i][j]: (H[i], H[j'], M[i][j] / DZ_H(H[j']), H[i] * H[j']) with j' the position of variable j in H */ \
    std::vector<FIELD> at_row, at_col, scaled_val, product;                                                                \
    at_row.reserve(K); at_col.reserve(K); scaled_val.reserve(K); product.reserve(K);                                       \
    for (std::size_t i = 0; i < this->matrix_->num_rows(); ++i) {                                                          \
        const linear_combination<FIELD> entries = this->matrix_->get_row(i);                                               \
        for (const auto &entry : entries.terms) {                                                                          \
            const std::size_t j = this->matrix_domain_.reindex_by_subset(this->input_variable_dim_, entry.index_);         \
            at_row.push_back(H[i]);                                                                                        \
            at_col.push_back(H[j]);                                                                                        \
            product.push_back(H[i] * H[j]);                                                                                \
            scaled_val.push_back(entry.coeff_ * inv_dZ[j]);                                                                \
        }                                                                                                                  \
    }                                                                                                                      \
    /* padding up to |K| (fractal_indexer.tcc:104-121): row / col at H[0] with value 0; the product column is padded with the  \
       square of the first element of K; and the index is for the transpose, so row and col trade places */               \
    const FIELD k0 = this->index_domain_.element_by_index(0);                                                              \
    at_row.resize(K, H[0]);                                                                                                \
    at_col.resize(K, H[0]);                                                                                                \
    scaled_val.resize(K, FIELD::zero());                                                                                   \
    product.resize(K, k0.squared());                                                                                       \
    return {std::move(at_col), std::move(at_row), std::move(scaled_val), std::move(product)};                              \
}
B200_FOR_EACH_FIELD(B200_MATRIX_INDEXER_SPECIALISE)
#undef B200_MATRIX_INDEXER_SPECIALISE
} // namespace libiop
