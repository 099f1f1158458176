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
    const std::vector<FIELD> H_elems = this->matrix_domain_.all_elements();                                                \
    const std::vector<FIELD> inv_derivative = b200_glue::inverse_vanishing_derivative_over<FIELD>(this->matrix_domain_);   \
    std::vector<FIELD> row_evals, col_evals, val_evals, row_times_col_evals;                                               \
    row_evals.reserve(K); col_evals.reserve(K); val_evals.reserve(K); row_times_col_evals.reserve(K);                      \
    std::size_t num_nonzero_cnt = 0;                                                                                       \
    for (std::size_t i = 0; i < this->matrix_->num_rows(); i++) {                                                          \
        const linear_combination<FIELD> row = this->matrix_->get_row(i);                                                   \
        const FIELD &row_index_elem = H_elems[i];                                                                          \
        for (auto &term : row.terms) {                                                                                     \
            num_nonzero_cnt++;                                                                                             \
            const std::size_t col_index = this->matrix_domain_.reindex_by_subset(this->input_variable_dim_, term.index_);  \
            row_evals.emplace_back(row_index_elem);                                                                        \
            col_evals.emplace_back(H_elems[col_index]);                                                                    \
            row_times_col_evals.emplace_back(row_index_elem * H_elems[col_index]);                                         \
            val_evals.emplace_back(term.coeff_ * inv_derivative[col_index]);                                               \
        }                                                                                                                  \
    }                                                                                                                      \
    for (std::size_t i = num_nonzero_cnt; i < K; i++) {     /* padding entries: val = 0 at an arbitrary element of H */     \
        row_evals.emplace_back(H_elems[0]);                                                                                \
        col_evals.emplace_back(H_elems[0]);                                                                                \
        val_evals.emplace_back(FIELD::zero());                                                                             \
    }                                                                                                                      \
    row_evals.swap(col_evals);                              /* the index is for the transpose (fractal_indexer.tcc:110-111) */ \
    const FIELD k0 = this->index_domain_.element_by_index(0);                                                              \
    row_evals.resize(K, k0);                                                                                               \
    col_evals.resize(K, k0);                                                                                               \
    val_evals.resize(K, FIELD::zero());                                                                                    \
    row_times_col_evals.resize(K, k0.squared());                                                                           \
    return {row_evals, col_evals, val_evals, row_times_col_evals};                                                         \
}
B200_FOR_EACH_FIELD(B200_MATRIX_INDEXER_SPECIALISE)
#undef B200_MATRIX_INDEXER_SPECIALISE
} // namespace libiop
