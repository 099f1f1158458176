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
#include "b200_matrix_views.hpp"
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

namespace b200_glue {
/* The four index oracles over K on the device, from the flattened entry list of a registered constraint system: per entry k
   with row i and variable j at position j' of H,  row = H[i], col = H[j'], val = M[i][j] / DZ_H(H[j']), row*col — i.e. two
   table gathers, one gather + product and one product over the nnz entries, constants in the padding up to |K|
   (fractal_indexer.tcc:52-123).  The host loop this replaces copies every row by value and takes 16 ms per matrix at
   N = 2^18, three times in the indexer and three times again in the prover (holographic_lincheck.tcc:451-468).
   Order of the result as the reference's: { col, row, val, row*col } of the transposed index.  false = not available. */
template<typename FieldT>
bool index_oracles_on_device(const sparse_matrix<FieldT> &matrix, const field_subset<FieldT> &matrix_domain,
                             const field_subset<FieldT> &index_domain, std::size_t input_variable_dim,
                             std::vector<std::vector<FieldT>> &result)
{
    using namespace libiop_b200;
    const std::size_t K = index_domain.num_elements(), n_H = matrix_domain.num_elements();
    if (K < residency::min_elems || n_H >= (std::size_t(1) << 32) || matrix.num_rows() > n_H || matrix.num_columns() > n_H) return false;
    const flat_matrix *M = device_matrix_of<FieldT>(matrix, false);
    if (!M || M->nnz > K) return false;
    const std::size_t nnz = M->nnz;
    /* tables over H: the elements, and 1 / DZ_H (|H| x^(|H|-1) on a coset; the constant DZ_H(0)... on a subspace) */
    device_buffer H(n_H), inv_dZ(n_H);
    domain_elements_affine<FieldT>(H, matrix_domain, false, FieldT::zero());
    if (matrix_domain.type() == affine_subspace_type) {
        const vanishing_polynomial<FieldT> Z(matrix_domain);
        const FieldT c = Z.formal_derivative_at_point(matrix_domain.element_by_index(0)).inverse();   /* constant on a subspace */
        throw_status(context(), b200_fill(context(), inv_dZ.get(), 0, limbs(c), n_H));
    } else {
        domain_element_powers<FieldT>(inv_dZ, matrix_domain, n_H - 1, FieldT(n_H));
        vec_inverse<FieldT>(inv_dZ, inv_dZ, n_H);
    }
    const b200_bytes position = reindex_table<FieldT>(matrix_domain, input_variable_dim, matrix.num_columns());
    device_buffer at_row(K), at_col(K), val(K), product(K);
    throw_status(context(), b200_gather_dev(context(), at_row.get(), H.get(), M->row->get(), 0, nnz));
    throw_status(context(), b200_gather_dev(context(), at_col.get(), H.get(), M->col->get(), position, nnz));
    throw_status(context(), b200_gather_dev(context(), val.get(), inv_dZ.get(), M->col->get(), position, nnz));
    if (nnz) {
        vec_mul_add<FieldT>(val, val, &M->coeff, nullptr, nullptr, false, nnz);
        vec_mul_add<FieldT>(product, at_row, &at_col, nullptr, nullptr, false, nnz);
    }
    if (K > nnz) {
        const FieldT h0 = matrix_domain.element_by_index(0), k0sq = index_domain.element_by_index(0).squared(), zero = FieldT::zero();
        throw_status(context(), b200_fill(context(), at_row.get(), nnz, limbs(h0), K - nnz));
        throw_status(context(), b200_fill(context(), at_col.get(), nnz, limbs(h0), K - nnz));
        throw_status(context(), b200_fill(context(), val.get(), nnz, limbs(zero), K - nnz));
        throw_status(context(), b200_fill(context(), product.get(), nnz, limbs(k0sq), K - nnz));
    }
    result.clear();
    result.emplace_back(make_lazy_vector<FieldT>(std::move(at_col), K));
    result.emplace_back(make_lazy_vector<FieldT>(std::move(at_row), K));
    result.emplace_back(make_lazy_vector<FieldT>(std::move(val), K));
    result.emplace_back(make_lazy_vector<FieldT>(std::move(product), K));
    return true;
}
}

#define B200_MATRIX_INDEXER_SPECIALISE(FIELD)                                                                              \
template<> inline std::vector<std::vector<FIELD>> matrix_indexer<FIELD>::compute_oracles_over_K()                          \
{                                                                                                                          \
    this->bivariate_lagrange_poly_ = bivariate_lagrange_polynomial<FIELD>(this->matrix_domain_);                           \
    {                                                                                                                      \
        std::vector<std::vector<FIELD>> on_device;                                                                         \
        if (b200_glue::index_oracles_on_device<FIELD>(*this->matrix_, this->matrix_domain_, this->index_domain_,           \
                                                      this->input_variable_dim_, on_device)) return on_device;             \
    }                                                                                                                      \
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
