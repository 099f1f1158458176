/* Shadow of libiop/protocols/encoded/lincheck/common.tcc for the drop-in build.
   compute_p_alpha_M (common.tcc:3-36) is the holographic lincheck's transposed sparse product
       p_alpha_M[pos(c)] = sum_m r_m sum_{i: M_m[i][c] != 0} M_m[i][c] * p_alpha[i]
   followed by an interpolation over the summation domain — per entry one by-value get_row() and two multiplications on the
   host (44 ms of a Fractal proof at N = 2^18).  The overloads below (preferred over the reference's template for the served
   fields) run it on the device when the matrices are views of a registered constraint system: one gather product per matrix
   over its TRANSPOSE (b200_constraint_matrices.hpp; prime fields have no atomic addition to scatter with), the random linear
   combination, the placement pos(c) = reindex_by_subset as an index table, and the IFFT; p_alpha, produced on the device by
   lagrange_polynomial::evaluations_over_field_subset, is used where it lies and the coefficients stay there.  Other matrices
   get the reference's loop over a host copy of p_alpha fetched with one transfer. */
#include "b200_dropin_common.hpp"
#include "b200_matrix_views.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/lincheck/common.tcc)

namespace libiop {
namespace b200_glue {
template<typename FieldT>
bool p_alpha_M_on_device(const std::size_t input_variable_dim, const field_subset<FieldT> &summation_domain,
                         const std::vector<FieldT> &p_alpha_over_H, const std::vector<FieldT> &r_Mz,
                         const std::vector<std::shared_ptr<sparse_matrix<FieldT>>> &matrices, std::vector<FieldT> &coefficients_out)
{
    using namespace libiop_b200;
    const std::size_t n_S = summation_domain.num_elements();
    if (matrices.empty() || r_Mz.size() != matrices.size() || p_alpha_over_H.size() != n_S || n_S >= (std::size_t(1) << 32)) return false;
    std::vector<const flat_matrix *> T;
    for (auto &M : matrices) {
        if (M->num_rows() != n_S || M->num_columns() > n_S) return false;
        const flat_matrix *t = device_matrix_of<FieldT>(*M, true);
        if (!t) return false;
        T.push_back(t);
    }
    const std::size_t n_cols = T[0]->n_rows;
    const device_view d_p_alpha(p_alpha_over_H);
    std::vector<device_buffer> products;
    std::vector<const device_buffer *> product_ptrs;
    products.reserve(T.size());
    for (const flat_matrix *t : T) {
        if (t->n_rows != n_cols) return false;
        products.emplace_back(n_cols);
        throw_status(context(), b200_spmv_csr_dev(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), n_cols,
                                                  t->row_ptr->get(), t->col->get(), t->coeff.get(), d_p_alpha.get(), products.back().get()));
    }
    for (auto &b : products) product_ptrs.push_back(&b);
    device_buffer combined(n_cols), over_H(n_S), coefficients(n_S);
    vec_lincomb<FieldT>(combined, n_cols, r_Mz, product_ptrs);
    throw_status(context(), b200_fill_zero(context(), over_H.get(), 0, n_S));
    throw_status(context(), b200_scatter_dev(context(), over_H.get(), combined.get(), reindex_table<FieldT>(summation_domain, input_variable_dim, n_cols), n_cols));
    libff::enter_block("multi_lincheck IFFT p_alpha_M");
    IFFT_over_field_subset_device<FieldT>(over_H, summation_domain, coefficients);
    libff::leave_block("multi_lincheck IFFT p_alpha_M");
    coefficients_out = make_lazy_vector<FieldT>(std::move(coefficients), n_S);
    return true;
}
}

#define B200_P_ALPHA_M_OVERLOAD(FIELD)                                                                                     \
inline std::vector<FIELD> compute_p_alpha_M(const size_t input_variable_dim, const field_subset<FIELD> &summation_domain,  \
                                            const std::vector<FIELD> &p_alpha_over_H, const std::vector<FIELD> &r_Mz,      \
                                            const std::vector<std::shared_ptr<sparse_matrix<FIELD>>> &matrices)            \
{                                                                                                                          \
    std::vector<FIELD> coefficients;                                                                                       \
    if (b200_glue::p_alpha_M_on_device<FIELD>(input_variable_dim, summation_domain, p_alpha_over_H, r_Mz, matrices, coefficients)) \
        return coefficients;                                                                                               \
    const libiop_b200::host_readable<FIELD> p_alpha(p_alpha_over_H);                                                       \
    return compute_p_alpha_M<FIELD>(input_variable_dim, summation_domain, p_alpha.get(), r_Mz, matrices);                  \
}
B200_P_ALPHA_M_OVERLOAD(libff::gf192)
B200_P_ALPHA_M_OVERLOAD(libff::edwards_Fr)
#undef B200_P_ALPHA_M_OVERLOAD
} // namespace libiop
