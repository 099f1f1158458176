/* Shadow of libiop/relations/r1cs.tcc for the drop-in build.
   r1cs_constraint_system::create_Az_Bz_Cz_from_variable_assignment (r1cs.tcc:236-266): the three sparse matrix-vector
   products A z, B z, C z.
   A constraint system registered by r1cs_sparse_matrix's constructor (every system the provers hand to a protocol) has its
   matrices flattened and kept on the device (b200_constraint_matrices.hpp, shared with the lincheck's transposed products):
   each product is one b200_spmv_csr_dev call on a device-resident z and the results stay on the device for the extension.
   (The Aurora prover does not even come here without zero knowledge: r1cs_rs_iop.tcc's shadow interleaves the flattening of
   one matrix with the product and extension of the previous one.)
   Systems nobody registered: b200_spmv_csr with host arrays from three terms per row on average — flattening costs about as
   much as the reference's own pass over the terms, so below that the host loop stays; small systems always keep it. */
#include "b200_dropin_common.hpp"
#include "b200_constraint_matrices.hpp"
#include B200_REF_FILE(libiop/relations/r1cs.tcc)

namespace libiop {
namespace b200_glue {
inline std::size_t spmv_min_terms_per_row()          /* B200_SPMV_MIN_TERMS_PER_ROW overrides the measured crossover (tests: 0) */
{
    static const std::size_t v = std::getenv("B200_SPMV_MIN_TERMS_PER_ROW") ? std::strtoull(std::getenv("B200_SPMV_MIN_TERMS_PER_ROW"), nullptr, 10) : 3;
    return v;
}
template<typename FieldT, typename SelectLC>
void spmv_rows(const std::vector<r1cs_constraint<FieldT>> &constraints, SelectLC lc_of, const libiop_b200::device_buffer &d_z, std::size_t n_z,
               std::vector<FieldT> &out)
{
    using namespace libiop_b200;
    const std::size_t n = constraints.size();
    std::vector<uint64_t> row_ptr(n + 1, 0);
    std::size_t nnz = 0;
    for (std::size_t i = 0; i < n; ++i) { nnz += lc_of(constraints[i]).terms.size(); row_ptr[i + 1] = nnz; }
    std::vector<uint32_t> col(nnz);
    std::vector<FieldT> coeff = make_uninitialized_vector<FieldT>(nnz);
    std::size_t k = 0;
    for (std::size_t i = 0; i < n; ++i)
        for (auto &lt : lc_of(constraints[i]).terms) { col[k] = (uint32_t)lt.index_; coeff[k] = lt.coeff_; ++k; }
    device_buffer d_y(n);
    throw_status(context(), b200_spmv_csr(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), n, row_ptr.data(), col.data(),
                                          reinterpret_cast<const uint64_t *>(coeff.data()), d_z.get(), n_z, d_y.get()));
    const std::vector<FieldT> y = d_y.template download<FieldT>(n);
    out.insert(out.end(), y.begin(), y.end());
}
}

#define B200_ABCZ_SPECIALISE(FIELD)                                                                                        \
template<> inline void r1cs_constraint_system<FIELD>::create_Az_Bz_Cz_from_variable_assignment(                            \
    const r1cs_variable_assignment<FIELD> &variable_assignment,                                                            \
    std::vector<FIELD> &Az_out, std::vector<FIELD> &Bz_out, std::vector<FIELD> &Cz_out) const                              \
{                                                                                                                          \
    if (const b200_glue::flat_matrix *M = b200_glue::constraint_matrices<FIELD>::instance().get(this)) {                 \
        /* the matrices are (or now get) resident on the device: three gather products, results stay there for the LDE */  \
        using namespace libiop_b200;                                                                                       \
        const device_view d_z(variable_assignment);                                                                        \
        std::vector<FIELD> *outs[3] = { &Az_out, &Bz_out, &Cz_out };                                                       \
        for (int m = 0; m < 3; ++m) {                                                                                      \
            device_buffer d_y(M[m].n_rows);                                                                                \
            throw_status(context(), b200_spmv_csr_dev(context(), field_traits<FIELD>::kind, field_traits<FIELD>::fp(), M[m].n_rows,  \
                                                      M[m].row_ptr->get(), M[m].col->get(), M[m].coeff.get(), d_z.get(), d_y.get())); \
            if (outs[m]->empty()) *outs[m] = make_lazy_vector<FIELD>(std::move(d_y), M[m].n_rows);                         \
            else { const std::vector<FIELD> y = d_y.template download<FIELD>(M[m].n_rows); outs[m]->insert(outs[m]->end(), y.begin(), y.end()); } \
        }                                                                                                                  \
        return;                                                                                                            \
    }                                                                                                                      \
    std::size_t n_terms = 0;                                                                                               \
    for (auto &cons : this->constraints_) n_terms += cons.a_.terms.size() + cons.b_.terms.size() + cons.c_.terms.size();   \
    if (this->constraints_.size() < 16384 || n_terms < 3 * b200_glue::spmv_min_terms_per_row() * this->constraints_.size()) { /* r1cs.tcc:243-265 */                                 \
        for (auto &cons : this->constraints_) {                                                                            \
            FIELD a = FIELD::zero(), b = FIELD::zero(), c = FIELD::zero();                                                 \
            for (auto &lt : cons.a_.terms) a += variable_assignment[lt.index_] * lt.coeff_;                                \
            for (auto &lt : cons.b_.terms) b += variable_assignment[lt.index_] * lt.coeff_;                                \
            for (auto &lt : cons.c_.terms) c += variable_assignment[lt.index_] * lt.coeff_;                                \
            Az_out.emplace_back(a); Bz_out.emplace_back(b); Cz_out.emplace_back(c);                                        \
        }                                                                                                                  \
        return;                                                                                                            \
    }                                                                                                                      \
    using namespace libiop_b200;                                                                                           \
    device_buffer d_z(variable_assignment.size());                                                                         \
    d_z.upload(variable_assignment);                                                                                       \
    b200_glue::spmv_rows<FIELD>(this->constraints_, [](const r1cs_constraint<FIELD> &c) -> const linear_combination<FIELD> & { return c.a_; }, \
                                d_z, variable_assignment.size(), Az_out);                                                  \
    b200_glue::spmv_rows<FIELD>(this->constraints_, [](const r1cs_constraint<FIELD> &c) -> const linear_combination<FIELD> & { return c.b_; }, \
                                d_z, variable_assignment.size(), Bz_out);                                                  \
    b200_glue::spmv_rows<FIELD>(this->constraints_, [](const r1cs_constraint<FIELD> &c) -> const linear_combination<FIELD> & { return c.c_; }, \
                                d_z, variable_assignment.size(), Cz_out);                                                  \
}
B200_FOR_EACH_FIELD(B200_ABCZ_SPECIALISE)
#undef B200_ABCZ_SPECIALISE

/* A_matrix / B_matrix / C_matrix (r1cs.tcc:175-236): the row-of-maps form Ligero's interleaved protocol builds three times per
   proof.  The reference copies every row's term vector before inserting its entries into the map; the specialisations read
   the terms in place and size the outer vector once.  Same maps. */
#define B200_NAIVE_MATRIX_SPECIALISE(FIELD, NAME, MEMBER)                                                                  \
template<> inline naive_sparse_matrix<FIELD> r1cs_constraint_system<FIELD>::NAME() const                                   \
{                                                                                                                          \
    naive_sparse_matrix<FIELD> matrix(this->num_constraints());                                                            \
    for (std::size_t i = 0; i < matrix.size(); ++i)                                                                        \
        for (const linear_term<FIELD> &term : this->constraints_[i].MEMBER.terms)                                          \
            matrix[i].insert(std::pair<std::size_t, FIELD>(term.index_, term.coeff_));                                     \
    return matrix;                                                                                                         \
}
#define B200_NAIVE_MATRICES(FIELD) B200_NAIVE_MATRIX_SPECIALISE(FIELD, A_matrix, a_) B200_NAIVE_MATRIX_SPECIALISE(FIELD, B_matrix, b_) \
                                   B200_NAIVE_MATRIX_SPECIALISE(FIELD, C_matrix, c_)
B200_FOR_EACH_FIELD(B200_NAIVE_MATRICES)
#undef B200_NAIVE_MATRICES
#undef B200_NAIVE_MATRIX_SPECIALISE
} // namespace libiop
