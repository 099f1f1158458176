/* Shadow of libiop/protocols/encoded/lincheck/basic_lincheck_aux.tcc for the drop-in build.
   multi_lincheck_virtual_oracle::evaluated_contents (basic_lincheck_aux.tcc:102-144):
       result = (sum_m r_m * f_Mz_m) * p_alpha_prime|L  -  f_z * p_alpha_ABC|L.
   On the B200 the two LDEs stay on the device, the random linear combination and the final expression are two passes,
   and only the constituent oracles go up / the result comes down.
   multi_lincheck_virtual_oracle::evaluation_at_point (:147-181) evaluates both degree-|H| polynomials by Horner for every
   query position (2 x 2^20 multiplications per position at N = 2^20: 2 s of bcs_prover::get_transcript); here the
   coefficient vectors are cached on the device and evaluated by b200_poly_eval — and, when the point is an element of the
   codeword domain (every prover-side query), the value is simply read out of the low-degree extension that
   evaluated_contents left on the device. */
#include "b200_dropin_common.hpp"
#include "b200_matrix_views.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/lincheck/basic_lincheck_aux.tcc)

namespace libiop {
namespace b200_glue {
/* The low-degree extensions of p_alpha' and p_alpha_ABC over the codeword domain, left on the device by
   evaluated_contents.  The prover's query phase asks for both polynomials at ~100 points that all lie in the codeword
   domain (iop.tcc:670-724 passes the position along with the point): those values are already in these arrays, so a query
   costs one 24-byte gather instead of a degree-2^20 Horner evaluation.  The arrays are attached to the coefficient vector's
   own device registration (libiop_b200/host/b200_residency.hpp): they live exactly as long as that vector and cannot be
   mistaken for another polynomial's. */
template<typename FieldT>
bool evaluate_from_lde(const std::vector<FieldT> &coefficients, const field_subset<FieldT> &domain, std::size_t position, const FieldT &x,
                       FieldT &value)
{
    if (position >= domain.num_elements()) return false;
    const b200_buf lde = libiop_b200::residency::instance().attachment(coefficients.data(), "lde", domain.num_elements());
    if (!lde || domain.element_by_index(position) != x) return false;
    libiop_b200::throw_status(libiop_b200::context(), b200_gather(libiop_b200::context(), reinterpret_cast<uint64_t *>(&value), lde, &position, 1));
    return true;
}

template<typename FieldT> FieldT evaluate_polynomial(const std::vector<FieldT> &coefficients, const FieldT &x)
{
    const bool resident = libiop_b200::residency::instance().is_resident(coefficients.data());
    if (coefficients.size() < 4096 && !resident) {   /* not worth a launch: same Horner recurrence as polynomial.tcc:98-109 */
        FieldT r = FieldT::zero();
        for (auto it = coefficients.rbegin(); it != coefficients.rend(); ++it) { r *= x; r += *it; }
        return r;
    }
    const libiop_b200::device_view c(coefficients);        /* set_challenge leaves the coefficients on the device */
    return libiop_b200::poly_eval<FieldT>(c.buf(), coefficients.size(), x);
}

/* Coefficients of the polynomial with evaluations `evals` over `domain`, as a vector whose interpolation (one IFFT) is put
   off until something reads the coefficients; the evaluations stay attached to it under "evals". */
template<typename FieldT>
std::vector<FieldT> deferred_interpolation(libiop_b200::device_buffer &&evals, const field_subset<FieldT> &domain)
{
    using namespace libiop_b200;
    const std::size_t n = domain.num_elements();
    std::map<std::string, device_buffer> inputs;
    inputs.emplace("evals", std::move(evals));
    return make_deferred_vector<FieldT>(n, std::move(inputs), [domain](device_buffer &out, std::map<std::string, device_buffer> &in) {
        IFFT_over_field_subset_device<FieldT>(in.at("evals"), domain, out);
    });
}
/* p over the codeword domain: straight from its evaluations over the summation domain when they are still attached to the
   coefficient vector and the pair of domains has a fused extension, else from the coefficients */
template<typename FieldT>
void extend_p_alpha(const std::vector<FieldT> &coefficients, const field_subset<FieldT> &summation_domain,
                    const field_subset<FieldT> &codeword_domain, libiop_b200::device_buffer &out)
{
    using namespace libiop_b200;
    const std::size_t n_S = summation_domain.num_elements();
    if (coefficients.size() == n_S) {
        if (const b200_buf evals = residency::instance().attachment(coefficients.data(), "evals", n_S))
            if (LDE_over_field_subsets_device<FieldT>(device_buffer::alias(evals, n_S), summation_domain, codeword_domain, out)) return;
    }
    const device_view c(coefficients);
    FFT_over_field_subset_device<FieldT>(c.buf(), coefficients.size(), codeword_domain, out);
}

/* calls f(row index, term) for every entry of M */
template<typename FieldT, typename Fn>
void for_each_entry(const sparse_matrix<FieldT> &M, Fn f)
{
    if (const r1cs_sparse_matrix<FieldT> *rm = dynamic_cast<const r1cs_sparse_matrix<FieldT> *>(&M)) {
        const r1cs_constraint_system<FieldT> &cs = r1cs_matrix_peek<FieldT>::system(*rm);
        const r1cs_sparse_matrix_type t = r1cs_matrix_peek<FieldT>::type(*rm);
        for (std::size_t i = 0; i < cs.constraints_.size(); ++i) {
            const linear_combination<FieldT> &row = t == r1cs_sparse_matrix_A ? cs.constraints_[i].a_
                                                   : t == r1cs_sparse_matrix_B ? cs.constraints_[i].b_ : cs.constraints_[i].c_;
            for (const auto &term : row.terms) f(i, term);
        }
        return;
    }
    for (std::size_t i = 0; i < M.num_rows(); ++i) {
        const linear_combination<FieldT> row = M.get_row(i);
        for (const auto &term : row.terms) f(i, term);
    }
}
}

#define B200_LINCHECK_SPECIALISE(FIELD)                                                                                    \
template<> inline std::shared_ptr<std::vector<FIELD>> multi_lincheck_virtual_oracle<FIELD>::evaluated_contents(            \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    libff::enter_block("multi_lincheck evaluated contents");                                                               \
    if (constituent_oracle_evaluations.size() != this->matrices_.size() + 1)                                               \
        throw std::invalid_argument("multi_lincheck uses more constituent oracles than what was provided.");               \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    device_buffer d_p1(n), d_p2(n), d_comb(n), d_out(n);                                                                   \
    b200_glue::extend_p_alpha<FIELD>(this->p_alpha_prime_.coefficients(), this->summation_domain_, this->codeword_domain_, d_p1); \
    b200_glue::extend_p_alpha<FIELD>(this->p_alpha_ABC_.coefficients(), this->summation_domain_, this->codeword_domain_, d_p2); \
    device_view d_fz = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                           \
    std::vector<device_view> d_Mz;                                                                                         \
    std::vector<const device_buffer *> Mz_ptrs;                                                                            \
    d_Mz.reserve(this->matrices_.size());                                                                                  \
    for (std::size_t m = 0; m < this->matrices_.size(); ++m) {                                                             \
        if (constituent_oracle_evaluations[m + 1]->size() != n) throw std::invalid_argument("Vectors of mismatched size."); \
        d_Mz.emplace_back(b200_glue::to_device(*constituent_oracle_evaluations[m + 1]));                                   \
    }                                                                                                                      \
    for (auto &b : d_Mz) Mz_ptrs.push_back(&b.buf());                                                                            \
    vec_lincomb<FIELD>(d_comb, n, this->r_Mz_, Mz_ptrs);                                                                   \
    vec_mul_add<FIELD>(d_out, d_comb, &d_p1, &d_fz.buf(), &d_p2, true, n);                                                       \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    /* keep both extensions next to their (device-resident) coefficient vectors for the query phase */                     \
    residency::instance().attach(this->p_alpha_prime_.coefficients().data(), "lde", std::move(d_p1));                      \
    residency::instance().attach(this->p_alpha_ABC_.coefficients().data(), "lde", std::move(d_p2));                        \
    libff::leave_block("multi_lincheck evaluated contents");                                                               \
    return result;                                                                                                         \
}                                                                                                                          \
/* set_challenge (basic_lincheck_aux.tcc:30-99): alpha powers over the constraint domain, p_alpha' = those powers placed in  \
   the summation domain, p_alpha_ABC = (sum_m r_m M_m)^T applied to them, both interpolated over the summation domain.      \
   The reference spends 0.13 s of an N = 2^20 proof here (3 * 2^20 get_row() copies, two multiplications per entry); on    \
   the B200 the powers, the placement, the transposed product (a scatter-accumulate over the flattened entries) and the    \
   two IFFTs run on the device and the coefficient vectors stay there. */                                                  \
template<> inline void multi_lincheck_virtual_oracle<FIELD>::set_challenge(const FIELD &alpha, const std::vector<FIELD> r_Mz) \
{                                                                                                                          \
    if (r_Mz.size() != this->matrices_.size())                                                                             \
        throw std::invalid_argument("Not enough random linear combination coefficients were provided");                    \
    if (this->use_lagrange_) throw std::logic_error("B200 drop-in: the Lagrange variant of multi_lincheck is not wired"); \
    this->r_Mz_ = r_Mz;                                                                                                    \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n_H = this->constraint_domain_.num_elements(), n_S = this->summation_domain_.num_elements();         \
    libff::enter_block("multi_lincheck compute alpha powers");                                                             \
    device_buffer d_alpha(n_H);                                                                                            \
    powers_of<FIELD>(d_alpha, n_H, alpha);                                                                                 \
    libff::leave_block("multi_lincheck compute alpha powers");                                                             \
    libff::enter_block("multi_lincheck compute p_alpha_prime");                                                            \
    /* element i of the constraint domain sits at position i * stride of the summation domain (identity for subspaces,     \
       every (|S|/|H|)-th element for cosets: field_subset::reindex_by_subset, subgroup.tcc:150-162) */                    \
    const std::size_t stride = n_H > 1 ? this->summation_domain_.reindex_by_subset(this->constraint_domain_.dimension(), 1) : 1; \
    device_buffer d_prime(n_S), d_abc(n_S);                                                                                \
    throw_status(context(), b200_fill_zero(context(), d_prime.get(), 0, n_S));                                             \
    throw_status(context(), b200_copy_strided(context(), d_prime.get(), stride, d_alpha.get(), 0, 1, n_H));               \
    libff::leave_block("multi_lincheck compute p_alpha_prime");                                                            \
    libff::enter_block("multi_lincheck compute p_alpha_ABC");                                                              \
    throw_status(context(), b200_fill_zero(context(), d_abc.get(), 0, n_S));                                               \
    const auto summation_index = [this](std::size_t column) {                                                              \
        return this->summation_domain_.reindex_by_subset(this->variable_domain_.dimension(),                               \
                                                         this->variable_domain_.reindex_by_subset(this->input_variable_dim_, column)); \
    };                                                                                                                     \
    if (field_traits<FIELD>::kind == B200_FIELD_GF192) {                                                                   \
        for (std::size_t m = 0; m < this->matrices_.size(); ++m) {                                                         \
            /* the constraint system's matrices already on the device (shared with create_Az_Bz_Cz): column index = position \
               in the summation domain for subspaces (reindex_by_subset is the identity there, field_subset.tcc:130-142) */  \
            if (const r1cs_sparse_matrix<FIELD> *rm = dynamic_cast<const r1cs_sparse_matrix<FIELD> *>(this->matrices_[m].get())) { \
                const b200_glue::flat_matrix *M = this->summation_domain_.type() == affine_subspace_type                   \
                    ? b200_glue::constraint_matrices<FIELD>::instance().get(&b200_glue::r1cs_matrix_peek<FIELD>::system(*rm)) : nullptr; \
                if (M && this->variable_domain_.num_elements() <= n_S) {                                                   \
                    const b200_glue::flat_matrix &F = M[(int)b200_glue::r1cs_matrix_peek<FIELD>::type(*rm) - 1];           \
                    throw_status(context(), b200_gf192_scatter_mul_add_dev(context(), d_abc.get(), F.nnz, F.col->get(), F.row->get(), \
                                                                           F.coeff.get(), d_alpha.get(), limbs(this->r_Mz_[m]))); \
                    continue;                                                                                              \
                }                                                                                                          \
            }                                                                                                              \
            std::vector<uint32_t> dst, src;                                                                                \
            std::vector<FIELD> coeff;                                                                                      \
            const std::size_t guess = this->matrices_[m]->num_rows() + (this->matrices_[m]->num_rows() >> 3);              \
            dst.reserve(guess); src.reserve(guess); coeff.reserve(guess);                                                  \
            b200_glue::for_each_entry<FIELD>(*this->matrices_[m], [&](std::size_t row, const linear_term<FIELD> &term) {   \
                dst.push_back((uint32_t)summation_index(term.index_)); src.push_back((uint32_t)row); coeff.push_back(term.coeff_); \
            });                                                                                                            \
            scatter_mul_add<FIELD>(d_abc, n_S, dst, src, coeff, d_alpha, n_H, this->r_Mz_[m]);                             \
        }                                                                                                                  \
    } else {                                                                                                               \
        /* no atomic modular addition: accumulate on the host from the downloaded powers (rows walked by reference) */      \
        const std::vector<FIELD> alpha_powers = d_alpha.template download<FIELD>(n_H);                                     \
        std::vector<FIELD> p_abc(n_S, FIELD::zero());                                                                      \
        for (std::size_t m = 0; m < this->matrices_.size(); ++m) {                                                         \
            const FIELD r = this->r_Mz_[m];                                                                                \
            b200_glue::for_each_entry<FIELD>(*this->matrices_[m], [&](std::size_t row, const linear_term<FIELD> &term) {   \
                p_abc[summation_index(term.index_)] += r * term.coeff_ * alpha_powers[row];                                \
            });                                                                                                            \
        }                                                                                                                  \
        d_abc.upload(p_abc);                                                                                               \
    }                                                                                                                      \
    libff::leave_block("multi_lincheck compute p_alpha_ABC");                                                              \
    libff::enter_block("multi_lincheck IFFT p_alphas");                                                                    \
    /* the coefficient vectors are DEFERRED: the prover only ever needs both polynomials over the codeword domain, which the   \
       fused LDE produces from these evaluations (extend_p_alpha); the interpolation runs if somebody asks for coefficients */ \
    this->p_alpha_ABC_ = polynomial<FIELD>(b200_glue::deferred_interpolation<FIELD>(std::move(d_abc), this->summation_domain_)); \
    this->p_alpha_prime_ = polynomial<FIELD>(b200_glue::deferred_interpolation<FIELD>(std::move(d_prime), this->summation_domain_)); \
    libff::leave_block("multi_lincheck IFFT p_alphas");                                                                    \
}                                                                                                                          \
template<> inline FIELD multi_lincheck_virtual_oracle<FIELD>::evaluation_at_point(                                         \
    const std::size_t evaluation_position, const FIELD evaluation_point,                                                   \
    const std::vector<FIELD> &constituent_oracle_evaluations) const                                                        \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != this->matrices_.size() + 1)                                               \
        throw std::invalid_argument("multi_lincheck uses more constituent oracles than what was provided.");               \
    if (this->use_lagrange_) throw std::logic_error("B200 drop-in: the Lagrange variant of multi_lincheck is not wired"); \
    FIELD p_alpha_prime_X, p_alpha_ABC_X;                                                                                  \
    if (!b200_glue::evaluate_from_lde<FIELD>(this->p_alpha_prime_.coefficients(), this->codeword_domain_, evaluation_position,  \
                                             evaluation_point, p_alpha_prime_X))                                           \
        p_alpha_prime_X = b200_glue::evaluate_polynomial<FIELD>(this->p_alpha_prime_.coefficients(), evaluation_point);    \
    if (!b200_glue::evaluate_from_lde<FIELD>(this->p_alpha_ABC_.coefficients(), this->codeword_domain_, evaluation_position,    \
                                             evaluation_point, p_alpha_ABC_X))                                             \
        p_alpha_ABC_X = b200_glue::evaluate_polynomial<FIELD>(this->p_alpha_ABC_.coefficients(), evaluation_point);        \
    FIELD f_combined_Mz_x = FIELD::zero();                                                                                 \
    for (std::size_t i = 0; i < this->r_Mz_.size(); i++) f_combined_Mz_x += this->r_Mz_[i] * constituent_oracle_evaluations[i + 1]; \
    return f_combined_Mz_x * p_alpha_prime_X - constituent_oracle_evaluations[0] * p_alpha_ABC_X;                          \
}
B200_FOR_EACH_FIELD(B200_LINCHECK_SPECIALISE)
#undef B200_LINCHECK_SPECIALISE
} // namespace libiop
