/* Shadow of libiop/protocols/encoded/lincheck/basic_lincheck_aux.tcc for the drop-in build.
   multi_lincheck_virtual_oracle::evaluated_contents (basic_lincheck_aux.tcc:102-144):
       result = (sum_m r_m * f_Mz_m) * p_alpha_prime|L  -  f_z * p_alpha_ABC|L.
   On the B200 the two LDEs stay on the device, the random linear combination and the final expression are two passes,
   and only the constituent oracles go up / the result comes down.
   multi_lincheck_virtual_oracle::evaluation_at_point (:147-181) evaluates both degree-|H| polynomials by Horner for every
   query position (2 x 2^20 multiplications per position at N = 2^20: 2 s of bcs_prover::get_transcript); here the
   coefficient vectors are cached on the device and evaluated by b200_poly_eval. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/lincheck/basic_lincheck_aux.tcc)

namespace libiop {
namespace b200_glue {
template<typename FieldT> libiop_b200::device_vector_cache<FieldT> &polynomial_cache()
{
    static libiop_b200::device_vector_cache<FieldT> cache(4);
    return cache;
}
template<typename FieldT> FieldT evaluate_polynomial(const std::vector<FieldT> &coefficients, const FieldT &x)
{
    if (coefficients.size() < 4096) {           /* not worth a launch: same Horner recurrence as polynomial.tcc:98-109 */
        FieldT r = FieldT::zero();
        for (auto it = coefficients.rbegin(); it != coefficients.rend(); ++it) { r *= x; r += *it; }
        return r;
    }
    return libiop_b200::poly_eval<FieldT>(polynomial_cache<FieldT>().get(coefficients), coefficients.size(), x);
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
    {                                                                                                                      \
        device_view c1 = b200_glue::to_device(this->p_alpha_prime_.coefficients());                                            \
        FFT_over_field_subset_device<FIELD>(c1.buf(), this->p_alpha_prime_.coefficients().size(), this->codeword_domain_, d_p1); \
        device_view c2 = b200_glue::to_device(this->p_alpha_ABC_.coefficients());                                                \
        FFT_over_field_subset_device<FIELD>(c2.buf(), this->p_alpha_ABC_.coefficients().size(), this->codeword_domain_, d_p2);   \
        throw_status(context(), b200_sync(context()));                                                                     \
    }                                                                                                                      \
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
    libff::leave_block("multi_lincheck evaluated contents");                                                               \
    return result;                                                                                                         \
}                                                                                                                          \
template<> inline FIELD multi_lincheck_virtual_oracle<FIELD>::evaluation_at_point(                                         \
    const std::size_t evaluation_position, const FIELD evaluation_point,                                                   \
    const std::vector<FIELD> &constituent_oracle_evaluations) const                                                        \
{                                                                                                                          \
    libff::UNUSED(evaluation_position);                                                                                    \
    if (constituent_oracle_evaluations.size() != this->matrices_.size() + 1)                                               \
        throw std::invalid_argument("multi_lincheck uses more constituent oracles than what was provided.");               \
    if (this->use_lagrange_) throw std::logic_error("B200 drop-in: the Lagrange variant of multi_lincheck is not wired"); \
    const FIELD p_alpha_prime_X = b200_glue::evaluate_polynomial<FIELD>(this->p_alpha_prime_.coefficients(), evaluation_point); \
    const FIELD p_alpha_ABC_X = b200_glue::evaluate_polynomial<FIELD>(this->p_alpha_ABC_.coefficients(), evaluation_point); \
    FIELD f_combined_Mz_x = FIELD::zero();                                                                                 \
    for (std::size_t i = 0; i < this->r_Mz_.size(); i++) f_combined_Mz_x += this->r_Mz_[i] * constituent_oracle_evaluations[i + 1]; \
    return f_combined_Mz_x * p_alpha_prime_X - constituent_oracle_evaluations[0] * p_alpha_ABC_X;                          \
}
B200_FOR_EACH_FIELD(B200_LINCHECK_SPECIALISE)
#undef B200_LINCHECK_SPECIALISE
} // namespace libiop
