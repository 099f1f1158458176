/* Shadow of libiop/protocols/encoded/sumcheck/rational_sumcheck.tcc for the drop-in build (Fractal).
   sumcheck_constraint_oracle::evaluated_contents (rational_sumcheck.tcc:63-117), inputs (p, N, D):
     additive        q = ( D * (p + eps^-1 mu * x^(|H|-1)) - N ) / Z_H
     multiplicative  q = ( D * (x * p + |H|^-1 mu) - N ) / Z_H
   Z_H takes |L|/|H| distinct values over L (one per coset of H: additive = blocks of |H| consecutive positions,
   multiplicative = position mod |L|/|H|), so its inverses are a small host table like in rowcheck (rowcheck.tcc:36-84);
   the rest is a handful of device passes. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/sumcheck/rational_sumcheck.tcc)

namespace libiop {
#define B200_RATIONAL_SUMCHECK_SPECIALISE(FIELD)                                                                           \
template<> inline std::shared_ptr<std::vector<FIELD>> sumcheck_constraint_oracle<FIELD>::evaluated_contents(               \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != 3)                                                                        \
        throw std::invalid_argument("sumcheck_constraint_oracle has three constituent oracles");                           \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    const std::size_t order_H = this->summation_domain_.num_elements();                                                    \
    const std::vector<FIELD> Z_inv = batch_inverse(this->Z_.unique_evaluations_over_field_subset(this->codeword_domain_)); \
    if (Z_inv.size() != n / order_H) throw std::logic_error("rational sumcheck: unexpected number of distinct Z_H evaluations"); \
    const device_view d_p = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                      \
    const device_view d_N = b200_glue::to_device(*constituent_oracle_evaluations[1]);                                      \
    const device_view d_D = b200_glue::to_device(*constituent_oracle_evaluations[2]);                                      \
    device_buffer d_t(n), d_u(n), d_out(n);                                                                                \
    if (this->field_subset_type_ == affine_subspace_type) {                                                                \
        /* d_t = x^-1, d_u = mu' x^|H|  ->  d_u = p + mu' x^(|H|-1) */                                                     \
        const device_buffer &d_xinv = subspace_element_inverses<FIELD>(this->codeword_domain_);                            \
        domain_element_powers<FIELD>(d_u, this->codeword_domain_, order_H, this->eps_inv_times_claimed_sum_);              \
        vec_mul_add<FIELD>(d_u, d_u, &d_xinv, &d_p.buf(), nullptr, false, n);                                              \
        vec_mul_add<FIELD>(d_t, d_u, &d_D.buf(), &d_N.buf(), nullptr, true, n);          /* D * (...) - N */               \
        vec_mul_table<FIELD>(d_out, &d_t, Z_inv, (unsigned)libff::log2(order_H), ~std::size_t(0), n);                      \
    } else {                                                                                                               \
        /* d_t = x;  d_u = x * p + c;  d_t = D * d_u - N */                                                                \
        domain_element_powers<FIELD>(d_t, this->codeword_domain_, 1, FIELD::one());                                        \
        vec_mul_add<FIELD>(d_u, d_p.buf(), &d_t, nullptr, nullptr, false, n);                                              \
        vec_affine<FIELD>(d_u, d_u, nullptr, &this->order_H_inv_times_claimed_sum_, n);                                    \
        vec_mul_add<FIELD>(d_t, d_u, &d_D.buf(), &d_N.buf(), nullptr, true, n);                                            \
        vec_mul_table<FIELD>(d_out, &d_t, Z_inv, 0, n / order_H - 1, n);                                                   \
    }                                                                                                                      \
    return b200_glue::to_host<FIELD>(d_out, n);                                                                            \
}
B200_FOR_EACH_FIELD(B200_RATIONAL_SUMCHECK_SPECIALISE)
#undef B200_RATIONAL_SUMCHECK_SPECIALISE
} // namespace libiop
