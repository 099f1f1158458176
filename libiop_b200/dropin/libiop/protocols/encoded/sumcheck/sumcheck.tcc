/* Shadow of libiop/protocols/encoded/sumcheck/sumcheck.tcc for the drop-in build.
   sumcheck_g_oracle::evaluated_contents (sumcheck.tcc:58-121):
     additive        g = f - ( eps^-1 mu * x^(|H|-1) + Z_H(x) * h )            x^(|H|-1) = x^|H| * x^-1
     multiplicative  g = ( f - |H|^-1 mu - Z_H(x) * h ) * x^-1
   On the B200: Z_H over L and mu' * x^|H| are F2-affine in the position (generated on the device from dim(L)+1 host
   values each), x^-1 is an element-wise inversion (additive) or a geometric sequence (multiplicative), and the
   expression is two or three fused passes.  The reference's zero handling (batch_inverse_and_mul with has_zeroes,
   algebra/utils.tcc:73-103: the inverse of 0 is taken as 0) is what the device inversion x^(2^192-2) gives. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/sumcheck/sumcheck.tcc)

namespace libiop {
#define B200_SUMCHECK_G_SPECIALISE(FIELD)                                                                                  \
template<> inline std::shared_ptr<std::vector<FIELD>> sumcheck_g_oracle<FIELD>::evaluated_contents(                        \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != 2)                                                                        \
        throw std::invalid_argument("sumcheck_g_oracle has two constituent oracles");                                      \
    libff::enter_block("Sumcheck: g evaluated contents");                                                                  \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    device_view d_f = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                          \
    device_view d_h = b200_glue::to_device(*constituent_oracle_evaluations[1]);                                          \
    device_buffer d_Z(n), d_t(n), d_u(n), d_out(n);                                                                        \
    vanishing_polynomial_over_domain<FIELD>(d_Z, this->Z_, this->codeword_domain_);                                        \
    if (this->field_subset_type_ == affine_subspace_type) {                                                                \
        /* d_t = elements of L, then their inverses; d_u = eps^-1 mu * x^|H| */                                            \
        domain_element_powers<FIELD>(d_t, this->codeword_domain_, 1, FIELD::one());                                        \
        throw_status(context(), b200_gf192_inv(context(), d_t.get(), d_t.get(), n));                                       \
        domain_element_powers<FIELD>(d_u, this->codeword_domain_, this->summation_domain_.num_elements(),                  \
                                     this->eps_inv_times_claimed_sum_);                                                    \
        vec_mul_add<FIELD>(d_u, d_u, &d_t, &d_Z, &d_h.buf(), false, n);          /* mu' x^(|H|-1) + Z_H h */                     \
        vec_mul_add<FIELD>(d_out, d_f.buf(), nullptr, &d_u, nullptr, true, n);   /* f - (...) */                                 \
    } else {                                                                                                               \
        /* d_t = f - Z_H h; d_u = constant |H|^-1 mu ... folded: (f - Z_H h - c) * x^-1 */                                 \
        vec_mul_add<FIELD>(d_t, d_f.buf(), nullptr, &d_Z, &d_h.buf(), true, n);                                                        \
        geometric_sequence<FIELD>(d_u, n, this->codeword_domain_.shift().inverse(), this->codeword_domain_.generator().inverse(), \
                                  FIELD::zero());                                                                          \
        /* (t - c) * xinv = t*xinv - c*xinv: second term via the same sequence scaled by c */                              \
        device_buffer d_v(n);                                                                                              \
        geometric_sequence<FIELD>(d_v, n, this->codeword_domain_.shift().inverse() * this->order_H_inv_times_claimed_sum_, \
                                  this->codeword_domain_.generator().inverse(), FIELD::zero());                            \
        vec_mul_add<FIELD>(d_out, d_t, &d_u, &d_v, nullptr, true, n);                                                      \
    }                                                                                                                      \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    libff::leave_block("Sumcheck: g evaluated contents");                                                                  \
    return result;                                                                                                         \
}
B200_FOR_EACH_FIELD(B200_SUMCHECK_G_SPECIALISE)
#undef B200_SUMCHECK_G_SPECIALISE
} // namespace libiop
