/* Shadow of libiop/protocols/encoded/common/boundary_constraint.tcc for the drop-in build.
   single_boundary_constraint::evaluated_contents (boundary_constraint.tcc:20-62): (f(x) - claimed_eval) / (x - eval_point)
   over the codeword domain; on the device x - eval_point is generated from O(dim) host values and inverted element-wise. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/common/boundary_constraint.tcc)

namespace libiop {
#define B200_BOUNDARY_SPECIALISE(FIELD)                                                                                    \
template<> inline std::shared_ptr<std::vector<FIELD>> single_boundary_constraint<FIELD>::evaluated_contents(               \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != 1)                                                                        \
        throw std::invalid_argument("Single Boundary Constraint: Expected exactly 1 constituent oracle.");                 \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    const device_view d_f = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                      \
    device_buffer d_den(n), d_t(n), d_out(n);                                                                              \
    domain_elements_affine<FIELD>(d_den, this->codeword_domain_, false, -this->eval_point_);     /* x - eval_point */      \
    vec_inverse<FIELD>(d_den, d_den, n);                                                                                   \
    const FIELD minus_eval = -this->oracle_evaluation_;                                                                    \
    vec_affine<FIELD>(d_t, d_f.buf(), nullptr, &minus_eval, n);                                                            \
    vec_mul_add<FIELD>(d_out, d_t, &d_den, nullptr, nullptr, false, n);                                                    \
    return b200_glue::to_host<FIELD>(d_out, n);                                                                            \
}
B200_FOR_EACH_FIELD(B200_BOUNDARY_SPECIALISE)
#undef B200_BOUNDARY_SPECIALISE
} // namespace libiop
