/* Shadow of libiop/protocols/encoded/common/random_linear_combination.tcc for the drop-in build.
   random_linear_combination_oracle::evaluated_contents (random_linear_combination.tcc:25-57): sum_i r_i * f_i over the
   codeword domain — one b200_vec_lincomb pass instead of one host sweep per constituent oracle. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/common/random_linear_combination.tcc)

namespace libiop {
#define B200_RLC_SPECIALISE(FIELD)                                                                                         \
template<> inline std::shared_ptr<std::vector<FIELD>> random_linear_combination_oracle<FIELD>::evaluated_contents(         \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != this->num_oracles_)                                                       \
        throw std::invalid_argument("Random Linear Combination Oracle: Expected same number of evaluations as in registration."); \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = constituent_oracle_evaluations[0]->size();                                                       \
    std::vector<device_view> d_in;                                                                                         \
    d_in.reserve(this->num_oracles_);                                                                                      \
    for (auto &v : constituent_oracle_evaluations) {                                                                       \
        if (v->size() != n) throw std::invalid_argument("Vectors of mismatched size.");                                    \
        d_in.emplace_back(b200_glue::to_device(*v));                                                                       \
    }                                                                                                                      \
    std::vector<const device_buffer *> ptrs;                                                                               \
    for (auto &b : d_in) ptrs.push_back(&b.buf());                                                                               \
    device_buffer d_out(n);                                                                                                \
    vec_lincomb<FIELD>(d_out, n, this->random_coefficients_, ptrs);                                                        \
    return b200_glue::to_host<FIELD>(d_out, n);                                                                            \
}
B200_FOR_EACH_FIELD(B200_RLC_SPECIALISE)
#undef B200_RLC_SPECIALISE
} // namespace libiop
