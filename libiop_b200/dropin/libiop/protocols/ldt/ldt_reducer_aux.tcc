/* Shadow of libiop/protocols/ldt/ldt_reducer_aux.tcc for the drop-in build.
   combined_LDT_virtual_oracle::evaluated_contents (ldt_reducer_aux.tcc:40-129):
       result = sum_i c_i f_i  +  sum_{submaximal i} c'_i * x^(maxdeg - deg_i) * f_i
   (the reference groups the two terms of a submaximal oracle as (c_i + c'_i x^e) f_i; field arithmetic is exact, so the
   regrouping is bit-identical).  On the B200: one random-linear-combination pass over all oracles, and per submaximal
   oracle the bump factors c'_i x^e are generated on the device (one F2-affine factor per set bit of e in the additive
   case, a geometric sequence in the multiplicative case) and accumulated with one fused pass. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/ldt/ldt_reducer_aux.tcc)

namespace libiop {
#define B200_LDT_COMBINE_SPECIALISE(FIELD)                                                                                 \
template<> inline std::shared_ptr<std::vector<FIELD>> combined_LDT_virtual_oracle<FIELD>::evaluated_contents(              \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != this->num_input_oracles_)                                                 \
        throw std::invalid_argument("Expected same number of evaluations as in registration.");                            \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = constituent_oracle_evaluations[0]->size();                                                       \
    std::vector<device_view> d_in;                                                                                         \
    d_in.reserve(this->num_input_oracles_);                                                                                \
    for (auto &v : constituent_oracle_evaluations) {                                                                       \
        if (v->size() != n) throw std::invalid_argument("Vectors of mismatched size.");                                    \
        d_in.emplace_back(b200_glue::to_device(*v));                                                                       \
    }                                                                                                                      \
    std::vector<const device_buffer *> ptrs;                                                                               \
    for (auto &b : d_in) ptrs.push_back(&b.buf());                                                                               \
    const std::vector<FIELD> direct(this->coefficients_.begin(), this->coefficients_.begin() + this->num_input_oracles_);  \
    device_buffer d_out(n), d_bump(n);                                                                                     \
    vec_lincomb<FIELD>(d_out, n, direct, ptrs);                                                                            \
    for (std::size_t i = 0; i < this->submaximal_oracle_indices_.size(); ++i) {                                            \
        const std::size_t idx = this->submaximal_oracle_indices_[i];                                                       \
        domain_element_powers<FIELD>(d_bump, this->codeword_domain_, this->max_degree_ - this->input_oracle_degrees_[idx], \
                                     this->coefficients_[this->num_input_oracles_ + i]);                                   \
        vec_mul_add<FIELD>(d_out, d_out, nullptr, &d_bump, &d_in[idx].buf(), false, n);                                          \
    }                                                                                                                      \
    return b200_glue::to_host<FIELD>(d_out, n);                                                                            \
}
B200_FOR_EACH_FIELD(B200_LDT_COMBINE_SPECIALISE)
#undef B200_LDT_COMBINE_SPECIALISE
} // namespace libiop
