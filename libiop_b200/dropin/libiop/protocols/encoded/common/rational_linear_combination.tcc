/* Shadow of libiop/protocols/encoded/common/rational_linear_combination.tcc for the drop-in build (Fractal).
     combined_denominator::evaluated_contents (rational_linear_combination.tcc:12-37)      prod_i D_i
     combined_numerator::evaluated_contents   (:80-110)                                    sum_i c_i N_i prod_{k != i} D_k
     rational_linear_combination::evaluated_contents (:200-224)                            numerator / denominator
   as device passes (products, one random linear combination, one element-wise inversion). */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/common/rational_linear_combination.tcc)

namespace libiop {
namespace b200_glue {
/* out = sum_i coeff_i * N_i * prod_{k != i} D_k */
template<typename FieldT>
void combined_numerator_device(libiop_b200::device_buffer &out, std::size_t n, const std::vector<FieldT> &coeffs,
                               const std::vector<const libiop_b200::device_buffer *> &N, const std::vector<const libiop_b200::device_buffer *> &D)
{
    using namespace libiop_b200;
    const std::size_t r = N.size();
    std::vector<device_buffer> terms;
    terms.reserve(r);
    std::vector<const device_buffer *> term_ptrs;
    for (std::size_t i = 0; i < r; ++i) {
        terms.emplace_back(n);
        bool first = true;
        for (std::size_t k = 0; k < r; ++k) {
            if (k == i) continue;
            vec_mul_add<FieldT>(terms[i], first ? *N[i] : terms[i], D[k], nullptr, nullptr, false, n);
            first = false;
        }
        term_ptrs.push_back(first ? N[i] : &terms[i]);      /* a single rational: the term is N_0 itself */
    }
    vec_lincomb<FieldT>(out, n, coeffs, term_ptrs);
}
}

#define B200_RATIONAL_LC_SPECIALISE(FIELD)                                                                                 \
template<> inline std::shared_ptr<std::vector<FIELD>> combined_denominator<FIELD>::evaluated_contents(                     \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != this->num_rationals_)                                                     \
        throw std::invalid_argument("Expected same number of evaluations as in registration.");                            \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = constituent_oracle_evaluations[0]->size();                                                       \
    std::vector<device_view> d_in;                                                                                         \
    d_in.reserve(this->num_rationals_);                                                                                    \
    for (auto &v : constituent_oracle_evaluations) {                                                                       \
        if (v->size() != n) throw std::invalid_argument("Vectors of mismatched size.");                                    \
        d_in.emplace_back(b200_glue::to_device(*v));                                                                       \
    }                                                                                                                      \
    device_buffer d_out(n);                                                                                                \
    vec_mul_add<FIELD>(d_out, d_in[0].buf(), d_in.size() > 1 ? &d_in[1].buf() : nullptr, nullptr, nullptr, false, n);      \
    for (std::size_t i = 2; i < d_in.size(); ++i) vec_mul_add<FIELD>(d_out, d_out, &d_in[i].buf(), nullptr, nullptr, false, n); \
    return b200_glue::to_host<FIELD>(d_out, n);                                                                            \
}                                                                                                                          \
template<> inline std::shared_ptr<std::vector<FIELD>> combined_numerator<FIELD>::evaluated_contents(                       \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != 2 * this->num_rationals_)                                                 \
        throw std::invalid_argument("Expected same number of evaluations as in registration.");                            \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = constituent_oracle_evaluations[0]->size();                                                       \
    std::vector<device_view> d_in;                                                                                         \
    d_in.reserve(2 * this->num_rationals_);                                                                                \
    for (auto &v : constituent_oracle_evaluations) d_in.emplace_back(b200_glue::to_device(*v));                            \
    std::vector<const device_buffer *> N, D;                                                                               \
    for (std::size_t i = 0; i < this->num_rationals_; ++i) { N.push_back(&d_in[i].buf()); D.push_back(&d_in[this->num_rationals_ + i].buf()); } \
    device_buffer d_out(n);                                                                                                \
    b200_glue::combined_numerator_device<FIELD>(d_out, n, this->coefficients_, N, D);                                      \
    return b200_glue::to_host<FIELD>(d_out, n);                                                                            \
}                                                                                                                          \
template<> inline std::vector<FIELD> rational_linear_combination<FIELD>::evaluated_contents(                               \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &numerator_evals,                                               \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &denominator_evals) const                                       \
{                                                                                                                          \
    std::vector<std::shared_ptr<std::vector<FIELD>>> all_evals(numerator_evals.begin(), numerator_evals.begin() + this->num_rationals_); \
    all_evals.insert(all_evals.end(), denominator_evals.begin(), denominator_evals.begin() + this->num_rationals_);        \
    /* the two virtual oracles above run on the device and leave their results resident; so does batch_inverse */          \
    const std::shared_ptr<std::vector<FIELD>> den = this->denominator_->evaluated_contents(denominator_evals);             \
    const std::vector<FIELD> den_inv = batch_inverse<FIELD>(*den, false);                                                  \
    const std::shared_ptr<std::vector<FIELD>> num = this->numerator_->evaluated_contents(all_evals);                       \
    const std::size_t n = num->size();                                                                                     \
    if (n < libiop_b200::residency::min_elems) {                                                                           \
        std::vector<FIELD> small(*num);                                                                                    \
        for (std::size_t i = 0; i < n; ++i) small[i] *= den_inv[i];                                                        \
        return small;                                                                                                      \
    }                                                                                                                      \
    using namespace libiop_b200;                                                                                           \
    const device_view d_num = b200_glue::to_device(*num);                                                                  \
    const device_view d_inv = b200_glue::to_device(den_inv);                                                               \
    device_buffer d_out(n);                                                                                                \
    vec_mul_add<FIELD>(d_out, d_num.buf(), &d_inv.buf(), nullptr, nullptr, false, n);                                      \
    return download_and_keep<FIELD>(std::move(d_out), n);                                                                  \
}
B200_FOR_EACH_FIELD(B200_RATIONAL_LC_SPECIALISE)
#undef B200_RATIONAL_LC_SPECIALISE
} // namespace libiop
