/* Shadow of libiop/protocols/encoded/lincheck/holographic_lincheck_aux.tcc for the drop-in build (Fractal).
   holographic_multi_lincheck_virtual_oracle::evaluated_contents (holographic_lincheck_aux.tcc:30-67):
       (sum_m r_m f_Mz_m) * p_alpha'|L - f_z * p_alpha_M
   single_matrix_denominator::evaluated_contents (:115-140):  row_col - cq*row - rq*col + rq*cq
   Both are one or two device passes; p_alpha'|L comes from lagrange_polynomial::evaluations_over_field_subset (shadowed),
   whose result stays resident on the device. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/lincheck/holographic_lincheck_aux.tcc)

namespace libiop {
#define B200_HOLO_LINCHECK_SPECIALISE(FIELD)                                                                               \
template<> inline std::shared_ptr<std::vector<FIELD>> holographic_multi_lincheck_virtual_oracle<FIELD>::evaluated_contents( \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    libff::enter_block("multi_lincheck evaluated contents");                                                               \
    if (constituent_oracle_evaluations.size() != this->matrices_.size() + 2)                                               \
        throw std::invalid_argument("multi_lincheck uses more constituent oracles than what was provided.");               \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    const std::vector<FIELD> p_alpha_prime_over_codeword_domain = this->p_alpha_prime_.evaluations_over_field_subset(this->codeword_domain_); \
    const device_view d_p1 = b200_glue::to_device(p_alpha_prime_over_codeword_domain);                                     \
    const device_view d_fz = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                     \
    std::vector<device_view> d_Mz;                                                                                         \
    d_Mz.reserve(this->matrices_.size());                                                                                  \
    for (std::size_t m = 0; m < this->matrices_.size(); ++m) d_Mz.emplace_back(b200_glue::to_device(*constituent_oracle_evaluations[m + 1])); \
    std::vector<const device_buffer *> Mz_ptrs;                                                                            \
    for (auto &b : d_Mz) Mz_ptrs.push_back(&b.buf());                                                                      \
    const device_view d_pM = b200_glue::to_device(*constituent_oracle_evaluations[this->matrices_.size() + 1]);            \
    device_buffer d_comb(n), d_out(n);                                                                                     \
    vec_lincomb<FIELD>(d_comb, n, this->r_Mz_, Mz_ptrs);                                                                   \
    vec_mul_add<FIELD>(d_out, d_comb, &d_p1.buf(), &d_fz.buf(), &d_pM.buf(), true, n);                                     \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    libff::leave_block("multi_lincheck evaluated contents");                                                               \
    return result;                                                                                                         \
}                                                                                                                          \
template<> inline std::shared_ptr<std::vector<FIELD>> single_matrix_denominator<FIELD>::evaluated_contents(                \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != 3)                                                                        \
        throw std::invalid_argument("single_matrix_denominator was expecting row, col, row*col oracles as input");         \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = constituent_oracle_evaluations[0]->size();                                                       \
    const device_view d_row = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                    \
    const device_view d_col = b200_glue::to_device(*constituent_oracle_evaluations[1]);                                    \
    const device_view d_rc = b200_glue::to_device(*constituent_oracle_evaluations[2]);                                     \
    device_buffer d_out(n);                                                                                                \
    vec_lincomb<FIELD>(d_out, n, { -this->column_query_point_, -this->row_query_point_, FIELD::one() },                    \
                       { &d_row.buf(), &d_col.buf(), &d_rc.buf() });                                                       \
    const FIELD shift = this->row_query_point_ * this->column_query_point_;                                                \
    vec_affine<FIELD>(d_out, d_out, nullptr, &shift, n);                                                                   \
    return b200_glue::to_host<FIELD>(d_out, n);                                                                            \
}
B200_FOR_EACH_FIELD(B200_HOLO_LINCHECK_SPECIALISE)
#undef B200_HOLO_LINCHECK_SPECIALISE
} // namespace libiop
