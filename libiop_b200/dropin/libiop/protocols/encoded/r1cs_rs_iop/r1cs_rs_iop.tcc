/* Shadow of libiop/protocols/encoded/r1cs_rs_iop/r1cs_rs_iop.tcc for the drop-in build.
   fz_virtual_oracle::evaluated_contents (r1cs_rs_iop.tcc:181-226): f_z = f_w * Z_{1,v} + f_{1,v} over the codeword
   domain.  On the B200: Z_{1,v} over L is generated on the device from dim(L)+1 host values, f_{1,v} is LDE'd on the
   device from its |1,v| coefficients, and one fused pass forms the result; only f_w goes up and f_z comes down. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/r1cs_rs_iop/r1cs_rs_iop.tcc)

namespace libiop {
/* create_fw_prime_evals (r1cs_rs_iop.tcc:423-446) reads f_{1,v} over the whole variable domain element by element; that
   vector is an FFT result, i.e. device-resident.  These overloads (preferred over the reference's template for the served
   fields) hand the reference's loop a host copy fetched with one transfer. */
#define B200_FW_PRIME_OVERLOAD(FIELD)                                                                                      \
inline std::vector<FIELD> create_fw_prime_evals(const std::vector<FIELD> &f_1v_over_variable_domain, const std::size_t primary_input_size, \
                                                const r1cs_auxiliary_input<FIELD> &auxiliary_input, const field_subset<FIELD> &variable_domain) \
{                                                                                                                          \
    const libiop_b200::host_readable<FIELD> f_1v(f_1v_over_variable_domain);                                               \
    return create_fw_prime_evals<FIELD>(f_1v.get(), primary_input_size, auxiliary_input, variable_domain);                 \
}
B200_FOR_EACH_FIELD(B200_FW_PRIME_OVERLOAD)
#undef B200_FW_PRIME_OVERLOAD

/* encoded_aurora_protocol::compute_fprime_ABCz_over_codeword_domain (r1cs_rs_iop.tcc:449-479): Az, Bz, Cz over the constraint
   domain -> f_Az, f_Bz, f_Cz over the codeword domain.  The reference interpolates each (IFFT over H), adds Z_H * R in the
   zero-knowledge case, and evaluates over L (FFT).  Without zero knowledge nothing happens between the two transforms, and
   the B200 library extends evaluations over H straight to L (b200_gf192_lde: the basis-only halves of the two transforms
   cancel) — 3 of the proof's 8 codeword-size LDEs, with no coefficient vector crossing PCIe.  The zero-knowledge case keeps
   the reference's three steps. */
#define B200_ABCZ_SPECIALISE(FIELD)                                                                                        \
template<> inline void encoded_aurora_protocol<FIELD>::compute_fprime_ABCz_over_codeword_domain(                           \
    std::vector<FIELD> &Az, std::vector<FIELD> &Bz, std::vector<FIELD> &Cz)                                                \
{                                                                                                                          \
    if (!this->params_.make_zk()) {                                                                                        \
        this->fprime_Az_over_codeword_domain_ = libiop_b200::LDE_over_field_subsets<FIELD>(Az, this->constraint_domain_, this->codeword_domain_); \
        this->fprime_Bz_over_codeword_domain_ = libiop_b200::LDE_over_field_subsets<FIELD>(Bz, this->constraint_domain_, this->codeword_domain_); \
        this->fprime_Cz_over_codeword_domain_ = libiop_b200::LDE_over_field_subsets<FIELD>(Cz, this->constraint_domain_, this->codeword_domain_); \
        return;                                                                                                            \
    }                                                                                                                      \
    polynomial<FIELD> f_Az(IFFT_over_field_subset<FIELD>(Az, this->constraint_domain_));                                   \
    polynomial<FIELD> f_Bz(IFFT_over_field_subset<FIELD>(Bz, this->constraint_domain_));                                   \
    polynomial<FIELD> f_Cz(IFFT_over_field_subset<FIELD>(Cz, this->constraint_domain_));                                   \
    const vanishing_polynomial<FIELD> constraint_vp(this->constraint_domain_);                                             \
    f_Az += constraint_vp * this->R_Az_;                                                                                   \
    f_Bz += constraint_vp * this->R_Bz_;                                                                                   \
    f_Cz += constraint_vp * this->R_Cz_;                                                                                   \
    this->fprime_Az_over_codeword_domain_ = FFT_over_field_subset<FIELD>(f_Az.coefficients(), this->codeword_domain_);     \
    this->fprime_Bz_over_codeword_domain_ = FFT_over_field_subset<FIELD>(f_Bz.coefficients(), this->codeword_domain_);     \
    this->fprime_Cz_over_codeword_domain_ = FFT_over_field_subset<FIELD>(f_Cz.coefficients(), this->codeword_domain_);     \
}
B200_FOR_EACH_FIELD(B200_ABCZ_SPECIALISE)
#undef B200_ABCZ_SPECIALISE

#define B200_FZ_SPECIALISE(FIELD)                                                                                          \
template<> inline std::shared_ptr<std::vector<FIELD>> fz_virtual_oracle<FIELD>::evaluated_contents(                        \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    libff::enter_block("fz evaluated contents");                                                                           \
    if (constituent_oracle_evaluations.size() != 1)                                                                        \
        throw std::invalid_argument("fz_virtual_oracle has one constituent oracle.");                                      \
    if (this->primary_input_.size() != this->primary_input_size_)                                                          \
        throw std::logic_error("Evaluation requested before primary_input is set.");                                       \
    const std::shared_ptr<std::vector<FIELD>> &fw = constituent_oracle_evaluations[0];                                     \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    if (fw->size() != n)                                                                                                   \
        throw std::invalid_argument("Provided fw evaluations don't match the declared codeword domain size.");             \
    using namespace libiop_b200;                                                                                           \
    device_view d_fw = b200_glue::to_device(*fw); device_buffer d_Z(n), d_f1v(n), d_out(n);                                            \
    const vanishing_polynomial<FIELD> input_vp(this->input_variable_domain_);                                              \
    vanishing_polynomial_over_domain<FIELD>(d_Z, input_vp, this->codeword_domain_);                                        \
    std::vector<FIELD> f_1v_evaluations({ FIELD::one() });                                                                 \
    f_1v_evaluations.insert(f_1v_evaluations.end(), this->primary_input_.begin(), this->primary_input_.end());             \
    const std::vector<FIELD> f_1v_coefficients = IFFT_over_field_subset<FIELD>(f_1v_evaluations, this->input_variable_domain_); \
    device_view d_c = b200_glue::to_device(f_1v_coefficients);                                                             \
    FFT_over_field_subset_device<FIELD>(d_c.buf(), f_1v_coefficients.size(), this->codeword_domain_, d_f1v);                     \
    vec_mul_add<FIELD>(d_out, d_fw.buf(), &d_Z, &d_f1v, nullptr, false, n);                                                      \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    libff::leave_block("fz evaluated contents");                                                                           \
    return result;                                                                                                         \
}
B200_FOR_EACH_FIELD(B200_FZ_SPECIALISE)
#undef B200_FZ_SPECIALISE
} // namespace libiop
