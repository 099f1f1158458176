/* Shadow of libiop/protocols/encoded/common/rowcheck.tcc for the drop-in build.
   rowcheck_ABC_virtual_oracle::evaluated_contents (rowcheck.tcc:18-87): (Az*Bz - Cz) / Z_H over the codeword domain,
   where Z_H takes only |L|/|H| distinct values (one per coset of H in L; additive: consecutive blocks of |H|
   positions, multiplicative: position mod |L|/|H|).  The few inverses are computed on the host exactly like the
   reference does; the two passes over L run on the device. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/common/rowcheck.tcc)

namespace libiop {
#define B200_ROWCHECK_SPECIALISE(FIELD)                                                                                    \
template<> inline std::shared_ptr<std::vector<FIELD>> rowcheck_ABC_virtual_oracle<FIELD>::evaluated_contents(              \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    libff::enter_block("rowcheck evaluated contents");                                                                     \
    if (constituent_oracle_evaluations.size() != 3)                                                                        \
        throw std::invalid_argument("rowcheck_ABC has three constituent oracles.");                                        \
    using namespace libiop_b200;                                                                                           \
    const std::vector<FIELD> Z_inv = batch_inverse(this->Z_.unique_evaluations_over_field_subset(this->codeword_domain_)); \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    const std::size_t order_H = this->constraint_domain_.num_elements();                                                   \
    const std::size_t num_cosets_of_H = n / order_H;                                                                       \
    if (Z_inv.size() != num_cosets_of_H) throw std::logic_error("rowcheck: unexpected number of distinct Z_H evaluations"); \
    device_view d_A = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                          \
    device_view d_B = b200_glue::to_device(*constituent_oracle_evaluations[1]);                                          \
    device_view d_C = b200_glue::to_device(*constituent_oracle_evaluations[2]);                                          \
    device_buffer d_t(n), d_out(n);                                                                                        \
    vec_mul_add<FIELD>(d_t, d_A.buf(), &d_B.buf(), &d_C.buf(), nullptr, true, n);                                                            \
    if (this->codeword_domain_.type() == affine_subspace_type)                                                             \
        vec_mul_table<FIELD>(d_out, &d_t, Z_inv, (unsigned)libff::log2(order_H), ~std::size_t(0), n);                      \
    else                                                                                                                   \
        vec_mul_table<FIELD>(d_out, &d_t, Z_inv, 0, num_cosets_of_H - 1, n);                                               \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    libff::leave_block("rowcheck evaluated contents");                                                                     \
    return result;                                                                                                         \
}
B200_FOR_EACH_FIELD(B200_ROWCHECK_SPECIALISE)
#undef B200_ROWCHECK_SPECIALISE
} // namespace libiop
