/* Shadow of libiop/protocols/encoded/lincheck/common.tcc for the drop-in build.
   compute_p_alpha_M (common.tcc:3-36, the holographic lincheck's transposed sparse product) reads p_alpha over the summation
   domain element by element; that vector is produced on the device (lagrange_polynomial::evaluations_over_field_subset) and is
   device-resident.  The overloads below — preferred over the reference's template for the served fields — hand the
   reference's loop a host copy fetched with one transfer instead of letting the first read fault the vector in. */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_residency.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/lincheck/common.tcc)

namespace libiop {
#define B200_P_ALPHA_M_OVERLOAD(FIELD)                                                                                     \
inline std::vector<FIELD> compute_p_alpha_M(const size_t input_variable_dim, const field_subset<FIELD> &summation_domain,  \
                                            const std::vector<FIELD> &p_alpha_over_H, const std::vector<FIELD> &r_Mz,      \
                                            const std::vector<std::shared_ptr<sparse_matrix<FIELD>>> &matrices)            \
{                                                                                                                          \
    const libiop_b200::host_readable<FIELD> p_alpha(p_alpha_over_H);                                                       \
    return compute_p_alpha_M<FIELD>(input_variable_dim, summation_domain, p_alpha.get(), r_Mz, matrices);                  \
}
B200_P_ALPHA_M_OVERLOAD(libff::gf192)
B200_P_ALPHA_M_OVERLOAD(libff::edwards_Fr)
#undef B200_P_ALPHA_M_OVERLOAD
} // namespace libiop
