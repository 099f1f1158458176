/* Shadow of libiop/algebra/polynomials/lagrange_polynomial.tcc for the drop-in build.
   lagrange_polynomial::evaluations_over_field_subset (lagrange_polynomial.tcc:67-135): (Z_S(x) - Z_S(y)) / (x - y) for
   every y of the evaluation domain (p_alpha' of the holographic lincheck over the codeword domain, holographic_lincheck_aux.tcc:41-43).
   General position (x outside the evaluation domain, evaluation domain different from S) runs on the device: x - y and
   Z_S(y) are generated from O(dim) host values, one element-wise inversion, two fused passes.  The two special cases
   (x inside the domain / the domain is S itself: verifier-side and tiny) evaluate point by point on the host. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/algebra/polynomials/lagrange_polynomial.tcc)

namespace libiop {
#define B200_LAGRANGE_SPECIALISE(FIELD)                                                                                    \
template<> inline std::vector<FIELD> lagrange_polynomial<FIELD>::evaluations_over_field_subset(                            \
    const field_subset<FIELD> &evaldomain) const                                                                           \
{                                                                                                                          \
    const std::size_t n = evaldomain.num_elements();                                                                       \
    if (n < 16384 || evaldomain == this->S_ || evaldomain.element_in_subset(this->x_)) {                                   \
        std::vector<FIELD> result;                                                                                         \
        result.reserve(n);                                                                                                 \
        for (std::size_t i = 0; i < n; ++i) result.emplace_back(this->evaluation_at_point(evaldomain.element_by_index(i))); \
        return result;                                                                                                     \
    }                                                                                                                      \
    using namespace libiop_b200;                                                                                           \
    device_buffer d_den(n), d_Z(n), d_out(n);                                                                              \
    domain_elements_affine<FIELD>(d_den, evaldomain, true, this->x_);            /* x - y */                               \
    vec_inverse<FIELD>(d_den, d_den, n);                                                                                   \
    vanishing_polynomial_over_domain<FIELD>(d_Z, this->Z_S_, evaldomain);        /* Z_S(y) */                              \
    const FIELD minus_norm = -this->normalization_coefficient_;                                                            \
    const FIELD scaled_Zx = this->Z_S_at_x_ * this->normalization_coefficient_;                                            \
    vec_affine<FIELD>(d_Z, d_Z, &minus_norm, &scaled_Zx, n);                     /* (Z_S(x) - Z_S(y)) * norm */            \
    vec_mul_add<FIELD>(d_out, d_Z, &d_den, nullptr, nullptr, false, n);                                                    \
    return download_and_keep<FIELD>(std::move(d_out), n);                                                                  \
}
B200_FOR_EACH_FIELD(B200_LAGRANGE_SPECIALISE)
#undef B200_LAGRANGE_SPECIALISE
} // namespace libiop
