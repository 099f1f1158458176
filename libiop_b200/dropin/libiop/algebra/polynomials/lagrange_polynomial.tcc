/* Shadow of libiop/algebra/polynomials/lagrange_polynomial.tcc for the drop-in build.
   lagrange_polynomial::evaluations_over_field_subset (lagrange_polynomial.tcc:67-135): (Z_S(x) - Z_S(y)) / (x - y) for
   every y of the evaluation domain (p_alpha' of the holographic lincheck over the codeword domain, holographic_lincheck_aux.tcc:41-43).
   On the device: x - y and Z_S(y) are generated from O(dim) host values, one element-wise inversion, one or two fused
   passes.  When the evaluation domain is S itself Z_S(y) vanishes; when x lies in the evaluation domain the position
   y = x is located on the device and patched with the reference's own special-case value. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/algebra/polynomials/lagrange_polynomial.tcc)

namespace libiop {
#define B200_LAGRANGE_SPECIALISE(FIELD)                                                                                    \
template<> inline std::vector<FIELD> lagrange_polynomial<FIELD>::evaluations_over_field_subset(                            \
    const field_subset<FIELD> &evaldomain) const                                                                           \
{                                                                                                                          \
    const std::size_t n = evaldomain.num_elements();                                                                       \
    if (n < 16384) {                                                                                                       \
        std::vector<FIELD> result;                                                                                         \
        result.reserve(n);                                                                                                 \
        for (std::size_t i = 0; i < n; ++i) result.emplace_back(this->evaluation_at_point(evaldomain.element_by_index(i))); \
        return result;                                                                                                     \
    }                                                                                                                      \
    using namespace libiop_b200;                                                                                           \
    device_buffer d_den(n), d_out(n);                                                                                      \
    domain_elements_affine<FIELD>(d_den, evaldomain, true, this->x_);            /* x - y */                               \
    std::size_t x_position = n;                                                  /* where y = x, if x lies in the domain */ \
    if (evaldomain.element_in_subset(this->x_))                                                                            \
        throw_status(context(), b200_vec_find_zero(context(), d_den.get(), n, &x_position));                               \
    vec_inverse<FIELD>(d_den, d_den, n);                                         /* the zero at x_position stays zero */   \
    if (evaldomain == this->S_) {                                                /* Z_S(y) = 0 on S (lagrange_polynomial.tcc:105-117) */ \
        const FIELD k = this->Z_S_at_x_ * this->normalization_coefficient_;                                                \
        vec_affine<FIELD>(d_out, d_den, &k, nullptr, n);                                                                   \
    } else {                                                                                                               \
        device_buffer d_Z(n);                                                                                              \
        vanishing_polynomial_over_domain<FIELD>(d_Z, this->Z_S_, evaldomain);    /* Z_S(y) */                              \
        const FIELD minus_norm = -this->normalization_coefficient_;                                                        \
        const FIELD scaled_Zx = this->Z_S_at_x_ * this->normalization_coefficient_;                                        \
        vec_affine<FIELD>(d_Z, d_Z, &minus_norm, &scaled_Zx, n);                 /* (Z_S(x) - Z_S(y)) * norm */            \
        vec_mul_add<FIELD>(d_out, d_Z, &d_den, nullptr, nullptr, false, n);                                                \
    }                                                                                                                      \
    if (x_position < n) {                                                        /* the formula's 0/0: the derivative (or 1), :39-52 */ \
        std::vector<FIELD> result = d_out.template download<FIELD>(n);                                                     \
        result[x_position] = this->evaluation_at_point(this->x_);                                                          \
        return result;                                                                                                     \
    }                                                                                                                      \
    return download_and_keep<FIELD>(std::move(d_out), n);                                                                  \
}
B200_FOR_EACH_FIELD(B200_LAGRANGE_SPECIALISE)
#undef B200_LAGRANGE_SPECIALISE
} // namespace libiop
