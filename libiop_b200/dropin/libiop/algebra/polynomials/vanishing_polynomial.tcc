/* Shadow of libiop/algebra/polynomials/vanishing_polynomial.tcc for the drop-in build.
   polynomial_over_vanishing_polynomial (vanishing_polynomial.tcc:360-371): the affine-subspace branch goes through
   polynomial_over_linearized_polynomial (shadowed in linearized_polynomial.tcc); the multiplicative branch
   (Z = y^|H| - c, vanishing_polynomial.tcc:312-358) is the one-term case of the same device recurrence. */
#include "b200_dropin_common.hpp"

#define polynomial_over_vanishing_polynomial ref_cpu_polynomial_over_vanishing_polynomial
#include B200_REF_FILE(libiop/algebra/polynomials/vanishing_polynomial.tcc)
#undef polynomial_over_vanishing_polynomial

namespace libiop {
namespace b200_glue {
template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> divide_vanishing(const polynomial<FieldT> &f, const vanishing_polynomial<FieldT> &Z, std::true_type)
{
    if (Z.type() == affine_subspace_type)
        return polynomial_over_linearized_polynomial<FieldT>(f, Z.get_linearized_polynomial());
    const std::size_t deg = Z.degree();
    if (f.degree() < deg || f.coefficients().size() - deg < 4096 || (f.coefficients().size() - deg) / deg > 64)
        return ref_cpu_polynomial_over_vanishing_polynomial<FieldT>(f, Z);
    auto qr = libiop_b200::poly_divide_sparse<FieldT>(f.coefficients(), deg, FieldT::one(), { Z.constant_coefficient() }, { 0 });
    return std::make_pair(polynomial<FieldT>(std::move(qr.first)), polynomial<FieldT>(std::move(qr.second)));
}
template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> divide_vanishing(const polynomial<FieldT> &f, const vanishing_polynomial<FieldT> &Z, std::false_type)
{ return ref_cpu_polynomial_over_vanishing_polynomial<FieldT>(f, Z); }
}

template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> polynomial_over_vanishing_polynomial(const polynomial<FieldT> &f,
                                                                                       const vanishing_polynomial<FieldT> &Z)
{ return b200_glue::divide_vanishing<FieldT>(f, Z, libiop_b200::b200_enabled<FieldT>()); }
} // namespace libiop
