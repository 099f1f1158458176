/* Shadow of libiop/algebra/polynomials/linearized_polynomial.tcc for the drop-in build.
   polynomial_over_linearized_polynomial (linearized_polynomial.tcc:238-291) — the division of the sumcheck polynomial by
   Z_H that yields h (sumcheck.tcc:363-366) — clears one coefficient at a time on the host (|P|-|H| steps of log|H|+1
   multiplications each, loop-carried).  b200_poly_divide_sparse runs the same recurrence block-parallel on the device.
   Small inputs keep the reference loop. */
#include "b200_dropin_common.hpp"

#define polynomial_over_linearized_polynomial ref_cpu_polynomial_over_linearized_polynomial
#include B200_REF_FILE(libiop/algebra/polynomials/linearized_polynomial.tcc)
#undef polynomial_over_linearized_polynomial

namespace libiop {
namespace b200_glue {
template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> divide_linearized(const polynomial<FieldT> &P, const linearized_polynomial<FieldT> &Z, std::true_type)
{
    const std::size_t deg_Z = Z.degree();
    /* the device recurrence advances deg_Z/2 quotient coefficients per launch: a low-degree divisor under a long
       dividend (e.g. Z_{1,v} under f_w, r1cs_rs_iop.tcc) stays on the host loop */
    if (Z.num_terms() < 2 || P.degree() < deg_Z || P.coefficients().size() - deg_Z < 4096 ||
        (P.coefficients().size() - deg_Z) / (deg_Z - deg_Z / 2) > 64)
        return ref_cpu_polynomial_over_linearized_polynomial<FieldT>(P, Z);
    std::vector<FieldT> coefs;
    std::vector<std::size_t> pows;
    std::size_t p_pow = deg_Z / 2;                       /* Z[j] multiplies y^(2^(j-1)) for j >= 1, Z[0] is the constant */
    for (std::size_t j = Z.num_terms() - 1; j--; ) { coefs.push_back(Z[j]); pows.push_back(p_pow); p_pow /= 2; }
    auto qr = libiop_b200::poly_divide_sparse<FieldT>(P.coefficients(), deg_Z, Z[Z.num_terms() - 1], coefs, pows);
    return std::make_pair(polynomial<FieldT>(std::move(qr.first)), polynomial<FieldT>(std::move(qr.second)));
}
template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> divide_linearized(const polynomial<FieldT> &P, const linearized_polynomial<FieldT> &Z, std::false_type)
{ return ref_cpu_polynomial_over_linearized_polynomial<FieldT>(P, Z); }
}

template<typename FieldT>
std::pair<polynomial<FieldT>, polynomial<FieldT>> polynomial_over_linearized_polynomial(const polynomial<FieldT> &P,
                                                                                        const linearized_polynomial<FieldT> &Z)
{ return b200_glue::divide_linearized<FieldT>(P, Z, libiop_b200::b200_enabled<FieldT>()); }
} // namespace libiop
