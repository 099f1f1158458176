/* Shadow of libiop/algebra/polynomials/polynomial.tcc for the drop-in build.
   polynomial::random_polynomial (polynomial.tcc:230-243) draws its coefficients one FieldT::random_element() at a time.  For
   GF(2^192) that is randombytes_buf(&value_, 24) per coefficient (libff gf192.cpp:295-298) — 24 consecutive bytes of the
   generator's stream each — so ONE call for 24 * n bytes fills the coefficient vector with exactly the same bytes (the
   seeded parity stream and the ChaCha20 pool are both continuous streams), without 2^21 calls per masking polynomial of a
   zero-knowledge proof.  Prime fields keep the reference's loop (random_element is not a byte copy there). */
#include <sodium/randombytes.h>
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_device.hpp"
#include B200_REF_FILE(libiop/algebra/polynomials/polynomial.tcc)

namespace libiop {
template<> inline polynomial<libff::gf192> polynomial<libff::gf192>::random_polynomial(const size_t degree_bound)
{
    static_assert(sizeof(libff::gf192) == 24, "GF(2^192) elements are 3 x u64");
    std::vector<libff::gf192> random_coefficients = libiop_b200::make_uninitialized_vector<libff::gf192>(degree_bound);
    if (degree_bound) randombytes_buf(random_coefficients.data(), degree_bound * sizeof(libff::gf192));
    return polynomial<libff::gf192>(std::move(random_coefficients));
}
} // namespace libiop
