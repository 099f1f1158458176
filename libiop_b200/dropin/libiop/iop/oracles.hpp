/* Shadow of libiop/iop/oracles.hpp for the drop-in build ("this is where a device-resident handle
   must hide so vectors never round-trip").
   oracle<FieldT> keeps its contents as shared_ptr<vector<FieldT>> (oracles.hpp:21-52).  With the B200 adapter a codeword
   vector is usually a LAZY vector — address space whose contents live in a device buffer (libiop_b200/host/
   b200_residency.hpp) — and the two constructors that deep-copy their argument (oracles.hpp:29-31 and :36-38, used by the FRI
   commit loop fri_ldt.tcc:503 and by matrix_indexer::compute_oracles, fractal_indexer.tcc:134-155) would read every element
   on the host.  For the served fields the copy is made device to device and the new oracle stays device-resident; the
   reference's own header is included unmodified and everything else is its code. */
#ifndef B200_DROPIN_IOP_ORACLES_HPP_
#define B200_DROPIN_IOP_ORACLES_HPP_
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_residency.hpp"
#include B200_REF_FILE(libiop/iop/oracles.hpp)

namespace libiop {
#define B200_ORACLE_SPECIALISE(FIELD)                                                                                      \
template<> inline oracle<FIELD>::oracle(const std::vector<FIELD> &evaluated_contents) :                                   \
    evaluated_contents_(std::make_shared<std::vector<FIELD>>(libiop_b200::clone_vector<FIELD>(evaluated_contents))) {}     \
template<> inline oracle<FIELD>::oracle(const std::shared_ptr<std::vector<FIELD>> &evaluated_contents) :                  \
    evaluated_contents_(std::make_shared<std::vector<FIELD>>(libiop_b200::clone_vector<FIELD>(*evaluated_contents))) {}
B200_ORACLE_SPECIALISE(libff::gf192)
B200_ORACLE_SPECIALISE(libff::edwards_Fr)
#undef B200_ORACLE_SPECIALISE
} // namespace libiop
#endif
