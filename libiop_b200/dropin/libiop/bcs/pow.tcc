/* Shadow of libiop/bcs/pow.tcc for the drop-in build: the proof-of-work grind
   pow<FieldT, binary_hash_digest>::solve_pow (pow.tcc:71-102) runs on the B200 (b200_pow_solve) for the fields
   the adapter serves; the result is re-checked with the tree's own node hasher (verify_pow, pow.tcc:104-112). */
#include "b200_ref_traits.hpp"
#include B200_REF_FILE(libiop/bcs/pow.tcc)

namespace libiop {
#define B200_POW_SPECIALISE(FIELD)                                                                                         \
template<> inline binary_hash_digest pow<FIELD, binary_hash_digest>::solve_pow(                                            \
    const two_to_one_hash_function<binary_hash_digest> &node_hasher, const binary_hash_digest &challenge) const            \
{                                                                                                                          \
    if (challenge.size() != 32) return this->solve_pow_internal(node_hasher, challenge);                                   \
    binary_hash_digest sol(32, 'X');                                                                                       \
    libiop_b200::throw_status(libiop_b200::context(),                                                                      \
        b200_pow_solve(libiop_b200::context(), reinterpret_cast<const uint8_t *>(challenge.data()), 32,                    \
                       (uint32_t)this->parameters_.pow_bitlen(), this->parameters_.pow_upperbound(),                       \
                       reinterpret_cast<uint8_t *>(&sol[0])));                                                             \
    if (!this->verify_pow(node_hasher, challenge, sol))                                                                    \
        throw std::logic_error("B200 proof of work only implements the blake2b two-to-one hash");                          \
    return sol;                                                                                                            \
}
B200_POW_SPECIALISE(libff::gf192)
B200_POW_SPECIALISE(libff::edwards_Fr)
#undef B200_POW_SPECIALISE
} // namespace libiop
