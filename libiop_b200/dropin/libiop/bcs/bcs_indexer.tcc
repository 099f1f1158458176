/* Shadow of libiop/bcs/bcs_indexer.tcc for the drop-in build.
   bcs_indexer::signal_index_submissions_done (bcs_indexer.tcc:17-57) Merkle-hashes the indexed oracles of round 0 and then
   deep-copies each of them into indexed_oracles_ before erasing the original (:48-53): 12 x 192 MiB of host memcpy for the
   Fractal index at N = 2^18 (0.32 s of the 0.85 s indexer) — and a copy that would read every element of a
   device-resident oracle on the host.  For the served fields the copy is made device to device (clone_vector); the
   sequence of steps — round bookkeeping, one tree per domain, the order of indexed_oracles_ — is the reference's. */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_residency.hpp"
#include B200_REF_FILE(libiop/bcs/bcs_indexer.tcc)

namespace libiop {
#define B200_INDEXER_SPECIALISE(FIELD)                                                                                     \
template<> inline void bcs_indexer<FIELD, binary_hash_digest>::signal_index_submissions_done()                             \
{                                                                                                                          \
    libff::enter_block("Merkelize indexed oracles");                                                                       \
    iop_protocol<FIELD>::signal_prover_round_done();                                                                       \
    if (this->num_prover_rounds_done_ != 1) throw std::invalid_argument("Index submissions should be round 0");            \
    const domain_to_oracles_map by_domain = this->oracles_in_round_by_domain(0);                                           \
    const round_parameters<FIELD> round_params = this->get_round_parameters(0);                                            \
    for (auto &kv : by_domain) {                                                                                           \
        std::vector<std::shared_ptr<std::vector<FIELD>>> columns;                                                          \
        for (auto &h : kv.second) columns.emplace_back(this->oracles_[h.id()].evaluated_contents());                       \
        libff::enter_block("Construct Merkle tree");                                                                       \
        this->Merkle_trees_[this->MTs_processed_].construct_with_leaves_serialized_by_cosets(columns, round_params.quotient_map_size_); \
        libff::leave_block("Construct Merkle tree");                                                                       \
        ++this->MTs_processed_;                                                                                            \
        for (auto &h : kv.second) {                 /* the index keeps its own copy of every oracle */                     \
            this->indexed_oracles_.emplace_back(libiop_b200::clone_vector<FIELD>(*this->oracles_[h.id()].evaluated_contents())); \
            this->oracles_[h.id()].erase_contents();                                                                       \
        }                                                                                                                  \
    }                                                                                                                      \
    libff::leave_block("Merkelize indexed oracles");                                                                       \
}
B200_INDEXER_SPECIALISE(libff::gf192)
B200_INDEXER_SPECIALISE(libff::edwards_Fr)
#undef B200_INDEXER_SPECIALISE
} // namespace libiop
