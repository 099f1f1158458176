/* Shadow of libiop/iop/iop.tcc for the drop-in build.
   iop_protocol::get_oracle_evaluation_at_point (iop.tcc:670-714) answers a query by reading one element of the oracle's
   host vector (:687).  With the B200 adapter the oracle's contents stay on the device (a lazy vector, libiop_b200/host/
   b200_residency.hpp), so for the served fields the element is read with a gather instead — and because the prover's
   query phase (bcs_prover::get_transcript, bcs_prover.tcc:139-232) asks for ~10^4 elements one at a time, the first such
   read fetches, with ONE gather per oracle, every position the registered queries will touch:
     - the query positions are obtained in ascending query order, exactly the order in which get_transcript's loop
       (bcs_prover.tcc:155-160) obtains them, so the hash chain is squeezed identically (bcs_common.tcc:538-548);
     - a query to a virtual oracle reads its constituent oracles at the same position (iop.tcc:698-702);
     - get_transcript then reads every oracle of a Merkle tree at every position queried in that tree (:181-207).
   The function's own logic — bounds check, position recording, virtual-oracle cache — is the reference's. */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_residency.hpp"
#include B200_REF_FILE(libiop/iop/iop.tcc)

#include <set>

namespace libiop {
#define B200_IOP_SPECIALISE(FIELD)                                                                                         \
template<> inline FIELD iop_protocol<FIELD>::get_oracle_evaluation_at_point(                                              \
    const oracle_handle_ptr &handle, const std::size_t evaluation_position, const bool record)                             \
{                                                                                                                          \
    if (std::dynamic_pointer_cast<oracle_handle>(handle)) {                                                                \
        const std::size_t id = handle->id();                                                                               \
        if (this->domains_[this->oracle_registrations_[id].domain().id()].num_elements() <= evaluation_position)          \
            throw std::invalid_argument("evaluation position is outside of domain");                                       \
        if (record) this->oracle_id_to_query_positions_[id].insert(evaluation_position);   /* iop.tcc:681-685 */           \
        const std::shared_ptr<std::vector<FIELD>> contents = this->oracles_[id].evaluated_contents();                      \
        libiop_b200::residency &res = libiop_b200::residency::instance();                                                  \
        if (this->num_prover_rounds_done_ == this->num_interaction_rounds_ &&                                              \
            res.needs_fetch(contents->data(), evaluation_position) && res.begin_bulk_fetch(this)) {                        \
            std::map<std::size_t, std::set<std::size_t>> wanted;          /* oracle id -> positions */                       \
            for (std::size_t q = 0; q < this->query_registrations_.size(); ++q) {                                          \
                const std::size_t pos = this->obtain_query_position(this->query_registrations_[q].query_position());       \
                std::vector<oracle_handle_ptr> open(1, this->query_registrations_[q].oracle());                            \
                while (!open.empty()) {                                                                                    \
                    const oracle_handle_ptr cur = open.back();                                                             \
                    open.pop_back();                                                                                       \
                    if (std::dynamic_pointer_cast<oracle_handle>(cur)) wanted[cur->id()].insert(pos);                      \
                    else if (std::dynamic_pointer_cast<virtual_oracle_handle>(cur))                                        \
                        for (auto &c : this->virtual_oracle_registrations_[cur->id()].constituent_oracles()) open.push_back(c); \
                }                                                                                                          \
            }                                                                                                              \
            for (std::size_t round = 0; round < this->num_interaction_rounds_; ++round) {                                  \
                for (auto &kv : this->oracles_in_round_by_domain(round)) {                                                 \
                    std::set<std::size_t> all;                                                                             \
                    for (auto &o : kv.second) { auto w = wanted.find(o.id()); if (w != wanted.end()) all.insert(w->second.begin(), w->second.end()); } \
                    if (!all.empty()) for (auto &o : kv.second) wanted[o.id()].insert(all.begin(), all.end());            \
                }                                                                                                          \
            }                                                                                                              \
            for (auto &w : wanted) {                                                                                       \
                if (w.first >= this->oracles_.size() || !this->oracles_present_[w.first]) continue;                        \
                const std::shared_ptr<std::vector<FIELD>> v = this->oracles_[w.first].evaluated_contents();                \
                res.prefetch(v->data(), std::vector<std::size_t>(w.second.begin(), w.second.end()));                       \
            }                                                                                                              \
        }                                                                                                                  \
        return libiop_b200::resident_element<FIELD>(*contents, evaluation_position);                                       \
    }                                                                                                                      \
    if (std::dynamic_pointer_cast<virtual_oracle_handle>(handle)) {                                                        \
        std::map<std::size_t, FIELD> &cache = this->virtual_oracle_evaluation_cache_[handle->id()];                        \
        const auto hit = cache.find(evaluation_position);                                                                  \
        if (hit != cache.end()) return hit->second;                                                                        \
        const virtual_oracle_registration &reg = this->virtual_oracle_registrations_[handle->id()];                        \
        std::vector<FIELD> constituent_evaluations;                                                                        \
        for (auto &constituent_handle : reg.constituent_oracles())                                                         \
            constituent_evaluations.emplace_back(this->get_oracle_evaluation_at_point(constituent_handle, evaluation_position, record)); \
        const FIELD evaluation_point = this->get_domain(reg.domain()).element_by_index(evaluation_position);               \
        const FIELD result = this->virtual_oracles_[handle->id()]->evaluation_at_point(evaluation_position, evaluation_point, \
                                                                                       constituent_evaluations);          \
        cache[evaluation_position] = result;                                                                               \
        return result;                                                                                                     \
    }                                                                                                                      \
    throw std::invalid_argument("oracle type not supported");                                                              \
}
B200_IOP_SPECIALISE(libff::gf192)
B200_IOP_SPECIALISE(libff::edwards_Fr)
#undef B200_IOP_SPECIALISE
} // namespace libiop
