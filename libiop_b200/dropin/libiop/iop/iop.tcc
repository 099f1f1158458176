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

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <set>

namespace libiop {
namespace b200_glue {
/* What the query phase of ONE prover object reads from its device-resident oracles: per oracle the sorted positions and
   their values, fetched with one gather each (see the file comment).  A read is then an index into `columns` and a search in
   a short sorted array; get_transcript asks for ascending positions per oracle, which the `next` hint makes O(1). */
template<typename FieldT>
struct query_columns {
    struct column { std::vector<std::size_t> positions; std::vector<FieldT> values; std::size_t next = 0; bool resident = false; };
    const void *owner = nullptr;
    std::size_t generation = ~std::size_t(0);
    std::vector<column> columns;                 /* by oracle id */

    bool lookup(std::size_t oracle_id, std::size_t position, FieldT &value)
    {
        if (oracle_id >= columns.size()) return false;
        column &c = columns[oracle_id];
        if (!c.resident) return false;
        if (c.next < c.positions.size() && c.positions[c.next] == position) { value = c.values[c.next++]; return true; }
        const auto it = std::lower_bound(c.positions.begin(), c.positions.end(), position);
        if (it == c.positions.end() || *it != position) return false;
        c.next = (std::size_t)(it - c.positions.begin());
        value = c.values[c.next++];
        return true;
    }
};
template<typename FieldT> query_columns<FieldT> &query_cache()
{
    static thread_local query_columns<FieldT> cache;
    return cache;
}
}

#define B200_IOP_SPECIALISE(FIELD)                                                                                         \
template<> inline FIELD iop_protocol<FIELD>::get_oracle_evaluation_at_point(                                              \
    const oracle_handle_ptr &handle, const std::size_t evaluation_position, const bool record)                             \
{                                                                                                                          \
    if (std::dynamic_pointer_cast<oracle_handle>(handle)) {                                                                \
        const std::size_t id = handle->id();                                                                               \
        if (this->domains_[this->oracle_registrations_[id].domain().id()].num_elements() <= evaluation_position)          \
            throw std::invalid_argument("evaluation position is outside of domain");                                       \
        if (record) this->oracle_id_to_query_positions_[id].insert(evaluation_position);   /* iop.tcc:681-685 */           \
        const std::vector<FIELD> &contents = *this->oracles_[id].evaluated_contents();                                     \
        if (!libiop_b200::lazy_region::instance().contains(contents.data()))                                               \
            return contents[evaluation_position];                  /* an ordinary host vector: the reference's read */     \
        libiop_b200::residency &res = libiop_b200::residency::instance();                                                  \
        b200_glue::query_columns<FIELD> &cache = b200_glue::query_cache<FIELD>();                                          \
        FIELD value;                                                                                                       \
        if (cache.owner == this && cache.generation == res.generation() && cache.lookup(id, evaluation_position, value))   \
            return value;                                                                                                  \
        if (this->num_prover_rounds_done_ == this->num_interaction_rounds_ &&                                              \
            !(cache.owner == this && cache.generation == res.generation())) {                                              \
            /* first read of this prover's query phase: fetch everything the registered queries will touch */              \
            const auto b200_t0 = std::chrono::steady_clock::now();                                                        \
            std::vector<std::vector<std::size_t>> wanted(this->oracles_.size());       /* oracle id -> positions */             \
            std::vector<std::vector<std::size_t>> closure(this->virtual_oracle_registrations_.size());  /* virtual id -> oracle ids */ \
            std::vector<char> closed(this->virtual_oracle_registrations_.size(), 0);                                       \
            for (std::size_t q = 0; q < this->query_registrations_.size(); ++q) {                                          \
                const std::size_t pos = this->obtain_query_position(this->query_registrations_[q].query_position());       \
                const oracle_handle_ptr top = this->query_registrations_[q].oracle();                                      \
                if (dynamic_cast<const oracle_handle *>(top.get())) { wanted[top->id()].push_back(pos); continue; }        \
                if (!dynamic_cast<const virtual_oracle_handle *>(top.get())) continue;                                     \
                std::vector<std::size_t> &leaves = closure[top->id()];                                                     \
                if (!closed[top->id()]) {                         /* the oracles a virtual oracle is built from, once */   \
                    closed[top->id()] = 1;                                                                                 \
                    std::vector<oracle_handle_ptr> open(1, top);                                                           \
                    while (!open.empty()) {                                                                                \
                        const oracle_handle_ptr cur = open.back();                                                         \
                        open.pop_back();                                                                                   \
                        if (dynamic_cast<const oracle_handle *>(cur.get())) leaves.push_back(cur->id());                   \
                        else if (dynamic_cast<const virtual_oracle_handle *>(cur.get()))                                   \
                            for (auto &c : this->virtual_oracle_registrations_[cur->id()].constituent_oracles()) open.push_back(c); \
                    }                                                                                                      \
                    std::sort(leaves.begin(), leaves.end());                                                               \
                    leaves.erase(std::unique(leaves.begin(), leaves.end()), leaves.end());                                 \
                }                                                                                                          \
                for (std::size_t o : leaves) wanted[o].push_back(pos);                                                     \
            }                                                                                                              \
            for (std::size_t round = 0; round < this->num_interaction_rounds_; ++round) {                                  \
                for (auto &kv : this->oracles_in_round_by_domain(round)) {       /* one Merkle tree: every oracle at every queried position */ \
                    std::vector<std::size_t> all;                                                                          \
                    for (auto &o : kv.second) all.insert(all.end(), wanted[o.id()].begin(), wanted[o.id()].end());         \
                    std::sort(all.begin(), all.end());                                                                     \
                    all.erase(std::unique(all.begin(), all.end()), all.end());                                             \
                    for (auto &o : kv.second) wanted[o.id()] = all;                                                        \
                }                                                                                                          \
            }                                                                                                              \
            cache.owner = this; cache.generation = res.generation();                                                       \
            cache.columns.assign(this->oracles_.size(), b200_glue::query_columns<FIELD>::column());                        \
            for (std::size_t o = 0; o < wanted.size(); ++o) {                                                              \
                if (wanted[o].empty() || !this->oracles_present_[o]) continue;                                             \
                const std::shared_ptr<std::vector<FIELD>> v = this->oracles_[o].evaluated_contents();                      \
                b200_glue::query_columns<FIELD>::column &col = cache.columns[o];                                           \
                col.positions.swap(wanted[o]);                                                                             \
                std::sort(col.positions.begin(), col.positions.end());                                                     \
                col.positions.erase(std::unique(col.positions.begin(), col.positions.end()), col.positions.end());         \
                col.values.resize(col.positions.size());                                                                   \
                col.resident = res.gather(v->data(), col.positions, reinterpret_cast<uint64_t *>(col.values.data()));      \
            }                                                                                                              \
            if (std::getenv("B200_QUERY_TIMING")) fprintf(stderr, "[b200] bulk query fetch: %.3f ms, %zu queries\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - b200_t0).count(), this->query_registrations_.size()); \
            if (cache.lookup(id, evaluation_position, value)) return value;                                                \
        }                                                                                                                  \
        return libiop_b200::resident_element<FIELD>(contents, evaluation_position);   /* a position no registered query covers */ \
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
