/* Shadow of libiop/bcs/bcs_prover.tcc for the drop-in build (the query phase).
   bcs_prover::get_transcript (bcs_prover.tcc:138-233) learns WHERE every committed oracle is read by answering each
   registered query through the IOP layer — a std::map lookup or two, a std::set insertion and a shared_ptr cast per query,
   plus the evaluation of every virtual oracle, whose value the transcript does not contain — and then reads every oracle
   of a tree at every position queried in that tree one element at a time (a make_shared per read).  Ligero registers
   ~230 000 queries over ~10^3 oracles: 0.25 s of its 0.49 s proof; Aurora spends 7 ms of 79 ms there.
   The specialisation below produces the same transcript from the same inputs:
     1. query positions are obtained in ascending query order (obtain_query_position — the hash chain is squeezed exactly as by
        the reference's loop) and recorded per committed oracle; a query to a virtual oracle records the position for every
        committed oracle it is built from (what get_oracle_evaluation_at_point's recursion records, iop.tcc:670-714);
     2. per Merkle tree, in the reference's order: the sorted distinct positions, the sorted distinct leaf positions, and the
        values — ONE b200_gather_multi for all device-resident oracles of the tree, host reads for the others;
     3. the authentication paths from the tree (merkle_tree.tcc shadow: device-resident nodes).
   Virtual oracles are not evaluated: nothing in the transcript depends on their values. */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_residency.hpp"
#include B200_REF_FILE(libiop/bcs/bcs_prover.tcc)

#include <algorithm>

namespace libiop {
namespace b200_glue {
/* values[p][o] = oracle o at positions[p]; device-resident contents are fetched together */
template<typename FieldT>
std::vector<std::vector<FieldT>> read_oracle_columns(const std::vector<std::shared_ptr<std::vector<FieldT>>> &contents,
                                                     const std::vector<std::size_t> &positions)
{
    using namespace libiop_b200;
    const std::size_t n_o = contents.size(), n_p = positions.size();
    std::vector<std::vector<FieldT>> values(n_p, std::vector<FieldT>(n_o));
    if (!n_p) return values;
    std::vector<b200_buf> sources;
    std::vector<std::size_t> source_of;                       /* index into `contents` */
    for (std::size_t o = 0; o < n_o; ++o) {
        const std::vector<FieldT> &v = *contents[o];
        const b200_buf dev = lazy_region::instance().contains(v.data()) ? residency::instance().find(v.data(), v.size()) : 0;
        if (dev) { sources.push_back(dev); source_of.push_back(o); continue; }
        for (std::size_t p = 0; p < n_p; ++p) values[p][o] = v[positions[p]];
    }
    if (!sources.empty()) {
        std::vector<FieldT> fetched(sources.size() * n_p);
        throw_status(context(), b200_gather_multi(context(), reinterpret_cast<uint64_t *>(fetched.data()), sources.data(), sources.size(),
                                                  positions.data(), n_p));
        residency::instance().count_gather();
        for (std::size_t s = 0; s < sources.size(); ++s)
            for (std::size_t p = 0; p < n_p; ++p) values[p][source_of[s]] = fetched[s * n_p + p];
    }
    return values;
}
}

#define B200_GET_TRANSCRIPT_SPECIALISE(FIELD)                                                                              \
template<> inline bcs_transformation_transcript<FIELD, binary_hash_digest> bcs_prover<FIELD, binary_hash_digest>::get_transcript() \
{                                                                                                                          \
    bcs_transformation_transcript<FIELD, binary_hash_digest> result;                                                       \
    result.prover_messages_ = this->prover_messages_;                                                                      \
    for (auto &MT : this->Merkle_trees_) result.MT_roots_.emplace_back(MT.get_root());                                     \
    /* 1. where is every committed oracle read? */                                                                         \
    std::vector<std::vector<std::size_t>> read_at(this->oracles_.size());                                                  \
    std::vector<std::vector<std::size_t>> built_from(this->virtual_oracle_registrations_.size());                          \
    std::vector<char> known(this->virtual_oracle_registrations_.size(), 0);                                                \
    const auto record = [this, &read_at](std::size_t oracle_id, std::size_t position) {                                    \
        if (this->domains_[this->oracle_registrations_[oracle_id].domain().id()].num_elements() <= position)              \
            throw std::invalid_argument("evaluation position is outside of domain");                                       \
        read_at[oracle_id].push_back(position);                                                                            \
    };                                                                                                                     \
    /* a position handle is shared by many queries (Ligero: ~10^3 oracles per position): obtained once each, in first-use    \
       order, i.e. exactly when the reference's loop obtains it first */                                                     \
    std::vector<std::vector<std::size_t>> known_position(2);                                                               \
    const std::size_t unknown = ~std::size_t(0);                                                                           \
    for (std::size_t q = 0; q < this->query_registrations_.size(); ++q) {                                                  \
        const query_position_handle handle = this->query_registrations_[q].query_position();                              \
        std::vector<std::size_t> &memo = known_position[handle.type() == random_query_type ? 0 : 1];                       \
        if (memo.size() <= handle.id()) memo.resize(handle.id() + 1, unknown);                                             \
        if (memo[handle.id()] == unknown) memo[handle.id()] = this->obtain_query_position(handle);                         \
        const std::size_t position = memo[handle.id()];                                                                    \
        const oracle_handle_ptr top = this->query_registrations_[q].oracle();                                              \
        if (dynamic_cast<const oracle_handle *>(top.get())) { record(top->id(), position); continue; }                     \
        if (!dynamic_cast<const virtual_oracle_handle *>(top.get())) throw std::invalid_argument("oracle type not supported"); \
        std::vector<std::size_t> &leaves = built_from[top->id()];                                                          \
        if (!known[top->id()]) {                                                                                           \
            known[top->id()] = 1;                                                                                          \
            std::vector<oracle_handle_ptr> open(1, top);                                                                   \
            while (!open.empty()) {                                                                                        \
                const oracle_handle_ptr cur = open.back();                                                                 \
                open.pop_back();                                                                                           \
                if (dynamic_cast<const oracle_handle *>(cur.get())) leaves.push_back(cur->id());                           \
                else if (dynamic_cast<const virtual_oracle_handle *>(cur.get()))                                           \
                    for (auto &c : this->virtual_oracle_registrations_[cur->id()].constituent_oracles()) open.push_back(c); \
                else throw std::invalid_argument("oracle type not supported");                                             \
            }                                                                                                              \
            std::sort(leaves.begin(), leaves.end());                                                                       \
            leaves.erase(std::unique(leaves.begin(), leaves.end()), leaves.end());                                         \
        }                                                                                                                  \
        for (std::size_t o : leaves) record(o, position);                                                                  \
    }                                                                                                                      \
    /* 2., 3. per Merkle tree */                                                                                           \
    result.total_depth_without_pruning = 0;                                                                                \
    std::size_t MT_idx = 0;                                                                                                \
    for (std::size_t round = 0; round < this->num_interaction_rounds_; ++round) {                                          \
        const domain_to_oracles_map mapping = this->oracles_in_round_by_domain(round);                                     \
        const round_parameters<FIELD> round_params = this->get_round_parameters(round);                                    \
        for (auto &kv : mapping) {                                                                                         \
            const std::size_t num_leaves = this->domains_[kv.first.id()].num_elements() / round_params.quotient_map_size_; \
            std::vector<std::size_t> query_positions;                                                                      \
            std::vector<std::shared_ptr<std::vector<FIELD>>> contents;                                                     \
            const std::vector<std::size_t> *previous = nullptr;                                                            \
            for (auto &oracle_h : kv.second) {                                                                             \
                const std::vector<std::size_t> &mine = read_at[oracle_h.id()];                                             \
                if (!previous || mine != *previous)          /* oracles read by the same queries bring nothing new */       \
                    query_positions.insert(query_positions.end(), mine.begin(), mine.end());                               \
                previous = &mine;                                                                                          \
                contents.emplace_back(this->oracles_[oracle_h.id()].evaluated_contents());                                 \
            }                                                                                                              \
            std::sort(query_positions.begin(), query_positions.end());                                                     \
            query_positions.erase(std::unique(query_positions.begin(), query_positions.end()), query_positions.end());     \
            std::vector<std::size_t> MT_leaf_positions;                                                                    \
            MT_leaf_positions.reserve(query_positions.size());                                                             \
            for (std::size_t pos : query_positions)                                                                        \
                MT_leaf_positions.push_back(query_position_to_merkle_tree_position(pos, num_leaves, round_params));        \
            std::sort(MT_leaf_positions.begin(), MT_leaf_positions.end());                                                 \
            MT_leaf_positions.erase(std::unique(MT_leaf_positions.begin(), MT_leaf_positions.end()), MT_leaf_positions.end()); \
            result.total_depth_without_pruning += MT_leaf_positions.size() * this->Merkle_trees_[MT_idx].depth();          \
            result.query_responses_.emplace_back(b200_glue::read_oracle_columns<FIELD>(contents, query_positions));        \
            result.MT_set_membership_proofs_.emplace_back(this->Merkle_trees_[MT_idx].get_set_membership_proof(MT_leaf_positions)); \
            result.query_positions_.emplace_back(std::move(query_positions));                                              \
            result.MT_leaf_positions_.emplace_back(std::move(MT_leaf_positions));                                          \
            ++MT_idx;                                                                                                      \
        }                                                                                                                  \
    }                                                                                                                      \
    if (this->is_preprocessing_) this->remove_index_info_from_transcript(result);                                          \
    result.proof_of_work_ = this->pow_answer_;                                                                             \
    return result;                                                                                                         \
}
B200_GET_TRANSCRIPT_SPECIALISE(libff::gf192)
B200_GET_TRANSCRIPT_SPECIALISE(libff::edwards_Fr)
#undef B200_GET_TRANSCRIPT_SPECIALISE
} // namespace libiop
