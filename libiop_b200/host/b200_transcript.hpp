/*
 * b200_transcript.hpp — canonical byte serialisation of a BCS transcript (SURVEY.md §8c "proof bytes", §8f N4).
 *
 * libiop cannot serialise a GF(2^192)/BLAKE2b transcript itself (bcs_common.tcc:361-378), so "bit-exact proof bytes"
 * is defined as this dump of bcs_transformation_transcript's fields in declaration order (bcs_common.hpp:68-88):
 * prover_messages_ (24 B little-endian limbs per element), MT_roots_, query_positions_ (u64), query_responses_,
 * MT_leaf_positions_, MT_set_membership_proofs_ (auxiliary_hashes then randomness_hashes, merkle_tree.hpp:24-27),
 * proof_of_work_ — every vector prefixed by its u64 length.  Generic over anything with those member names.
 */
#ifndef LIBIOP_B200_TRANSCRIPT_HPP_
#define LIBIOP_B200_TRANSCRIPT_HPP_

#include <cstdint>
#include <vector>

namespace libiop_b200 {

class transcript_writer {
    std::vector<uint8_t> &out_;
public:
    explicit transcript_writer(std::vector<uint8_t> &out) : out_(out) {}
    void u64(uint64_t v) { raw(&v, 8); }
    void raw(const void *p, std::size_t n) { const uint8_t *q = static_cast<const uint8_t *>(p); out_.insert(out_.end(), q, q + n); }
    template<typename Digest> void digest(const Digest &d) { u64(d.size()); raw(d.data(), d.size()); }
    template<typename FieldT> void elements(const std::vector<FieldT> &v)
    {
        static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
        u64(v.size());
        if (!v.empty()) raw(v.data(), v.size() * 24);
    }
    void positions(const std::vector<std::size_t> &v) { u64(v.size()); for (std::size_t p : v) u64(p); }
};

template<typename TranscriptT>
std::vector<uint8_t> canonical_transcript_bytes(const TranscriptT &t)
{
    std::vector<uint8_t> bytes;
    transcript_writer w(bytes);
    w.u64(t.prover_messages_.size());
    for (auto &m : t.prover_messages_) w.elements(m);
    w.u64(t.MT_roots_.size());
    for (auto &r : t.MT_roots_) w.digest(r);
    w.u64(t.query_positions_.size());
    for (auto &q : t.query_positions_) w.positions(q);
    w.u64(t.query_responses_.size());
    for (auto &per_tree : t.query_responses_) {
        w.u64(per_tree.size());
        for (auto &column : per_tree) w.elements(column);
    }
    w.u64(t.MT_leaf_positions_.size());
    for (auto &q : t.MT_leaf_positions_) w.positions(q);
    w.u64(t.MT_set_membership_proofs_.size());
    for (auto &proof : t.MT_set_membership_proofs_) {
        w.u64(proof.auxiliary_hashes.size());
        for (auto &h : proof.auxiliary_hashes) w.digest(h);
        w.u64(proof.randomness_hashes.size());
        for (auto &h : proof.randomness_hashes) w.digest(h);
    }
    w.u64(t.proof_of_work_.size());
    w.raw(t.proof_of_work_.data(), t.proof_of_work_.size());
    return bytes;
}

} // namespace libiop_b200
#endif
