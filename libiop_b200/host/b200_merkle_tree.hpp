/*
 * b200_merkle_tree.hpp — merkle_tree<FieldT, hash_digest_type> with BLAKE2b hashing on the device.
 * Mirrors libiop/bcs/merkle_tree.hpp:45-122: same constructor checks and exceptions
 * (merkle_tree.tcc:13-35,93-108), same leaf layout (merkle_tree.tcc:123-134), heap-ordered nodes
 * (merkle_tree.tcc:145,201-229), same multi-proof extraction (merkle_tree.tcc:243-336).
 * The hash functions are fixed to the reference's blake2b_leafhash / blake2b_two_to_one_hash
 * (bcs/hashing/blake2b.tcc:121-160, blake2b.cpp:28-48) — the only ones the BASELINE configs use.
 * zk salts are supplied by the caller (the reference draws them from randombytes_buf in
 * merkle_tree.tcc:37-72; a bit-exact prover passes the same bytes).
 */
#ifndef LIBIOP_B200_MERKLE_TREE_HPP_
#define LIBIOP_B200_MERKLE_TREE_HPP_

#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "b200_device.hpp"
#include "b200_fields.hpp"

namespace libiop_b200 {

typedef std::string binary_hash_digest; /* hashing.hpp:21-22 */
typedef std::string zk_salt_type;

struct merkle_tree_set_membership_proof { /* merkle_tree.hpp:24-27 */
    std::vector<binary_hash_digest> auxiliary_hashes;
    std::vector<zk_salt_type> randomness_hashes;
};

template<typename FieldT>
class merkle_tree {
    bool constructed_ = false;
    std::size_t num_leaves_, digest_len_bytes_;
    bool make_zk_;
    std::size_t num_zk_bytes_;
    std::vector<zk_salt_type> zk_leaf_randomness_elements_;
    b200_bytes nodes_ = 0;
    mutable std::vector<uint8_t> host_nodes_;   /* filled lazily for proofs */
public:
    merkle_tree(std::size_t num_leaves, std::size_t digest_len_bytes, bool make_zk, std::size_t security_parameter)
        : num_leaves_(num_leaves), digest_len_bytes_(digest_len_bytes), make_zk_(make_zk),
          num_zk_bytes_((security_parameter * 2 + 7) / 8)
    {
        if (num_leaves < 2 || (num_leaves & (num_leaves - 1)))
            throw std::invalid_argument("Merkle tree size must be a power of two, and at least 2.");
    }
    merkle_tree(const merkle_tree &) = delete;
    ~merkle_tree() { if (nodes_) b200_free_bytes(context(), nodes_); }

    void set_leaf_randomness(const std::vector<zk_salt_type> &salts) { zk_leaf_randomness_elements_ = salts; }

    /* device-resident oracles (what a device-aware BCS prover passes) */
    void construct_with_leaves_serialized_by_cosets_device(const std::vector<const device_buffer *> &leaf_contents,
                                                           std::size_t n_elems, std::size_t coset_serialization_size)
    {
        if (constructed_) throw std::logic_error("Attempting to double-construct a Merkle tree.");
        if (n_elems / coset_serialization_size != num_leaves_)
            throw std::logic_error("Attempting to construct a Merkle tree with a constituent vector of wrong size");
        std::vector<b200_buf> bufs;
        for (const device_buffer *b : leaf_contents) bufs.push_back(b->get());
        std::vector<uint8_t> salts;
        if (make_zk_) {
            if (zk_leaf_randomness_elements_.size() != num_leaves_)
                throw std::logic_error("zk Merkle tree needs one salt per leaf (set_leaf_randomness)");
            salts.reserve(num_leaves_ * num_zk_bytes_);
            for (const zk_salt_type &s : zk_leaf_randomness_elements_) salts.insert(salts.end(), s.begin(), s.end());
        }
        nodes_ = b200_alloc_bytes(context(), (2 * num_leaves_ - 1) * digest_len_bytes_);
        if (!nodes_) throw std::runtime_error(b200_last_error(context()));
        throw_status(context(), b200_merkle_build(context(), bufs.data(), bufs.size(), n_elems, coset_serialization_size,
                                                  field_traits<FieldT>::kind == B200_FIELD_GF192 ? B200_DOMAIN_AFFINE_SUBSPACE
                                                                                                : B200_DOMAIN_MULTIPLICATIVE_COSET,
                                                  make_zk_ ? salts.data() : nullptr, num_zk_bytes_, digest_len_bytes_, nodes_));
        throw_status(context(), b200_sync(context()));
        constructed_ = true;
    }

    /* merkle_tree.hpp:88-90 */
    void construct_with_leaves_serialized_by_cosets(const std::vector<std::shared_ptr<std::vector<FieldT>>> &leaf_contents,
                                                    std::size_t coset_serialization_size)
    {
        if (constructed_) throw std::logic_error("Attempting to double-construct a Merkle tree.");
        for (auto &v : leaf_contents)
            if (v->size() / coset_serialization_size != num_leaves_)
                throw std::logic_error("Attempting to construct a Merkle tree with a constituent vector of wrong size");
        std::vector<device_view> bufs;                    /* the resident copies of device-resident oracles, uploads of the others */
        std::vector<const device_buffer *> ptrs;
        bufs.reserve(leaf_contents.size());
        for (auto &v : leaf_contents) bufs.emplace_back(*v);
        for (auto &b : bufs) ptrs.push_back(&b.buf());
        construct_with_leaves_serialized_by_cosets_device(ptrs, leaf_contents[0]->size(), coset_serialization_size);
    }
    void construct(const std::vector<std::shared_ptr<std::vector<FieldT>>> &leaf_contents) /* merkle_tree.tcc:85-90 */
    {
        construct_with_leaves_serialized_by_cosets(leaf_contents, 1);
    }

    binary_hash_digest get_root() const /* merkle_tree.tcc:232-240 */
    {
        if (!constructed_) throw std::logic_error("Attempting to obtain a Merkle tree root without constructing the tree first.");
        binary_hash_digest r(digest_len_bytes_, 'X');
        throw_status(context(), b200_merkle_root(context(), nodes_, reinterpret_cast<uint8_t *>(&r[0]), digest_len_bytes_));
        return r;
    }

    binary_hash_digest node(std::size_t heap_index) const
    {
        binary_hash_digest r(digest_len_bytes_, 'X');
        throw_status(context(), b200_merkle_nodes(context(), nodes_, &heap_index, 1, digest_len_bytes_, reinterpret_cast<uint8_t *>(&r[0])));
        return r;
    }

    /* merkle_tree.tcc:243-336: multi-proof, auxiliary hashes bottom-up, left-to-right, skipping known siblings */
    merkle_tree_set_membership_proof get_set_membership_proof(const std::vector<std::size_t> &positions) const
    {
        if (!constructed_) throw std::logic_error("Attempting to obtain a Merkle tree authentication path without constructing the tree first.");
        merkle_tree_set_membership_proof result;
        if (positions.empty()) return result;
        std::vector<std::size_t> S = positions;
        std::sort(S.begin(), S.end());
        S.erase(std::unique(S.begin(), S.end()), S.end());
        if (std::any_of(S.begin(), S.end(), [this](std::size_t pos) { return pos >= num_leaves_; }))
            throw std::invalid_argument("All positions must be between 0 and num_leaves-1.");
        if (make_zk_) for (auto pos : S) result.randomness_hashes.emplace_back(zk_leaf_randomness_elements_[pos]);
        for (auto &pos : S) pos += num_leaves_ - 1;
        std::vector<std::size_t> aux_idx;
        while (true) {
            auto it = S.begin();
            if (*it == 0 && it == --S.end()) break;
            std::vector<std::size_t> new_S;
            while (it != S.end()) {
                const std::size_t it_pos = *it;
                auto next_it = ++it;
                new_S.emplace_back((it_pos - 1) / 2);
                if ((it_pos & 1) == 0) aux_idx.push_back(it_pos - 1);
                else if (next_it == S.end() || *next_it != it_pos + 1) aux_idx.push_back(it_pos + 1);
                else ++next_it;
                it = next_it;
            }
            std::swap(S, new_S);
        }
        std::vector<uint8_t> raw(aux_idx.size() * digest_len_bytes_);
        throw_status(context(), b200_merkle_nodes(context(), nodes_, aux_idx.data(), aux_idx.size(), digest_len_bytes_, raw.data()));
        for (std::size_t i = 0; i < aux_idx.size(); ++i)
            result.auxiliary_hashes.emplace_back(reinterpret_cast<const char *>(raw.data()) + i * digest_len_bytes_, digest_len_bytes_);
        return result;
    }

    std::size_t num_leaves() const { return num_leaves_; }
    bool zk() const { return make_zk_; }
    b200_bytes device_nodes() const { return nodes_; }
};

} // namespace libiop_b200
#endif
