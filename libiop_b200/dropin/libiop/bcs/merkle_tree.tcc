/* Shadow of libiop/bcs/merkle_tree.tcc for the drop-in integration build.
   merkle_tree<F, binary_hash_digest>::construct_with_leaves_serialized_by_cosets (merkle_tree.hpp:88-90) is explicitly
   specialised for libff::gf192 and libff::edwards_Fr to build the tree on the B200.  The reference's own bookkeeping
   (zk salt sampling in stream order, merkle_tree.tcc:37-72,110-114) is kept.

   Node storage: the reference keeps one std::string per node (merkle_tree.hpp:49), i.e. 2^23 heap allocations for the
   first Aurora tree at N = 2^20 — 0.4 s of host time for a tree the device hashes in 3 ms.  A device-built tree is
   therefore stored FLAT: inner_nodes_ holds a single string with the (2*leaves-1)*digest_len heap-ordered bytes, and
   the two readers of inner_nodes_ (get_root, merkle_tree.tcc:230-240; get_set_membership_proof, :242-336) are
   specialised to slice it.  Trees built by the reference's other constructors keep the per-node layout; both layouts
   are recognised by the size of inner_nodes_. */
#include "b200_ref_traits.hpp"
#include B200_REF_FILE(libiop/bcs/merkle_tree.tcc)

#include <algorithm>

namespace libiop {
namespace b200_glue {

template<typename FieldT>
void gpu_merkle(const std::vector<std::shared_ptr<std::vector<FieldT>>> &leaf_contents, size_t coset, size_t num_leaves,
                size_t digest_len, const std::vector<zk_salt_type> *salts, std::vector<binary_hash_digest> &nodes_out)
{
    using namespace libiop_b200;
    std::vector<device_buffer> bufs;
    std::vector<b200_buf> handles;
    bufs.reserve(leaf_contents.size());
    for (auto &v : leaf_contents) { bufs.emplace_back(v->size()); bufs.back().upload_async(*v); handles.push_back(bufs.back().get()); }
    std::vector<uint8_t> flat;
    size_t salt_len = 0;
    if (salts) { salt_len = (*salts)[0].size(); flat.reserve(salt_len * salts->size()); for (auto &s : *salts) flat.insert(flat.end(), s.begin(), s.end()); }
    const size_t n_nodes = 2 * num_leaves - 1;
    b200_bytes dn = b200_alloc_bytes(context(), n_nodes * digest_len);
    if (!dn) throw std::runtime_error(b200_last_error(context()));
    const int kind = field_traits<FieldT>::kind == B200_FIELD_GF192 ? B200_DOMAIN_AFFINE_SUBSPACE : B200_DOMAIN_MULTIPLICATIVE_COSET;
    const int rc = b200_merkle_build(context(), handles.data(), handles.size(), leaf_contents[0]->size(), coset, kind,
                                     salts ? flat.data() : nullptr, salt_len, digest_len, dn);
    nodes_out.assign(1, binary_hash_digest());
    nodes_out[0].resize(n_nodes * digest_len);
    if (rc == 0) b200_download_bytes(context(), reinterpret_cast<uint8_t *>(&nodes_out[0][0]), dn, 0, n_nodes * digest_len);
    b200_free_bytes(context(), dn);
    throw_status(context(), rc);
}

/* node i of a tree in either layout */
inline binary_hash_digest node_at(const std::vector<binary_hash_digest> &nodes, size_t num_leaves, size_t digest_len, size_t i)
{
    if (nodes.size() == 2 * num_leaves - 1) return nodes[i];
    return nodes[0].substr(i * digest_len, digest_len);
}

/* Authentication paths for a set of leaves, pruned where two queried paths merge.  Same output order as
   merkle_tree.tcc:242-336: salts of the sorted distinct positions first, then level by level from the leaves up and
   left to right the sibling digests that the verifier cannot recompute from the queried nodes. */
inline merkle_tree_set_membership_proof<binary_hash_digest> membership_proof(
    const std::vector<binary_hash_digest> &nodes, size_t num_leaves, size_t digest_len, bool zk,
    const std::vector<zk_salt_type> &salts, const std::vector<std::size_t> &positions)
{
    merkle_tree_set_membership_proof<binary_hash_digest> proof;
    if (positions.empty()) return proof;
    std::vector<std::size_t> level(positions);
    std::sort(level.begin(), level.end());
    level.erase(std::unique(level.begin(), level.end()), level.end());
    if (level.back() >= num_leaves) throw std::invalid_argument("All positions must be between 0 and num_leaves-1.");
    if (zk) for (std::size_t pos : level) proof.randomness_hashes.emplace_back(salts[pos]);
    for (std::size_t &pos : level) pos += num_leaves - 1;          /* heap index of the leaf */
    while (!(level.size() == 1 && level[0] == 0)) {
        std::vector<std::size_t> parents;
        parents.reserve(level.size());
        for (std::size_t k = 0; k < level.size(); ++k) {
            const std::size_t idx = level[k];
            parents.push_back((idx - 1) / 2);
            if (idx % 2 == 0) {                                     /* right child whose left sibling is not queried */
                proof.auxiliary_hashes.emplace_back(node_at(nodes, num_leaves, digest_len, idx - 1));
            } else if (k + 1 < level.size() && level[k + 1] == idx + 1) {
                ++k;                                                /* both children known: nothing to send */
            } else {
                proof.auxiliary_hashes.emplace_back(node_at(nodes, num_leaves, digest_len, idx + 1));
            }
        }
        level.swap(parents);
    }
    return proof;
}
}

#define B200_MERKLE_SPECIALISE(FIELD)                                                                                              \
template<> inline void merkle_tree<FIELD, binary_hash_digest>::construct_with_leaves_serialized_by_cosets(                          \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &leaf_contents, const size_t coset_serialization_size)                  \
{                                                                                                                                  \
    if (this->constructed_) throw std::logic_error("Attempting to double-construct a Merkle tree.");                               \
    for (auto &v : leaf_contents)                                                                                                  \
        if ((v->size() / coset_serialization_size) != this->num_leaves_)                                                           \
            throw std::logic_error("Attempting to construct a Merkle tree with a constituent vector of wrong size");              \
    if (this->make_zk_) this->sample_leaf_randomness();                                                                            \
    b200_glue::gpu_merkle<FIELD>(leaf_contents, coset_serialization_size, this->num_leaves_, this->digest_len_bytes_,               \
                                 this->make_zk_ ? &this->zk_leaf_randomness_elements_ : nullptr, this->inner_nodes_);              \
    /* the device hashes with BLAKE2b: make sure that is what this tree was configured with */                                     \
    if (this->num_leaves_ > 1 &&                                                                                                   \
        this->node_hasher_(b200_glue::node_at(this->inner_nodes_, this->num_leaves_, this->digest_len_bytes_, 1),                   \
                           b200_glue::node_at(this->inner_nodes_, this->num_leaves_, this->digest_len_bytes_, 2),                   \
                           this->digest_len_bytes_) != b200_glue::node_at(this->inner_nodes_, this->num_leaves_, this->digest_len_bytes_, 0)) \
        throw std::logic_error("B200 Merkle build only implements the blake2b leaf / two-to-one hashes");                           \
    this->constructed_ = true;                                                                                                     \
}                                                                                                                                  \
template<> inline binary_hash_digest merkle_tree<FIELD, binary_hash_digest>::get_root() const                                      \
{                                                                                                                                  \
    if (!this->constructed_)                                                                                                       \
        throw std::logic_error("Attempting to obtain a Merkle tree root without constructing the tree first.");                    \
    return b200_glue::node_at(this->inner_nodes_, this->num_leaves_, this->digest_len_bytes_, 0);                                  \
}                                                                                                                                  \
template<> inline merkle_tree_set_membership_proof<binary_hash_digest>                                                             \
merkle_tree<FIELD, binary_hash_digest>::get_set_membership_proof(const std::vector<std::size_t> &positions) const                  \
{                                                                                                                                  \
    if (!this->constructed_)                                                                                                       \
        throw std::logic_error("Attempting to obtain a Merkle tree authentication path without constructing the tree first.");     \
    return b200_glue::membership_proof(this->inner_nodes_, this->num_leaves_, this->digest_len_bytes_, this->make_zk_,             \
                                       this->zk_leaf_randomness_elements_, positions);                                             \
}
B200_MERKLE_SPECIALISE(libff::gf192)
B200_MERKLE_SPECIALISE(libff::edwards_Fr)
#undef B200_MERKLE_SPECIALISE
} // namespace libiop
