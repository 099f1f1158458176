/* Shadow of libiop/bcs/merkle_tree.tcc for the drop-in integration build.
   merkle_tree<F, binary_hash_digest>::construct_with_leaves_serialized_by_cosets (merkle_tree.hpp:88-90) is explicitly
   specialised for libff::gf192 and libff::edwards_Fr to build the tree on the B200.  The reference's own bookkeeping
   (zk salts drawn from randombytes_buf with the call pattern of merkle_tree.tcc:37-72,110-114) is kept.

   Node storage: the reference keeps one std::string per node (merkle_tree.hpp:49), i.e. 2^23 heap allocations for the
   first Aurora tree at N = 2^20 — 0.4 s of host time for a tree the device hashes in 3 ms — and of those 256 MiB the
   prover later reads the root and ~50 authentication paths.  A device-built tree therefore leaves its nodes ON THE
   DEVICE: inner_nodes_ stays empty and the node array is owned by a decorator installed over the tree's leaf hasher
   (a shared_ptr member, so the array lives exactly as long as the tree and its copies).  The two readers of
   inner_nodes_ (get_root, merkle_tree.tcc:230-240; get_set_membership_proof, :242-336) are specialised to fetch the
   digests they need with one gather.  Trees built by the reference's other constructors keep the per-node layout. */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_residency.hpp"
#include B200_REF_FILE(libiop/bcs/merkle_tree.tcc)

#include <algorithm>

namespace libiop {
namespace b200_glue {

/* forwards leaf hashing to the tree's configured hasher and owns the device-resident node array */
template<typename FieldT>
class resident_nodes_leafhash : public leafhash<FieldT, binary_hash_digest> {
public:
    std::shared_ptr<leafhash<FieldT, binary_hash_digest>> inner;
    libiop_b200::device_bytes nodes;
    resident_nodes_leafhash(const std::shared_ptr<leafhash<FieldT, binary_hash_digest>> &h, std::size_t bytes) : inner(h), nodes(bytes) {}
    binary_hash_digest hash(const std::vector<FieldT> &leaf) { return inner->hash(leaf); }
    binary_hash_digest zk_hash(const std::vector<FieldT> &leaf, const zk_salt_type &zk_salt) { return inner->zk_hash(leaf, zk_salt); }
};

/* The zk salts of all leaves as ONE byte string, drawn from randombytes_buf with exactly the call sizes and order of
   merkle_tree::sample_leaf_randomness (merkle_tree.tcc:36-72: 1 MiB batches of which the first leafs_per_batch salts are
   used, then one call per remaining leaf), so the random stream is consumed identically — without the reference's one
   std::string per leaf (2^22 heap allocations for the first Aurora tree at N = 2^18 zk: 0.36 s). */
inline std::string sample_salts_flat(std::size_t num_leaves, std::size_t num_zk_bytes)
{
    libff::enter_block("BCS: Sample randomness");
    std::string salts(num_leaves * num_zk_bytes, '\0');
    const std::size_t batch_bytes = std::size_t(1) << 20;
    const std::size_t num_batches = (num_leaves * num_zk_bytes) / batch_bytes;
    const std::size_t leafs_per_batch = (num_batches == 0) ? 0 : num_leaves / num_batches;
    std::vector<uint8_t> batch(batch_bytes);
    std::size_t leaf = 0;
    for (std::size_t b = 0; b < num_batches; ++b) {
        randombytes_buf(batch.data(), batch_bytes);
        /* salt lengths that do not divide 2^20 make leafs_per_batch * num_zk_bytes exceed the batch (the reference reads past
           its buffer there, merkle_tree.tcc:52-58): the tail of such a batch is drawn with a further call instead */
        const std::size_t want = leafs_per_batch * num_zk_bytes, have = std::min(want, batch_bytes);
        std::memcpy(&salts[leaf * num_zk_bytes], batch.data(), have);
        if (want > have) randombytes_buf(&salts[leaf * num_zk_bytes + have], want - have);
        leaf += leafs_per_batch;
    }
    for (; leaf < num_leaves; ++leaf) randombytes_buf(&salts[leaf * num_zk_bytes], num_zk_bytes);
    libff::leave_block("BCS: Sample randomness");
    return salts;
}
/* salt of leaf `pos`: flat layout (one string for all leaves, trees built by the B200 path) or the reference's per-leaf strings */
inline zk_salt_type salt_at(const std::vector<zk_salt_type> &salts, std::size_t num_leaves, std::size_t pos)
{
    if (salts.size() == num_leaves) return salts[pos];
    const std::size_t len = salts[0].size() / num_leaves;
    return salts[0].substr(pos * len, len);
}

template<typename FieldT>
std::shared_ptr<resident_nodes_leafhash<FieldT>> gpu_merkle(const std::vector<std::shared_ptr<std::vector<FieldT>>> &leaf_contents, size_t coset,
                                                            size_t num_leaves, size_t digest_len, const std::vector<zk_salt_type> *salts,
                                                            const std::shared_ptr<leafhash<FieldT, binary_hash_digest>> &leaf_hasher)
{
    using namespace libiop_b200;
    std::vector<device_view> views;
    std::vector<b200_buf> handles;
    views.reserve(leaf_contents.size());
    for (auto &v : leaf_contents) { views.emplace_back(*v); handles.push_back(views.back().get()); }
    /* salts: one flat string (sample_salts_flat) */
    const size_t salt_len = salts ? (*salts)[0].size() / num_leaves : 0;
    const size_t n_nodes = 2 * num_leaves - 1;
    /* a tree rebuilt over an already decorated hasher keeps the original underneath */
    auto prev = std::dynamic_pointer_cast<resident_nodes_leafhash<FieldT>>(leaf_hasher);
    auto holder = std::make_shared<resident_nodes_leafhash<FieldT>>(prev ? prev->inner : leaf_hasher, n_nodes * digest_len);
    const int kind = field_traits<FieldT>::kind == B200_FIELD_GF192 ? B200_DOMAIN_AFFINE_SUBSPACE : B200_DOMAIN_MULTIPLICATIVE_COSET;
    throw_status(context(), b200_merkle_build(context(), handles.data(), handles.size(), leaf_contents[0]->size(), coset, kind,
                                              salts ? reinterpret_cast<const uint8_t *>((*salts)[0].data()) : nullptr, salt_len, digest_len,
                                              holder->nodes.get()));
    throw_status(context(), b200_sync(context()));       /* uploads of non-resident oracles and the salts are done */
    return holder;
}

/* digests of the heap nodes `idx`, whichever way the tree stores them */
template<typename FieldT>
std::vector<binary_hash_digest> nodes_at(const std::vector<binary_hash_digest> &inner_nodes,
                                         const std::shared_ptr<leafhash<FieldT, binary_hash_digest>> &leaf_hasher,
                                         size_t digest_len, const std::vector<std::size_t> &idx)
{
    std::vector<binary_hash_digest> out;
    out.reserve(idx.size());
    if (!inner_nodes.empty()) {
        for (std::size_t i : idx) out.emplace_back(inner_nodes[i]);
        return out;
    }
    auto holder = std::dynamic_pointer_cast<resident_nodes_leafhash<FieldT>>(leaf_hasher);
    if (!holder) throw std::logic_error("merkle_tree: no node storage");
    const std::string flat = holder->nodes.gather(idx, digest_len);
    for (std::size_t k = 0; k < idx.size(); ++k) out.emplace_back(flat.substr(k * digest_len, digest_len));
    return out;
}

/* Authentication paths for a set of leaves, pruned where two queried paths merge.  Same output order as
   merkle_tree.tcc:242-336: salts of the sorted distinct positions first, then level by level from the leaves up and
   left to right the sibling digests that the verifier cannot recompute from the queried nodes.  The heap indices of
   those siblings are collected first and fetched together. */
inline std::vector<std::size_t> membership_proof_nodes(size_t num_leaves, std::vector<std::size_t> level /* sorted distinct leaf positions */)
{
    std::vector<std::size_t> wanted;
    for (std::size_t &pos : level) pos += num_leaves - 1;          /* heap index of the leaf */
    while (!(level.size() == 1 && level[0] == 0)) {
        std::vector<std::size_t> parents;
        parents.reserve(level.size());
        for (std::size_t k = 0; k < level.size(); ++k) {
            const std::size_t idx = level[k];
            parents.push_back((idx - 1) / 2);
            if (idx % 2 == 0) {                                     /* right child whose left sibling is not queried */
                wanted.push_back(idx - 1);
            } else if (k + 1 < level.size() && level[k + 1] == idx + 1) {
                ++k;                                                /* both children known: nothing to send */
            } else {
                wanted.push_back(idx + 1);
            }
        }
        level.swap(parents);
    }
    return wanted;
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
    if (this->make_zk_)                                                                                                            \
        this->zk_leaf_randomness_elements_.assign(1, b200_glue::sample_salts_flat(this->num_leaves_, this->num_zk_bytes_));        \
    this->inner_nodes_.clear();                                                                                                    \
    this->leaf_hasher_ = b200_glue::gpu_merkle<FIELD>(leaf_contents, coset_serialization_size, this->num_leaves_,                  \
        this->digest_len_bytes_, this->make_zk_ ? &this->zk_leaf_randomness_elements_ : nullptr, this->leaf_hasher_);              \
    /* the device hashes with BLAKE2b: make sure that is what this tree was configured with */                                     \
    if (this->num_leaves_ > 1) {                                                                                                   \
        const std::vector<binary_hash_digest> top = b200_glue::nodes_at<FIELD>(this->inner_nodes_, this->leaf_hasher_,             \
                                                                              this->digest_len_bytes_, {0, 1, 2});                 \
        if (this->node_hasher_(top[1], top[2], this->digest_len_bytes_) != top[0])                                                 \
            throw std::logic_error("B200 Merkle build only implements the blake2b leaf / two-to-one hashes");                       \
    }                                                                                                                              \
    this->constructed_ = true;                                                                                                     \
}                                                                                                                                  \
template<> inline binary_hash_digest merkle_tree<FIELD, binary_hash_digest>::get_root() const                                      \
{                                                                                                                                  \
    if (!this->constructed_)                                                                                                       \
        throw std::logic_error("Attempting to obtain a Merkle tree root without constructing the tree first.");                    \
    return b200_glue::nodes_at<FIELD>(this->inner_nodes_, this->leaf_hasher_, this->digest_len_bytes_, {0})[0];                    \
}                                                                                                                                  \
template<> inline merkle_tree_set_membership_proof<binary_hash_digest>                                                             \
merkle_tree<FIELD, binary_hash_digest>::get_set_membership_proof(const std::vector<std::size_t> &positions) const                  \
{                                                                                                                                  \
    if (!this->constructed_)                                                                                                       \
        throw std::logic_error("Attempting to obtain a Merkle tree authentication path without constructing the tree first.");     \
    merkle_tree_set_membership_proof<binary_hash_digest> proof;                                                                    \
    if (positions.empty()) return proof;                                                                                           \
    std::vector<std::size_t> S(positions);                                                                                         \
    std::sort(S.begin(), S.end());                                                                                                 \
    S.erase(std::unique(S.begin(), S.end()), S.end());                                                                             \
    if (S.back() >= this->num_leaves_) throw std::invalid_argument("All positions must be between 0 and num_leaves-1.");           \
    if (this->make_zk_)                                                                                                            \
        for (std::size_t pos : S)                                                                                                  \
            proof.randomness_hashes.emplace_back(b200_glue::salt_at(this->zk_leaf_randomness_elements_, this->num_leaves_, pos));  \
    proof.auxiliary_hashes = b200_glue::nodes_at<FIELD>(this->inner_nodes_, this->leaf_hasher_, this->digest_len_bytes_,           \
                                                       b200_glue::membership_proof_nodes(this->num_leaves_, S));                   \
    return proof;                                                                                                                  \
}
B200_MERKLE_SPECIALISE(libff::gf192)
B200_MERKLE_SPECIALISE(libff::edwards_Fr)
#undef B200_MERKLE_SPECIALISE
} // namespace libiop
