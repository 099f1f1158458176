/* Shadow of libiop/bcs/merkle_tree.tcc for the drop-in integration build (TEST INFRASTRUCTURE):
   merkle_tree<F, binary_hash_digest>::construct_with_leaves_serialized_by_cosets (merkle_tree.hpp:88-90)
   is explicitly specialised for libff::gf192 and libff::edwards_Fr to build the tree on the B200.
   The reference's own bookkeeping (zk salt sampling in stream order, merkle_tree.tcc:37-72,110-114;
   inner_nodes_ as heap-ordered digests, merkle_tree.tcc:116,145) is kept so proofs and roots come from it. */
#include "b200_ref_traits.hpp"
#include B200_REF_FILE(libiop/bcs/merkle_tree.tcc)

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
    for (auto &v : leaf_contents) { bufs.emplace_back(v->size()); bufs.back().upload(*v); handles.push_back(bufs.back().get()); }
    std::vector<uint8_t> flat;
    size_t salt_len = 0;
    if (salts) { salt_len = (*salts)[0].size(); for (auto &s : *salts) flat.insert(flat.end(), s.begin(), s.end()); }
    const size_t n_nodes = 2 * num_leaves - 1;
    b200_bytes dn = b200_alloc_bytes(context(), n_nodes * digest_len);
    if (!dn) throw std::runtime_error(b200_last_error(context()));
    const int kind = field_traits<FieldT>::kind == B200_FIELD_GF192 ? B200_DOMAIN_AFFINE_SUBSPACE : B200_DOMAIN_MULTIPLICATIVE_COSET;
    const int rc = b200_merkle_build(context(), handles.data(), handles.size(), leaf_contents[0]->size(), coset, kind,
                                     salts ? flat.data() : nullptr, salt_len, digest_len, dn);
    std::vector<uint8_t> raw(n_nodes * digest_len);
    if (rc == 0) b200_download_bytes(context(), raw.data(), dn, 0, raw.size());
    b200_free_bytes(context(), dn);
    throw_status(context(), rc);
    nodes_out.resize(n_nodes);
    for (size_t i = 0; i < n_nodes; ++i) nodes_out[i].assign(reinterpret_cast<const char *>(raw.data()) + i * digest_len, digest_len);
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
    if (this->node_hasher_(this->inner_nodes_[1], this->inner_nodes_[2], this->digest_len_bytes_) != this->inner_nodes_[0])        \
        throw std::logic_error("B200 Merkle build only implements the blake2b leaf / two-to-one hashes");                           \
    this->constructed_ = true;                                                                                                     \
}
B200_MERKLE_SPECIALISE(libff::gf192)
B200_MERKLE_SPECIALISE(libff::edwards_Fr)
#undef B200_MERKLE_SPECIALISE
} // namespace libiop
