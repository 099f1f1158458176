/*
 * b200_matrix_views.hpp — from the reference's sparse_matrix views (relations/sparse_matrix.hpp) to the device-resident
 * matrices of b200_constraint_matrices.hpp, plus the index tables the protocols address them with.  Included by the shadows of
 * the lincheck and indexer files, after relations/sparse_matrix.hpp.
 */
#ifndef B200_DROPIN_MATRIX_VIEWS_HPP_
#define B200_DROPIN_MATRIX_VIEWS_HPP_
#include <list>
#include <memory>
#include <mutex>
#include <vector>

#include "libiop/relations/sparse_matrix.hpp"
#include "libiop/algebra/field_subset/field_subset.hpp"
#include "b200_constraint_matrices.hpp"

namespace libiop {
namespace b200_glue {
/* the constraint system behind an r1cs_sparse_matrix (a protected member: reached through a pointer to member formed in
   a derived class), so that rows can be walked by reference instead of through get_row()'s by-value copies */
template<typename FieldT>
struct r1cs_matrix_peek : public r1cs_sparse_matrix<FieldT> {
    static const r1cs_constraint_system<FieldT> &system(const r1cs_sparse_matrix<FieldT> &m) { return *(m.*(&r1cs_matrix_peek::constraint_system_)); }
    static r1cs_sparse_matrix_type type(const r1cs_sparse_matrix<FieldT> &m) { return m.*(&r1cs_matrix_peek::matrix_type_); }
};
/* the device-resident form of M (or of its transpose) when M is a view of a registered constraint system, else nullptr */
template<typename FieldT>
const flat_matrix *device_matrix_of(const sparse_matrix<FieldT> &M, bool transposed)
{
    const r1cs_sparse_matrix<FieldT> *rm = dynamic_cast<const r1cs_sparse_matrix<FieldT> *>(&M);
    if (!rm) return nullptr;
    const int m = (int)r1cs_matrix_peek<FieldT>::type(*rm) - 1;
    if (m < 0 || m > 2) return nullptr;
    const r1cs_constraint_system<FieldT> *cs = &r1cs_matrix_peek<FieldT>::system(*rm);
    return transposed ? constraint_matrices<FieldT>::instance().get_transposed(cs, m) : constraint_matrices<FieldT>::instance().get_matrix(cs, m);
}

/* field_subset::reindex_by_subset(subset_dim, i) for i < n as a device array of u32 (field_subset.tcc:126-142: the identity for
   subspaces, subgroup.tcc:150-173 for cosets); the last few tables stay on the device */
template<typename FieldT>
b200_bytes reindex_table(const field_subset<FieldT> &domain, std::size_t subset_dim, std::size_t n)
{
    struct entry { int type; std::size_t dim, subset_dim, n; std::unique_ptr<libiop_b200::device_bytes> table; };
    static std::list<entry> *cache = new std::list<entry>();
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (auto it = cache->begin(); it != cache->end(); ++it)
        if (it->type == (int)domain.type() && it->dim == domain.dimension() && it->subset_dim == subset_dim && it->n == n) {
            cache->splice(cache->begin(), *cache, it);
            return cache->front().table->get();
        }
    std::vector<uint32_t> host(n ? n : 1);
    for (std::size_t i = 0; i < n; ++i) host[i] = (uint32_t)domain.reindex_by_subset(subset_dim, i);
    std::unique_ptr<libiop_b200::device_bytes> dev(new libiop_b200::device_bytes(host.size() * sizeof(uint32_t)));
    libiop_b200::throw_status(libiop_b200::context(), b200_upload_bytes(libiop_b200::context(), dev->get(), reinterpret_cast<const uint8_t *>(host.data()),
                                                                        host.size() * sizeof(uint32_t)));
    libiop_b200::throw_status(libiop_b200::context(), b200_sync(libiop_b200::context()));
    while (cache->size() >= 4) cache->pop_back();
    cache->push_front(entry{ (int)domain.type(), domain.dimension(), subset_dim, n, std::move(dev) });
    return cache->front().table->get();
}

} // namespace b200_glue
} // namespace libiop
#endif
