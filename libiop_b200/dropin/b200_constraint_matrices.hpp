/*
 * b200_constraint_matrices.hpp — the R1CS matrices A, B, C of a constraint system, flattened ONCE and kept on the device.
 *
 * The reference walks the constraint system's rows twice per proof with one heap-allocated row object per step:
 *   r1cs_constraint_system::create_Az_Bz_Cz_from_variable_assignment (relations/r1cs.tcc:236-266)   A z, B z, C z
 *   multi_lincheck_virtual_oracle::set_challenge (lincheck/basic_lincheck_aux.tcc:63-84)             (sum_m r_m M_m)^T alpha-powers
 * Both are products with the same three sparse matrices.  The first of the two that runs flattens them in one sequential
 * pass over the constraints into CSR arrays (row pointers, column indices, explicit row indices, coefficients), uploads
 * them, and both products run on the device from there (b200_spmv_csr_dev, b200_gf192_scatter_mul_add_dev).
 *
 * Lifetime: an entry belongs to one r1cs_constraint_system OBJECT.  r1cs_sparse_matrix's constructor (shadowed in
 * relations/sparse_matrix.tcc) registers the shared_ptr it is given; the entry holds a weak_ptr and is valid only while
 * that pointer is alive and still points at the queried address, so a constraint system that dies — or another one that is
 * later allocated at the same address — can never be served stale matrices.  Systems nobody registered (no shared_ptr
 * seen) keep the host loops.
 */
#ifndef B200_DROPIN_CONSTRAINT_MATRICES_HPP_
#define B200_DROPIN_CONSTRAINT_MATRICES_HPP_
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_elementwise.hpp"
#include "libiop_b200/host/b200_residency.hpp"

namespace libiop {
namespace b200_glue {

struct flat_matrix {
    std::size_t n_rows = 0, nnz = 0;
    std::unique_ptr<libiop_b200::device_bytes> row_ptr, col, row;   /* u64[n_rows+1], u32[nnz], u32[nnz] */
    libiop_b200::device_buffer coeff;                               /* nnz elements */
};

template<typename FieldT>
class constraint_matrices {
    struct entry {
        std::weak_ptr<r1cs_constraint_system<FieldT>> alive;
        bool built = false;
        flat_matrix M[3];
    };
    std::map<const void *, entry> entries_;
    std::mutex mu_;

    static void upload_bytes(std::unique_ptr<libiop_b200::device_bytes> &dst, const void *src, std::size_t bytes)
    {
        using namespace libiop_b200;
        dst.reset(new device_bytes(bytes ? bytes : 8));
        if (bytes) throw_status(context(), b200_upload_bytes(context(), dst->get(), static_cast<const uint8_t *>(src), bytes));
    }
    static void build(entry &e, const r1cs_constraint_system<FieldT> &cs)
    {
        using namespace libiop_b200;
        const std::size_t n = cs.constraints_.size();
        std::vector<uint64_t> row_ptr[3];
        std::vector<uint32_t> col[3], row[3];
        std::vector<FieldT> coeff[3];
        for (int m = 0; m < 3; ++m) { row_ptr[m].reserve(n + 1); row_ptr[m].push_back(0); col[m].reserve(n + n / 8); row[m].reserve(n + n / 8); coeff[m].reserve(n + n / 8); }
        for (std::size_t i = 0; i < n; ++i) {
            const linear_combination<FieldT> *lc[3] = { &cs.constraints_[i].a_, &cs.constraints_[i].b_, &cs.constraints_[i].c_ };
            for (int m = 0; m < 3; ++m) {
                for (const auto &t : lc[m]->terms) { col[m].push_back((uint32_t)t.index_); row[m].push_back((uint32_t)i); coeff[m].push_back(t.coeff_); }
                row_ptr[m].push_back(col[m].size());
            }
        }
        for (int m = 0; m < 3; ++m) {
            flat_matrix &F = e.M[m];
            F.n_rows = n; F.nnz = col[m].size();
            upload_bytes(F.row_ptr, row_ptr[m].data(), row_ptr[m].size() * sizeof(uint64_t));
            upload_bytes(F.col, col[m].data(), F.nnz * sizeof(uint32_t));
            upload_bytes(F.row, row[m].data(), F.nnz * sizeof(uint32_t));
            F.coeff = device_buffer(F.nnz ? F.nnz : 1);
            F.coeff.upload(coeff[m]);                        /* synchronises: the host arrays die with this scope */
        }
        throw_status(context(), b200_sync(context()));
        e.built = true;
    }

public:
    static constraint_matrices &instance() { static constraint_matrices *r = new constraint_matrices(); return *r; }

    /* called by r1cs_sparse_matrix's constructor: this constraint system is owned by `cs` */
    void remember(const std::shared_ptr<r1cs_constraint_system<FieldT>> &cs)
    {
        if (!cs) return;
        std::lock_guard<std::mutex> lock(mu_);
        for (auto it = entries_.begin(); it != entries_.end();)           /* systems that died release their device arrays */
            if (it->second.alive.expired()) it = entries_.erase(it); else ++it;
        entry &e = entries_[cs.get()];
        const std::shared_ptr<r1cs_constraint_system<FieldT>> cur = e.alive.lock();
        if (cur.get() != cs.get()) { e = entry(); e.alive = cs; }
    }
    /* the flattened matrices of the constraint system at `cs` (built on first use), or nullptr if that object is not a
       registered, still-living system large enough to be worth it */
    const flat_matrix *get(const r1cs_constraint_system<FieldT> *cs)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = entries_.find(cs);
        if (it == entries_.end()) return nullptr;
        const std::shared_ptr<r1cs_constraint_system<FieldT>> cur = it->second.alive.lock();
        if (cur.get() != cs) { entries_.erase(it); return nullptr; }
        if (cs->constraints_.size() < 16384 || cs->num_variables() + 1 >= (std::size_t(1) << 32)) return nullptr;
        if (!it->second.built) build(it->second, *cs);
        return it->second.M;
    }
};

} // namespace b200_glue
} // namespace libiop
#endif
