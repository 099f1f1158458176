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
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <sched.h>

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
    /* host threads for the fill pass: B200_HOST_THREADS, else up to 4 of the cores this process may run on */
    static unsigned fill_threads(std::size_t n)
    {
        if (const char *e = std::getenv("B200_HOST_THREADS")) { const long v = std::atol(e); return v < 1 ? 1u : (v > 64 ? 64u : (unsigned)v); }
        if (n < (std::size_t(1) << 16)) return 1;
        cpu_set_t set; CPU_ZERO(&set);
        const unsigned avail = sched_getaffinity(0, sizeof set, &set) == 0 ? (unsigned)CPU_COUNT(&set) : std::thread::hardware_concurrency();
        return avail >= 4 ? 4u : (avail ? avail : 1u);
    }
    /* Two passes over the constraints.  The first reads only the rows' term counts (the constraint array itself, sequential)
       and lays down the three row-pointer arrays; the second copies indices and coefficients into arrays of the exact size,
       row ranges split over a few threads (the pass is bound by the pointer chase into one heap block per linear
       combination: 24.7 ms single-threaded with push_back at N = 2^20, profiles/r2_host_gaps_N20_v1.txt). */
    static void build(entry &e, const r1cs_constraint_system<FieldT> &cs)
    {
        using namespace libiop_b200;
        const std::size_t n = cs.constraints_.size();
        std::vector<uint64_t> row_ptr[3];
        for (int m = 0; m < 3; ++m) row_ptr[m].resize(n + 1);
        uint64_t run[3] = { 0, 0, 0 };
        for (std::size_t i = 0; i < n; ++i) {
            const r1cs_constraint<FieldT> &c = cs.constraints_[i];
            row_ptr[0][i] = run[0]; run[0] += c.a_.terms.size();
            row_ptr[1][i] = run[1]; run[1] += c.b_.terms.size();
            row_ptr[2][i] = run[2]; run[2] += c.c_.terms.size();
        }
        struct raw_free { void operator()(void *p) const { ::operator delete(p); } };
        std::unique_ptr<uint32_t[]> col[3], row[3];
        std::unique_ptr<void, raw_free> coeff_store[3];
        FieldT *coeff[3];
        for (int m = 0; m < 3; ++m) {
            row_ptr[m][n] = run[m];
            col[m].reset(new uint32_t[run[m] ? run[m] : 1]);
            row[m].reset(new uint32_t[run[m] ? run[m] : 1]);
            coeff_store[m].reset(::operator new((run[m] ? run[m] : 1) * sizeof(FieldT)));      /* every slot is written below */
            coeff[m] = static_cast<FieldT *>(coeff_store[m].get());
        }
        auto fill = [&](std::size_t i0, std::size_t i1) {
            for (std::size_t i = i0; i < i1; ++i) {
                const linear_combination<FieldT> *lc[3] = { &cs.constraints_[i].a_, &cs.constraints_[i].b_, &cs.constraints_[i].c_ };
                for (int m = 0; m < 3; ++m) {
                    uint64_t k = row_ptr[m][i];
                    for (const auto &t : lc[m]->terms) { col[m][k] = (uint32_t)t.index_; row[m][k] = (uint32_t)i; new (&coeff[m][k]) FieldT(t.coeff_); ++k; }
                }
            }
        };
        const unsigned T = fill_threads(n);
        if (T <= 1) fill(0, n);
        else {
            std::vector<std::thread> workers;
            for (unsigned t = 1; t < T; ++t) workers.emplace_back(fill, n * t / T, n * (t + 1) / T);
            fill(0, n / T);
            for (auto &w : workers) w.join();
        }
        for (int m = 0; m < 3; ++m) {
            flat_matrix &F = e.M[m];
            F.n_rows = n; F.nnz = run[m];
            upload_bytes(F.row_ptr, row_ptr[m].data(), row_ptr[m].size() * sizeof(uint64_t));
            upload_bytes(F.col, col[m].get(), F.nnz * sizeof(uint32_t));
            upload_bytes(F.row, row[m].get(), F.nnz * sizeof(uint32_t));
            F.coeff = device_buffer(F.nnz ? F.nnz : 1);
            if (F.nnz) throw_status(context(), b200_upload(context(), F.coeff.get(), 0, reinterpret_cast<const uint64_t *>(coeff[m]), F.nnz));
        }
        throw_status(context(), b200_sync(context()));        /* the host arrays die with this scope */
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
