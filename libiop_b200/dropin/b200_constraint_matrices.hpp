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
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <future>
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
    struct host_arrays;
    struct entry {
        std::weak_ptr<r1cs_constraint_system<FieldT>> alive;
        bool built = false;
        flat_matrix M[3];
        std::future<std::shared_ptr<host_arrays>> prep;       /* a flatten running beside the caller (prefetch) */
    };
    std::map<const void *, entry> entries_;
    std::mutex mu_;

    static void upload_bytes(std::unique_ptr<libiop_b200::device_bytes> &dst, const void *src, std::size_t bytes)
    {
        using namespace libiop_b200;
        dst.reset(new device_bytes(bytes ? bytes : 8));
        if (bytes) throw_status(context(), b200_upload_bytes(context(), dst->get(), static_cast<const uint8_t *>(src), bytes));
    }
    /* host threads for the flatten: B200_HOST_THREADS, else up to 8 of the cores this process may run on */
    static unsigned fill_threads(std::size_t n)
    {
        if (const char *e = std::getenv("B200_HOST_THREADS")) { const long v = std::atol(e); return v < 1 ? 1u : (v > 64 ? 64u : (unsigned)v); }
        if (n < (std::size_t(1) << 16)) return 1;
        cpu_set_t set; CPU_ZERO(&set);
        const unsigned avail = sched_getaffinity(0, sizeof set, &set) == 0 ? (unsigned)CPU_COUNT(&set) : std::thread::hardware_concurrency();
        return avail >= 8 ? 8u : (avail ? avail : 1u);
    }
    struct raw_free { void operator()(void *p) const { ::operator delete(p); } };
    struct host_arrays {                                     /* the flattened matrices in (page-locked) host memory */
        std::size_t n = 0;
        uint64_t nnz[3] = { 0, 0, 0 };
        std::unique_ptr<uint64_t[]> row_ptr[3];             /* n + 1 each */
        std::unique_ptr<uint32_t[]> col[3], row[3];
        std::unique_ptr<void, raw_free> coeff[3];
    };
    /* runs f(t) for t = 0 .. T-1, t = 0 on the calling thread */
    template<typename Fn> static void on_threads(unsigned T, Fn f)
    {
        std::vector<std::thread> workers;
        for (unsigned t = 1; t < T; ++t) workers.emplace_back(f, t);
        f(0u);
        for (auto &w : workers) w.join();
    }
    /* Two passes over the constraints, row ranges split over a few threads.  The first reads only the rows' term counts (the
       constraint array itself, sequential) and lays down the row pointers relative to the range's start; the second — after
       the ranges' totals are prefix-summed and arrays of the exact size allocated — rebases them and copies indices and
       coefficients.  The work is bound by memory traffic (~300 bytes per constraint, one heap block per linear combination:
       24.7 ms single-threaded with push_back at N = 2^20, profiles/r2_host_gaps_N20_v1.txt).  Host work only: safe to run
       beside the thread that owns the device context. */
    static std::shared_ptr<host_arrays> flatten(const r1cs_constraint_system<FieldT> &cs)
    {
        std::shared_ptr<host_arrays> H = std::make_shared<host_arrays>();
        const std::size_t n = cs.constraints_.size();
        const bool trace = std::getenv("B200_FLATTEN_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        H->n = n;
        for (int m = 0; m < 3; ++m) { H->row_ptr[m].reset(new uint64_t[n + 1]); }
        const unsigned T = fill_threads(n);
        std::vector<std::array<uint64_t, 3>> range_total(T);
        host_arrays *h = H.get();
        on_threads(T, [&cs, h, n, T, &range_total](unsigned t) {
            uint64_t run[3] = { 0, 0, 0 };
            for (std::size_t i = n * t / T, end = n * (t + 1) / T; i < end; ++i) {
                const r1cs_constraint<FieldT> &c = cs.constraints_[i];
                h->row_ptr[0][i] = run[0]; run[0] += c.a_.terms.size();
                h->row_ptr[1][i] = run[1]; run[1] += c.b_.terms.size();
                h->row_ptr[2][i] = run[2]; run[2] += c.c_.terms.size();
            }
            range_total[t] = { run[0], run[1], run[2] };
        });
        std::vector<std::array<uint64_t, 3>> range_base(T);
        uint64_t run[3] = { 0, 0, 0 };
        for (unsigned t = 0; t < T; ++t) for (int m = 0; m < 3; ++m) { range_base[t][m] = run[m]; run[m] += range_total[t][m]; }
        FieldT *coeff[3];
        for (int m = 0; m < 3; ++m) {
            H->row_ptr[m][n] = H->nnz[m] = run[m];
            H->col[m].reset(new uint32_t[run[m] ? run[m] : 1]);
            H->row[m].reset(new uint32_t[run[m] ? run[m] : 1]);
            H->coeff[m].reset(::operator new((run[m] ? run[m] : 1) * sizeof(FieldT)));      /* every slot is written below */
            coeff[m] = static_cast<FieldT *>(H->coeff[m].get());
        }
        const auto t1 = std::chrono::steady_clock::now();
        on_threads(T, [&cs, h, n, T, &range_base, &coeff](unsigned t) {
            for (std::size_t i = n * t / T, end = n * (t + 1) / T; i < end; ++i) {
                const linear_combination<FieldT> *lc[3] = { &cs.constraints_[i].a_, &cs.constraints_[i].b_, &cs.constraints_[i].c_ };
                if (i + 12 < end) {                            /* the term blocks are separate heap objects: ask for them early */
                    const r1cs_constraint<FieldT> &ahead = cs.constraints_[i + 12];
                    __builtin_prefetch(ahead.a_.terms.data()); __builtin_prefetch(ahead.b_.terms.data()); __builtin_prefetch(ahead.c_.terms.data());
                }
                for (int m = 0; m < 3; ++m) {
                    uint64_t k = (h->row_ptr[m][i] += range_base[t][m]);
                    for (const auto &term : lc[m]->terms) { h->col[m][k] = (uint32_t)term.index_; h->row[m][k] = (uint32_t)i; new (&coeff[m][k]) FieldT(term.coeff_); ++k; }
                }
            }
        });
        if (trace) {
            const auto t2 = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[b200] flatten: %zu rows, count pass %.2f ms, fill pass %.2f ms on %u threads\n", n,
                         std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(), T);
        }
        return H;
    }
    static void upload(entry &e, const host_arrays &H)
    {
        using namespace libiop_b200;
        for (int m = 0; m < 3; ++m) {
            flat_matrix &F = e.M[m];
            F.n_rows = H.n; F.nnz = H.nnz[m];
            upload_bytes(F.row_ptr, H.row_ptr[m].get(), (H.n + 1) * sizeof(uint64_t));
            upload_bytes(F.col, H.col[m].get(), F.nnz * sizeof(uint32_t));
            upload_bytes(F.row, H.row[m].get(), F.nnz * sizeof(uint32_t));
            F.coeff = device_buffer(F.nnz ? F.nnz : 1);
            if (F.nnz) throw_status(context(), b200_upload(context(), F.coeff.get(), 0, static_cast<const uint64_t *>(H.coeff[m].get()), F.nnz));
        }
        throw_status(context(), b200_wait_uploads(context()));        /* the host arrays die with the caller's scope */
        e.built = true;
    }
    static bool worth_it(const r1cs_constraint_system<FieldT> &cs)
    {
        return cs.constraints_.size() >= 16384 && cs.num_variables() + 1 < (std::size_t(1) << 32);
    }
    /* the live entry of `cs` under the lock, or nullptr */
    entry *live_entry(const r1cs_constraint_system<FieldT> *cs, std::shared_ptr<r1cs_constraint_system<FieldT>> &keep)
    {
        auto it = entries_.find(cs);
        if (it == entries_.end()) return nullptr;
        keep = it->second.alive.lock();
        if (keep.get() != cs) { entries_.erase(it); return nullptr; }
        return &it->second;
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
        /* the same OWNER (control block), not merely the same address: a caller's system borrowed again for the next proof
           (aurora_iop.tcc shadow) is a new registration and is flattened anew — it may have changed in between */
        const bool same_owner = !e.alive.expired() && !e.alive.owner_before(cs) && !cs.owner_before(e.alive);
        if (!same_owner) { e = entry(); e.alive = cs; }
    }
    /* the flattened matrices of the constraint system at `cs` (built on first use), or nullptr if that object is not a
       registered, still-living system large enough to be worth it */
    const flat_matrix *get(const r1cs_constraint_system<FieldT> *cs)
    {
        std::lock_guard<std::mutex> lock(mu_);
        std::shared_ptr<r1cs_constraint_system<FieldT>> keep;
        entry *e = live_entry(cs, keep);
        if (!e || !worth_it(*cs)) return nullptr;
        if (!e->built) {
            const std::shared_ptr<host_arrays> H = e->prep.valid() ? e->prep.get() : flatten(*cs);
            upload(*e, *H);
        }
        return e->M;
    }
    /* Start flattening `cs` on another thread (host work only) if its matrices are going to be asked for and are not there
       yet: the prover calls this before the witness codeword keeps the device — and this thread — busy, so that get()
       finds the arrays ready.  The worker holds the owning pointer: the system cannot die under it. */
    void prefetch(const r1cs_constraint_system<FieldT> *cs)
    {
        std::lock_guard<std::mutex> lock(mu_);
        std::shared_ptr<r1cs_constraint_system<FieldT>> keep;
        entry *e = live_entry(cs, keep);
        if (!e || e->built || e->prep.valid() || !worth_it(*cs)) return;
        if (const char *off = std::getenv("B200_FLATTEN_ASYNC")) if (off[0] == '0') return;
        e->prep = std::async(std::launch::async, [keep]() { return flatten(*keep); });
    }
    /* wait for a flatten that prefetch() started and nobody consumed (error paths) */
    void settle(const r1cs_constraint_system<FieldT> *cs)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = entries_.find(cs);
        if (it != entries_.end() && it->second.prep.valid()) it->second.prep.wait();
    }
};

} // namespace b200_glue
} // namespace libiop
#endif
