/*
 * b200_constraint_matrices.hpp — the R1CS matrices A, B, C of a constraint system, flattened ONCE and kept on the device.
 *
 * The reference walks the constraint system's rows twice per proof with one heap-allocated row object per step:
 *   r1cs_constraint_system::create_Az_Bz_Cz_from_variable_assignment (relations/r1cs.tcc:236-266)   A z, B z, C z
 *   multi_lincheck_virtual_oracle::set_challenge (lincheck/basic_lincheck_aux.tcc:63-84)             (sum_m r_m M_m)^T alpha-powers
 * Both are products with the same three sparse matrices.  The first of the two that asks for a matrix flattens it into CSR
 * arrays (row pointers, column indices, explicit row indices, coefficients: two multi-threaded passes over the
 * constraints, no arithmetic), uploads them, and both products run on the device from there (b200_spmv_csr_dev,
 * b200_gf192_scatter_mul_add_dev).
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
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>

#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_elementwise.hpp"
#include "libiop_b200/host/b200_residency.hpp"

namespace libiop {
namespace b200_glue {

/* A few persistent host threads for the flatten passes.  Threads are created once per process (on first use, more when a
   larger fan-out is asked for) and parked on a condition variable: no thread creation — and no thread-exit loss of the small-
   block pool's per-thread free lists — per proof.  run(T, f) executes f(0) .. f(T-1), the caller taking part; one run at a
   time.  A forked child starts with an empty pool (threads do not survive fork). */
class host_workers {
    std::mutex mu_, run_mu_;
    std::condition_variable cv_work_, cv_done_;
    unsigned n_threads_ = 0;
    const std::function<void(unsigned)> *job_ = nullptr;
    unsigned n_tasks_ = 0, next_task_ = 0, done_tasks_ = 0;
    unsigned long long epoch_ = 0;

    void work()
    {
        for (;;) {
            unsigned t;
            const std::function<void(unsigned)> *f;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!job_ || next_task_ >= n_tasks_) return;
                t = next_task_++; f = job_;
            }
            (*f)(t);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (++done_tasks_ == n_tasks_) cv_done_.notify_all();
            }
        }
    }
    void loop()
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
            }
            work();
        }
    }
    static void forget_in_child() { new (&instance()) host_workers(); }          /* fresh mutexes, no threads */

public:
    static host_workers &instance()
    {
        static host_workers *w = [] { host_workers *p = new host_workers(); pthread_atfork(nullptr, nullptr, &host_workers::forget_in_child); return p; }();
        return *w;
    }
    void run(unsigned T, const std::function<void(unsigned)> &f)
    {
        if (T <= 1) { f(0u); return; }
        std::lock_guard<std::mutex> one_at_a_time(run_mu_);
        while (n_threads_ + 1 < T) { std::thread(&host_workers::loop, this).detach(); ++n_threads_; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &f; n_tasks_ = T; next_task_ = 0; done_tasks_ = 0; ++epoch_;
        }
        cv_work_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return done_tasks_ == n_tasks_; });
        job_ = nullptr;
    }
};

struct flat_matrix {
    std::size_t n_rows = 0, nnz = 0;
    std::unique_ptr<libiop_b200::device_bytes> row_ptr, col, row;   /* u64[n_rows+1], u32[nnz], u32[nnz] */
    libiop_b200::device_buffer coeff;                               /* nnz elements */
};

template<typename FieldT>
class constraint_matrices {
    /* One pass over the constraint array — 72 bytes per constraint, the only time it is read — leaves, per matrix, the row
       pointers and the address of each row's term block; the per-matrix fill passes then read 16 bytes per row instead of the
       constraint again (the flatten is bound by host memory traffic, and N ranks share the host's bandwidth). */
    struct row_index {
        std::size_t n = 0;
        uint64_t nnz[3] = { 0, 0, 0 };
        std::unique_ptr<uint64_t[]> row_ptr[3];                                  /* n + 1 each */
        std::unique_ptr<const linear_term<FieldT> *[]> terms[3];                 /* n each */
    };
    struct entry {
        std::weak_ptr<r1cs_constraint_system<FieldT>> alive;
        bool built[3] = { false, false, false };
        flat_matrix M[3];
        bool built_t[3] = { false, false, false };
        flat_matrix T[3];                                     /* the transposes, CSR over the columns of M */
        std::shared_ptr<row_index> rows;                      /* where every row's terms lie, for all three matrices (first use) */
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
    /* runs f(t) for t = 0 .. T-1 on the process's persistent host workers, the calling thread taking part */
    template<typename Fn> static void on_threads(unsigned T, Fn f)
    {
        const std::function<void(unsigned)> job(f);
        host_workers::instance().run(T, job);
    }
    static const linear_combination<FieldT> &row_of(const r1cs_constraint<FieldT> &c, int m) { return m == 0 ? c.a_ : m == 1 ? c.b_ : c.c_; }
    struct raw_free { void operator()(void *p) const { ::operator delete(p); } };

    /* Matrix m of `cs` (0 = A, 1 = B, 2 = C) flattened and uploaded.  The first matrix asked for pays the pass over the
       constraint array (index_rows: row pointers and term-block addresses of all three matrices, row ranges split over a few
       threads, range totals prefix-summed); every matrix then has its column indices and coefficients copied into arrays of
       the exact size (page-locked: the uploads are DMA in place).  The work is bound by memory
       traffic (one heap block per linear combination).  One matrix at a time so that the caller can start the device on
       A z while B is being flattened (r1cs_rs_iop.tcc shadow); profiles/r2_host_gaps_N20_v*.txt. */
    static std::shared_ptr<row_index> index_rows(const r1cs_constraint_system<FieldT> &cs)
    {
        std::shared_ptr<row_index> R = std::make_shared<row_index>();
        const std::size_t n = cs.constraints_.size();
        R->n = n;
        for (int m = 0; m < 3; ++m) { R->row_ptr[m].reset(new uint64_t[n + 1]); R->terms[m].reset(new const linear_term<FieldT> *[n ? n : 1]); }
        const unsigned T = fill_threads(n);
        std::vector<std::array<uint64_t, 3>> range_total(T);
        row_index *r = R.get();
        on_threads(T, [&cs, r, n, T, &range_total](unsigned t) {
            uint64_t run[3] = { 0, 0, 0 };
            for (std::size_t i = n * t / T, end = n * (t + 1) / T; i < end; ++i)
                for (int m = 0; m < 3; ++m) {
                    const std::vector<linear_term<FieldT>> &terms = row_of(cs.constraints_[i], m).terms;
                    r->row_ptr[m][i] = run[m]; r->terms[m][i] = terms.data(); run[m] += terms.size();
                }
            range_total[t] = { run[0], run[1], run[2] };
        });
        uint64_t base[3] = { 0, 0, 0 };
        std::vector<std::array<uint64_t, 3>> range_base(T);
        for (unsigned t = 0; t < T; ++t) for (int m = 0; m < 3; ++m) { range_base[t][m] = base[m]; base[m] += range_total[t][m]; }
        on_threads(T, [r, n, T, &range_base](unsigned t) {                        /* rebase: relative -> absolute row pointers */
            for (std::size_t i = n * t / T, end = n * (t + 1) / T; i < end; ++i)
                for (int m = 0; m < 3; ++m) r->row_ptr[m][i] += range_base[t][m];
        });
        for (int m = 0; m < 3; ++m) { R->row_ptr[m][n] = R->nnz[m] = base[m]; }
        return R;
    }
    static void build(entry &e, const r1cs_constraint_system<FieldT> &cs, int m)
    {
        using namespace libiop_b200;
        const std::size_t n = cs.constraints_.size();
        const bool trace = std::getenv("B200_FLATTEN_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        if (!e.rows) e.rows = index_rows(cs);
        const row_index &R = *e.rows;
        const uint64_t nnz = R.nnz[m];
        const uint64_t *rp = R.row_ptr[m].get();
        const linear_term<FieldT> *const *tp = R.terms[m].get();
        const unsigned T = fill_threads(n);
        std::unique_ptr<uint32_t[]> col(new uint32_t[nnz ? nnz : 1]);
        std::unique_ptr<void, raw_free> coeff_store(::operator new((nnz ? nnz : 1) * sizeof(FieldT)));      /* every slot is written below */
        FieldT *coeff = static_cast<FieldT *>(coeff_store.get());
        uint32_t *colp = col.get();
        const auto t1 = std::chrono::steady_clock::now();
        on_threads(T, [rp, tp, colp, coeff, n, T](unsigned t) {
            for (std::size_t i = n * t / T, end = n * (t + 1) / T; i < end; ++i) {
                const linear_term<FieldT> *term = tp[i];
                for (uint64_t k = rp[i], stop = rp[i + 1]; k < stop; ++k, ++term) { colp[k] = (uint32_t)term->index_; new (&coeff[k]) FieldT(term->coeff_); }
            }
        });
        const auto t2 = std::chrono::steady_clock::now();
        flat_matrix &F = e.M[m];
        F.n_rows = n; F.nnz = nnz;
        upload_bytes(F.row_ptr, rp, (n + 1) * sizeof(uint64_t));
        upload_bytes(F.col, colp, nnz * sizeof(uint32_t));
        F.row.reset(new device_bytes((nnz ? nnz : 1) * sizeof(uint32_t)));       /* the explicit row of every entry: expanded on the device */
        throw_status(context(), b200_csr_row_indices_dev(context(), n, F.row_ptr->get(), F.row->get()));
        F.coeff = device_buffer(nnz ? nnz : 1);
        if (nnz) throw_status(context(), b200_upload(context(), F.coeff.get(), 0, reinterpret_cast<const uint64_t *>(coeff), nnz));
        throw_status(context(), b200_wait_uploads(context()));        /* the host arrays die with this scope */
        e.built[m] = true;
        if (e.built[0] && e.built[1] && e.built[2]) e.rows.reset();
        if (trace)
            std::fprintf(stderr, "[b200] flatten %c: %zu rows, %llu entries, row index %.2f ms, fill pass %.2f ms on %u threads, upload %.2f ms\n",
                         "ABC"[m], n, (unsigned long long)nnz, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                         std::chrono::duration<double, std::milli>(t2 - t1).count(), T,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count());
    }
    /* The transpose of matrix m as CSR over M's columns (num_variables + 1 of them): n_rows = columns of M, `col` = the row of
       each entry.  For fields without an atomic addition on the device the transposed product (compute_p_alpha_M,
       lincheck/common.tcc:3-36) is then the same gather product as M z.  Made on the device from the flattened matrix
       (b200_csr_transpose_dev: counting sort by column; a host counting sort took 4 ms per matrix at N = 2^18). */
    static void build_transposed(entry &e, const r1cs_constraint_system<FieldT> &cs, int m)
    {
        using namespace libiop_b200;
        if (!e.built[m]) build(e, cs, m);
        const flat_matrix &S = e.M[m];
        const std::size_t n_cols = cs.num_variables() + 1;
        flat_matrix &F = e.T[m];
        F.n_rows = n_cols; F.nnz = S.nnz;
        F.row_ptr.reset(new device_bytes((n_cols + 1) * sizeof(uint64_t)));
        F.col.reset(new device_bytes((S.nnz ? S.nnz : 1) * sizeof(uint32_t)));
        F.row.reset();
        F.coeff = device_buffer(S.nnz ? S.nnz : 1);
        throw_status(context(), b200_csr_transpose_dev(context(), n_cols, S.nnz, S.col->get(), S.row->get(), S.coeff.get(),
                                                       F.row_ptr->get(), F.col->get(), F.coeff.get()));
        e.built_t[m] = true;
    }
    static bool worth_it(const r1cs_constraint_system<FieldT> &cs)
    {
        return cs.constraints_.size() >= 16384 && cs.num_variables() + 1 < (std::size_t(1) << 32);
    }
    /* the live entry of `cs` (call under the lock), or nullptr */
    entry *live_entry(const r1cs_constraint_system<FieldT> *cs)
    {
        auto it = entries_.find(cs);
        if (it == entries_.end()) return nullptr;
        const std::shared_ptr<r1cs_constraint_system<FieldT>> cur = it->second.alive.lock();
        if (cur.get() != cs) { entries_.erase(it); return nullptr; }
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
    /* will get_matrix() / get() serve this constraint system? */
    bool serves(const r1cs_constraint_system<FieldT> *cs)
    {
        std::lock_guard<std::mutex> lock(mu_);
        return live_entry(cs) != nullptr && worth_it(*cs);
    }
    /* matrix m (0 = A, 1 = B, 2 = C) of the constraint system at `cs`, flattened on first use, or nullptr if that object is
       not a registered, still-living system large enough to be worth it */
    const flat_matrix *get_matrix(const r1cs_constraint_system<FieldT> *cs, int m)
    {
        std::lock_guard<std::mutex> lock(mu_);
        entry *e = live_entry(cs);
        if (!e || !worth_it(*cs)) return nullptr;
        if (!e->built[m]) build(*e, *cs, m);
        return &e->M[m];
    }
    /* the transpose of matrix m (see build_transposed), or nullptr like get_matrix */
    const flat_matrix *get_transposed(const r1cs_constraint_system<FieldT> *cs, int m)
    {
        std::lock_guard<std::mutex> lock(mu_);
        entry *e = live_entry(cs);
        if (!e || !worth_it(*cs)) return nullptr;
        if (!e->built_t[m]) build_transposed(*e, *cs, m);
        return &e->T[m];
    }
    /* all three (an array of 3) */
    const flat_matrix *get(const r1cs_constraint_system<FieldT> *cs)
    {
        std::lock_guard<std::mutex> lock(mu_);
        entry *e = live_entry(cs);
        if (!e || !worth_it(*cs)) return nullptr;
        for (int m = 0; m < 3; ++m) if (!e->built[m]) build(*e, *cs, m);
        return e->M;
    }
};

} // namespace b200_glue
} // namespace libiop
#endif
