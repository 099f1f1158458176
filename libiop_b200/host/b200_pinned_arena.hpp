/*
 * b200_pinned_arena.hpp — host memory of the prover's large vectors: a page-locked arena and a region of LAZY blocks.
 *
 * Why: the reference hands every oracle around as std::vector<FieldT> (iop/oracles.hpp:21-52,
 * SURVEY.md §8b).  With pageable vectors a 192 MiB codeword costs ~50-100 ms per host<->device
 * move (driver staging copies plus first-touch page faults of the fresh vector) while the kernel
 * that produced it ran for 8 ms.  Serving the large blocks of the process from cudaHostAlloc'd
 * slabs makes every such vector DMA-able in place at PCIe line rate and removes the page faults.
 *
 * How: B200_DEFINE_PINNED_OPERATOR_NEW (below) replaces the global allocation functions of the
 * binary that includes it (C++ [replacement.functions]); blocks >= B200_PINNED_MIN_BYTES (32 KiB) come from
 * the arena, everything else from malloc.  operator delete recognises arena blocks by address, so
 * memory allocated by either path can be released by the other translation units of the process.
 * If page-locking fails (no device, ulimit) the arena declines and malloc serves the request:
 * only speed changes, never results.
 *
 * Environment: B200_PINNED_ARENA=0 disables the arena, B200_PINNED_SLAB_MB sets the slab size
 * (default 1024).  The arena is thread-safe (one mutex); slabs are never returned to the OS.
 *
 * Lazy blocks (device-resident oracles).  The reference keeps every oracle as a std::vector<FieldT>
 * (iop/oracles.hpp:21-52) but reads only ~10^2 positions of it on the host (the query phase).  make_lazy_vector() (see
 * b200_residency.hpp) therefore hands back a vector whose storage is address space only: a block of `lazy_region`, mapped
 * PROT_NONE, while the contents live in the device buffer registered for it.  Adapter calls that receive the vector use
 * the device buffer; the query phase gathers the queried positions; and ANY other host access — code the drop-in does not
 * shadow, a by-value copy, an in-place update — faults, the SIGSEGV handler downloads the whole vector into the block,
 * forgets the device copy and resumes the access.  A device copy is thus tied to the block by ownership (it exists exactly
 * while the block is live and untouched by the host), not by comparing contents.  B200_LAZY=0 turns lazy blocks off
 * (every result is downloaded), B200_LAZY_TRACE=1 prints a backtrace for every materialising fault.
 *
 * A shared library that expands B200_DEFINE_PINNED_OPERATOR_NEW must link its C++ runtime statically and keep the
 * operators out of its dynamic symbol table (libiop_b200/dropin/Makefile does: -static-libstdc++, exports.map), so that
 * every allocation and release made by code of the library goes through the same pair of functions.
 */
#ifndef LIBIOP_B200_PINNED_ARENA_HPP_
#define LIBIOP_B200_PINNED_ARENA_HPP_

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include <execinfo.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

namespace libiop_b200 {

#ifndef B200_PINNED_MIN_BYTES
#define B200_PINNED_MIN_BYTES (std::size_t(32) << 10)   /* Ligero's 2^11-element rows (48 KiB) included */
#endif

/* First-fit free-list allocator over slabs obtained from `slab_alloc` (cudaHostAlloc in production, malloc in the
   CPU unit test).  Bookkeeping lives in malloc'd std::map nodes, never inside the pinned memory. */
class pinned_arena {
public:
    typedef void *(*slab_alloc_fn)(std::size_t);

    explicit pinned_arena(slab_alloc_fn slab_alloc, std::size_t slab_bytes) : slab_alloc_(slab_alloc), slab_bytes_(slab_bytes) {}

    void *allocate(std::size_t bytes)
    {
        bytes = (bytes + kAlign - 1) & ~(kAlign - 1);
        std::lock_guard<std::mutex> lock(mu_);
        void *p = take(bytes);
        if (!p && grow(bytes)) p = take(bytes);
        return p;
    }

    /* size of the live block at p (0 if p is not the start of one) */
    std::size_t block_bytes(const void *p)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = used_.find((uintptr_t)p);
        return it == used_.end() ? 0 : it->second;
    }

    /* true if p belonged to the arena (and has been released) */
    bool release(void *p)
    {
        const uintptr_t a = (uintptr_t)p;
        if (!owns(a)) return false;
        std::lock_guard<std::mutex> lock(mu_);
        auto it = used_.find(a);
        if (it != used_.end()) {
            std::size_t len = it->second;
            used_.erase(it);
            uintptr_t start = a;
            /* coalesce with the free neighbours (never across slabs: slabs are separate mappings) */
            auto next = free_.lower_bound(start);
            if (next != free_.end() && next->first == start + len && same_slab(start, next->first)) {
                len += next->second;
                next = free_.erase(next);
            }
            if (next != free_.begin()) {
                auto prev = std::prev(next);
                if (prev->first + prev->second == start && same_slab(prev->first, start)) {
                    start = prev->first;
                    len += prev->second;
                    free_.erase(prev);
                }
            }
            free_[start] = len;
        }
        return true;
    }

    bool owns(uintptr_t a) const
    {
        /* called for every operator delete of the process: one range test rejects malloc'd pointers */
        if (a < __atomic_load_n(&lo_, __ATOMIC_RELAXED) || a >= __atomic_load_n(&hi_, __ATOMIC_RELAXED)) return false;
        /* slabs_ only grows and is read without the lock: entries are published with their count last */
        const std::size_t n = __atomic_load_n(&n_slabs_, __ATOMIC_ACQUIRE);
        for (std::size_t i = 0; i < n; ++i)
            if (a >= slabs_[i].base && a < slabs_[i].base + slabs_[i].bytes) return true;
        return false;
    }

    std::size_t slab_count() const { return n_slabs_; }
    std::size_t reserved_bytes() const
    {
        std::size_t t = 0;
        for (std::size_t i = 0; i < n_slabs_; ++i) t += slabs_[i].bytes;
        return t;
    }
    std::size_t free_bytes()
    {
        std::lock_guard<std::mutex> lock(mu_);
        std::size_t t = 0;
        for (auto &kv : free_) t += kv.second;
        return t;
    }

private:
    static constexpr std::size_t kAlign = 4096;
    static constexpr std::size_t kMaxSlabs = 256;
    struct slab { uintptr_t base; std::size_t bytes; };

    bool same_slab(uintptr_t a, uintptr_t b) const
    {
        for (std::size_t i = 0; i < n_slabs_; ++i) {
            const bool ia = a >= slabs_[i].base && a < slabs_[i].base + slabs_[i].bytes;
            const bool ib = b >= slabs_[i].base && b < slabs_[i].base + slabs_[i].bytes;
            if (ia || ib) return ia && ib;
        }
        return false;
    }

    void *take(std::size_t bytes)
    {
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second < bytes) continue;
            const uintptr_t a = it->first;
            const std::size_t rest = it->second - bytes;
            free_.erase(it);
            if (rest) free_[a + bytes] = rest;
            used_[a] = bytes;
            return (void *)a;
        }
        return nullptr;
    }

    bool grow(std::size_t bytes)
    {
        if (n_slabs_ >= kMaxSlabs) return false;
        std::size_t want = bytes > slab_bytes_ ? bytes : slab_bytes_;
        void *p = slab_alloc_(want);
        if (!p && want > bytes) { want = bytes; p = slab_alloc_(want); }
        if (!p) return false;
        slabs_[n_slabs_].base = (uintptr_t)p;
        slabs_[n_slabs_].bytes = want;
        __atomic_store_n(&n_slabs_, n_slabs_ + 1, __ATOMIC_RELEASE);
        if ((uintptr_t)p < lo_) __atomic_store_n(&lo_, (uintptr_t)p, __ATOMIC_RELAXED);
        if ((uintptr_t)p + want > hi_) __atomic_store_n(&hi_, (uintptr_t)p + want, __ATOMIC_RELAXED);
        free_[(uintptr_t)p] = want;
        return true;
    }

    slab_alloc_fn slab_alloc_;
    std::size_t slab_bytes_;
    std::mutex mu_;
    slab slabs_[kMaxSlabs];
    std::size_t n_slabs_ = 0;
    uintptr_t lo_ = ~uintptr_t(0), hi_ = 0;          /* address range spanned by all slabs */
    std::map<uintptr_t, std::size_t> free_, used_;
};

/* Called with the address of every arena block just before it is released (the residency cache uses it to drop the
   device copy of a host vector the moment the vector dies).  One hook per process. */
typedef void (*arena_release_hook_fn)(void *);
inline arena_release_hook_fn &arena_release_hook()
{
    static arena_release_hook_fn hook = nullptr;
    return hook;
}

/* What the next operator new of exactly `bytes` bytes on this thread should return (set by make_lazy_vector around the one
   allocation it makes; consumed by the replaced operator new). */
struct alloc_hint { enum kind_t { none, lazy } kind; std::size_t bytes; };
inline alloc_hint &current_alloc_hint()
{
    static thread_local alloc_hint h = {alloc_hint::none, 0};
    return h;
}

/* Called with a faulting address inside the lazy region; returns true when the block was materialised and the access can
   be retried.  One hook per process (the residency table installs it). */
typedef bool (*lazy_fault_hook_fn)(void *);
inline lazy_fault_hook_fn &lazy_fault_hook()
{
    static lazy_fault_hook_fn hook = nullptr;
    return hook;
}

/* Address space for lazy blocks: reservations of PROT_NONE, MAP_NORESERVE memory cut into page-aligned blocks by the same
   free-list allocator as the pinned slabs.  Pages exist only for blocks the host touched (see the file comment). */
class lazy_region {
public:
    static lazy_region &instance() { static lazy_region *r = new lazy_region(); return *r; }   /* never destroyed */

    bool enabled() const { return enabled_; }
    bool contains(const void *p) const { return enabled_ && blocks_.owns((uintptr_t)p); }

    void *allocate(std::size_t bytes)
    {
        if (!enabled_) return nullptr;
        install_handler();
        return blocks_.allocate(bytes);
    }
    /* give a block back; pages the host touched are returned to the kernel and the range is protected again */
    void release(void *p)
    {
        const std::size_t len = blocks_.block_bytes(p);
        if (len && is_touched(p)) {
            madvise(p, len, MADV_DONTNEED);
            mprotect(p, len, PROT_NONE);
            set_touched(p, false);
        }
        blocks_.release(p);
    }
    /* make the block at p ordinary read-write memory (the caller fills it) */
    bool open(void *p)
    {
        const std::size_t len = blocks_.block_bytes(p);
        if (!len || mprotect(p, len, PROT_READ | PROT_WRITE) != 0) return false;
        set_touched(p, true);
        return true;
    }
    std::size_t reserved_bytes() const { return blocks_.reserved_bytes(); }
    std::size_t faults() const { return faults_; }

private:
    static void *reserve(std::size_t bytes)
    {
        void *p = mmap(nullptr, bytes, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        return p == MAP_FAILED ? nullptr : p;
    }
    lazy_region() : blocks_(&reserve, reservation_bytes())
    {
        const char *e = std::getenv("B200_LAZY");
        enabled_ = !(e && e[0] == '0');
        const char *t = std::getenv("B200_LAZY_TRACE");
        trace_ = t && t[0] == '1';
    }
    static std::size_t reservation_bytes()
    {
        const char *g = std::getenv("B200_LAZY_RESERVE_GB");
        return (std::size_t)(g ? std::strtoull(g, nullptr, 10) : 64) << 30;
    }
    bool is_touched(void *p) { std::lock_guard<std::mutex> l(mu_); return touched_.count((uintptr_t)p) != 0; }
    void set_touched(void *p, bool v)
    {
        std::lock_guard<std::mutex> l(mu_);
        if (v) touched_[(uintptr_t)p] = true; else touched_.erase((uintptr_t)p);
    }

    static void on_fault(int sig, siginfo_t *si, void *uctx)
    {
        lazy_region &r = instance();
        void *addr = si ? si->si_addr : nullptr;
        if (addr && r.contains(addr) && lazy_fault_hook()) {
            if (r.trace_) {
                void *frames[32];
                const int n = backtrace(frames, 32);
                static const char msg[] = "[b200 lazy] host access materialises a device-resident vector:\n";
                if (write(2, msg, sizeof msg - 1) < 0) {}
                backtrace_symbols_fd(frames, n, 2);
            }
            if (lazy_fault_hook()(addr)) { ++r.faults_; return; }
        }
        /* not ours: hand the signal to whoever had it before (Python's faulthandler, a debugger, the default action) */
        if (r.previous_.sa_flags & SA_SIGINFO) {
            if (r.previous_.sa_sigaction) { r.previous_.sa_sigaction(sig, si, uctx); return; }
        } else if (r.previous_.sa_handler != SIG_DFL && r.previous_.sa_handler != SIG_IGN) {
            r.previous_.sa_handler(sig);
            return;
        }
        signal(sig, SIG_DFL);          /* returning re-executes the access and takes the default action */
    }
    void install_handler()
    {
        std::lock_guard<std::mutex> l(mu_);
        if (handler_installed_) return;
        struct sigaction sa;
        std::memset(&sa, 0, sizeof sa);
        sa.sa_sigaction = &lazy_region::on_fault;
        sa.sa_flags = SA_SIGINFO | SA_NODEFER;
        sigemptyset(&sa.sa_mask);
        sigaction(SIGSEGV, &sa, &previous_);
        handler_installed_ = true;
    }

    pinned_arena blocks_;
    bool enabled_ = true, trace_ = false, handler_installed_ = false;
    std::mutex mu_;
    std::map<uintptr_t, bool> touched_;
    std::size_t faults_ = 0;
    struct sigaction previous_;
};

/* Small-block pool.  The reference builds its relations and IOP bookkeeping out of millions of tiny heap objects — the R1CS
   deep copy inside aurora_snark_prover alone is 3 * 2^20 vectors of one 32-byte term (aurora_iop's constructor,
   relations/r1cs.hpp) plus the std::map nodes of every registered query — and at N = 2^20 their allocation and release is
   ~0.2 s of a 0.45 s proof once the data-parallel work is on the device.  Blocks of up to 256 bytes are therefore served
   from per-thread free lists over 64 KiB pages of one address-space reservation: a pop / push of a singly linked list
   instead of a trip through malloc.  A page holds blocks of ONE size class, so operator delete finds the class from the
   address alone (unsized delete works), and a block may be released by any thread (it joins that thread's list).
   B200_SMALL_POOL=0 turns the pool off; pages are never returned to the OS (they are reused through the lists). */
class small_pool {
public:
    static constexpr std::size_t kMaxBytes = 256, kStep = 16, kClasses = kMaxBytes / kStep, kPageLog = 16;
    static small_pool &instance() { static small_pool *p = create(); return *p; }
    bool enabled() const { return base_ != nullptr; }
    std::size_t max_bytes() const { return max_bytes_; }
    bool contains(const void *p) const { return (uintptr_t)p - (uintptr_t)base_ < bytes_; }

    void *allocate(std::size_t n)
    {
        const std::size_t cls = (n + kStep - 1) / kStep - (n ? 1 : 0);      /* 1..16 -> 0, 17..32 -> 1, ... */
        cache &c = local();
        node *head = c.free_[cls];
        if (head) { c.free_[cls] = head->next; return head; }
        return refill(c, cls);
    }
    void release(void *p)
    {
        const std::size_t cls = page_class_[((uintptr_t)p - (uintptr_t)base_) >> kPageLog];
        cache &c = local();
        node *nd = static_cast<node *>(p);
        nd->next = c.free_[cls];
        c.free_[cls] = nd;
    }

private:
    struct node { node *next; };
    struct cache { node *free_[kClasses]; };
    /* initial-exec TLS: a plain %fs-relative access instead of a __tls_get_addr call per allocation (128 bytes of static TLS) */
    static cache &local() { static thread_local cache c __attribute__((tls_model("initial-exec"))) = {}; return c; }

    static small_pool *create()
    {
        void *mem = std::malloc(sizeof(small_pool));
        small_pool *p = new (mem) small_pool();
        const char *off = std::getenv("B200_SMALL_POOL");
        if (off && off[0] == '0') return p;
        const char *g = std::getenv("B200_SMALL_POOL_GB");
        const std::size_t bytes = (std::size_t)(g ? std::strtoull(g, nullptr, 10) : 32) << 30;
        void *r = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (r == MAP_FAILED) return p;
        p->page_class_ = static_cast<unsigned char *>(std::malloc(bytes >> kPageLog));
        if (!p->page_class_) { munmap(r, bytes); return p; }
        p->base_ = static_cast<char *>(r); p->bytes_ = bytes;
        const char *mx = std::getenv("B200_SMALL_POOL_MAX");          /* largest block served (bytes, <= 256) */
        if (mx) p->max_bytes_ = std::min<std::size_t>(kMaxBytes, std::strtoull(mx, nullptr, 10));
        return p;
    }
    /* a fresh page for class `cls`, cut into blocks that go onto the calling thread's list; nullptr when the reservation is used up */
    void *refill(cache &c, std::size_t cls)
    {
        const std::size_t page = __atomic_fetch_add(&next_page_, 1, __ATOMIC_RELAXED);
        if (((page + 1) << kPageLog) > bytes_) return nullptr;
        page_class_[page] = (unsigned char)cls;
        char *start = base_ + (page << kPageLog);
        const std::size_t bs = (cls + 1) * kStep, count = (std::size_t(1) << kPageLog) / bs;
        for (std::size_t i = count - 1; i >= 1; --i) {
            node *nd = reinterpret_cast<node *>(start + i * bs);
            nd->next = c.free_[cls];
            c.free_[cls] = nd;
        }
        return start;
    }
    char *base_ = nullptr;
    std::size_t bytes_ = 0, next_page_ = 0, max_bytes_ = kMaxBytes;
    unsigned char *page_class_ = nullptr;
};

/* std::vector<T> of n elements whose storage is NOT value-initialised (the caller overwrites all of it, e.g. by a
   device->host copy).  std::vector offers no such constructor; with libstdc++ / libc++ (three-pointer layout
   {begin, end, end_of_storage}) the end pointer is advanced over reserved storage.  Any other library takes the
   portable route (value-initialising constructor).  T must be trivially destructible. */
template<typename T>
std::vector<T> make_uninitialized_vector(std::size_t n)
{
#if defined(__GLIBCXX__) || defined(_LIBCPP_VERSION)
    static_assert(sizeof(std::vector<T>) == 3 * sizeof(T *), "unexpected std::vector layout");
    std::vector<T> v;
    v.reserve(n);
    T **raw = reinterpret_cast<T **>(&v);
    if (raw[0] == v.data() && raw[1] == v.data()) { raw[1] = v.data() + n; return v; }
#endif
    return std::vector<T>(n);
}

} // namespace libiop_b200

/*
 * Replacement global allocation functions; expand in exactly ONE translation unit of the final binary, after
 * including b200_device.hpp (the slabs come from b200_host_alloc = cudaHostAlloc on the adapter's context).
 */
#define B200_DEFINE_PINNED_OPERATOR_NEW                                                                                  \
    namespace libiop_b200 { namespace pinned_new_detail {                                                                \
    inline void *cuda_slab(std::size_t bytes)                                                                            \
    {                                                                                                                    \
        b200_ctx *c = nullptr;                                                                                           \
        try { c = context(); } catch (...) { return nullptr; }                                                          \
        return b200_host_alloc(c, bytes);                                                                                \
    }                                                                                                                    \
    inline pinned_arena *make_arena()              /* malloc + placement new: must not re-enter operator new */            \
    {                                                                                                                    \
        const char *off = std::getenv("B200_PINNED_ARENA");                                                              \
        if (off && off[0] == '0') return nullptr;                                                                        \
        const char *mb = std::getenv("B200_PINNED_SLAB_MB");                                                             \
        const std::size_t slab = (mb ? std::strtoull(mb, nullptr, 10) : 1024) << 20;                                     \
        void *mem = std::malloc(sizeof(pinned_arena));                                                                   \
        return mem ? new (mem) pinned_arena(&cuda_slab, slab) : nullptr;                                                 \
    }                                                                                                                    \
    inline pinned_arena *arena()                                                                                         \
    {                                                                                                                    \
        static pinned_arena *a = make_arena();     /* thread-safe one-time initialisation, never destroyed */            \
        return a;                                                                                                        \
    }                                                                                                                    \
    inline void *alloc(std::size_t n)                                                                                    \
    {                                                                                                                    \
        static thread_local bool inside = false; /* context creation / cudaHostAlloc may allocate: do not recurse */    \
        if (n <= small_pool::kMaxBytes) {                     /* the common case first: tiny objects (never re-enters) */  \
            small_pool &sp = small_pool::instance();                                                                     \
            if (sp.enabled() && n <= sp.max_bytes()) { void *p = sp.allocate(n); if (p) return p; }                      \
        }                                                                                                                \
        alloc_hint &hint = current_alloc_hint();                                                                         \
        if (hint.kind == alloc_hint::lazy && hint.bytes == n && !inside) {                                               \
            hint.kind = alloc_hint::none;                                                                                \
            inside = true;                                                                                               \
            void *p = lazy_region::instance().allocate(n);                                                               \
            inside = false;                                                                                              \
            if (p) return p;                                                                                             \
        }                                                                                                                \
        if (n >= B200_PINNED_MIN_BYTES && !inside) {                                                                     \
            inside = true;                                                                                               \
            pinned_arena *a = arena();                                                                                   \
            void *p = a ? a->allocate(n) : nullptr;                                                                      \
            inside = false;                                                                                              \
            if (p) return p;                                                                                             \
        }                                                                                                                \
        return std::malloc(n ? n : 1);                                                                                   \
    }                                                                                                                    \
    inline void dealloc(void *p) noexcept                                                                                \
    {                                                                                                                    \
        if (!p) return;                                                                                                  \
        if (small_pool::instance().contains(p)) { small_pool::instance().release(p); return; }                           \
        if (lazy_region::instance().contains(p)) {                                                                       \
            if (arena_release_hook()) arena_release_hook()(p);                                                           \
            lazy_region::instance().release(p);                                                                          \
            return;                                                                                                      \
        }                                                                                                                \
        pinned_arena *a = arena();                                                                                       \
        if (a && a->owns((uintptr_t)p)) {                                                                                \
            if (arena_release_hook()) arena_release_hook()(p);                                                           \
            a->release(p);                                                                                               \
            return;                                                                                                      \
        }                                                                                                                \
        std::free(p);                                                                                                    \
    }                                                                                                                    \
    } }                                                                                                                  \
    void *operator new(std::size_t n) { void *p = libiop_b200::pinned_new_detail::alloc(n); if (!p) throw std::bad_alloc(); return p; } \
    void *operator new[](std::size_t n) { void *p = libiop_b200::pinned_new_detail::alloc(n); if (!p) throw std::bad_alloc(); return p; } \
    void *operator new(std::size_t n, const std::nothrow_t &) noexcept { return libiop_b200::pinned_new_detail::alloc(n); } \
    void *operator new[](std::size_t n, const std::nothrow_t &) noexcept { return libiop_b200::pinned_new_detail::alloc(n); } \
    void operator delete(void *p) noexcept { libiop_b200::pinned_new_detail::dealloc(p); }                              \
    void operator delete[](void *p) noexcept { libiop_b200::pinned_new_detail::dealloc(p); }                            \
    void operator delete(void *p, std::size_t) noexcept { libiop_b200::pinned_new_detail::dealloc(p); }                 \
    void operator delete[](void *p, std::size_t) noexcept { libiop_b200::pinned_new_detail::dealloc(p); }               \
    void operator delete(void *p, const std::nothrow_t &) noexcept { libiop_b200::pinned_new_detail::dealloc(p); }      \
    void operator delete[](void *p, const std::nothrow_t &) noexcept { libiop_b200::pinned_new_detail::dealloc(p); }

#endif
