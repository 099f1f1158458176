/*
 * b200_pinned_arena.hpp — page-locked host arena for the prover's large vectors.
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
 */
#ifndef LIBIOP_B200_PINNED_ARENA_HPP_
#define LIBIOP_B200_PINNED_ARENA_HPP_

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

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
