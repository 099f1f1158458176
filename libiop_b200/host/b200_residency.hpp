/*
 * b200_residency.hpp — device-resident oracles behind the reference's std::vector interfaces.
 *
 * The reference's IOP layer owns every oracle as a host std::vector<FieldT> (iop/oracles.hpp:21-52) and hands the same
 * vectors to the next stage: an LDE result goes to the Merkle tree, to two or three virtual oracles and to the FRI fold
 * (SURVEY.md §3.1); the host itself reads ~10^2 positions of each (the query phase, iop.tcc:670-688).  When the adapter
 * produces such a vector it therefore does NOT download it: make_lazy_vector() returns a vector whose storage is a block
 * of the lazy region (b200_pinned_arena.hpp: address space only, PROT_NONE) and registers the device buffer that holds
 * the contents under the block's address.
 *   - An adapter call that receives the vector (device_view) finds the buffer and uses it: no PCIe traffic.
 *   - The query phase reads single positions through resident_element() / prefetch_elements(): one gather per oracle.
 *   - Any OTHER host access faults; the handler downloads the whole vector into the block, drops the registration (the
 *     host may now change the data) and the access resumes.  Code the drop-in does not know about stays correct.
 *   - operator delete of the block drops the registration, so the buffer lives exactly as long as the vector.
 * A registration is thus valid by construction — it exists only while the block is live and has never been visible to
 * the host — and no contents are compared.  Without the replaced operator new (or with B200_LAZY=0) make_lazy_vector()
 * simply downloads.  The table is guarded by a mutex; the context and its stream remain single-threaded
 * (SURVEY.md §8b "Threading").
 */
#ifndef LIBIOP_B200_RESIDENCY_HPP_
#define LIBIOP_B200_RESIDENCY_HPP_

#include <array>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "b200_device.hpp"

namespace libiop_b200 {

class residency {
    struct entry {
        std::size_t n;                                        /* elements */
        device_buffer buf;
        std::unordered_map<std::size_t, std::array<uint64_t, 3>> fetched;   /* positions already read by the query phase */
        std::map<std::string, device_buffer> attached;       /* device arrays derived from this vector (e.g. its extension over L) */
        std::function<void(device_buffer &, std::map<std::string, device_buffer> &)> pending;   /* contents not computed yet (put_deferred) */
        b200_buf contents()                                   /* runs the deferred computation on first use */
        {
            if (pending) { auto f = std::move(pending); pending = nullptr; f(buf, attached); }
            return buf.get();
        }
    };
    std::map<uintptr_t, entry> blocks_;                       /* by block address, ordered: lookups by contained address */
    std::mutex mu_;
    std::size_t bytes_ = 0, hits_ = 0, misses_ = 0, faults_ = 0, fault_bytes_ = 0, gathers_ = 0;
    std::atomic<std::size_t> generation_{0};                  /* bumped by every registration */
    const void *bulk_owner_ = nullptr;                        /* last object that ran a bulk fetch, and when */
    std::size_t bulk_generation_ = ~std::size_t(0);

    residency()
    {
        arena_release_hook() = &residency::host_block_released;
        lazy_fault_hook() = &residency::host_block_touched;
    }
    static void host_block_released(void *p) { instance().drop(p); }
    static bool host_block_touched(void *addr) { return instance().materialise(addr); }

    /* entry whose block contains [a, a + bytes) */
    std::map<uintptr_t, entry>::iterator containing(uintptr_t a, std::size_t bytes)
    {
        auto it = blocks_.upper_bound(a);
        if (it == blocks_.begin()) return blocks_.end();
        --it;
        if (a >= it->first && a + bytes <= it->first + it->second.n * 24) return it;
        return blocks_.end();
    }

public:
    static residency &instance() { static residency *r = new residency(); return *r; }   /* never destroyed: outlives every vector */
    static constexpr std::size_t min_elems = 4096;            /* smaller results are simply downloaded */

    /* `buf` (n elements) holds the contents of the lazy block at `host` */
    void put(void *host, std::size_t n, device_buffer &&buf)
    {
        std::lock_guard<std::mutex> lock(mu_);
        entry &e = blocks_[(uintptr_t)host];
        e.n = n; e.buf = std::move(buf); e.fetched.clear();
        bytes_ += n * 24;
        ++generation_;
    }

    /* The same, with the contents still to be computed: `fill(buf, attachments)` runs — once, under the table's lock, so it
       must not call back into the table — the first time anything asks for the contents (a device view, a query, a host
       touch).  Consumers that only need something the producer attached (e.g. the evaluations the coefficients would be
       interpolated from) never trigger it.  `inputs` become the entry's attachments and are what `fill` reads. */
    void put_deferred(void *host, std::size_t n, device_buffer &&buf, std::map<std::string, device_buffer> &&inputs,
                      std::function<void(device_buffer &, std::map<std::string, device_buffer> &)> fill)
    {
        std::lock_guard<std::mutex> lock(mu_);
        entry &e = blocks_[(uintptr_t)host];
        e.n = n; e.buf = std::move(buf); e.fetched.clear(); e.attached = std::move(inputs); e.pending = std::move(fill);
        bytes_ += n * 24;
        ++generation_;
    }
    bool is_pending(const void *host)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        return it != blocks_.end() && (bool)it->second.pending;
    }

    /* the block died (operator delete) */
    void drop(const void *host)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        if (it == blocks_.end()) return;
        bytes_ -= it->second.n * 24;
        blocks_.erase(it);
    }

    /* device address of the n elements at `host` if they lie in a registered block; 0 otherwise */
    b200_buf find(const void *host, std::size_t n)
    {
        if (n == 0) return 0;
        std::lock_guard<std::mutex> lock(mu_);
        auto it = containing((uintptr_t)host, n * 24);
        if (it == blocks_.end()) { ++misses_; return 0; }
        const std::size_t off = (uintptr_t)host - it->first;
        if (off % 24) { ++misses_; return 0; }
        ++hits_;
        return it->second.contents() + off;
    }
    bool is_resident(const void *host) { std::lock_guard<std::mutex> lock(mu_); return containing((uintptr_t)host, 1) != blocks_.end(); }

    /* the host touched a lazy block: bring the contents over, the block is an ordinary host vector from now on */
    bool materialise(void *addr)
    {
        std::unique_lock<std::mutex> lock(mu_);
        auto it = containing((uintptr_t)addr, 1);
        if (it == blocks_.end()) return false;
        void *base = (void *)it->first;
        const std::size_t n = it->second.n;
        it->second.contents();
        device_buffer buf = std::move(it->second.buf);
        bytes_ -= n * 24;
        blocks_.erase(it);
        ++faults_; fault_bytes_ += n * 24;
        lock.unlock();
        if (!lazy_region::instance().open(base)) return false;
        return b200_download(context(), reinterpret_cast<uint64_t *>(base), buf.get(), 0, n) == B200_OK;
    }

    /* Query phase: element `position` of the vector at `host` (iop.tcc:687 reads evaluated_contents()[position]).
       Returns false if the vector is not device-resident (the caller then reads the host memory). */
    bool element(const void *host, std::size_t position, uint64_t out[3])
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        if (it == blocks_.end() || position >= it->second.n) return false;
        auto f = it->second.fetched.find(position);
        if (f == it->second.fetched.end()) {
            std::array<uint64_t, 3> v;
            throw_status(context(), b200_gather(context(), v.data(), it->second.contents(), &position, 1));
            ++gathers_;
            f = it->second.fetched.emplace(position, v).first;
        }
        std::memcpy(out, f->second.data(), 24);
        return true;
    }
    /* Keep `buf` next to the registered vector at `host` under `key`: it is released with the vector.  Returns false (and
       drops `buf`) when the vector is not device-resident. */
    bool attach(const void *host, const std::string &key, device_buffer &&buf)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        if (it == blocks_.end()) return false;
        it->second.attached[key] = std::move(buf);
        return true;
    }
    /* device address of the attachment `key` (of exactly n elements) of the vector at `host`, or 0 */
    b200_buf attachment(const void *host, const std::string &key, std::size_t n)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        if (it == blocks_.end()) return 0;
        auto a = it->second.attached.find(key);
        return (a != it->second.attached.end() && a->second.size() == n) ? a->second.get() : 0;
    }
    /* values of `positions` of the vector at `host` with one gather; false (nothing written) if it is not resident */
    bool gather(const void *host, const std::vector<std::size_t> &positions, uint64_t *out)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        if (it == blocks_.end()) return false;
        for (std::size_t p : positions) if (p >= it->second.n) return false;
        if (!positions.empty()) {
            throw_status(context(), b200_gather(context(), out, it->second.contents(), positions.data(), positions.size()));
            ++gathers_;
        }
        return true;
    }
    /* bumped by every registration: readers that cache fetched values compare it to know their cache is current */
    std::size_t generation() const { return generation_.load(std::memory_order_acquire); }
    /* true when `position` of the vector at `host` would need a gather (resident and not fetched yet) */
    bool needs_fetch(const void *host, std::size_t position)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        return it != blocks_.end() && position < it->second.n && !it->second.fetched.count(position);
    }
    /* A reader that is about to fetch its whole query set asks first: true once per (reader, state of the table), so a
       position the bulk fetch did not cover costs one single gather, not another walk over all queries. */
    bool begin_bulk_fetch(const void *owner)
    {
        std::lock_guard<std::mutex> lock(mu_);
        if (bulk_owner_ == owner && bulk_generation_ == generation_.load()) return false;
        bulk_owner_ = owner; bulk_generation_ = generation_.load();
        return true;
    }
    /* fetch many positions of one vector with a single gather (no-op for vectors that are not resident) */
    void prefetch(const void *host, const std::vector<std::size_t> &positions)
    {
        std::lock_guard<std::mutex> lock(mu_);
        auto it = blocks_.find((uintptr_t)host);
        if (it == blocks_.end()) return;
        std::vector<std::size_t> want;
        for (std::size_t p : positions)
            if (p < it->second.n && !it->second.fetched.count(p)) want.push_back(p);
        if (want.empty()) return;
        std::vector<uint64_t> vals(3 * want.size());
        throw_status(context(), b200_gather(context(), vals.data(), it->second.contents(), want.data(), want.size()));
        ++gathers_;
        for (std::size_t i = 0; i < want.size(); ++i) {
            std::array<uint64_t, 3> v = {vals[3 * i], vals[3 * i + 1], vals[3 * i + 2]};
            it->second.fetched.emplace(want[i], v);
        }
    }

    std::size_t hits() const { return hits_; }
    std::size_t misses() const { return misses_; }
    std::size_t faults() const { return faults_; }
    std::size_t fault_bytes() const { return fault_bytes_; }
    std::size_t gathers() const { return gathers_; }
    void count_gather() { std::lock_guard<std::mutex> lock(mu_); ++gathers_; }       /* a fetch made by the caller itself (b200_gather_multi) */
    std::size_t resident_bytes() const { return bytes_; }
};

/* a device view of host elements: the resident copy if there is one, otherwise a fresh (asynchronous) upload; the host
   data must stay alive and unmodified until the next synchronising call on the context */
class device_view {
    device_buffer owned_;      /* an upload, or a non-owning alias of the resident buffer */
    bool uploaded_ = false;
public:
    device_view(const void *host, std::size_t n)
    {
        const b200_buf resident = residency::instance().find(host, n);
        if (resident) { owned_ = device_buffer::alias(resident, n); return; }
        owned_ = device_buffer(n);
        if (n) { throw_status(context(), b200_upload(context(), owned_.get(), 0, reinterpret_cast<const uint64_t *>(host), n)); uploaded_ = true; }
    }
    device_view(device_view &&o) noexcept : owned_(std::move(o.owned_)), uploaded_(o.uploaded_) { o.uploaded_ = false; }
    device_view(const device_view &) = delete;
    device_view &operator=(const device_view &) = delete;
    /* the host data may be released by the caller once the view is gone: wait for the copy (not for the kernels behind it) */
    ~device_view() { if (uploaded_) b200_wait_uploads(context()); }
    template<typename FieldT> explicit device_view(const std::vector<FieldT> &v) : device_view(v.data(), v.size())
    {
        static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
    }
    const device_buffer &buf() const { return owned_; }
    b200_buf get() const { return owned_.get(); }
};

/* The host-side face of a device buffer: a vector of n elements that costs no transfer until the host reads it (see the
   file comment).  Falls back to a plain download when lazy blocks are unavailable or the result is small. */
template<typename FieldT>
std::vector<FieldT> make_lazy_vector(device_buffer &&buf, std::size_t n)
{
    static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
    if (n >= residency::min_elems && lazy_region::instance().enabled()) {
        alloc_hint &hint = current_alloc_hint();
        hint.kind = alloc_hint::lazy; hint.bytes = n * sizeof(FieldT);
        std::vector<FieldT> v = make_uninitialized_vector<FieldT>(n);
        hint.kind = alloc_hint::none;
        if (v.size() == n && lazy_region::instance().contains(v.data())) {
            residency::instance().put(v.data(), n, std::move(buf));
            return v;
        }
        /* the hint was not honoured (operator new is not replaced in this binary, or the reservation failed) */
        if (n) throw_status(context(), b200_download(context(), reinterpret_cast<uint64_t *>(v.data()), buf.get(), 0, n));
        return v;
    }
    return buf.template download<FieldT>(n);
}
/* A lazy vector whose contents are not even computed until somebody asks for them (residency::put_deferred): `fill(out,
   inputs)` writes the n elements into `out` from the device arrays in `inputs`, which stay attached to the vector under
   their keys.  Without lazy blocks the computation runs at once and the result is downloaded (the inputs are dropped). */
template<typename FieldT>
std::vector<FieldT> make_deferred_vector(std::size_t n, std::map<std::string, device_buffer> &&inputs,
                                         std::function<void(device_buffer &, std::map<std::string, device_buffer> &)> fill)
{
    static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
    device_buffer buf(n);
    if (n >= residency::min_elems && lazy_region::instance().enabled()) {
        alloc_hint &hint = current_alloc_hint();
        hint.kind = alloc_hint::lazy; hint.bytes = n * sizeof(FieldT);
        std::vector<FieldT> v = make_uninitialized_vector<FieldT>(n);
        hint.kind = alloc_hint::none;
        if (v.size() == n && lazy_region::instance().contains(v.data())) {
            residency::instance().put_deferred(v.data(), n, std::move(buf), std::move(inputs), std::move(fill));
            return v;
        }
        fill(buf, inputs);
        if (n) throw_status(context(), b200_download(context(), reinterpret_cast<uint64_t *>(v.data()), buf.get(), 0, n));
        return v;
    }
    fill(buf, inputs);
    return buf.template download<FieldT>(n);
}
/* historical name */
template<typename FieldT>
std::vector<FieldT> download_and_keep(device_buffer &&buf, std::size_t n) { return make_lazy_vector<FieldT>(std::move(buf), n); }

/* a copy of `src` that stays on the device when `src` does (the reference deep-copies oracles in a few places, e.g.
   oracle(const shared_ptr&), iop/oracles.hpp:36-38; bcs_indexer.tcc:48-53) */
template<typename FieldT>
std::vector<FieldT> clone_vector(const std::vector<FieldT> &src)
{
    const b200_buf resident = residency::instance().find(src.data(), src.size());
    if (!resident) return std::vector<FieldT>(src);
    device_buffer copy(src.size());
    throw_status(context(), b200_copy(context(), copy.get(), 0, resident, 0, src.size()));
    return make_lazy_vector<FieldT>(std::move(copy), src.size());
}

/* `v` in a form host code may read: `v` itself, or — for a device-resident vector — a downloaded copy in page-locked
   memory (a plain transfer at PCIe rate; reading `v` directly would fault and take the slow whole-vector path) */
template<typename FieldT>
class host_readable {
    const std::vector<FieldT> *ref_;
    std::vector<FieldT> copy_;
public:
    explicit host_readable(const std::vector<FieldT> &v) : ref_(&v)
    {
        const b200_buf resident = residency::instance().find(v.data(), v.size());
        if (!resident) return;
        copy_ = make_uninitialized_vector<FieldT>(v.size());
        throw_status(context(), b200_download(context(), reinterpret_cast<uint64_t *>(copy_.data()), resident, 0, v.size()));
        ref_ = &copy_;
    }
    const std::vector<FieldT> &get() const { return *ref_; }
};

/* element `position` of `v`, without materialising a device-resident vector */
template<typename FieldT>
FieldT resident_element(const std::vector<FieldT> &v, std::size_t position)
{
    FieldT out;
    if (residency::instance().element(v.data(), position, reinterpret_cast<uint64_t *>(&out))) return out;
    return v[position];
}

/* RAII owner of a device byte array (Merkle nodes that stay on the device: a tree of 2^23 nodes is 256 MiB of which
   the prover later reads ~50 authentication paths, merkle_tree.tcc:242-336) */
class device_bytes {
    b200_bytes dev_ = 0;
    std::size_t bytes_ = 0;
public:
    explicit device_bytes(std::size_t bytes) : bytes_(bytes)
    {
        dev_ = b200_alloc_bytes(context(), bytes);
        if (!dev_) throw std::runtime_error(b200_last_error(context()));
    }
    device_bytes(const device_bytes &) = delete;
    device_bytes &operator=(const device_bytes &) = delete;
    ~device_bytes() { if (dev_) b200_free_bytes(context(), dev_); }
    b200_bytes get() const { return dev_; }
    std::size_t size() const { return bytes_; }
    /* the digests of nodes `idx`, concatenated */
    std::string gather(const std::vector<std::size_t> &idx, std::size_t digest_len) const
    {
        std::string out(idx.size() * digest_len, '\0');
        if (!idx.empty())
            throw_status(context(), b200_merkle_nodes(context(), dev_, idx.data(), idx.size(), digest_len, reinterpret_cast<uint8_t *>(&out[0])));
        return out;
    }
};

} // namespace libiop_b200
#endif
