/*
 * b200_residency.hpp — keeps device copies of the prover's large host vectors so they do not cross PCIe twice.
 *
 * The reference's IOP layer owns every oracle as a host std::vector<FieldT> (iop/oracles.hpp:21-52) and hands the same
 * vectors to the next stage: an LDE result goes to the Merkle tree, to two or three virtual oracles and to the FRI fold
 * (SURVEY.md §3.1).  The adapter still returns host vectors (the reference's query phase reads them element-wise,
 * iop.tcc:670-688), but whenever it produces one from a device buffer it parks that buffer here, keyed by the vector's
 * data pointer; when a later call receives the same vector it finds the device copy and skips the upload.
 *
 * Safety: a hit requires the same address AND length AND an equal sample of 32 elements (first, last, evenly strided)
 * read from the host vector at lookup time, so an allocation recycled for different contents is never mistaken for
 * the vector that was registered.  Only codeword-domain evaluation vectors are registered (immutable once submitted,
 * iop.tcc:272-307).  When the process uses the pinned arena (b200_pinned_arena.hpp) the cache is told the moment a
 * host block is released and drops the entry, so device memory follows the lifetime of the host vectors.
 * B200_RESIDENCY_VERIFY=1 additionally compares the complete contents on every hit (debug / tests),
 * B200_RESIDENCY=0 disables the cache.  The cache is LRU with a byte budget (B200_RESIDENT_GB, default 24).
 */
#ifndef LIBIOP_B200_RESIDENCY_HPP_
#define LIBIOP_B200_RESIDENCY_HPP_

#include <cstdlib>
#include <cstring>
#include <list>
#include <unordered_map>
#include <string>
#include <vector>

#include "b200_device.hpp"

namespace libiop_b200 {

class residency {
    struct entry { const void *host; std::size_t n; unsigned char sample[32 * 24]; std::size_t sample_bytes; device_buffer buf; };
    std::list<entry> lru_;                       /* front = most recent */
    std::unordered_map<const void *, std::list<entry>::iterator> by_host_;   /* one entry per host address */
    std::size_t bytes_ = 0, budget_, hits_ = 0, misses_ = 0, stale_ = 0;
    bool enabled_, verify_;

    static std::size_t take_sample(const void *host, std::size_t n, unsigned char *out)
    {
        const unsigned char *p = static_cast<const unsigned char *>(host);
        const std::size_t k = n < 32 ? n : 32;
        for (std::size_t i = 0; i < k; ++i) std::memcpy(out + 24 * i, p + 24 * ((k > 1) ? (i * (n - 1)) / (k - 1) : 0), 24);
        return 24 * k;
    }
    residency()
    {
        const char *e = std::getenv("B200_RESIDENCY");
        enabled_ = !(e && e[0] == '0');
        const char *v = std::getenv("B200_RESIDENCY_VERIFY");
        verify_ = v && v[0] == '1';
        const char *g = std::getenv("B200_RESIDENT_GB");
        budget_ = (std::size_t)(g ? std::strtoull(g, nullptr, 10) : 24) << 30;
        arena_release_hook() = &residency::host_block_released;
    }
    /* the pinned arena is about to recycle this host block: its device copy is dead weight from now on */
    static void host_block_released(void *p) { instance().drop(p); }
public:
    static residency &instance() { static residency *r = new residency(); return *r; }   /* never destroyed: outlives every vector */
    void drop(const void *host)
    {
        auto found = by_host_.find(host);
        if (found == by_host_.end()) return;
        bytes_ -= found->second->n * 24;
        lru_.erase(found->second);
        by_host_.erase(found);
    }
    static constexpr std::size_t min_elems = 4096;

    /* park `buf` (n elements) as the device copy of the host array at `host` */
    void put(const void *host, std::size_t n, device_buffer &&buf)
    {
        if (!enabled_ || n < min_elems) return;
        drop(host);                                               /* one entry per address */
        lru_.emplace_front();
        entry &e = lru_.front();
        e.host = host; e.n = n; e.buf = std::move(buf);
        e.sample_bytes = take_sample(host, n, e.sample);
        by_host_[host] = lru_.begin();
        bytes_ += n * 24;
        while (bytes_ > budget_ && lru_.size() > 1) drop(lru_.back().host);
    }

    /* device copy of the n-element host array at `host`, or nullptr */
    const device_buffer *find(const void *host, std::size_t n)
    {
        if (!enabled_ || n < min_elems) return nullptr;
        auto found = by_host_.find(host);
        if (found != by_host_.end()) {
            auto it = found->second;
            unsigned char now[32 * 24];
            const std::size_t nb = take_sample(host, n, now);
            if (it->n != n || nb != it->sample_bytes || std::memcmp(now, it->sample, nb) != 0) {
                drop(host); ++stale_; ++misses_;
                return nullptr;
            }
            if (verify_) {
                std::vector<unsigned char> dev(n * 24);
                throw_status(context(), b200_download(context(), reinterpret_cast<uint64_t *>(dev.data()), it->buf.get(), 0, n));
                if (std::memcmp(dev.data(), host, n * 24) != 0)
                    throw std::logic_error("libiop_b200::residency: device copy differs from the host vector (B200_RESIDENCY_VERIFY)");
            }
            lru_.splice(lru_.begin(), lru_, it);
            ++hits_;
            return &lru_.front().buf;
        }
        ++misses_;
        return nullptr;
    }
    void clear() { lru_.clear(); by_host_.clear(); bytes_ = 0; }
    std::size_t hits() const { return hits_; }
    std::size_t misses() const { return misses_; }
    std::size_t stale() const { return stale_; }
    std::size_t resident_bytes() const { return bytes_; }
};

/* a device view of a host vector: the resident copy if there is one, otherwise a fresh (asynchronous) upload */
class device_view {
    const device_buffer *resident_ = nullptr;
    device_buffer owned_;
public:
    template<typename FieldT> explicit device_view(const std::vector<FieldT> &v)
    {
        resident_ = residency::instance().find(v.data(), v.size());
        if (!resident_) { owned_ = device_buffer(v.size()); owned_.upload_async(v); }
    }
    const device_buffer &buf() const { return resident_ ? *resident_ : owned_; }
    b200_buf get() const { return buf().get(); }
};

/* download `buf` into a fresh host vector and park the buffer as its device copy */
template<typename FieldT>
std::vector<FieldT> download_and_keep(device_buffer &&buf, std::size_t n)
{
    std::vector<FieldT> v = buf.template download<FieldT>(n);
    residency::instance().put(v.data(), n, std::move(buf));
    return v;
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
