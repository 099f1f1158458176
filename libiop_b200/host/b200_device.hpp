/*
 * b200_device.hpp — process-wide device context + RAII device buffers for the C++ adapter.
 * The adapter is header-only C++14 (like the reference) over the C ABI of include/b200_iop.h.
 * Errors of the C ABI are rethrown as the exception types the reference uses at the same place:
 * B200_ERR_INVALID -> std::invalid_argument (fft.tcc:333,368; fri_aux.tcc:33),
 * B200_ERR_LOGIC   -> std::logic_error     (merkle_tree.tcc:98-108), anything else std::runtime_error.
 */
#ifndef LIBIOP_B200_DEVICE_HPP_
#define LIBIOP_B200_DEVICE_HPP_

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200_iop.h"
#include "b200_pinned_arena.hpp"

namespace libiop_b200 {

inline void throw_status(b200_ctx *ctx, int rc)
{
    if (rc == B200_OK) return;
    const std::string msg = b200_last_error(ctx);
    if (rc == B200_ERR_INVALID) throw std::invalid_argument(msg);
    if (rc == B200_ERR_LOGIC) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}

/* one context per process (the reference is single-threaded, SURVEY.md §8b "Threading") */
inline b200_ctx *make_context()
{
    /* device ordinal: B200_DEVICE, else LOCAL_RANK (one process per GPU under torchrun / mpirun), else 0 */
    const char *d = std::getenv("B200_DEVICE");
    if (!d) d = std::getenv("LOCAL_RANK");
    b200_ctx *ctx = b200_ctx_create(d ? std::atoi(d) : 0);
    if (!ctx) throw std::runtime_error("libiop_b200: no CUDA device (the hot path has no CPU fallback)");
    return ctx;
}
inline b200_ctx *context()
{
    static b200_ctx *ctx = make_context();       /* C++11 one-time initialisation: safe against concurrent first calls */
    return ctx;
}

class device_buffer {
    b200_buf buf_ = 0;
    std::size_t n_ = 0;
    bool owned_ = true;
public:
    device_buffer() = default;
    /* a non-owning handle on n elements of somebody else's buffer (device_view of a resident vector) */
    static device_buffer alias(b200_buf b, std::size_t n) { device_buffer d; d.buf_ = b; d.n_ = n; d.owned_ = false; return d; }
    explicit device_buffer(std::size_t n_elems) : n_(n_elems)
    {
        buf_ = b200_alloc(context(), n_elems ? n_elems : 1);
        if (!buf_) throw std::runtime_error(b200_last_error(context()));
    }
    device_buffer(const device_buffer &) = delete;
    device_buffer &operator=(const device_buffer &) = delete;
    device_buffer(device_buffer &&o) noexcept : buf_(o.buf_), n_(o.n_), owned_(o.owned_) { o.buf_ = 0; o.n_ = 0; o.owned_ = true; }
    device_buffer &operator=(device_buffer &&o) noexcept
    {
        if (this != &o) { release(); buf_ = o.buf_; n_ = o.n_; owned_ = o.owned_; o.buf_ = 0; o.n_ = 0; o.owned_ = true; }
        return *this;
    }
    ~device_buffer() { release(); }
    void release() { if (buf_ && owned_) b200_free(context(), buf_); buf_ = 0; owned_ = true; }
    bool owns() const { return owned_; }
    b200_buf get() const { return buf_; }
    std::size_t size() const { return n_; }

    /* FieldT must be a 24-byte trivially copyable element (gf192: gf192.hpp:98-99, Fp_model<3>: fp.hpp:42) */
    template<typename FieldT> void upload(const std::vector<FieldT> &v, std::size_t offset = 0)
    {
        static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
        if (!v.empty())
            throw_status(context(), b200_upload(context(), buf_, offset, reinterpret_cast<const uint64_t *>(v.data()), v.size()));
        throw_status(context(), b200_sync(context()));
    }
    /* stream-ordered upload without the trailing synchronisation: `v` must stay alive and unmodified until the next
       synchronising call on the context (download, b200_sync) */
    template<typename FieldT> void upload_async(const std::vector<FieldT> &v, std::size_t offset = 0)
    {
        static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
        if (!v.empty())
            throw_status(context(), b200_upload(context(), buf_, offset, reinterpret_cast<const uint64_t *>(v.data()), v.size()));
    }
    template<typename FieldT> std::vector<FieldT> download(std::size_t n, std::size_t offset = 0) const
    {
        static_assert(sizeof(FieldT) == 24, "field elements must be 3 x u64");
        std::vector<FieldT> v = make_uninitialized_vector<FieldT>(n);   /* every element is overwritten below */
        if (n) throw_status(context(), b200_download(context(), reinterpret_cast<uint64_t *>(v.data()), buf_, offset, n));
        return v;
    }
};

} // namespace libiop_b200
#endif
