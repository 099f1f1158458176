/* ctx.h — internal context shared by the translation units of libb200iop.so */
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200_iop.h"

namespace b200 {
struct AdditivePlan;
struct NttPlan;
}

struct b200_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_upload = nullptr;   /* recorded after every host->device copy (b200_wait_uploads) */
    bool upload_pending = false;
    /* cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remembered per context, not per process */
    bool attr_additive[2] = {false, false};
    bool attr_ntt = false;
    char err[512] = {0};
    /* grow-only device scratch (work arrays of the FFTs), keyed by slot */
    void *scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    /* caching device allocator behind b200_alloc / b200_free: freed blocks are kept in exact-size bins and handed out
       again (a prover allocates and releases the same few work-array sizes hundreds of times; cudaMalloc/cudaFree cost
       milliseconds each and cudaFree synchronises the device).  Reuse is safe without synchronisation because every
       use of a block is ordered on the context's single stream. */
    std::multimap<size_t, void *> free_blocks;
    std::map<void *, size_t> live_blocks;
    size_t cached_bytes = 0;
    /* pinned staging for small host<->device transfers */
    void *staging = nullptr;
    size_t staging_bytes = 0;
    /* cached per-domain plans */
    std::map<std::string, std::shared_ptr<b200::AdditivePlan>> additive_plans;
    std::map<std::string, std::shared_ptr<b200::NttPlan>> ntt_plans;
    /* instrumentation: every kernel launch is counted; with profiling on, tile passes are bracketed by events */
    unsigned long long n_launches = 0;
    unsigned long long h2d_bytes = 0, d2h_bytes = 0;   /* bytes moved by upload / download / gather calls */
    bool trace = false;   /* B200_TRACE=1: every C-ABI compute/transfer call prints its synchronised wall time to stderr */
    bool profiling = false;
    struct ProfRec { std::string label; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
};

namespace b200 {

inline int set_err(b200_ctx *c, int code, const char *fmt, ...)
{
    if (c) {
        va_list ap; va_start(ap, fmt);
        vsnprintf(c->err, sizeof c->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

/* Every C-ABI entry point makes its context's device current first: a process may hold one context per GPU (or share the
   thread with a framework that switches devices), and allocations / launches follow the CURRENT device. */
#define B200_ENTER(ctx) do { if (ctx) cudaSetDevice((ctx)->device); } while (0)

#define B200_CUDA(ctx, call)                                                                      \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return b200::set_err((ctx), B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call,            \
                                 cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

#define B200_LAUNCH_CHECK(ctx)                                                                    \
    do {                                                                                          \
        (ctx)->n_launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess)                                                                   \
            return b200::set_err((ctx), B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)",        \
                                 cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

/* grow-only scratch buffer */
inline void *scratch_get(b200_ctx *c, int slot, size_t bytes)
{
    if (c->scratch_bytes[slot] >= bytes) return c->scratch[slot];
    if (c->scratch[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->scratch[slot]); c->scratch[slot] = nullptr; c->scratch_bytes[slot] = 0; }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    c->scratch[slot] = p; c->scratch_bytes[slot] = bytes;
    return p;
}

inline void *staging_get(b200_ctx *c, size_t bytes)
{
    if (c->staging_bytes >= bytes) return c->staging;
    if (c->staging) { cudaStreamSynchronize(c->stream); cudaFreeHost(c->staging); c->staging = nullptr; c->staging_bytes = 0; }
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    c->staging = p; c->staging_bytes = bytes;
    return p;
}

/* RAII event bracket around one launch when profiling is enabled */
struct ProfScope {
    b200_ctx *c; size_t idx; bool on;
    ProfScope(b200_ctx *ctx, const std::string &label) : c(ctx), idx(0), on(ctx->profiling)
    {
        if (!on) return;
        b200_ctx::ProfRec r; r.label = label;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, c->stream);
        c->prof.push_back(r); idx = c->prof.size() - 1;
    }
    ~ProfScope() { if (on) cudaEventRecord(c->prof[idx].b, c->stream); }
};

/* B200_TRACE=1 scope: prints "b200 <name> <detail> <ms>" with the stream drained on both sides */
struct ApiTrace {
    b200_ctx *c; const char *name; size_t detail; std::chrono::steady_clock::time_point t0;
    ApiTrace(b200_ctx *ctx, const char *n, size_t d) : c(ctx), name(n), detail(d)
    {
        if (c && c->trace) { cudaStreamSynchronize(c->stream); t0 = std::chrono::steady_clock::now(); }
    }
    ~ApiTrace()
    {
        if (c && c->trace) {
            cudaStreamSynchronize(c->stream);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            static const std::chrono::steady_clock::time_point epoch = std::chrono::steady_clock::now();
            const double at = std::chrono::duration<double, std::milli>(t0 - epoch).count();
            fprintf(stderr, "b200 %-28s %12zu %10.3f ms  @%10.3f\n", name, detail, ms, at);
        }
    }
};
#define B200_TRACE_SCOPE(ctx, name, detail) b200::ApiTrace trace_scope__((ctx), (name), (size_t)(detail))

inline uint32_t ceil_log2(size_t n)
{
    uint32_t r = 0;
    while (((size_t)1 << r) < n) ++r;
    return r;
}

/* implemented in the per-subsystem translation units */
int additive_fft(b200_ctx *c, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf out, size_t batch);
int additive_ifft(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf out, size_t batch);
int additive_lde(b200_ctx *c, const b200_domain *H, const b200_domain *L, b200_buf evals, b200_buf out, size_t batch);
int additive_lde_interpolate(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf coeffs_out, size_t batch);
int additive_lde_evaluate(b200_ctx *c, const b200_domain *L, uint32_t d, b200_buf coeffs, b200_buf out, size_t batch);
int gf192_mul_vec(b200_ctx *c, b200_buf out, b200_buf a, b200_buf b, size_t n);
int gf192_inv_vec(b200_ctx *c, b200_buf out, b200_buf a, size_t n);
int fp_ntt(b200_ctx *c, const b200_fp_params *P, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf out);
int fp_intt(b200_ctx *c, const b200_fp_params *P, const b200_domain *H, b200_buf evals, b200_buf out);
int fp_binop_vec(b200_ctx *c, const b200_fp_params *P, int op, b200_buf out, b200_buf a, b200_buf b, size_t n);
int fri_fold(b200_ctx *c, int field, const b200_fp_params *fp, const b200_domain *L, uint32_t eta,
             const uint64_t x[3], b200_buf f, b200_buf out);
int merkle_build(b200_ctx *c, const b200_buf *oracles, size_t n_oracles, size_t n_elems, size_t coset,
                 int kind, const uint8_t *salts, size_t salt_len, size_t digest_len, b200_bytes nodes);
int blake2b_batch(b200_ctx *c, b200_bytes in, size_t msg_len, size_t n_msgs, size_t digest_len, b200_bytes out);
int squeeze_field_elements(b200_ctx *c, int field, const b200_fp_params *fp, const uint8_t *root_plus_index, size_t msg_len,
                           size_t n, b200_buf out);
int pow_solve(b200_ctx *c, const uint8_t *challenge, size_t digest_len, uint32_t pow_bitlen, uint64_t pow_upperbound, uint8_t *pow_out);
void plans_clear(b200_ctx *c);

} // namespace b200
