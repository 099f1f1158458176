/*
 * api.cu — C ABI of libb200iop.so (declared in include/b200_iop.h): context, memory, and thin
 * forwarding to the kernels.  No CPU fallback: every compute entry point launches CUDA kernels.
 */
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.h"

namespace b200 {
void plans_clear(b200_ctx *c) { c->additive_plans.clear(); c->ntt_plans.clear(); }
}

using namespace b200;

/* dst[i * dst_stride] = src[i * src_stride] over 24-byte elements (one thread per element, three u64 words) */
__global__ void copy_strided_kernel(uint64_t *dst, size_t dst_stride, const uint64_t *src, size_t src_stride, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t *q = src + 3 * i * src_stride;
        uint64_t *d = dst + 3 * i * dst_stride;
        d[0] = q[0]; d[1] = q[1]; d[2] = q[2];
    }
}

extern "C" {

const char *b200_version(void) { return "libb200iop 0.1 (sm_100a)"; }

b200_ctx *b200_ctx_create(int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return nullptr; }
    if (device < 0 || device >= count) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    b200_ctx *c = new (std::nothrow) b200_ctx();
    if (!c) return nullptr;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return nullptr; }
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError(); delete c; return nullptr;
    }
    const char *tr = getenv("B200_TRACE");
    c->trace = tr && tr[0] == '1';
    return c;
}

void b200_ctx_destroy(b200_ctx *c)
{
    B200_ENTER(c);
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    plans_clear(c);
    b200_trim(c);
    for (auto &kv : c->live_blocks) cudaFree(kv.first);
    for (int i = 0; i < 4; ++i) if (c->scratch[i]) cudaFree(c->scratch[i]);
    if (c->staging) cudaFreeHost(c->staging);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev_upload);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char *b200_last_error(b200_ctx *c) { B200_ENTER(c); return c ? c->err : "no context (is a CUDA device present?)"; }

int b200_sync(b200_ctx *c)
{
    B200_ENTER(c);
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
void *b200_stream(b200_ctx *c) { B200_ENTER(c); return (void *)c->stream; }

int b200_device_info(b200_ctx *c, int *sm_count, size_t *total_mem, int *cc_major, int *cc_minor)
{
    B200_ENTER(c);
    cudaDeviceProp prop;
    B200_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

int b200_trim(b200_ctx *c)
{
    B200_ENTER(c);
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto &kv : c->free_blocks) cudaFree(kv.second);
    c->free_blocks.clear();
    c->cached_bytes = 0;
    return 0;
}
b200_buf b200_alloc(b200_ctx *c, size_t n_elems) { B200_ENTER(c); return b200_alloc_bytes(c, n_elems * 24); }
void b200_free(b200_ctx *c, b200_buf b) { B200_ENTER(c); b200_free_bytes(c, b); }
b200_bytes b200_alloc_bytes(b200_ctx *c, size_t n_bytes)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "alloc", n_bytes);
    const size_t want = ((n_bytes ? n_bytes : 8) + 511) & ~(size_t)511;
    auto hit = c->free_blocks.find(want);
    if (hit != c->free_blocks.end()) {
        void *p = hit->second;
        c->free_blocks.erase(hit);
        c->cached_bytes -= want;
        c->live_blocks[p] = want;
        return (b200_bytes)(uintptr_t)p;
    }
    void *p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        cudaGetLastError();
        b200_trim(c);                                   /* give the cached blocks back and retry once */
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            set_err(c, B200_ERR_NOMEM, "cudaMalloc of %zu bytes failed", n_bytes);
            return 0;
        }
    }
    c->live_blocks[p] = want;
    return (b200_bytes)(uintptr_t)p;
}
void b200_free_bytes(b200_ctx *c, b200_bytes b)
{
    B200_ENTER(c);
    if (!b) return;
    void *p = (void *)(uintptr_t)b;
    auto it = c->live_blocks.find(p);
    if (it == c->live_blocks.end()) { cudaStreamSynchronize(c->stream); cudaFree(p); return; }   /* not ours: plain release */
    c->free_blocks.emplace(it->second, p);
    c->cached_bytes += it->second;
    c->live_blocks.erase(it);
}
void *b200_host_alloc(b200_ctx *c, size_t n_bytes)
{
    B200_ENTER(c);
    void *p = nullptr;
    if (cudaMallocHost(&p, n_bytes ? n_bytes : 8) != cudaSuccess) {
        cudaGetLastError();
        set_err(c, B200_ERR_NOMEM, "cudaMallocHost of %zu bytes failed", n_bytes);
        return nullptr;
    }
    return p;
}
void b200_host_free(b200_ctx *, void *p) { if (p) cudaFreeHost(p); }

int b200_upload(b200_ctx *c, b200_buf dst, size_t dst_offset, const uint64_t *src, size_t n_elems)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "upload", n_elems * 24);
    c->h2d_bytes += n_elems * 24;
    B200_CUDA(c, cudaMemcpyAsync((char *)(uintptr_t)dst + dst_offset * 24, src, n_elems * 24, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));
    c->upload_pending = true;
    return 0;
}
int b200_wait_uploads(b200_ctx *c)
{
    B200_ENTER(c);
    if (!c->upload_pending) return 0;
    B200_CUDA(c, cudaEventSynchronize(c->ev_upload));
    c->upload_pending = false;
    return 0;
}
int b200_download(b200_ctx *c, uint64_t *dst, b200_buf src, size_t src_offset, size_t n_elems)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "download", n_elems * 24);
    c->d2h_bytes += n_elems * 24;
    B200_CUDA(c, cudaMemcpyAsync(dst, (const char *)(uintptr_t)src + src_offset * 24, n_elems * 24, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_upload_bytes(b200_ctx *c, b200_bytes dst, const uint8_t *src, size_t n_bytes)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "upload_bytes", n_bytes);
    c->h2d_bytes += n_bytes;
    B200_CUDA(c, cudaMemcpyAsync((void *)(uintptr_t)dst, src, n_bytes, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaEventRecord(c->ev_upload, c->stream));
    c->upload_pending = true;
    return 0;
}
int b200_download_bytes(b200_ctx *c, uint8_t *dst, b200_bytes src, size_t offset, size_t n_bytes)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "download_bytes", n_bytes);
    c->d2h_bytes += n_bytes;
    B200_CUDA(c, cudaMemcpyAsync(dst, (const char *)(uintptr_t)src + offset, n_bytes, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_copy(b200_ctx *c, b200_buf dst, size_t dst_offset, b200_buf src, size_t src_offset, size_t n_elems)
{
    B200_ENTER(c);
    B200_CUDA(c, cudaMemcpyAsync((char *)(uintptr_t)dst + dst_offset * 24, (const char *)(uintptr_t)src + src_offset * 24,
                                 n_elems * 24, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
int b200_copy_strided(b200_ctx *c, b200_buf dst, size_t dst_stride, b200_buf src, size_t src_offset, size_t src_stride, size_t n_elems)
{
    B200_ENTER(c);
    if (!n_elems) return 0;
    const unsigned grid = (unsigned)((n_elems + 255) / 256 < 148 * 8 ? (n_elems + 255) / 256 : 148 * 8);
    copy_strided_kernel<<<grid, 256, 0, c->stream>>>((uint64_t *)(uintptr_t)dst, dst_stride, (const uint64_t *)(uintptr_t)src + 3 * src_offset,
                                                     src_stride, n_elems);
    B200_LAUNCH_CHECK(c);
    return 0;
}
int b200_fill_zero(b200_ctx *c, b200_buf dst, size_t dst_offset, size_t n_elems)
{
    B200_ENTER(c);
    B200_CUDA(c, cudaMemsetAsync((char *)(uintptr_t)dst + dst_offset * 24, 0, n_elems * 24, c->stream));
    return 0;
}

/* out[k] = table[map ? map[idx[k]] : idx[k]]; out[map[c]] = x[c]; dst[i] = v — 24-byte elements as three u64 words */
__global__ void gather_dev_kernel(uint64_t *out, const uint64_t *table, const uint32_t *idx, const uint32_t *map, size_t n)
{
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        size_t j = idx[k];
        if (map) j = map[j];
        const uint64_t *q = table + 3 * j;
        uint64_t *d = out + 3 * k;
        d[0] = q[0]; d[1] = q[1]; d[2] = q[2];
    }
}
__global__ void scatter_dev_kernel(uint64_t *out, const uint64_t *x, const uint32_t *map, size_t n)
{
    for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
        const uint64_t *q = x + 3 * c;
        uint64_t *d = out + 3 * (size_t)map[c];
        d[0] = q[0]; d[1] = q[1]; d[2] = q[2];
    }
}
__global__ void fill_kernel(uint64_t *dst, uint64_t v0, uint64_t v1, uint64_t v2, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t *d = dst + 3 * i;
        d[0] = v0; d[1] = v1; d[2] = v2;
    }
}

__global__ void gather_kernel(uint64_t *out, const uint64_t *src, const size_t *idx, size_t k)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t < k * 3) out[t] = src[idx[t / 3] * 3 + t % 3];
}

/* ---- transposition of an entry list by counting sort: histogram, exclusive scan, fill through per-column cursors ---- */
__global__ void csr_col_histogram_kernel(const uint32_t *col, size_t nnz, unsigned int *count)
{
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nnz; k += (size_t)gridDim.x * blockDim.x) atomicAdd(&count[col[k]], 1u);
}
/* one CTA: col_ptr[c] = sum_{c' < c} count[c'] for c <= n, and count[c] becomes the 32-bit cursor of column c */
__global__ void __launch_bounds__(1024) csr_scan_kernel(unsigned int *count, unsigned long long *col_ptr, size_t n)
{
    __shared__ unsigned long long part[1024];
    const size_t chunk = (n + blockDim.x - 1) / blockDim.x;
    const size_t lo = threadIdx.x * chunk < n ? threadIdx.x * chunk : n, hi = lo + chunk < n ? lo + chunk : n;
    unsigned long long sum = 0;
    for (size_t c = lo; c < hi; ++c) sum += count[c];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (unsigned t = 0; t < blockDim.x; ++t) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        col_ptr[n] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (size_t c = lo; c < hi; ++c) { const unsigned int v = count[c]; col_ptr[c] = run; count[c] = (unsigned int)run; run += v; }
}
__global__ void csr_transpose_fill_kernel(const uint32_t *col, const uint32_t *row, const uint64_t *coeff, size_t nnz, unsigned int *cursor,
                                          uint32_t *src_out, uint64_t *coeff_out)
{
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nnz; k += (size_t)gridDim.x * blockDim.x) {
        const size_t at = atomicAdd(&cursor[col[k]], 1u);
        src_out[at] = row[k];
        const uint64_t *q = coeff + 3 * k;
        uint64_t *d = coeff_out + 3 * at;
        d[0] = q[0]; d[1] = q[1]; d[2] = q[2];
    }
}

static unsigned copy_grid(size_t n) { return (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8); }
int b200_gather_dev(b200_ctx *c, b200_buf out, b200_buf table, b200_bytes idx, b200_bytes map, size_t n)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gather_dev", n);
    if (!n) return 0;
    gather_dev_kernel<<<copy_grid(n), 256, 0, c->stream>>>((uint64_t *)(uintptr_t)out, (const uint64_t *)(uintptr_t)table,
                                                          (const uint32_t *)(uintptr_t)idx, (const uint32_t *)(uintptr_t)map, n);
    B200_LAUNCH_CHECK(c);
    return 0;
}
int b200_scatter_dev(b200_ctx *c, b200_buf out, b200_buf x, b200_bytes map, size_t n)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "scatter_dev", n);
    if (!n) return 0;
    if (!map) return set_err(c, B200_ERR_INVALID, "b200_scatter_dev needs an index map");
    scatter_dev_kernel<<<copy_grid(n), 256, 0, c->stream>>>((uint64_t *)(uintptr_t)out, (const uint64_t *)(uintptr_t)x, (const uint32_t *)(uintptr_t)map, n);
    B200_LAUNCH_CHECK(c);
    return 0;
}
int b200_fill(b200_ctx *c, b200_buf dst, size_t dst_offset, const uint64_t value[3], size_t n)
{
    B200_ENTER(c);
    if (!n) return 0;
    fill_kernel<<<copy_grid(n), 256, 0, c->stream>>>((uint64_t *)(uintptr_t)dst + 3 * dst_offset, value[0], value[1], value[2], n);
    B200_LAUNCH_CHECK(c);
    return 0;
}

__global__ void csr_row_indices_kernel(const unsigned long long *row_ptr, uint32_t *row_out, size_t n_rows)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_rows; i += (size_t)gridDim.x * blockDim.x)
        for (unsigned long long k = row_ptr[i], e = row_ptr[i + 1]; k < e; ++k) row_out[k] = (uint32_t)i;
}
int b200_csr_row_indices_dev(b200_ctx *c, size_t n_rows, b200_bytes row_ptr, b200_bytes row_out)
{
    B200_ENTER(c);
    if (!n_rows) return 0;
    if (n_rows >= ((size_t)1 << 32)) return set_err(c, B200_ERR_INVALID, "b200_csr_row_indices_dev: 32-bit row indices");
    csr_row_indices_kernel<<<copy_grid(n_rows), 256, 0, c->stream>>>((const unsigned long long *)(uintptr_t)row_ptr, (uint32_t *)(uintptr_t)row_out, n_rows);
    B200_LAUNCH_CHECK(c);
    return 0;
}
int b200_csr_transpose_dev(b200_ctx *c, size_t n_cols, size_t nnz, b200_bytes col, b200_bytes row, b200_buf coeff,
                           b200_bytes col_ptr_out, b200_bytes src_out, b200_buf coeff_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "csr_transpose_dev", nnz);
    if (nnz >= ((size_t)1 << 32) || n_cols >= ((size_t)1 << 32)) return set_err(c, B200_ERR_INVALID, "b200_csr_transpose_dev: 32-bit indices");
    unsigned int *count = (unsigned int *)scratch_get(c, 3, (n_cols + 1) * sizeof(unsigned int));
    if (!count) return set_err(c, B200_ERR_NOMEM, "cannot allocate the column counters");
    B200_CUDA(c, cudaMemsetAsync(count, 0, (n_cols + 1) * sizeof(unsigned int), c->stream));
    if (nnz) {
        csr_col_histogram_kernel<<<copy_grid(nnz), 256, 0, c->stream>>>((const uint32_t *)(uintptr_t)col, nnz, count);
        B200_LAUNCH_CHECK(c);
    }
    csr_scan_kernel<<<1, 1024, 0, c->stream>>>(count, (unsigned long long *)(uintptr_t)col_ptr_out, n_cols);
    B200_LAUNCH_CHECK(c);
    if (nnz) {
        csr_transpose_fill_kernel<<<copy_grid(nnz), 256, 0, c->stream>>>((const uint32_t *)(uintptr_t)col, (const uint32_t *)(uintptr_t)row,
                                                                        (const uint64_t *)(uintptr_t)coeff, nnz, count,
                                                                        (uint32_t *)(uintptr_t)src_out, (uint64_t *)(uintptr_t)coeff_out);
        B200_LAUNCH_CHECK(c);
    }
    return 0;
}

int b200_gather(b200_ctx *c, uint64_t *dst, b200_buf src, const size_t *idx, size_t k)
{
    B200_ENTER(c);
    if (!k) return 0;
    const size_t bytes = k * sizeof(size_t) + k * 24;
    char *dev = (char *)scratch_get(c, 3, bytes);
    if (!dev) return set_err(c, B200_ERR_NOMEM, "cannot allocate gather buffer");
    size_t *didx = (size_t *)dev;
    uint64_t *dout = (uint64_t *)(dev + k * sizeof(size_t));
    B200_CUDA(c, cudaMemcpyAsync(didx, idx, k * sizeof(size_t), cudaMemcpyHostToDevice, c->stream));
    gather_kernel<<<(unsigned)((k * 3 + 255) / 256), 256, 0, c->stream>>>(dout, (const uint64_t *)(uintptr_t)src, didx, k);
    B200_LAUNCH_CHECK(c);
    B200_CUDA(c, cudaMemcpyAsync(dst, dout, k * 24, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d_bytes += k * sizeof(size_t); c->d2h_bytes += k * 24;
    return 0;
}

__global__ void gather_multi_kernel(uint64_t *out, const uint64_t *const *srcs, const size_t *idx, size_t n_srcs, size_t k)
{
    const size_t total = n_srcs * k * 3;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t e = t / 3, s = e / k, j = e % k;
        out[t] = srcs[s][idx[j] * 3 + t % 3];
    }
}
int b200_gather_multi(b200_ctx *c, uint64_t *dst, const b200_buf *srcs, size_t n_srcs, const size_t *idx, size_t k)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gather_multi", n_srcs * k);
    if (!k || !n_srcs) return 0;
    const size_t b_src = n_srcs * sizeof(uint64_t), b_idx = k * sizeof(size_t), b_out = n_srcs * k * 24;
    char *dev = (char *)scratch_get(c, 3, b_src + b_idx + b_out);
    if (!dev) return set_err(c, B200_ERR_NOMEM, "cannot allocate gather buffer");
    const uint64_t **dsrc = (const uint64_t **)dev;
    size_t *didx = (size_t *)(dev + b_src);
    uint64_t *dout = (uint64_t *)(dev + b_src + b_idx);
    B200_CUDA(c, cudaMemcpyAsync(dsrc, srcs, b_src, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaMemcpyAsync(didx, idx, b_idx, cudaMemcpyHostToDevice, c->stream));
    gather_multi_kernel<<<copy_grid(n_srcs * k * 3), 256, 0, c->stream>>>(dout, dsrc, didx, n_srcs, k);
    B200_LAUNCH_CHECK(c);
    B200_CUDA(c, cudaMemcpyAsync(dst, dout, b_out, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d_bytes += b_src + b_idx; c->d2h_bytes += b_out;
    return 0;
}

/* ---- instrumentation ---- */
unsigned long long b200_launch_count(b200_ctx *c) { B200_ENTER(c); return c->n_launches; }
void b200_transfer_bytes(b200_ctx *c, unsigned long long *h2d, unsigned long long *d2h)
{
    B200_ENTER(c);
    if (h2d) *h2d = c->h2d_bytes;
    if (d2h) *d2h = c->d2h_bytes;
}
int b200_profile_begin(b200_ctx *c)
{
    B200_ENTER(c);
    for (auto &r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    c->prof.clear();
    c->profiling = true;
    return 0;
}
/* writes "label\tms\n" lines into buf (truncated to cap); returns the number of records */
int b200_profile_end(b200_ctx *c, char *buf, size_t cap)
{
    B200_ENTER(c);
    c->profiling = false;
    cudaStreamSynchronize(c->stream);
    size_t off = 0;
    int n = 0;
    for (auto &r : c->prof) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        if (buf && off < cap) off += (size_t)snprintf(buf + off, cap - off, "%s\t%.6f\n", r.label.c_str(), ms);
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        ++n;
    }
    c->prof.clear();
    return n;
}

int b200_timer_start(b200_ctx *c) { B200_ENTER(c); B200_CUDA(c, cudaEventRecord(c->ev0, c->stream)); return 0; }
int b200_timer_stop(b200_ctx *c, float *ms)
{
    B200_ENTER(c);
    B200_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    B200_CUDA(c, cudaEventSynchronize(c->ev1));
    B200_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}

int b200_gf192_fft(b200_ctx *c, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf evals_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_fft", (size_t)1 << L->dim);
    return additive_fft(c, L, coeffs, n_coeffs, evals_out, 1);
}
int b200_gf192_ifft(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf coeffs_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_ifft", (size_t)1 << H->dim);
    return additive_ifft(c, H, evals, coeffs_out, 1);
}
int b200_gf192_fft_batch(b200_ctx *c, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, size_t batch, b200_buf evals_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_fft_batch", batch << L->dim);
    return additive_fft(c, L, coeffs, n_coeffs, evals_out, batch);
}
int b200_gf192_ifft_batch(b200_ctx *c, const b200_domain *H, b200_buf evals, size_t batch, b200_buf coeffs_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_ifft_batch", batch << H->dim);
    return additive_ifft(c, H, evals, coeffs_out, batch);
}
int b200_gf192_lde(b200_ctx *c, const b200_domain *H, const b200_domain *L, b200_buf evals, b200_buf evals_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_lde", (size_t)1 << (L ? L->dim : 0));
    return additive_lde(c, H, L, evals, evals_out, 1);
}
int b200_gf192_lde_batch(b200_ctx *c, const b200_domain *H, const b200_domain *L, b200_buf evals, size_t batch, b200_buf evals_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_lde_batch", batch << (L ? L->dim : 0));
    return additive_lde(c, H, L, evals, evals_out, batch);
}
int b200_gf192_lde_interpolate(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf coeffs_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_lde_interpolate", (size_t)1 << (H ? H->dim : 0));
    return additive_lde_interpolate(c, H, evals, coeffs_out, 1);
}
int b200_gf192_lde_evaluate(b200_ctx *c, const b200_domain *L, uint32_t coeff_dim, b200_buf coeffs, b200_buf evals_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_lde_evaluate", (size_t)1 << (L ? L->dim : 0));
    return additive_lde_evaluate(c, L, coeff_dim, coeffs, evals_out, 1);
}
int b200_gf192_mul(b200_ctx *c, b200_buf out, b200_buf a, b200_buf b, size_t n) { B200_ENTER(c); return gf192_mul_vec(c, out, a, b, n); }
int b200_gf192_inv(b200_ctx *c, b200_buf out, b200_buf a, size_t n)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "gf192_inv", n);
    return gf192_inv_vec(c, out, a, n);
}

int b200_fp_ntt(b200_ctx *c, const b200_fp_params *P, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf evals_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "fp_ntt", (size_t)1 << L->dim);
    return fp_ntt(c, P, L, coeffs, n_coeffs, evals_out);
}
int b200_fp_intt(b200_ctx *c, const b200_fp_params *P, const b200_domain *H, b200_buf evals, b200_buf coeffs_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "fp_intt", (size_t)1 << H->dim);
    return fp_intt(c, P, H, evals, coeffs_out);
}
int b200_fp_mul(b200_ctx *c, const b200_fp_params *P, b200_buf out, b200_buf a, b200_buf b, size_t n) { B200_ENTER(c); return fp_binop_vec(c, P, 0, out, a, b, n); }
int b200_fp_add(b200_ctx *c, const b200_fp_params *P, b200_buf out, b200_buf a, b200_buf b, size_t n) { B200_ENTER(c); return fp_binop_vec(c, P, 1, out, a, b, n); }
int b200_fp_sub(b200_ctx *c, const b200_fp_params *P, b200_buf out, b200_buf a, b200_buf b, size_t n) { B200_ENTER(c); return fp_binop_vec(c, P, 2, out, a, b, n); }
int b200_fp_inv(b200_ctx *c, const b200_fp_params *P, b200_buf out, b200_buf a, size_t n) { B200_ENTER(c); return fp_binop_vec(c, P, 3, out, a, 0, n); }

int b200_fri_fold(b200_ctx *c, int field, const b200_fp_params *fp, const b200_domain *L_i, uint32_t eta,
                  const uint64_t x_i[3], b200_buf f_i, b200_buf f_next_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "fri_fold", (size_t)1 << L_i->dim);
    return fri_fold(c, field, fp, L_i, eta, x_i, f_i, f_next_out);
}

int b200_merkle_build(b200_ctx *c, const b200_buf *oracles, size_t n_oracles, size_t n_elems, size_t coset_size,
                      int domain_kind, const uint8_t *zk_salts_host, size_t salt_len, size_t digest_len, b200_bytes nodes_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "merkle_build", n_elems * n_oracles);
    return merkle_build(c, oracles, n_oracles, n_elems, coset_size, domain_kind, zk_salts_host, salt_len, digest_len, nodes_out);
}
int b200_merkle_root(b200_ctx *c, b200_bytes nodes, uint8_t *root_out, size_t digest_len)
{
    B200_ENTER(c);
    return b200_download_bytes(c, root_out, nodes, 0, digest_len);
}

__global__ void node_gather_kernel(uint64_t *out, const uint64_t *nodes, const size_t *idx, size_t k, uint32_t dw)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t < k * dw) out[t] = nodes[idx[t / dw] * dw + t % dw];
}

int b200_merkle_nodes(b200_ctx *c, b200_bytes nodes, const size_t *node_idx, size_t k, size_t digest_len, uint8_t *out)
{
    B200_ENTER(c);
    if (!k) return 0;
    if (digest_len % 8) return set_err(c, B200_ERR_INVALID, "digest length must be a multiple of 8");
    const uint32_t dw = (uint32_t)(digest_len / 8);
    char *dev = (char *)scratch_get(c, 3, k * sizeof(size_t) + k * digest_len);
    if (!dev) return set_err(c, B200_ERR_NOMEM, "cannot allocate gather buffer");
    size_t *didx = (size_t *)dev;
    uint64_t *dout = (uint64_t *)(dev + k * sizeof(size_t));
    B200_CUDA(c, cudaMemcpyAsync(didx, node_idx, k * sizeof(size_t), cudaMemcpyHostToDevice, c->stream));
    node_gather_kernel<<<(unsigned)((k * dw + 255) / 256), 256, 0, c->stream>>>(dout, (const uint64_t *)(uintptr_t)nodes, didx, k, dw);
    B200_LAUNCH_CHECK(c);
    B200_CUDA(c, cudaMemcpyAsync(out, dout, k * digest_len, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int b200_pow_solve(b200_ctx *c, const uint8_t *challenge, size_t digest_len, uint32_t pow_bitlen, uint64_t pow_upperbound, uint8_t *pow_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "pow_solve", pow_bitlen);
    return pow_solve(c, challenge, digest_len, pow_bitlen, pow_upperbound, pow_out);
}

int b200_squeeze_field_elements(b200_ctx *c, int field, const b200_fp_params *fp, const uint8_t *root_plus_index, size_t msg_len,
                                size_t n, b200_buf out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "squeeze_field_elements", n);
    return squeeze_field_elements(c, field, fp, root_plus_index, msg_len, n, out);
}

int b200_blake2b_batch(b200_ctx *c, b200_bytes in, size_t msg_len, size_t n_msgs, size_t digest_len, b200_bytes out)
{
    B200_ENTER(c);
    return blake2b_batch(c, in, msg_len, n_msgs, digest_len, out);
}

} /* extern "C" */
