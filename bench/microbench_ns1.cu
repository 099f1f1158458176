/*
 * microbench_ns1.cu — the kernel techniques BASELINE.json's north star names, measured against the shipped design on the
 * shipped arithmetic (libiop_b200/csrc/gf192.cuh), so that keeping or dropping each one is a measurement, not an argument:
 *
 *   stage64   tile staged with 64-bit loads / stores, all butterfly stages through shared memory   (round-1 kernel shape)
 *   stage128  the same with 128-bit loads / stores (LDG.E.128 / STG.E.128)                          (shipped in round 2)
 *   tma       tile staged by ONE thread with cp.async.bulk (TMA bulk copy, UBLKCP in SASS) + mbarrier, written back with
 *             cp.async.bulk.global.shared; the tile stays in the array's own AoS layout in shared memory
 *   shfl      the six innermost butterfly stages (partner distance < 64) executed in registers with warp shuffles: a lane
 *             holds two elements, one __shfl_xor per stage re-pairs them; the outer stages go through shared memory
 *
 * Work per CTA = what additive_tile_pass does in its dominant pass: a tile of 4096 elements of GF(2^192), 8 butterfly
 * stages (a' = a + tw*b, b' = b + a'), twiddle = XOR of table entries selected by the butterfly's index bits.  All four
 * kernels compute the same function; the program checks that their outputs are identical and prints the time of each
 * over 2^23 elements (2048 CTAs), CUDA events, best of 5.
 *
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o microbench_ns1 microbench_ns1.cu
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../libiop_b200/csrc/gf192.cuh"

using namespace b200;

#define TILE_BITS 12
#define TILE (1u << TILE_BITS)
#define THREADS 256
#define STAGES 8
#define PAD 4

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ gf192 from_u64(uint64_t a, uint64_t b, uint64_t c)
{
    gf192 r;
    r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32); r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32); r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
    return r;
}
__device__ __forceinline__ uint64_t lo64(const gf192 &x, int i) { return (uint64_t)x.w[2 * i] | ((uint64_t)x.w[2 * i + 1] << 32); }

/* twiddle of the butterfly at in-tile index L0 (bit `pos` clear) in stage `pos`: table[pos][low bits of L0] ^ tau */
__device__ __forceinline__ gf192 twiddle(const gf192 *tab, uint32_t pos, uint32_t L0, const gf192 &tau)
{
    const uint32_t tix = L0 & ((1u << pos) - 1u);
    gf192 tw = tau;
    const uint64_t *q = reinterpret_cast<const uint64_t *>(tab + pos * 256 + (tix & 255u));
    const gf192 e = from_u64(__ldg(q), __ldg(q + 1), __ldg(q + 2));
    gf_addto(tw, e);
    return tw;
}
__device__ __forceinline__ void butterfly(gf192 &a, gf192 &b, const gf192 &tw)
{
    gf_addto(a, gf_mul(b, tw));
    gf_addto(b, a);
}

/* ---- SoA tile in shared memory (three u64 planes), the shipped layout ---- */
struct SoA {
    uint64_t *p0, *p1, *p2;
    __device__ __forceinline__ gf192 ld(uint32_t L) const { return from_u64(p0[L], p1[L], p2[L]); }
    __device__ __forceinline__ void st(uint32_t L, const gf192 &x) const { p0[L] = lo64(x, 0); p1[L] = lo64(x, 1); p2[L] = lo64(x, 2); }
};
/* ---- AoS tile (the array's own layout: what a bulk copy delivers) ---- */
struct AoS {
    uint64_t *p;
    __device__ __forceinline__ gf192 ld(uint32_t L) const { return from_u64(p[3 * L], p[3 * L + 1], p[3 * L + 2]); }
    __device__ __forceinline__ void st(uint32_t L, const gf192 &x) const { p[3 * L] = lo64(x, 0); p[3 * L + 1] = lo64(x, 1); p[3 * L + 2] = lo64(x, 2); }
};

template <typename TileT>
__device__ __forceinline__ void smem_stages(const TileT &t, const gf192 *tab, const gf192 &tau, uint32_t first, uint32_t last)
{
    for (uint32_t pos = first; pos < last; ++pos) {
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < TILE / 2; k += THREADS) {
            const uint32_t low = k & ((1u << pos) - 1u);
            const uint32_t L0 = ((k >> pos) << (pos + 1)) | low, L1 = L0 | (1u << pos);
            gf192 a = t.ld(L0), b = t.ld(L1);
            butterfly(a, b, twiddle(tab, pos, L0, tau));
            t.st(L0, a); t.st(L1, b);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 2) k_stage64(uint64_t *data, const gf192 *tab, gf192 tau)
{
    extern __shared__ __align__(128) uint64_t smem[];
    SoA t{smem, smem + TILE + PAD, smem + 2 * (TILE + PAD)};
    uint64_t *g = data + (size_t)blockIdx.x * TILE * 3;
    for (uint32_t L = threadIdx.x; L < TILE; L += THREADS) { t.p0[L] = g[3 * L]; t.p1[L] = g[3 * L + 1]; t.p2[L] = g[3 * L + 2]; }
    smem_stages(t, tab, tau, 0, STAGES);
    for (uint32_t L = threadIdx.x; L < TILE; L += THREADS) { g[3 * L] = t.p0[L]; g[3 * L + 1] = t.p1[L]; g[3 * L + 2] = t.p2[L]; }
}

__device__ __forceinline__ void load128(const SoA &t, const uint64_t *g)
{
    for (uint32_t L = 2 * threadIdx.x; L < TILE; L += 2 * THREADS) {
        const uint4 *q = reinterpret_cast<const uint4 *>(g + 3 * L);
        const uint4 x = q[0], y = q[1], z = q[2];
        *reinterpret_cast<uint4 *>(t.p0 + L) = make_uint4(x.x, x.y, y.z, y.w);
        *reinterpret_cast<uint4 *>(t.p1 + L) = make_uint4(x.z, x.w, z.x, z.y);
        *reinterpret_cast<uint4 *>(t.p2 + L) = make_uint4(y.x, y.y, z.z, z.w);
    }
}
__device__ __forceinline__ void store128(const SoA &t, uint64_t *g)
{
    for (uint32_t L = 2 * threadIdx.x; L < TILE; L += 2 * THREADS) {
        const uint4 a = *reinterpret_cast<const uint4 *>(t.p0 + L), b = *reinterpret_cast<const uint4 *>(t.p1 + L), c = *reinterpret_cast<const uint4 *>(t.p2 + L);
        uint4 *q = reinterpret_cast<uint4 *>(g + 3 * L);
        q[0] = make_uint4(a.x, a.y, b.x, b.y); q[1] = make_uint4(c.x, c.y, a.z, a.w); q[2] = make_uint4(b.z, b.w, c.z, c.w);
    }
}
__global__ void __launch_bounds__(THREADS, 2) k_stage128(uint64_t *data, const gf192 *tab, gf192 tau)
{
    extern __shared__ __align__(128) uint64_t smem[];
    SoA t{smem, smem + TILE + PAD, smem + 2 * (TILE + PAD)};
    uint64_t *g = data + (size_t)blockIdx.x * TILE * 3;
    load128(t, g);
    smem_stages(t, tab, tau, 0, STAGES);
    store128(t, g);
}

/* ---- TMA bulk copy (cp.async.bulk, non-tensor): one thread moves the whole 96 KiB tile in four 24 KiB copies ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(THREADS, 2) k_tma(uint64_t *data, const gf192 *tab, gf192 tau)
{
    extern __shared__ __align__(128) uint64_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    AoS t{smem};
    uint64_t *g = data + (size_t)blockIdx.x * TILE * 3;
    const uint32_t bytes = TILE * 24, chunk = bytes / 4;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
        for (uint32_t c = 0; c < 4; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(reinterpret_cast<char *>(smem) + c * chunk)), "l"(reinterpret_cast<char *>(g) + c * chunk), "r"(chunk),
                           "r"(smem_u32(&mbar)) : "memory");
    }
    {   /* every thread waits for phase 0 of the barrier */
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    smem_stages(t, tab, tau, 0, STAGES);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* generic-proxy writes visible to the async proxy */
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t c = 0; c < 4; ++c)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(reinterpret_cast<char *>(g) + c * chunk), "r"(smem_u32(reinterpret_cast<char *>(smem) + c * chunk)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

/* ---- warp-shuffle inner stages: a lane holds two elements of a 64-element block, indices l and l + 32 ---- */
__device__ __forceinline__ gf192 shfl_xor_elem(const gf192 &x, int mask)
{
    gf192 r;
#pragma unroll
    for (int i = 0; i < 6; ++i) r.w[i] = __shfl_xor_sync(0xffffffffu, x.w[i], mask);
    return r;
}
__global__ void __launch_bounds__(THREADS, 2) k_shfl(uint64_t *data, const gf192 *tab, gf192 tau)
{
    extern __shared__ __align__(128) uint64_t smem[];
    SoA t{smem, smem + TILE + PAD, smem + 2 * (TILE + PAD)};
    uint64_t *g = data + (size_t)blockIdx.x * TILE * 3;
    load128(t, g);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    /* stages 0..5 inside 64-element blocks: each warp owns TILE / 64 / 8 blocks */
    for (uint32_t blk = warp; blk < TILE / 64; blk += THREADS / 32) {
        const uint32_t base = blk * 64;
        gf192 A = t.ld(base + lane), B = t.ld(base + lane + 32);          /* indices differ in bit 5 */
#pragma unroll
        for (uint32_t pos = 0; pos < 5; ++pos) {
            /* re-pair on bit `pos`: lanes with that bit clear keep A and receive the partner's A; the others keep B */
            const bool upper = (lane >> pos) & 1u;
            const gf192 recv = shfl_xor_elem(upper ? A : B, 1 << pos);
            gf192 x0 = upper ? recv : A, x1 = upper ? B : recv;           /* x0: bit pos clear, x1: bit pos set */
            const uint32_t L0 = base + ((lane & ~(1u << pos)) | (upper ? 32u : 0u));
            butterfly(x0, x1, twiddle(tab, pos, L0, tau));
            /* back to the canonical pairing */
            const gf192 back = shfl_xor_elem(upper ? x0 : x1, 1 << pos);
            A = upper ? back : x0; B = upper ? x1 : back;
        }
        butterfly(A, B, twiddle(tab, 5, base + lane, tau));               /* stage 5: the pair the lane already holds */
        t.st(base + lane, A); t.st(base + lane + 32, B);
    }
    smem_stages(t, tab, tau, 6, STAGES);
    store128(t, g);
}

int main(int argc, char **argv)
{
    const size_t n = (size_t)1 << 23, n_tiles = n / TILE;
    std::vector<uint64_t> h(n * 3);
    uint64_t s = 88172645463325252ull;
    for (auto &v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = s; }
    std::vector<uint64_t> htab(STAGES * 256 * 3);
    for (auto &v : htab) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = s; }
    gf192 tau; for (int i = 0; i < 6; ++i) tau.w[i] = 0x9e3779b9u * (i + 1);
    uint64_t *d_in, *d_work; gf192 *d_tab;
    CK(cudaMalloc(&d_in, n * 24)); CK(cudaMalloc(&d_work, n * 24)); CK(cudaMalloc(&d_tab, htab.size() * 8));
    CK(cudaMemcpy(d_in, h.data(), n * 24, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tab, htab.data(), htab.size() * 8, cudaMemcpyHostToDevice));
    const size_t smem_soa = 3 * (TILE + PAD) * 8, smem_aos = TILE * 24;
    CK(cudaFuncSetAttribute(k_stage64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_soa));
    CK(cudaFuncSetAttribute(k_stage128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_soa));
    CK(cudaFuncSetAttribute(k_shfl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_soa));
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_aos));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct variant { const char *name; void (*fn)(uint64_t *, const gf192 *, gf192); size_t smem; } vs[] = {
        {"stage64", k_stage64, smem_soa}, {"stage128", k_stage128, smem_soa}, {"tma", k_tma, smem_aos}, {"shfl", k_shfl, smem_soa}};
    std::vector<uint64_t> ref, out(n * 3);
    int bad = 0;
    for (auto &v : vs) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaMemcpy(d_work, d_in, n * 24, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e0));
            v.fn<<<(unsigned)n_tiles, THREADS, v.smem>>>(d_work, d_tab, tau);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaMemcpy(out.data(), d_work, n * 24, cudaMemcpyDeviceToHost));
        if (ref.empty()) ref = out;
        const bool same = memcmp(ref.data(), out.data(), n * 24) == 0;
        bad += !same;
        const double mults = (double)STAGES * (n / 2);
        printf("%-9s %8.3f ms  %6.2f G mult/s  %7.1f GB/s (2 x 24 B x 2^23)  output %s\n", v.name, best, mults / best / 1e6, 2.0 * 24 * n / best / 1e6,
               same ? "identical" : "DIFFERENT");
    }
    return bad;
}
