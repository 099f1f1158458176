// bench/microbench.cu — pipe-level microbenchmarks for the integer kernels (run under gpurun).
// Prints one line per test: name, Gops/s.  Build: see bench/Makefile (two builds: GF spacing 4 and 5).
#include <cstdio>
#include <cuda_runtime.h>
#include "../libiop_b200/csrc/gf192.cuh"
#include "../libiop_b200/csrc/fp192.cuh"
using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void k_imad_wide(uint64_t *out, int iters, uint32_t seed)
{
    uint64_t acc[ILP]; uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = (uint64_t)(a + i) * (uint64_t)((uint32_t)acc[i] ^ b) + acc[i];
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int ILP>
__global__ void k_imad_lo(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t acc[ILP]; uint32_t a = seed + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = i + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * (a + i) + acc[i];
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int ILP>
__global__ void k_lop3(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t acc[ILP]; uint32_t a = seed + threadIdx.x, b = seed ^ blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = i + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = (acc[i] ^ a) & (b | acc[(i + 1) % ILP]);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mixed: IMAD.WIDE and LOP3 interleaved 1:1
template <int ILP>
__global__ void k_mixed(uint64_t *out, int iters, uint32_t seed)
{
    uint64_t acc[ILP]; uint32_t x[ILP]; uint32_t a = seed + threadIdx.x, b = seed ^ blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = i; x[i] = i + b; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            acc[i] = (uint64_t)(a + i) * (uint64_t)x[i] + acc[i];
            x[i] = (x[i] ^ a) & (b | x[(i + 1) % ILP]);
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= acc[i] ^ x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_gfmul(gf192 *out, int iters, gf192 seed)
{
    gf192 x = seed, y = seed;
    x.w[0] ^= threadIdx.x; y.w[1] ^= blockIdx.x;
    for (int it = 0; it < iters; ++it) { x = gf_mul(x, y); y.w[0] ^= x.w[5]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_gfmul2(gf192 *out, int iters, gf192 seed)   // two independent chains per thread
{
    gf192 x = seed, y = seed, u = seed, v = seed;
    x.w[0] ^= threadIdx.x; y.w[1] ^= blockIdx.x; u.w[2] ^= threadIdx.x * 7; v.w[3] ^= blockIdx.x * 5;
    for (int it = 0; it < iters; ++it) { x = gf_mul(x, y); u = gf_mul(u, v); y.w[0] ^= x.w[5]; v.w[0] ^= u.w[5]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = gf_add(x, u);
}
__global__ void k_gfmul_prep(gf192 *out, int iters, gf192 seed)  // one operand prepared once (twiddle reuse)
{
    gf192 x = seed, y = seed;
    x.w[0] ^= threadIdx.x; y.w[1] ^= blockIdx.x;
    const gf192_prepared py = gf_prepare(y);
    for (int it = 0; it < iters; ++it) { x = gf_mul(x, py); x.w[0] ^= it; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_fpmul(fp192 *out, int iters, fp192 seed, fp_dev_params P)
{
    fp192 x = seed, y = seed;
    x.w[0] ^= threadIdx.x; y.w[1] ^= blockIdx.x;
    for (int it = 0; it < iters; ++it) { x = fp_mul(x, y, P); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename F>
static float time_ms(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d spacing %d\n", prop.name, prop.multiProcessorCount, B200_GF_SPACING);
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    void *buf; CK(cudaMalloc(&buf, (size_t)blocks * threads * 32));
    const double nthreads = (double)blocks * threads;
    {
        const int iters = 4096; constexpr int ILP = 8;
        float ms = time_ms([&] { k_imad_wide<ILP><<<blocks, threads>>>((uint64_t *)buf, iters, 12345); });
        printf("imad_wide   %.1f Gop/s (%.1f per clk per SM @1.9GHz)\n", nthreads * iters * ILP / ms * 1e-6, nthreads * iters * ILP / ms * 1e-6 / prop.multiProcessorCount / 1.9);
        ms = time_ms([&] { k_imad_lo<ILP><<<blocks, threads>>>((uint32_t *)buf, iters, 12345); });
        printf("imad_lo     %.1f Gop/s (%.1f per clk per SM @1.9GHz)\n", nthreads * iters * ILP / ms * 1e-6, nthreads * iters * ILP / ms * 1e-6 / prop.multiProcessorCount / 1.9);
        ms = time_ms([&] { k_lop3<ILP><<<blocks, threads>>>((uint32_t *)buf, iters, 12345); });
        printf("lop3 (2/it) %.1f Gop/s\n", nthreads * iters * ILP * 2 / ms * 1e-6);
        ms = time_ms([&] { k_mixed<ILP><<<blocks, threads>>>((uint64_t *)buf, iters, 12345); });
        printf("mixed imad.wide+2lop3: %.1f G(imad)/s\n", nthreads * iters * ILP / ms * 1e-6);
    }
    {
        gf192 seed; for (int i = 0; i < 6; ++i) seed.w[i] = 0x9E3779B9u * (i + 1);
        const int iters = 512;
        float ms = time_ms([&] { k_gfmul<<<blocks, threads>>>((gf192 *)buf, iters, seed); });
        printf("gf_mul chain     %.2f Gmul/s\n", nthreads * iters / ms * 1e-6);
        ms = time_ms([&] { k_gfmul2<<<blocks, threads>>>((gf192 *)buf, iters, seed); });
        printf("gf_mul 2 chains  %.2f Gmul/s\n", nthreads * iters * 2 / ms * 1e-6);
        ms = time_ms([&] { k_gfmul_prep<<<blocks, threads>>>((gf192 *)buf, iters, seed); });
        printf("gf_mul prepared  %.2f Gmul/s\n", nthreads * iters / ms * 1e-6);
    }
    {
        // edwards_Fr
        const uint64_t mod[3] = {0x4F1DE42600000001ull, 0, 0};
        (void)mod;
        fp_dev_params P;
        const uint32_t m32[6] = {0x00000001u, 0x4f1de426u, 0x5c1a3d4cu, 0x81205e0bu, 0x26be5e6bu, 0x00103541u}; // not the real modulus: timing only
        for (int i = 0; i < 6; ++i) { P.mod[i] = m32[i]; P.one[i] = i; P.r2[i] = i * 3 + 1; }
        P.inv32 = 0xffffffffu;
        fp192 seed; for (int i = 0; i < 6; ++i) seed.w[i] = 0x12345u * (i + 1);
        seed.w[5] &= 0xffff;
        const int iters = 2048;
        float ms = time_ms([&] { k_fpmul<<<blocks, threads>>>((fp192 *)buf, iters, seed, P); });
        printf("fp_mul chain     %.2f Gmul/s\n", nthreads * iters / ms * 1e-6);
    }
    return 0;
}
