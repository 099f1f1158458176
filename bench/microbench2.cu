// microbench2.cu — gf_mul throughput with fresh operands from shared memory at different occupancies.
#include <cstdio>
#include <cuda_runtime.h>
#include "../libiop_b200/csrc/gf192.cuh"
using namespace b200;

template <int MINB, int ILP>
__global__ void __launch_bounds__(256, MINB) k_fresh(gf192 *out, int iters, uint32_t seed)
{
    extern __shared__ uint64_t sm[];
    const int T = 2048;
    uint64_t *p0 = sm, *p1 = sm + T, *p2 = sm + 2 * T;
    for (int i = threadIdx.x; i < T; i += 256) { p0[i] = seed * 0x9E3779B97F4A7C15ull + i; p1[i] = (uint64_t)i * seed; p2[i] = ~(uint64_t)i ^ seed; }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        gf192 a[ILP], b[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const int ia = (threadIdx.x + 256 * u + 37 * it) & (T - 1), ib = (ia + 1024) & (T - 1);
            uint64_t x = p0[ia], y = p1[ia], z = p2[ia];
            a[u].w[0] = x; a[u].w[1] = x >> 32; a[u].w[2] = y; a[u].w[3] = y >> 32; a[u].w[4] = z; a[u].w[5] = z >> 32;
            x = p0[ib]; y = p1[ib]; z = p2[ib];
            b[u].w[0] = x; b[u].w[1] = x >> 32; b[u].w[2] = y; b[u].w[3] = y >> 32; b[u].w[4] = z; b[u].w[5] = z >> 32;
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const gf192 r = gf_mul(a[u], b[u]);
            const int ia = (threadIdx.x + 256 * u + 37 * it) & (T - 1);
            p0[ia] = (uint64_t)r.w[0] | ((uint64_t)r.w[1] << 32);
            p1[ia] = (uint64_t)r.w[2] | ((uint64_t)r.w[3] << 32);
            p2[ia] = (uint64_t)r.w[4] | ((uint64_t)r.w[5] << 32);
        }
        __syncthreads();
    }
    gf192 r; r.w[0] = p0[threadIdx.x]; r.w[1] = p1[threadIdx.x]; r.w[2] = p2[threadIdx.x]; r.w[3] = r.w[4] = r.w[5] = 0;
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

template <typename K>
static void run(const char *name, K kern, int blocks_per_sm, int ilp, int sms, gf192 *buf)
{
    const int iters = 256;
    const int smem = 3 * 2048 * 8;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // pad smem so that exactly blocks_per_sm fit
    const int pad = (227 * 1024 / blocks_per_sm) - 2048;
    const int use = pad > smem ? pad : smem;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, use);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    const int blocks = sms * occ * 4;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<blocks, 256, use>>>(buf, iters, 7); cudaDeviceSynchronize();
    cudaEventRecord(a); kern<<<blocks, 256, use>>>(buf, iters, 7); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-18s regs %3d blocks/SM %d warps/SM %2d ILP %d : %.2f Gmul/s  (%s)\n", name, fa.numRegs, occ, occ * 8, ilp,
           (double)blocks * 256 * iters * ilp / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    gf192 *buf; cudaMalloc(&buf, (size_t)prop.multiProcessorCount * 64 * 256 * sizeof(gf192));
    const int sms = prop.multiProcessorCount;
    run("minb1 ilp1", k_fresh<1, 1>, 1, 1, sms, buf);
    run("minb2 ilp1", k_fresh<2, 1>, 2, 1, sms, buf);
    run("minb3 ilp1", k_fresh<3, 1>, 3, 1, sms, buf);
    run("minb4 ilp1", k_fresh<4, 1>, 4, 1, sms, buf);
    run("minb5 ilp1", k_fresh<5, 1>, 5, 1, sms, buf);
    run("minb6 ilp1", k_fresh<6, 1>, 6, 1, sms, buf);
    run("minb8 ilp1", k_fresh<8, 1>, 8, 1, sms, buf);
    run("minb1 ilp2", k_fresh<1, 2>, 1, 2, sms, buf);
    run("minb2 ilp2", k_fresh<2, 2>, 2, 2, sms, buf);
    run("minb3 ilp2", k_fresh<3, 2>, 3, 2, sms, buf);
    run("minb4 ilp2", k_fresh<4, 2>, 4, 2, sms, buf);
    run("minb2 ilp4", k_fresh<2, 4>, 2, 4, sms, buf);
    return 0;
}
