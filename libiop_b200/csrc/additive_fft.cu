/*
 * additive_fft.cu — Gao–Mateer additive FFT / IFFT over GF(2^192) on affine subspaces, sm_100a.
 *
 * Replaces additive_FFT (libiop/algebra/fft.tcc:39-124) and additive_IFFT (fft.tcc:126-204) behind
 * FFT_over_field_subset / IFFT_over_field_subset (fft.tcc:414-433).  Outputs are the evaluations of
 * the unique polynomial on the points `shift + sum_k bit_k(idx) basis[k]` in index order
 * (subspace.tcc:54-71,161-170), so they are bit-identical to the reference whatever the schedule.
 *
 * B200 design (not the reference's ~m^2/2 full-array sweeps):
 *   - The array lives in HBM/L2; every kernel launch ("tile pass") stages a tile of 2^TB elements in
 *     shared memory (SoA, 3 x u64 planes), runs as many consecutive algorithm steps as touch only
 *     the index bits of that tile, and writes the tile back.  Steps are twist (TW), Taylor/radix
 *     conversion at one stride (TAY) and butterfly stage (BF); a host-side scheduler packs them.
 *   - The forward half (monomial -> "novel" basis: twists + Taylor expansions) depends only on the
 *     basis, not on the shift.  An FFT of n_coeffs <= 2^d coefficients onto a 2^m-point domain is
 *     therefore done as ONE forward half on 2^d points followed by 2^(m-d) independent butterfly
 *     halves, one per coset of span(basis[0..d)) (contiguous output blocks), instead of the
 *     reference's size-2^m transform of a zero-padded vector: 100 instead of 276 field multiplications
 *     per input coefficient for the Aurora LDE shape (d = 20, m = 23).
 *   - twist factors beta_j^e are cached per domain in HBM (the prover reuses its domains for every
 *     oracle); butterfly twiddles `shift + subset-sum` are rebuilt per butterfly from 256-entry
 *     subset-sum tables, so no per-shift twiddle array exists.
 */
#include <algorithm>
#include <cstring>

#include "ctx.h"
#include <cooperative_groups.h>
#include <cstdlib>
#include <type_traits>

#include "gf192.cuh"

namespace b200 {

enum : uint32_t { OP_TW = 1, OP_TAY = 2, OP_BF = 3 };
#define B200_MAX_OPS 160
#ifndef B200_ILP
#define B200_ILP 1   /* independent butterflies in flight per thread; with the 4-class multiplier 1 is 1-1.5 % faster than 2 (bench/build_variants.sh) */
#endif
#ifndef B200_TILE_THREADS
#define B200_TILE_THREADS 256
#endif
#ifndef B200_TILE_MINBLOCKS
#define B200_TILE_MINBLOCKS 2
#endif
#define B200_SMEM_PAD 4 /* u64 words of padding between planes */
#ifndef B200_TILE_VEC128
#define B200_TILE_VEC128 1 /* 128-bit tile loads / stores (LDG.E.128 / STG.E.128): measured against the 64-bit variant, profiles/r2_variants.md */
#endif
#define B200_SUM_CHUNKS 4 /* 8-bit chunks of the butterfly index: stages up to 2^32 */

struct PassParams {
    const gf192 *src;
    gf192 *dst;
    const gf192 *tw;    /* twist tables (forward or inverse) */
    const gf192 *sums;  /* [d][B200_SUM_CHUNKS][256] subset sums of the recursed bases */
    const gf192 *taus;  /* [(1 + n_extra)][d] recursed shifts: row 0 for the domain shift, 1+k for basis[d+k] */
    uint32_t d, lo, hi, c, ctop;
    uint32_t cb;        /* log2(cosets per tile): small transforms pack several cosets into one CTA */
    uint32_t bitrev;    /* load src at bitrev_d(i) (out-of-place first butterfly pass) */
    uint32_t n_extra;
    uint32_t src_coset_stride_log; /* src advances by (coset >> src_coset_shr) << this (use 32 for "shared source") */
    uint32_t src_coset_shr;        /* batched transforms: the 2^shr cosets of one batch row read the same source row */
    uint32_t n_ops;
    uint32_t cl_log;    /* log2(CTAs per tile) for additive_cluster_pass (thread-block cluster, tile spread over DSMEM) */
    uint32_t ops[B200_MAX_OPS];
};

struct AdditivePlan {
    uint32_t d = 0;
    std::vector<gf192> basis, beta, betainv;
    std::vector<std::vector<gf192>> gamma; /* gamma[j][b], b < d-1-j : recursed_betas of fft.tcc:85-92 */
    gf192 *tw = nullptr, *itw = nullptr, *sums = nullptr;
    /* recursed shifts per (shift, extra basis vectors), kept on the device: a prover calls the same few domains thousands
       of times (Ligero: 2570 row transforms over two domains) and recomputing them costs a stream synchronisation */
    std::map<std::string, gf192 *> taus_cache;
    ~AdditivePlan() { cudaFree(tw); cudaFree(itw); cudaFree(sums); for (auto &kv : taus_cache) cudaFree(kv.second); }
};

__host__ __device__ __forceinline__ size_t tw_offset(uint32_t d, uint32_t j)
{
    return ((size_t)2 << d) - ((size_t)2 << (d - j));
}

/* ---------------------------------------------------------------------------------------------- */
/* shared-memory tile access (SoA: three planes of u64)                                           */
struct Tile {
    uint64_t *p0, *p1, *p2;
    __device__ __forceinline__ gf192 ld(uint32_t L) const
    {
        const uint64_t a = p0[L], b = p1[L], c = p2[L];
        gf192 r;
        r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
        r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
        r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
        return r;
    }
    __device__ __forceinline__ void st(uint32_t L, const gf192 &x) const
    {
        p0[L] = (uint64_t)x.w[0] | ((uint64_t)x.w[1] << 32);
        p1[L] = (uint64_t)x.w[2] | ((uint64_t)x.w[3] << 32);
        p2[L] = (uint64_t)x.w[4] | ((uint64_t)x.w[5] << 32);
    }
    __device__ __forceinline__ void xor_into(uint32_t Ldst, uint32_t Lsrc) const
    {
        p0[Ldst] ^= p0[Lsrc]; p1[Ldst] ^= p1[Lsrc]; p2[Ldst] ^= p2[Lsrc];
    }
};

/* the array's own layout (24-byte elements): what a bulk copy delivers.  64-bit accesses at a 24-byte stride are
   conflict-free (16 lanes x 8 bytes cover the 32 banks once). */
struct TileAoS {
    uint64_t *p;
    __device__ __forceinline__ gf192 ld(uint32_t L) const
    {
        const uint64_t a = p[3 * L], b = p[3 * L + 1], c = p[3 * L + 2];
        gf192 r;
        r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
        r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
        r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
        return r;
    }
    __device__ __forceinline__ void st(uint32_t L, const gf192 &x) const
    {
        p[3 * L] = (uint64_t)x.w[0] | ((uint64_t)x.w[1] << 32);
        p[3 * L + 1] = (uint64_t)x.w[2] | ((uint64_t)x.w[3] << 32);
        p[3 * L + 2] = (uint64_t)x.w[4] | ((uint64_t)x.w[5] << 32);
    }
    __device__ __forceinline__ void xor_into(uint32_t Ldst, uint32_t Lsrc) const
    {
        p[3 * Ldst] ^= p[3 * Lsrc]; p[3 * Ldst + 1] ^= p[3 * Lsrc + 1]; p[3 * Ldst + 2] ^= p[3 * Lsrc + 2];
    }
};
__device__ __forceinline__ uint32_t smem_u32(const void *q) { return (uint32_t)__cvta_generic_to_shared(q); }

__device__ __forceinline__ gf192 ldg_elem(const gf192 *p)
{
    const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
    const uint64_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    gf192 r;
    r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
    r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
    r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
    return r;
}

__device__ __forceinline__ uint32_t insert_zero(uint32_t k, uint32_t pos)
{
    const uint32_t low = k & ((1u << pos) - 1u);
    return ((k >> pos) << (pos + 1)) | low;
}

/* NT threads per CTA, at least MB resident CTAs per SM.  Two instantiations per direction: 256 x 2 for grids that fill the
   machine (two 96 KiB tiles per SM) and 512 x 1 for small transforms whose whole pass is a handful of CTAs (a 2^11-point
   Ligero row is ONE tile: the chain of ~70 dependent steps then runs twice as fast on 16 warps). */
/* TMA = true: the tile is one contiguous run of the array on both sides (P.lo = P.c = P.ctop = P.cb = 0, no bit reversal: the
   12-stage passes of the fused extension and of its halves).  One thread stages it with cp.async.bulk (TMA bulk copy,
   UBLKCP in SASS) completing on an mbarrier, the tile keeps the array's AoS layout in shared memory, and it is written back
   with cp.async.bulk.global.shared.  Measured against the 128-bit staging: profiles/r2_variants.md 1. */
template <bool INV, int NT, int MB, bool TMA = false>
__global__ void __launch_bounds__(NT, MB) additive_tile_pass(const PassParams P)
{
    extern __shared__ __align__(128) uint64_t smem[];
    __shared__ __align__(8) uint64_t mbar;

    const uint32_t mainb = P.hi - P.lo;
    const uint32_t tbw = mainb + P.c + P.ctop;              /* index bits of one coset inside the tile */
    const uint32_t tb = tbw + P.cb;
    const uint32_t T = 1u << tb;
    const uint32_t wmask = (1u << tbw) - 1u;
    typename std::conditional<TMA, TileAoS, Tile>::type t;
    if constexpr (TMA) { t.p = smem; }
    else { t.p0 = smem; t.p1 = smem + T + B200_SMEM_PAD; t.p2 = smem + 2 * (T + B200_SMEM_PAD); }
    gf192 *tau_s = reinterpret_cast<gf192 *>(smem + 3 * (T + B200_SMEM_PAD));   /* [2^cb][d] */

    const uint32_t d = P.d;
    const uint32_t tilebits = d - tbw;                      /* tiles per coset = 2^tilebits */
    const uint32_t coset_base = (blockIdx.x >> tilebits) << P.cb;
    const uint32_t tile = blockIdx.x & ((1u << tilebits) - 1u);
    const uint32_t midbits = P.lo - P.c;
    const uint32_t mid = tile & ((1u << midbits) - 1u);
    const uint32_t top = tile >> midbits;

    const uint32_t tile_bytes = T * (uint32_t)sizeof(gf192);
    if constexpr (TMA) {
        /* contiguous tile: element tile * T of the coset's source row */
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const size_t cofs = P.src_coset_stride_log >= 32 ? 0 : ((size_t)(coset_base >> P.src_coset_shr) << P.src_coset_stride_log);
            const char *g = reinterpret_cast<const char *>(P.src + cofs + (size_t)tile * T);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(tile_bytes) : "memory");
            const uint32_t chunk = tile_bytes > 24576u ? 24576u : tile_bytes;
            for (uint32_t ofs = 0; ofs < tile_bytes; ofs += chunk)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(reinterpret_cast<char *>(smem) + ofs)), "l"(g + ofs), "r"(chunk), "r"(smem_u32(&mbar)) : "memory");
        }
    }

    /* recursed shift of each coset for every depth: F2-linear in the shift, so
       tau(shift + sum_k q_k basis[d+k]) = tau(shift) + sum_k q_k tau(basis[d+k]) */
    for (uint32_t e = threadIdx.x; e < (d << P.cb); e += blockDim.x) {
        const uint32_t cl = e / d, dep = e - cl * d;
        const uint32_t coset = coset_base + cl;
        gf192 v = ldg_elem(P.taus + dep);
        for (uint32_t k = 0; k < P.n_extra; ++k)
            if ((coset >> k) & 1u) gf_addto(v, ldg_elem(P.taus + (size_t)(1 + k) * d + dep));
        tau_s[e] = v;
    }

    const bool shared_src = P.src_coset_stride_log >= 32;

    auto global_index = [&](uint32_t L) -> uint32_t {
        const uint32_t Lw = L & wmask;
        const uint32_t w = Lw & ((1u << P.c) - 1u);
        const uint32_t u = (Lw >> P.c) & ((1u << mainb) - 1u);
        const uint32_t tbv = Lw >> (P.c + mainb);
        return w | (mid << P.c) | (u << P.lo) | (top << P.hi) | (P.ctop ? (tbv << (d - P.ctop)) : 0u);
    };

    if constexpr (TMA) {
        __syncthreads();                                   /* the barrier's initialisation is visible to every waiter */
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    } else {
#if B200_TILE_VEC128
    /* 128-bit variant: a thread moves two consecutive elements (48 bytes) with three 16-byte loads and three 16-byte
       shared-memory stores.  Index bit 0 of the tile is index bit 0 of the array in every tile shape, so elements L and L+1
       (L even) are adjacent in memory unless the load is the bit-reversing gather. */
    const bool vec_in = P.bitrev != 1 && tbw >= 1 && ((reinterpret_cast<uintptr_t>(P.src) & 15u) == 0);
    if (vec_in) {
        for (uint32_t L = 2 * threadIdx.x; L < T; L += 2 * blockDim.x) {
            const uint32_t i = global_index(L);
            const size_t cofs = shared_src ? 0 : ((size_t)((coset_base + (L >> tbw)) >> P.src_coset_shr) << P.src_coset_stride_log);
            const uint4 *q = reinterpret_cast<const uint4 *>(P.src + cofs + i);
            const uint4 x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);     /* a0 b0 | c0 a1 | b1 c1 (64-bit words) */
            *reinterpret_cast<uint4 *>(t.p0 + L) = make_uint4(x.x, x.y, y.z, y.w);
            *reinterpret_cast<uint4 *>(t.p1 + L) = make_uint4(x.z, x.w, z.x, z.y);
            *reinterpret_cast<uint4 *>(t.p2 + L) = make_uint4(y.x, y.y, z.z, z.w);
        }
    } else
#endif
    for (uint32_t L = threadIdx.x; L < T; L += blockDim.x) {
        const uint32_t i = global_index(L);
        const uint32_t a = (P.bitrev == 1) ? (__brev(i) >> (32 - d)) : i;
        const size_t cofs = shared_src ? 0 : ((size_t)((coset_base + (L >> tbw)) >> P.src_coset_shr) << P.src_coset_stride_log);
        t.st(L, ldg_elem(P.src + cofs + a));
    }
    }

    for (uint32_t o = 0; o < P.n_ops; ++o) {
#ifndef B200_EXP_NOSYNC
        __syncthreads();
#else
        if (o == 0) __syncthreads();
#endif
        const uint32_t kind = P.ops[o] >> 8, bit = P.ops[o] & 0xffu;
        if (kind == OP_TW) {
            /* twist at depth j = bit: S[i] *= beta_j^(i >> j)     (fft.tcc:61-70 / 190-200) */
            const gf192 *tab = P.tw + tw_offset(d, bit);
            for (uint32_t L = threadIdx.x; L < T; L += blockDim.x * B200_ILP) {
                gf192 v[B200_ILP], f[B200_ILP];
                bool act[B200_ILP];
#pragma unroll
                for (int u = 0; u < B200_ILP; ++u) {
                    const uint32_t LL = L + u * blockDim.x;
                    const uint32_t e = (LL < T) ? (global_index(LL) >> bit) : 0u;
                    act[u] = e != 0;
                    if (act[u]) { v[u] = t.ld(LL); f[u] = ldg_elem(tab + e); }
                }
#pragma unroll
                for (int u = 0; u < B200_ILP; ++u)
                    if (act[u]) t.st(L + u * blockDim.x, gf_mul(v[u], f[u]));
            }
        } else if (kind == OP_TAY) {
            /* radix conversion at linear stride 2^bit: [A B C D] -> [A, B+C+D, C+D, D] (fft.tcc:72-83),
               inverse [A, B+C, C+D, D] (fft.tcc:175-188) */
            const uint32_t pos = bit - P.lo + P.c;
            for (uint32_t k = threadIdx.x; k < (T >> 2); k += blockDim.x) {
                const uint32_t low = k & ((1u << pos) - 1u);
                const uint32_t base = ((k >> pos) << (pos + 2)) | low;
                const uint32_t B = base | (1u << pos), C = base | (2u << pos), D = base | (3u << pos);
                if (!INV) { t.xor_into(C, D); t.xor_into(B, C); }
                else      { t.xor_into(B, C); t.xor_into(C, D); }
            }
        } else { /* OP_BF: butterfly stage jj = bit, twiddle = tau + subset sum over index bits < jj */
            const uint32_t jj = bit, pos = jj - P.lo + P.c;
            const uint32_t depth = d - 1 - jj;
            const gf192 *stab = P.sums + (size_t)jj * B200_SUM_CHUNKS * 256;
            for (uint32_t k = threadIdx.x; k < (T >> 1); k += blockDim.x * B200_ILP) {
                gf192 a[B200_ILP], b[B200_ILP], tw[B200_ILP];
                uint32_t L0[B200_ILP];
                bool act[B200_ILP];
#pragma unroll
                for (int u = 0; u < B200_ILP; ++u) {
                    const uint32_t kk = k + u * blockDim.x;
                    act[u] = kk < (T >> 1);
                    if (act[u]) {
                        L0[u] = insert_zero(kk, pos);
                        const uint32_t tix = global_index(L0[u]) & ((1u << jj) - 1u);
                        tw[u] = tau_s[(L0[u] >> tbw) * d + depth];
#ifndef B200_EXP_NOTW
                        if (jj > 0) gf_addto(tw[u], ldg_elem(stab + (tix & 255u)));
                        if (jj > 8) gf_addto(tw[u], ldg_elem(stab + 256 + ((tix >> 8) & 255u)));
                        if (jj > 16) gf_addto(tw[u], ldg_elem(stab + 512 + ((tix >> 16) & 255u)));
                        if (jj > 24) gf_addto(tw[u], ldg_elem(stab + 768 + (tix >> 24)));
#else
                        tw[u].w[0] ^= tix;
#endif
                        a[u] = t.ld(L0[u]); b[u] = t.ld(L0[u] | (1u << pos));
                        if (INV) gf_addto(b[u], a[u]);            /* fft.tcc:164 */
                    }
                }
#pragma unroll
                for (int u = 0; u < B200_ILP; ++u) {
                    if (act[u]) {
#ifndef B200_EXP_NOMUL
                        gf_addto(a[u], gf_mul(b[u], tw[u]));      /* fft.tcc:116 / 165 */
#else
                        gf_addto(a[u], gf_add(b[u], tw[u]));
#endif
                        if (!INV) gf_addto(b[u], a[u]);           /* fft.tcc:117 */
                        t.st(L0[u], a[u]); t.st(L0[u] | (1u << pos), b[u]);
                    }
                }
            }
        }
    }
    if constexpr (TMA) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* generic-proxy writes visible to the async proxy */
        __syncthreads();
        if (threadIdx.x == 0) {
            char *g = reinterpret_cast<char *>(P.dst + ((size_t)coset_base << d) + (size_t)tile * T);
            const uint32_t chunk = tile_bytes > 24576u ? 24576u : tile_bytes;
            for (uint32_t ofs = 0; ofs < tile_bytes; ofs += chunk)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(g + ofs), "r"(smem_u32(reinterpret_cast<char *>(smem) + ofs)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    } else {
    __syncthreads();

    uint64_t *out = reinterpret_cast<uint64_t *>(P.dst);
#if B200_TILE_VEC128
    const bool vec_out = P.bitrev != 2 && tbw >= 1 && ((reinterpret_cast<uintptr_t>(P.dst) & 15u) == 0);
    if (vec_out) {
        for (uint32_t L = 2 * threadIdx.x; L < T; L += 2 * blockDim.x) {
            const uint32_t i = global_index(L);
            const uint4 a = *reinterpret_cast<const uint4 *>(t.p0 + L), b = *reinterpret_cast<const uint4 *>(t.p1 + L),
                        c2 = *reinterpret_cast<const uint4 *>(t.p2 + L);
            uint4 *q = reinterpret_cast<uint4 *>(out + (((size_t)(coset_base + (L >> tbw)) << d) + i) * 3);
            q[0] = make_uint4(a.x, a.y, b.x, b.y);
            q[1] = make_uint4(c2.x, c2.y, a.z, a.w);
            q[2] = make_uint4(b.z, b.w, c2.z, c2.w);
        }
        return;
    }
#endif
    for (uint32_t L = threadIdx.x; L < T; L += blockDim.x) {
        const uint32_t i = global_index(L);
        const uint32_t a = (P.bitrev == 2) ? (__brev(i) >> (32 - d)) : i;
        uint64_t *q = out + (((size_t)(coset_base + (L >> tbw)) << d) + a) * 3;
        q[0] = t.p0[L]; q[1] = t.p1[L]; q[2] = t.p2[L];
    }
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* Small transforms: one tile spread over a thread-block cluster.
   A transform of <= 2^12 points is a single tile, and additive_tile_pass runs its chain of dependent steps (11 twists, 55
   radix-conversion strides and 11 butterfly stages for a 2^11-point Ligero row) on ONE SM: ~0.3 ms during which the other
   147 SMs idle, and Ligero makes 2570 such calls per proof.  Here the tile is split over the shared memories of a cluster
   of up to 8 CTAs (distributed shared memory): element L lives in CTA L >> lb at offset L & (2^lb - 1); twists touch only
   the CTA's own slice, radix-conversion and butterfly steps whose stride crosses slices read and write the partner CTA's
   shared memory directly.  Steps that stay inside a slice are separated by __syncthreads, steps that cross (or follow a
   crossing step) by a cluster barrier.  Same step semantics as additive_tile_pass. */
struct ClusterTile {
    uint64_t *base[8];   /* plane 0 of every CTA of the cluster (generic addresses into distributed shared memory) */
    uint32_t lb, S1;     /* log2(elements per CTA), plane stride in u64 */
    __device__ __forceinline__ uint64_t *at(uint32_t L) const { return base[L >> lb] + (L & ((1u << lb) - 1u)); }
    __device__ __forceinline__ gf192 ld(uint32_t L) const
    {
        const uint64_t *q = at(L);
        const uint64_t a = q[0], b = q[S1], c = q[2 * S1];
        gf192 r;
        r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
        r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
        r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
        return r;
    }
    __device__ __forceinline__ void st(uint32_t L, const gf192 &x) const
    {
        uint64_t *q = at(L);
        q[0] = (uint64_t)x.w[0] | ((uint64_t)x.w[1] << 32);
        q[S1] = (uint64_t)x.w[2] | ((uint64_t)x.w[3] << 32);
        q[2 * S1] = (uint64_t)x.w[4] | ((uint64_t)x.w[5] << 32);
    }
    __device__ __forceinline__ void xor_into(uint32_t Ldst, uint32_t Lsrc) const
    {
        uint64_t *qd = at(Ldst);
        const uint64_t *qs = at(Lsrc);
        qd[0] ^= qs[0]; qd[S1] ^= qs[S1]; qd[2 * S1] ^= qs[2 * S1];
    }
};

template <bool INV>
__global__ void __launch_bounds__(256, 1) additive_cluster_pass(const PassParams P)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ uint64_t smem[];

    const uint32_t rank = cluster.block_rank();
    const uint32_t C = 1u << P.cl_log;
    const uint32_t mainb = P.hi - P.lo;
    const uint32_t tbw = mainb + P.c + P.ctop;
    const uint32_t tb = tbw + P.cb;
    const uint32_t T = 1u << tb;
    const uint32_t lb = tb - P.cl_log, slice = 1u << lb;
    const uint32_t wmask = (1u << tbw) - 1u;
    ClusterTile t;
    t.lb = lb; t.S1 = slice + B200_SMEM_PAD;
    for (uint32_t r = 0; r < 8; ++r) t.base[r] = (r < C) ? cluster.map_shared_rank(smem, r) : smem;
    gf192 *tau_s = reinterpret_cast<gf192 *>(smem + 3 * (slice + B200_SMEM_PAD));   /* [2^cb][d], one copy per CTA */

    const uint32_t d = P.d;
    const uint32_t tilebits = d - tbw;
    const uint32_t tile_id = blockIdx.x >> P.cl_log;
    const uint32_t coset_base = (tile_id >> tilebits) << P.cb;
    const uint32_t tile = tile_id & ((1u << tilebits) - 1u);
    const uint32_t midbits = P.lo - P.c;
    const uint32_t mid = tile & ((1u << midbits) - 1u);
    const uint32_t top = tile >> midbits;

    for (uint32_t e = threadIdx.x; e < (d << P.cb); e += blockDim.x) {
        const uint32_t cl = e / d, dep = e - cl * d;
        const uint32_t coset = coset_base + cl;
        gf192 v = ldg_elem(P.taus + dep);
        for (uint32_t k = 0; k < P.n_extra; ++k)
            if ((coset >> k) & 1u) gf_addto(v, ldg_elem(P.taus + (size_t)(1 + k) * d + dep));
        tau_s[e] = v;
    }

    const bool shared_src = P.src_coset_stride_log >= 32;
    auto global_index = [&](uint32_t L) -> uint32_t {
        const uint32_t Lw = L & wmask;
        const uint32_t w = Lw & ((1u << P.c) - 1u);
        const uint32_t u = (Lw >> P.c) & ((1u << mainb) - 1u);
        const uint32_t tbv = Lw >> (P.c + mainb);
        return w | (mid << P.c) | (u << P.lo) | (top << P.hi) | (P.ctop ? (tbv << (d - P.ctop)) : 0u);
    };

    /* every CTA fills its own slice */
    for (uint32_t o = threadIdx.x; o < slice; o += blockDim.x) {
        const uint32_t L = rank * slice + o;
        const uint32_t i = global_index(L);
        const uint32_t a = (P.bitrev == 1) ? (__brev(i) >> (32 - d)) : i;
        const size_t cofs = shared_src ? 0 : ((size_t)((coset_base + (L >> tbw)) >> P.src_coset_shr) << P.src_coset_stride_log);
        t.st(L, ldg_elem(P.src + cofs + a));
    }

    bool prev_crosses = true;                      /* the first barrier must be cluster-wide: remote slices are being filled */
    for (uint32_t o = 0; o < P.n_ops; ++o) {
        const uint32_t kind = P.ops[o] >> 8, bit = P.ops[o] & 0xffu;
        const uint32_t pos = bit - P.lo + P.c;
        const bool crosses = (kind == OP_TAY) ? (pos + 2 > lb) : (kind == OP_BF) ? (pos + 1 > lb) : false;
        if (prev_crosses || crosses) cluster.sync(); else __syncthreads();
        prev_crosses = crosses;
        if (kind == OP_TW) {
            const gf192 *tab = P.tw + tw_offset(d, bit);
            for (uint32_t q = threadIdx.x; q < slice; q += blockDim.x) {
                const uint32_t L = rank * slice + q;
                const uint32_t e = global_index(L) >> bit;
                if (e != 0) t.st(L, gf_mul(t.ld(L), ldg_elem(tab + e)));
            }
        } else if (kind == OP_TAY) {
            const uint32_t per = (T >> 2) >> P.cl_log;               /* quads per CTA; contiguous k blocks stay in the slice when the stride does */
            for (uint32_t kl = threadIdx.x; kl < per; kl += blockDim.x) {
                const uint32_t k = rank * per + kl;
                const uint32_t low = k & ((1u << pos) - 1u);
                const uint32_t base = ((k >> pos) << (pos + 2)) | low;
                const uint32_t B = base | (1u << pos), Cq = base | (2u << pos), D = base | (3u << pos);
                if (!INV) { t.xor_into(Cq, D); t.xor_into(B, Cq); }
                else      { t.xor_into(B, Cq); t.xor_into(Cq, D); }
            }
        } else {
            const uint32_t jj = bit;
            const uint32_t depth = d - 1 - jj;
            const gf192 *stab = P.sums + (size_t)jj * B200_SUM_CHUNKS * 256;
            const uint32_t per = (T >> 1) >> P.cl_log;
            for (uint32_t kl = threadIdx.x; kl < per; kl += blockDim.x) {
                const uint32_t k = rank * per + kl;
                const uint32_t L0 = insert_zero(k, pos), L1 = L0 | (1u << pos);
                const uint32_t tix = global_index(L0) & ((1u << jj) - 1u);
                gf192 tw = tau_s[(L0 >> tbw) * d + depth];
                if (jj > 0) gf_addto(tw, ldg_elem(stab + (tix & 255u)));
                if (jj > 8) gf_addto(tw, ldg_elem(stab + 256 + ((tix >> 8) & 255u)));
                if (jj > 16) gf_addto(tw, ldg_elem(stab + 512 + ((tix >> 16) & 255u)));
                if (jj > 24) gf_addto(tw, ldg_elem(stab + 768 + (tix >> 24)));
                gf192 a = t.ld(L0), b = t.ld(L1);
                if (INV) gf_addto(b, a);
                gf_addto(a, gf_mul(b, tw));
                if (!INV) gf_addto(b, a);
                t.st(L0, a); t.st(L1, b);
            }
        }
    }
    cluster.sync();                                /* remote writes of the last step have landed; nobody touches a peer afterwards */

    uint64_t *out = reinterpret_cast<uint64_t *>(P.dst);
    const uint64_t *mine = smem;
    const uint32_t S1 = slice + B200_SMEM_PAD;
    for (uint32_t o = threadIdx.x; o < slice; o += blockDim.x) {
        const uint32_t L = rank * slice + o;
        const uint32_t i = global_index(L);
        const uint32_t a = (P.bitrev == 2) ? (__brev(i) >> (32 - d)) : i;
        uint64_t *q = out + (((size_t)(coset_base + (L >> tbw)) << d) + a) * 3;
        q[0] = mine[o]; q[1] = mine[o + S1]; q[2] = mine[o + 2 * S1];
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* plan construction                                                                              */

/* powers by doubling: tab[2^b + e] = tab[e] * base^(2^b) */
__global__ void power_table_step(gf192 *tab, gf192 factor, uint32_t len)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < len) tab[len + e] = gf_mul(tab[e], factor);
}

static int build_power_tables(b200_ctx *c, gf192 **out, const std::vector<gf192> &base, uint32_t d)
{
    const size_t total = ((size_t)2 << d);
    gf192 *tab = nullptr;
    if (cudaMalloc(&tab, total * sizeof(gf192)) != cudaSuccess) {
        cudaGetLastError();
        return set_err(c, B200_ERR_NOMEM, "cannot allocate %zu bytes of twist tables", total * sizeof(gf192));
    }
    for (uint32_t j = 0; j < d; ++j) {
        gf192 *t = tab + tw_offset(d, j);
        const gf192 one = gf_one();
        B200_CUDA(c, cudaMemcpyAsync(t, &one, sizeof one, cudaMemcpyHostToDevice, c->stream));
        gf192 f = base[j];
        for (uint32_t b = 0; b < d - j; ++b) {
            const uint32_t len = 1u << b;
            power_table_step<<<(len + 255) / 256, 256, 0, c->stream>>>(t, f, len);
            f = gf_sqr(f);
        }
    }
    B200_LAUNCH_CHECK(c);
    B200_CUDA(c, cudaStreamSynchronize(c->stream)); /* `one` is a stack temporary */
    *out = tab;
    return 0;
}

static int get_plan(b200_ctx *c, const gf192 *basis, uint32_t d, bool need_inverse,
                    std::shared_ptr<AdditivePlan> *out, bool need_twists = true)
{
    std::string key((const char *)basis, (size_t)d * sizeof(gf192));
    key.push_back((char)d);
    auto it = c->additive_plans.find(key);
    std::shared_ptr<AdditivePlan> pl;
    if (it != c->additive_plans.end()) pl = it->second;
    else {
        pl = std::make_shared<AdditivePlan>();
        pl->d = d;
        pl->basis.assign(basis, basis + d);
        pl->beta.resize(d); pl->betainv.resize(d); pl->gamma.resize(d);
        std::vector<gf192> betas2(pl->basis);
        for (uint32_t j = 0; j < d; ++j) {                 /* fft.tcc:56-97 (basis recursion only) */
            const gf192 beta = betas2[d - 1 - j];
            if (gf_is_zero(beta)) return set_err(c, B200_ERR_INVALID, "subspace basis is linearly dependent");
            const gf192 binv = gf_inv(beta);
            pl->beta[j] = beta; pl->betainv[j] = binv;
            pl->gamma[j].resize(d - 1 - j);
            for (uint32_t i = 0; i + 1 + j < d; ++i) {
                const gf192 nb = gf_mul(betas2[i], binv);
                pl->gamma[j][i] = nb;
                betas2[i] = gf_add(gf_sqr(nb), nb);
            }
        }
        /* subset-sum tables per butterfly stage jj (depth d-1-jj), three 8-bit chunks */
        std::vector<gf192> sums((size_t)(d ? d : 1) * B200_SUM_CHUNKS * 256, gf_zero());
        for (uint32_t jj = 0; jj < d; ++jj) {
            const std::vector<gf192> &g = pl->gamma[d - 1 - jj]; /* jj elements */
            for (uint32_t q = 0; q < B200_SUM_CHUNKS; ++q)
                for (uint32_t v = 1; v < 256; ++v) {
                    const uint32_t lowbit = v & (~v + 1u);
                    const uint32_t b = 8 * q + (31 - __builtin_clz(lowbit));
                    gf192 s = sums[((size_t)jj * B200_SUM_CHUNKS + q) * 256 + (v ^ lowbit)];
                    if (b < jj) gf_addto(s, g[b]);
                    sums[((size_t)jj * B200_SUM_CHUNKS + q) * 256 + v] = s;
                }
        }
        if (cudaMalloc(&pl->sums, sums.size() * sizeof(gf192)) != cudaSuccess) {
            cudaGetLastError();
            return set_err(c, B200_ERR_NOMEM, "cannot allocate subset-sum tables");
        }
        B200_CUDA(c, cudaMemcpy(pl->sums, sums.data(), sums.size() * sizeof(gf192), cudaMemcpyHostToDevice));
        c->additive_plans[key] = pl;
    }
    if (need_twists && !need_inverse && !pl->tw && d > 0) { int rc = build_power_tables(c, &pl->tw, pl->beta, d); if (rc) return rc; }
    if (need_twists && need_inverse && !pl->itw && d > 0) { int rc = build_power_tables(c, &pl->itw, pl->betainv, d); if (rc) return rc; }
    *out = pl;
    return 0;
}

void plans_clear_additive(b200_ctx *c) { c->additive_plans.clear(); }

/* recursed shifts tau_j(s), j < d   (fft.tcc:94-96) */
static void compute_taus(const AdditivePlan &pl, gf192 s, gf192 *out)
{
    for (uint32_t j = 0; j < pl.d; ++j) {
        const gf192 ns = gf_mul(s, pl.betainv[j]);
        out[j] = ns;
        s = gf_add(gf_sqr(ns), ns);
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* scheduler: pack consecutive steps into tile passes                                             */
struct Pass { uint32_t lo, hi, c, ctop, bitrev, cb = 0; std::vector<uint32_t> ops; };

static const uint32_t kTileBits = 12;  /* 4096 elements = 96 KiB of shared memory, 2 CTAs per SM */
static const uint32_t kCoalesceBits = 4; /* strided tiles keep runs of 16 elements = 384 B */

static void finish_pass(std::vector<Pass> &out, std::vector<uint32_t> &ops, int bmin, int bmax, uint32_t d)
{
    if (ops.empty()) return;
    Pass p; p.ctop = 0; p.bitrev = 0; p.ops = ops;
    const uint32_t tb = std::min(kTileBits, d);
    if (bmin < 0) { bmin = 0; bmax = 0; } /* only element-wise steps */
    uint32_t lo = (uint32_t)bmin, hi = (uint32_t)bmax + 1;
    if (lo <= kCoalesceBits || hi <= tb) { p.c = 0; lo = 0; if (hi < tb) hi = tb; }
    else {
        p.c = kCoalesceBits;
        while (hi - lo + p.c < tb) { if (lo > p.c) --lo; else if (hi < d) ++hi; else break; }
        if (lo == p.c) { lo = 0; p.c = 0; }
    }
    p.lo = lo; p.hi = hi;
    out.push_back(p);
    ops.clear();
}

static bool fits(int bmin, int bmax, uint32_t d)
{
    const uint32_t tb = std::min(kTileBits, d);
    const uint32_t lo = (uint32_t)bmin, hi = (uint32_t)bmax + 1;
    if (hi <= tb) return true;                       /* contiguous tile [0, tb) */
    if (lo <= kCoalesceBits) return false;
    return hi - lo + kCoalesceBits <= tb;
}

/* steps: sequence of (kind, bit) with the index bits each needs */
static std::vector<Pass> schedule(const std::vector<uint32_t> &steps, uint32_t d)
{
    std::vector<Pass> passes;
    std::vector<uint32_t> cur;
    int bmin = -1, bmax = -1;
    for (uint32_t s : steps) {
        const uint32_t kind = s >> 8, bit = s & 0xff;
        int nmin = bmin, nmax = bmax;
        if (kind == OP_TAY) { nmin = (bmin < 0) ? (int)bit : std::min(bmin, (int)bit); nmax = std::max(bmax, (int)bit + 1); }
        else if (kind == OP_BF) { nmin = (bmin < 0) ? (int)bit : std::min(bmin, (int)bit); nmax = std::max(bmax, (int)bit); }
        if ((nmin >= 0 && !fits(nmin, nmax, d)) || cur.size() >= B200_MAX_OPS) {
            finish_pass(passes, cur, bmin, bmax, d);
            bmin = bmax = -1;
            if (kind == OP_TAY) { bmin = (int)bit; bmax = (int)bit + 1; }
            else if (kind == OP_BF) { bmin = bmax = (int)bit; }
        } else { bmin = nmin; bmax = nmax; }
        cur.push_back(s);
    }
    finish_pass(passes, cur, bmin, bmax, d);
    return passes;
}

static bool tile_tma_enabled()
{
    static const bool on = !(getenv("B200_TILE_TMA") && getenv("B200_TILE_TMA")[0] == '0');     /* ablation switch */
    return on;
}
static bool cluster_disabled()
{
    static const bool off = getenv("B200_NO_CLUSTER") != nullptr;     /* ablation switch */
    return off;
}

template <bool INV>
static int launch_pass(b200_ctx *c, const Pass &p, PassParams P, uint32_t d, uint32_t n_cosets)
{
    P.lo = p.lo; P.hi = p.hi; P.c = p.c; P.ctop = p.ctop; P.bitrev = p.bitrev; P.d = d; P.cb = p.cb;
    P.n_ops = (uint32_t)p.ops.size();
    std::copy(p.ops.begin(), p.ops.end(), P.ops);
    const uint32_t tbw = (p.hi - p.lo) + p.c + p.ctop;
    const uint32_t tb = tbw + p.cb;
    const size_t smem = 3 * (((size_t)1 << tb) + B200_SMEM_PAD) * sizeof(uint64_t) + ((size_t)d << p.cb) * sizeof(gf192);
    bool *attr_set = c->attr_additive;
    if (!attr_set[INV]) {
        const int max_smem = (int)(3 * (((size_t)1 << 13) + B200_SMEM_PAD) * sizeof(uint64_t) + 8192);
        B200_CUDA(c, cudaFuncSetAttribute(additive_tile_pass<INV, B200_TILE_THREADS, B200_TILE_MINBLOCKS>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        B200_CUDA(c, cudaFuncSetAttribute(additive_tile_pass<INV, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        B200_CUDA(c, cudaFuncSetAttribute(additive_tile_pass<INV, B200_TILE_THREADS, B200_TILE_MINBLOCKS, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set[INV] = true;
    }
    const uint64_t nblocks = ((uint64_t)1 << (d - tbw)) * (n_cosets >> p.cb);
    if (nblocks > 0x7fffffffull) return set_err(c, B200_ERR_INVALID, "grid too large");
    const bool few_ctas = nblocks <= (uint64_t)c->sm_count && tb >= 10;   /* less than one CTA per SM: latency-bound, widen the CTA */
    /* a handful of tiles: spread each over a cluster of CTAs (distributed shared memory), see additive_cluster_pass */
    const bool cluster_off = cluster_disabled();
    uint32_t cl_log = 0;
    if (!cluster_off && tb >= 9 && nblocks * 8 <= (uint64_t)c->sm_count) cl_log = 3;
    else if (!cluster_off && tb >= 9 && nblocks * 4 <= (uint64_t)c->sm_count) cl_log = 2;
    uint32_t n_tw = 0, n_tay = 0, n_bf = 0;
    for (uint32_t o : p.ops) { const uint32_t k = o >> 8; n_tw += k == OP_TW; n_tay += k == OP_TAY; n_bf += k == OP_BF; }
    char label[160];
    snprintf(label, sizeof label, "additive_tile_pass<%d> d=%u cosets=%u grid=%llu tile=2^%u tw=%u tay=%u bf=%u", (int)INV, d, n_cosets,
             (unsigned long long)nblocks, tb, n_tw, n_tay, n_bf);
    ProfScope prof(c, label);
    if (cl_log) {
        P.cl_log = cl_log;
        const size_t slice = (size_t)1 << (tb - cl_log);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3((unsigned)(nblocks << cl_log), 1, 1);
        cfg.blockDim = dim3(256, 1, 1);
        cfg.dynamicSmemBytes = 3 * (slice + B200_SMEM_PAD) * sizeof(uint64_t) + ((size_t)d << p.cb) * sizeof(gf192);
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1u << cl_log; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        B200_CUDA(c, cudaLaunchKernelEx(&cfg, additive_cluster_pass<INV>, P));
        B200_LAUNCH_CHECK(c);
        return 0;
    }
    /* contiguous tile on both sides -> TMA bulk staging (16-byte aligned runs of a multiple of 16 bytes) */
    const bool contiguous = p.lo == 0 && p.c == 0 && p.ctop == 0 && p.cb == 0 && p.bitrev == 0 && tb >= 1 &&
                            ((reinterpret_cast<uintptr_t>(P.src) | reinterpret_cast<uintptr_t>(P.dst)) & 15u) == 0;
    if (few_ctas) additive_tile_pass<INV, 512, 1><<<(unsigned)nblocks, 512, smem, c->stream>>>(P);
    else if (contiguous && tile_tma_enabled())
        additive_tile_pass<INV, B200_TILE_THREADS, B200_TILE_MINBLOCKS, true><<<(unsigned)nblocks, B200_TILE_THREADS, smem, c->stream>>>(P);
    else additive_tile_pass<INV, B200_TILE_THREADS, B200_TILE_MINBLOCKS><<<(unsigned)nblocks, B200_TILE_THREADS, smem, c->stream>>>(P);
    B200_LAUNCH_CHECK(c);
    return 0;
}

__global__ void fill_kernel(gf192 *dst, const gf192 *src_elem, size_t n)
{
    const gf192 v = src_elem ? *src_elem : gf_zero();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

static int upload_taus(b200_ctx *c, const AdditivePlan &pl, const b200_domain *L, uint32_t d, uint32_t m, gf192 **dev)
{
    const uint32_t rows = 1 + (m - d);
    const size_t bytes = (size_t)rows * std::max(d, 1u) * sizeof(gf192);
    const gf192 *basis = (const gf192 *)L->basis;
    std::string key((const char *)L->shift, sizeof(gf192));
    if (m > d) key.append((const char *)(basis + d), (size_t)(m - d) * sizeof(gf192));
    AdditivePlan &plan = const_cast<AdditivePlan &>(pl);
    auto hit = plan.taus_cache.find(key);
    if (hit != plan.taus_cache.end()) { *dev = hit->second; return 0; }
    gf192 *h = (gf192 *)staging_get(c, std::max(bytes, (size_t)4096));
    if (!h) return set_err(c, B200_ERR_NOMEM, "cannot allocate tau staging");
    B200_CUDA(c, cudaStreamSynchronize(c->stream)); /* staging buffer may still be in flight */
    if (plan.taus_cache.size() >= 256) {             /* bounded: drop everything (the stream is idle here) */
        for (auto &kv : plan.taus_cache) cudaFree(kv.second);
        plan.taus_cache.clear();
    }
    gf192 *dv = nullptr;
    if (cudaMalloc(&dv, std::max(bytes, (size_t)256)) != cudaSuccess) { cudaGetLastError(); return set_err(c, B200_ERR_NOMEM, "cannot allocate tau table"); }
    gf192 s; memcpy(&s, L->shift, sizeof s);
    compute_taus(pl, s, h);
    for (uint32_t k = 0; k < m - d; ++k) compute_taus(pl, basis[d + k], h + (size_t)(1 + k) * d);
    B200_CUDA(c, cudaMemcpyAsync(dv, h, bytes, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));  /* the staging buffer is reused by the next miss */
    plan.taus_cache[key] = dv;
    *dev = dv;
    return 0;
}

/* `batch` independent polynomials over the same domain (Ligero's interleaved rows): row r reads coeffs + r * n_coeffs and
   writes out + r * 2^m.  Every pass treats the rows as further cosets of the tile grid, so a batch of small transforms
   fills the machine in the same handful of launches one transform needs. */
int additive_fft(b200_ctx *c, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf out, size_t batch)
{
    if (!L || L->kind != B200_DOMAIN_AFFINE_SUBSPACE)
        return set_err(c, B200_ERR_INVALID, "additive FFT needs an affine subspace domain");
    const uint32_t m = L->dim;
    if (m > 31) return set_err(c, B200_ERR_INVALID, "domain dimension %u too large", m);
    const size_t n = (size_t)1 << m;
    if (n_coeffs > n) return set_err(c, B200_ERR_INVALID, "more coefficients (%zu) than domain elements (%zu)", n_coeffs, n);
    gf192 *o = (gf192 *)out;
    if (batch == 0) return 0;
    if (batch > ((size_t)1 << 20) || (batch << m) > ((size_t)1 << 34))
        return set_err(c, B200_ERR_INVALID, "batch of %zu transforms of 2^%u points is too large", batch, m);
    if (n_coeffs <= 1) { /* constant polynomials */
        for (size_t r = 0; r < batch; ++r) {
            fill_kernel<<<std::min<size_t>((n + 255) / 256, 1184), 256, 0, c->stream>>>(o + r * n, n_coeffs ? (const gf192 *)coeffs + r : nullptr, n);
            B200_LAUNCH_CHECK(c);
        }
        return 0;
    }
    const uint32_t d = ceil_log2(n_coeffs);
    std::shared_ptr<AdditivePlan> pl;
    int rc = get_plan(c, (const gf192 *)L->basis, d, false, &pl);
    if (rc) return rc;
    gf192 *taus = nullptr;
    rc = upload_taus(c, *pl, L, d, m, &taus);
    if (rc) return rc;

    const size_t nd = (size_t)1 << d;
    const uint32_t B = (uint32_t)batch;
    gf192 *X = (gf192 *)scratch_get(c, 0, batch * nd * sizeof(gf192));
    if (!X) return set_err(c, B200_ERR_NOMEM, "cannot allocate FFT work array");
    if (n_coeffs == nd || batch == 1) {
        B200_CUDA(c, cudaMemcpyAsync(X, (const void *)coeffs, batch * n_coeffs * sizeof(gf192), cudaMemcpyDeviceToDevice, c->stream));
        if (n_coeffs < nd) B200_CUDA(c, cudaMemsetAsync(X + n_coeffs, 0, (nd - n_coeffs) * sizeof(gf192), c->stream));
    } else {                                         /* zero-pad every row to 2^d coefficients */
        B200_CUDA(c, cudaMemsetAsync(X, 0, batch * nd * sizeof(gf192), c->stream));
        B200_CUDA(c, cudaMemcpy2DAsync(X, nd * sizeof(gf192), (const void *)coeffs, n_coeffs * sizeof(gf192), n_coeffs * sizeof(gf192),
                                       batch, cudaMemcpyDeviceToDevice, c->stream));
    }

    PassParams P; memset(&P, 0, sizeof P);
    P.tw = pl->tw; P.sums = pl->sums; P.taus = taus; P.n_extra = m - d;

    /* forward half on X (shared by all cosets): fft.tcc:56-83 */
    std::vector<uint32_t> steps;
    for (uint32_t j = 0; j < d; ++j) {
        steps.push_back((OP_TW << 8) | j);
        for (int s = (int)d - 2; s >= (int)j; --s) steps.push_back((OP_TAY << 8) | (uint32_t)s);
    }
    for (const Pass &p : schedule(steps, d)) {
        P.src = X; P.dst = X; P.src_coset_stride_log = (B > 1) ? d : 32; P.src_coset_shr = 0; P.n_extra = 0;
        rc = launch_pass<false>(c, p, P, d, B);
        if (rc) return rc;
    }
    /* butterfly half per coset: bit-reversed load from X (fft.tcc:99), stages 0..d-1 (fft.tcc:101-120) */
    P.n_extra = m - d;
    const uint32_t n_cosets = B << (m - d);          /* coset index = row * 2^(m-d) + coset of the row */
    const uint32_t tb = std::min(kTileBits, d);
    uint32_t first_hi = d <= tb ? d : tb - kCoalesceBits;
    {
        Pass p; p.lo = 0; p.hi = first_hi; p.c = 0; p.ctop = d <= tb ? 0 : kCoalesceBits; p.bitrev = 1;
        const bool spread = !cluster_disabled() && d >= 9 && (uint64_t)n_cosets * 8 <= (uint64_t)c->sm_count;
        if (d < kTileBits && !spread) {           /* small transform: several cosets per CTA, tau table <= 8 KiB */
            p.cb = std::min(kTileBits - d, m - d);    /* (with few cosets each one is spread over a cluster instead) */
            while (p.cb > 0 && ((size_t)d << p.cb) * sizeof(gf192) > 8192) --p.cb;
        }
        for (uint32_t jj = 0; jj < first_hi; ++jj) p.ops.push_back((OP_BF << 8) | jj);
        P.src = X; P.dst = o;
        if (B > 1) { P.src_coset_stride_log = d; P.src_coset_shr = m - d; }   /* the cosets of a row share the row's X */
        else { P.src_coset_stride_log = 32; P.src_coset_shr = 0; }
        rc = launch_pass<false>(c, p, P, d, n_cosets);
        if (rc) return rc;
    }
    steps.clear();
    for (uint32_t jj = first_hi; jj < d; ++jj) steps.push_back((OP_BF << 8) | jj);
    for (const Pass &p : schedule(steps, d)) {
        P.src = o; P.dst = o; P.src_coset_stride_log = d; P.src_coset_shr = 0;
        rc = launch_pass<false>(c, p, P, d, n_cosets);
        if (rc) return rc;
    }
    return 0;
}

int additive_ifft(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf out, size_t batch)
{
    if (!H || H->kind != B200_DOMAIN_AFFINE_SUBSPACE)
        return set_err(c, B200_ERR_INVALID, "additive IFFT needs an affine subspace domain");
    const uint32_t m = H->dim;
    if (m > 31) return set_err(c, B200_ERR_INVALID, "domain dimension %u too large", m);
    const size_t n = (size_t)1 << m;
    if (batch == 0) return 0;
    if (batch > ((size_t)1 << 20) || (batch << m) > ((size_t)1 << 34))
        return set_err(c, B200_ERR_INVALID, "batch of %zu transforms of 2^%u points is too large", batch, m);
    if (m == 0) {
        B200_CUDA(c, cudaMemcpyAsync((void *)out, (const void *)evals, batch * sizeof(gf192), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    const uint32_t B = (uint32_t)batch;
    std::shared_ptr<AdditivePlan> pl;
    int rc = get_plan(c, (const gf192 *)H->basis, m, true, &pl);
    if (rc) return rc;
    gf192 *taus = nullptr;
    rc = upload_taus(c, *pl, H, m, m, &taus);
    if (rc) return rc;

    gf192 *X = (gf192 *)scratch_get(c, 0, batch * n * sizeof(gf192));
    if (!X) return set_err(c, B200_ERR_NOMEM, "cannot allocate IFFT work array");
    gf192 *o = (gf192 *)out;
    PassParams P; memset(&P, 0, sizeof P);
    P.tw = pl->itw; P.sums = pl->sums; P.taus = taus; P.n_extra = 0;
    P.src_coset_stride_log = (B > 1) ? m : 32;       /* batch rows are the "cosets" of the tile grid */

    /* inverse butterflies, large stride first (fft.tcc:139-168): evals -> X */
    const uint32_t tb = std::min(kTileBits, m);
    std::vector<uint32_t> steps;
    const uint32_t last_hi = m <= tb ? m : tb - kCoalesceBits; /* stages [0, last_hi) go into the bit-reversing pass */
    for (int jj = (int)m - 1; jj >= (int)last_hi; --jj) steps.push_back((OP_BF << 8) | (uint32_t)jj);
    const gf192 *cur = (const gf192 *)evals;
    for (const Pass &p : schedule(steps, m)) {
        P.src = cur; P.dst = X;
        rc = launch_pass<true>(c, p, P, m, B);
        if (rc) return rc;
        cur = X;
    }
    {   /* last inverse stages, then store bit-reversed (fft.tcc:170) into `out` */
        Pass p; p.lo = 0; p.hi = last_hi; p.c = 0; p.ctop = m <= tb ? 0 : kCoalesceBits; p.bitrev = 2;
        for (int jj = (int)last_hi - 1; jj >= 0; --jj) p.ops.push_back((OP_BF << 8) | (uint32_t)jj);
        P.src = cur; P.dst = o;
        rc = launch_pass<true>(c, p, P, m, B);
        if (rc) return rc;
    }
    /* radix combinations + inverse twists, deepest level first (fft.tcc:172-201) */
    steps.clear();
    for (int jd = (int)m - 1; jd >= 0; --jd) {
        for (int s = jd; s <= (int)m - 2; ++s) steps.push_back((OP_TAY << 8) | (uint32_t)s);
        steps.push_back((OP_TW << 8) | (uint32_t)jd);
    }
    for (const Pass &p : schedule(steps, m)) {
        P.src = o; P.dst = o;
        rc = launch_pass<true>(c, p, P, m, B);
        if (rc) return rc;
    }
    return 0;
}

/* Fused low-degree extension: evaluations over H (dimension d) -> evaluations over L (dimension m >= d) of the same
   polynomial, when L's first d basis vectors are H's — the reference's IFFT_over_field_subset(H) followed by
   FFT_over_field_subset(L) (r1cs_rs_iop.tcc:449-479) without ever forming the monomial coefficients.
   additive_IFFT is (inverse butterflies, bit reversal, radix combinations + inverse twists) and additive_FFT is (twists +
   radix conversions, bit reversal, butterflies); the two middle parts depend only on the basis and are exact inverses of each
   other, and the two bit reversals cancel.  What remains: d inverse butterfly stages with H's recursed shifts, then d
   butterfly stages per coset of span(basis[0..d)) in L with that coset's recursed shifts.  90 * 2^d instead of 130 * 2^d
   multiplications for d = 20, m = 23, no twist tables, and — no bit-reversed gather to coalesce — 12 stages per
   contiguous tile pass: 2 + 2 launches instead of 69.  Same batch convention as additive_fft. */
/* The two halves of the extension as separate steps, X = the basis-dependent intermediate coefficients (2^d per row; what the
   reference holds after additive_IFFT's inverse butterflies, fft.tcc:139-168, before its bit reversal).  X is linear in the
   polynomial, so callers may combine several interpolations before one evaluation (the sumcheck quotient). */
static int lde_inverse_half(b200_ctx *c, const std::shared_ptr<AdditivePlan> &pl, const b200_domain *H, const gf192 *evals, gf192 *X, uint32_t B)
{
    const uint32_t d = H->dim;
    gf192 *taus_h = nullptr;
    int rc = upload_taus(c, *pl, H, d, d, &taus_h);
    if (rc) return rc;
    PassParams P; memset(&P, 0, sizeof P);
    P.sums = pl->sums;
    const uint32_t tb = std::min(kTileBits, d);
    /* inverse butterflies over H, large stride first (fft.tcc:139-168): evals -> X, natural order throughout */
    P.taus = taus_h; P.n_extra = 0; P.src_coset_stride_log = (B > 1) ? d : 32; P.src_coset_shr = 0;
    std::vector<uint32_t> steps;
    for (int jj = (int)d - 1; jj >= (int)tb; --jj) steps.push_back((OP_BF << 8) | (uint32_t)jj);
    const gf192 *cur = evals;
    for (const Pass &p : schedule(steps, d)) {
        P.src = cur; P.dst = X;
        rc = launch_pass<true>(c, p, P, d, B);
        if (rc) return rc;
        cur = X;
    }
    {   /* stages tb-1 .. 0 in one contiguous tile */
        Pass p; p.lo = 0; p.hi = tb; p.c = 0; p.ctop = 0; p.bitrev = 0;
        for (int jj = (int)tb - 1; jj >= 0; --jj) p.ops.push_back((OP_BF << 8) | (uint32_t)jj);
        P.src = cur; P.dst = X;
        rc = launch_pass<true>(c, p, P, d, B);
        if (rc) return rc;
    }
    return 0;
}
static int lde_forward_half(b200_ctx *c, const std::shared_ptr<AdditivePlan> &pl, uint32_t d, const b200_domain *L, const gf192 *X, gf192 *o, uint32_t B)
{
    const uint32_t m = L->dim;
    gf192 *taus_l = nullptr;
    int rc = upload_taus(c, *pl, L, d, m, &taus_l);
    if (rc) return rc;
    PassParams P; memset(&P, 0, sizeof P);
    P.sums = pl->sums;
    const uint32_t tb = std::min(kTileBits, d);
    /* butterflies per coset of L (fft.tcc:101-120): X -> out */
    P.taus = taus_l; P.n_extra = m - d;
    const uint32_t n_cosets = B << (m - d);
    {
        Pass p; p.lo = 0; p.hi = tb; p.c = 0; p.ctop = 0; p.bitrev = 0;
        const bool spread = !cluster_disabled() && d >= 9 && (uint64_t)n_cosets * 8 <= (uint64_t)c->sm_count;
        if (d < kTileBits && !spread) {               /* small transform: several cosets per CTA, tau table <= 8 KiB */
            p.cb = std::min(kTileBits - d, m - d);
            while (p.cb > 0 && ((size_t)d << p.cb) * sizeof(gf192) > 8192) --p.cb;
        }
        for (uint32_t jj = 0; jj < tb; ++jj) p.ops.push_back((OP_BF << 8) | jj);
        P.src = X; P.dst = o;
        if (B > 1) { P.src_coset_stride_log = d; P.src_coset_shr = m - d; }   /* the cosets of a row share the row's X */
        else { P.src_coset_stride_log = 32; P.src_coset_shr = 0; }
        rc = launch_pass<false>(c, p, P, d, n_cosets);
        if (rc) return rc;
    }
    std::vector<uint32_t> steps;
    for (uint32_t jj = tb; jj < d; ++jj) steps.push_back((OP_BF << 8) | jj);
    for (const Pass &p : schedule(steps, d)) {
        P.src = o; P.dst = o; P.src_coset_stride_log = d; P.src_coset_shr = 0;
        rc = launch_pass<false>(c, p, P, d, n_cosets);
        if (rc) return rc;
    }
    return 0;
}
static int lde_check(b200_ctx *c, const b200_domain *L, uint32_t d, size_t batch)
{
    if (!L || L->kind != B200_DOMAIN_AFFINE_SUBSPACE)
        return set_err(c, B200_ERR_INVALID, "the fused extension needs affine subspace domains");
    const uint32_t m = L->dim;
    if (m > 31 || d > m) return set_err(c, B200_ERR_INVALID, "extension from dimension %u to %u is not possible", d, m);
    if (batch > ((size_t)1 << 20) || (batch << m) > ((size_t)1 << 34))
        return set_err(c, B200_ERR_INVALID, "batch of %zu extensions to 2^%u points is too large", batch, m);
    return 0;
}

int additive_lde(b200_ctx *c, const b200_domain *H, const b200_domain *L, b200_buf evals, b200_buf out, size_t batch)
{
    if (!H || H->kind != B200_DOMAIN_AFFINE_SUBSPACE) return set_err(c, B200_ERR_INVALID, "the fused extension needs two affine subspace domains");
    const uint32_t d = H->dim;
    int rc = lde_check(c, L, d, batch);
    if (rc) return rc;
    const uint32_t m = L->dim;
    if (d && memcmp(H->basis, L->basis, (size_t)d * sizeof(gf192)) != 0)
        return set_err(c, B200_ERR_INVALID, "the codeword domain does not extend the basis of the interpolation domain");
    if (batch == 0) return 0;
    const size_t nd = (size_t)1 << d, n = (size_t)1 << m;
    gf192 *o = (gf192 *)out;
    if (d == 0) {                                    /* constant polynomials */
        for (size_t r = 0; r < batch; ++r) {
            fill_kernel<<<std::min<size_t>((n + 255) / 256, 1184), 256, 0, c->stream>>>(o + r * n, (const gf192 *)evals + r, n);
            B200_LAUNCH_CHECK(c);
        }
        return 0;
    }
    std::shared_ptr<AdditivePlan> pl;
    rc = get_plan(c, (const gf192 *)H->basis, d, false, &pl, /*need_twists=*/false);
    if (rc) return rc;
    gf192 *X = (gf192 *)scratch_get(c, 0, batch * nd * sizeof(gf192));
    if (!X) return set_err(c, B200_ERR_NOMEM, "cannot allocate extension work array");
    rc = lde_inverse_half(c, pl, H, (const gf192 *)evals, X, (uint32_t)batch);
    if (rc) return rc;
    return lde_forward_half(c, pl, d, L, X, o, (uint32_t)batch);
}
/* interpolation half alone: evaluations over H -> intermediate coefficients (2^H->dim per row) */
int additive_lde_interpolate(b200_ctx *c, const b200_domain *H, b200_buf evals, b200_buf coeffs_out, size_t batch)
{
    if (!H || H->kind != B200_DOMAIN_AFFINE_SUBSPACE) return set_err(c, B200_ERR_INVALID, "the fused extension needs affine subspace domains");
    const uint32_t d = H->dim;
    if (d > 31 || batch > ((size_t)1 << 20) || (batch << d) > ((size_t)1 << 34)) return set_err(c, B200_ERR_INVALID, "interpolation of %zu rows of 2^%u is too large", batch, d);
    if (batch == 0) return 0;
    if (d == 0) { B200_CUDA(c, cudaMemcpyAsync((void *)coeffs_out, (const void *)evals, batch * sizeof(gf192), cudaMemcpyDeviceToDevice, c->stream)); return 0; }
    if (evals == coeffs_out) return set_err(c, B200_ERR_INVALID, "the interpolation half does not run in place");
    std::shared_ptr<AdditivePlan> pl;
    int rc = get_plan(c, (const gf192 *)H->basis, d, false, &pl, /*need_twists=*/false);
    if (rc) return rc;
    return lde_inverse_half(c, pl, H, (const gf192 *)evals, (gf192 *)coeffs_out, (uint32_t)batch);
}
/* evaluation half alone: intermediate coefficients w.r.t. L's first d basis vectors -> evaluations over L */
int additive_lde_evaluate(b200_ctx *c, const b200_domain *L, uint32_t d, b200_buf coeffs, b200_buf out, size_t batch)
{
    int rc = lde_check(c, L, d, batch);
    if (rc) return rc;
    if (batch == 0) return 0;
    const size_t n = (size_t)1 << L->dim;
    if (d == 0) {
        for (size_t r = 0; r < batch; ++r) {
            fill_kernel<<<std::min<size_t>((n + 255) / 256, 1184), 256, 0, c->stream>>>((gf192 *)out + r * n, (const gf192 *)coeffs + r, n);
            B200_LAUNCH_CHECK(c);
        }
        return 0;
    }
    if (coeffs == out) return set_err(c, B200_ERR_INVALID, "the evaluation half does not run in place");
    std::shared_ptr<AdditivePlan> pl;
    rc = get_plan(c, (const gf192 *)L->basis, d, false, &pl, /*need_twists=*/false);
    if (rc) return rc;
    return lde_forward_half(c, pl, d, L, (const gf192 *)coeffs, (gf192 *)out, (uint32_t)batch);
}

/* ---------------------------------------------------------------------------------------------- */
/* element-wise helpers                                                                           */
__global__ void gf192_mul_kernel(gf192 *out, const gf192 *a, const gf192 *b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = gf_mul(a[i], b[i]);
}
/* Element-wise inversion with Montgomery's trick per thread (the reference's batch_inverse, algebra/utils.tcc:38-69, does
   the same over the whole vector sequentially): RUN consecutive elements share one Itoh-Tsujii inversion (191 squarings +
   12 multiplications) at the price of 3 multiplications each (the run lives in local memory: L1-resident).  Zero maps to zero like x^(2^192-2) and like the
   reference's has_zeroes path (utils.tcc:73-103). */
#ifndef B200_INV_RUN
#define B200_INV_RUN 32          /* measured at 2^23 elements: 8 -> 4.17 ms, 16 -> 3.21 ms, 32 -> 2.35 ms (profiles/r2_time_ops_v4.txt) */
#endif
__global__ void __launch_bounds__(128) gf192_inv_kernel(gf192 *out, const gf192 *a, size_t n)
{
    const size_t runs = (n + B200_INV_RUN - 1) / B200_INV_RUN;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < runs; t += (size_t)gridDim.x * blockDim.x) {
        const size_t base = t * B200_INV_RUN;
        const int cnt = (int)((n - base < B200_INV_RUN) ? n - base : B200_INV_RUN);
        gf192 x[B200_INV_RUN], pre[B200_INV_RUN];
        gf192 acc = gf_one();
#pragma unroll 1
        for (int e = 0; e < cnt; ++e) {
            x[e] = a[base + e];
            pre[e] = acc;                                      /* product of the non-zero elements before e */
            if (!gf_is_zero(x[e])) acc = gf_mul(acc, x[e]);
        }
        gf192 inv = gf_inv(acc);
#pragma unroll 1
        for (int e = cnt - 1; e >= 0; --e) {
            if (gf_is_zero(x[e])) { out[base + e] = gf_zero(); continue; }
            out[base + e] = gf_mul(inv, pre[e]);
            inv = gf_mul(inv, x[e]);
        }
    }
}

int gf192_mul_vec(b200_ctx *c, b200_buf out, b200_buf a, b200_buf b, size_t n)
{
    if (!n) return 0;
    gf192_mul_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, c->stream>>>(
        (gf192 *)out, (const gf192 *)a, (const gf192 *)b, n);
    B200_LAUNCH_CHECK(c);
    return 0;
}
int gf192_inv_vec(b200_ctx *c, b200_buf out, b200_buf a, size_t n)
{
    if (!n) return 0;
    const size_t runs = (n + B200_INV_RUN - 1) / B200_INV_RUN;
    gf192_inv_kernel<<<(unsigned)std::min<size_t>((runs + 127) / 128, 148 * 16), 128, 0, c->stream>>>(
        (gf192 *)out, (const gf192 *)a, n);
    B200_LAUNCH_CHECK(c);
    return 0;
}

} // namespace b200
