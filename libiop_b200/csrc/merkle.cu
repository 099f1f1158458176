/*
 * merkle.cu — BLAKE2b Merkle-tree build for the BCS transformation, sm_100a.
 *
 * Replaces merkle_tree::construct_with_leaves_serialized_by_cosets (libiop/bcs/merkle_tree.tcc:93-151),
 * compute_inner_nodes (merkle_tree.tcc:201-229) and the hash callables injected into it:
 * blake2b_field_element_hash (libiop/bcs/hashing/blake2b.tcc:141-160), blake2b_leafhash::zk_hash
 * (blake2b.tcc:127-136) and blake2b_two_to_one_hash (libiop/bcs/hashing/blake2b.cpp:28-48).
 * BLAKE2b itself is RFC 7693 as provided by libsodium's crypto_generichash_blake2b (unkeyed,
 * digest_length = outlen, fanout = depth = 1).
 *
 * Leaf i hashes, for oracle k (registration order) and j < coset_size, the raw 24-byte element at
 * pos(i,j) = i*coset + j (affine subspace, subspace.tcc:87-91) or i + j*leaves (multiplicative coset,
 * subgroup.tcc:192-197), laid out slice[j + k*coset] (merkle_tree.tcc:127-134).  Nodes are stored in
 * heap order, root at 0, leaf i at leaves-1+i (merkle_tree.tcc:145).
 *
 * B200 design: one thread per hash (BLAKE2b is 64-bit ARX: no cross-lane traffic is needed, the
 * 16-word state and the 16-word message block live in registers).  The leaf kernel gathers the
 * column slice straight from the oracle arrays; because the tree is heap-ordered, a CTA that hashes
 * 2^k consecutive leaves also owns the k levels above them, so the lower levels are reduced inside
 * shared memory in the same launch and only the top of the tree runs as small per-level launches.
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctx.h"

namespace b200 {

__constant__ uint64_t c_blake_iv[8] = {
    0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};

#define BLAKE_IV0 0x6a09e667f3bcc908ULL
#define BLAKE_IV1 0xbb67ae8584caa73bULL
#define BLAKE_IV2 0x3c6ef372fe94f82bULL
#define BLAKE_IV3 0xa54ff53a5f1d36f1ULL
#define BLAKE_IV4 0x510e527fade682d1ULL
#define BLAKE_IV5 0x9b05688c2b3e6c1fULL
#define BLAKE_IV6 0x1f83d9abfb41bd6bULL
#define BLAKE_IV7 0x5be0cd19137e2179ULL

__device__ __forceinline__ uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
__device__ __forceinline__ uint64_t rotr32x(uint64_t x) { return (x >> 32) | (x << 32); }
/* byte-granular rotations as byte permutes */
__device__ __forceinline__ uint64_t rotr24(uint64_t x)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return (uint64_t)__byte_perm(lo, hi, 0x6543) | ((uint64_t)__byte_perm(lo, hi, 0x2107) << 32);
}
__device__ __forceinline__ uint64_t rotr16(uint64_t x)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return (uint64_t)__byte_perm(lo, hi, 0x5432) | ((uint64_t)__byte_perm(lo, hi, 0x1076) << 32);
}

/* BLAKE2b is 64-bit add / xor / rotate: on this SM all three run on the ALU pipe (IADD3 + IADD3.X, LOP3, SHF/PRMT), which the
   hash kernels saturate (ncu, round 2: pow_grind_kernel ALU pipe 97.6 %, FMA pipe 9.6 %).  B200_BLAKE_IMAD moves 64-bit
   additions to the idle FMA pipe: x + y = mad.wide.u32(y.lo, 1, x) then hi += y.hi as mad.lo.u32, with the multiplier 1
   read from constant memory so that ptxas keeps the multiply-adds (IMAD.WIDE.U32 + IMAD).  Level 1: the two c += d of each
   G; level 2: all four additions.  Measured in profiles/r2_variants.md. */
#ifndef B200_BLAKE_IMAD
#define B200_BLAKE_IMAD 0
#endif
__constant__ uint32_t c_blake_one = 1;
__device__ __forceinline__ uint64_t add64_fma(uint64_t x, uint64_t y, uint32_t one)
{
    uint64_t r;
    asm("{\n\t.reg .u32 ylo, yhi, tlo, thi, rhi;\n\t.reg .u64 t;\n\t"
        "mov.b64 {ylo, yhi}, %2;\n\t"
        "mad.wide.u32 t, ylo, %3, %1;\n\t"
        "mov.b64 {tlo, thi}, t;\n\t"
        "mad.lo.u32 rhi, yhi, %3, thi;\n\t"
        "mov.b64 %0, {tlo, rhi};\n\t}"
        : "=l"(r) : "l"(x), "l"(y), "r"(one));
    return r;
}
#if B200_BLAKE_IMAD >= 2
#define BLAKE_ADD3(a, b, x) add64_fma(add64_fma(a, b, blake_one), (x), blake_one)
#else
#define BLAKE_ADD3(a, b, x) ((a) + (b) + (x))
#endif
#if B200_BLAKE_IMAD >= 1
#define BLAKE_ADD2(c, d) add64_fma(c, d, blake_one)
#define BLAKE_ONE_DECL const uint32_t blake_one = c_blake_one;
#else
#define BLAKE_ADD2(c, d) ((c) + (d))
#define BLAKE_ONE_DECL
#endif

#define BLAKE_G(a, b, c, d, x, y)                              \
    do {                                                       \
        a = BLAKE_ADD3(a, b, x); d = rotr32x(d ^ a);           \
        c = BLAKE_ADD2(c, d);    b = rotr24(b ^ c);            \
        a = BLAKE_ADD3(a, b, y); d = rotr16(d ^ a);            \
        c = BLAKE_ADD2(c, d);    b = rotr64(b ^ c, 63);        \
    } while (0)

#define BLAKE_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                      \
        BLAKE_G(v0, v4, v8, v12, m[s0], m[s1]);                                               \
        BLAKE_G(v1, v5, v9, v13, m[s2], m[s3]);                                               \
        BLAKE_G(v2, v6, v10, v14, m[s4], m[s5]);                                              \
        BLAKE_G(v3, v7, v11, v15, m[s6], m[s7]);                                              \
        BLAKE_G(v0, v5, v10, v15, m[s8], m[s9]);                                              \
        BLAKE_G(v1, v6, v11, v12, m[s10], m[s11]);                                            \
        BLAKE_G(v2, v7, v8, v13, m[s12], m[s13]);                                             \
        BLAKE_G(v3, v4, v9, v14, m[s14], m[s15]);                                             \
    } while (0)

/* one BLAKE2b compression; m[] must be indexable with compile-time constants (registers) */
__device__ __forceinline__ void blake2b_compress(uint64_t h[8], const uint64_t m[16], uint64_t t, bool last)
{
    BLAKE_ONE_DECL
    uint64_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint64_t v8 = BLAKE_IV0, v9 = BLAKE_IV1, v10 = BLAKE_IV2, v11 = BLAKE_IV3;
    uint64_t v12 = BLAKE_IV4 ^ t, v13 = BLAKE_IV5, v14 = last ? ~BLAKE_IV6 : BLAKE_IV6, v15 = BLAKE_IV7;
    BLAKE_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    BLAKE_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    BLAKE_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    BLAKE_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    BLAKE_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    BLAKE_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    BLAKE_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    BLAKE_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    BLAKE_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    BLAKE_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    BLAKE_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    BLAKE_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

__device__ __forceinline__ void blake2b_init(uint64_t h[8], uint32_t outlen)
{
    h[0] = BLAKE_IV0 ^ 0x01010000ULL ^ (uint64_t)outlen;
    h[1] = BLAKE_IV1; h[2] = BLAKE_IV2; h[3] = BLAKE_IV3; h[4] = BLAKE_IV4; h[5] = BLAKE_IV5; h[6] = BLAKE_IV6; h[7] = BLAKE_IV7;
}

/* Hash a message delivered word by word: word(w) returns the w-th little-endian u64 of the message,
   n_words its length in whole words (all messages on this path are multiples of 8 bytes). */
template <typename WordFn>
__device__ __forceinline__ void blake2b_words(uint64_t h[8], uint32_t outlen, size_t n_words, WordFn word)
{
    blake2b_init(h, outlen);
    uint64_t m[16];
    size_t done = 0;
    for (;;) {
        const size_t remaining = n_words - done;
        const bool last = remaining <= 16;
#pragma unroll
        for (int t = 0; t < 16; ++t) m[t] = ((size_t)t < remaining) ? word(done + t) : 0ull;
        const size_t consumed = last ? remaining : 16;
        done += consumed;
        blake2b_compress(h, m, (uint64_t)done * 8, last);
        if (last) break;
    }
}

/* two-to-one hash of 2 x digest_words u64 (digest_len <= 64, multiple of 8) */
__device__ __forceinline__ void blake2b_two_to_one(uint64_t h[8], uint32_t outlen, const uint64_t *left, const uint64_t *right, uint32_t dw)
{
    blake2b_init(h, outlen);
    uint64_t m[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        uint64_t v = 0;
        if ((uint32_t)t < dw) v = left[t];
        else if ((uint32_t)t < 2 * dw) v = right[t - dw];
        m[t] = v;
    }
    blake2b_compress(h, m, (uint64_t)dw * 16, true);
}

struct LeafParams {
    const uint64_t *oracles[64];
    uint32_t n_oracles;
    uint32_t coset;        /* power of two */
    uint32_t coset_log;
    int kind;
    size_t leaves;
    const uint64_t *salts; /* device, leaves * salt_words, or null */
    uint32_t salt_words, digest_words, digest_len;
    uint64_t *nodes;       /* (2*leaves-1) * digest_words */
    uint32_t fused_levels; /* levels reduced inside the CTA (CTA covers 2^fused_levels leaves) */
    const uint64_t *const *oracle_table; /* device array when n_oracles > 64 */
};

#define MERKLE_CTA 256

__global__ void __launch_bounds__(MERKLE_CTA) merkle_leaf_kernel(const LeafParams P)
{
    __shared__ uint64_t sdig[MERKLE_CTA * 8];
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const uint32_t dw = P.digest_words;
    uint64_t h[8];
    if (i < P.leaves) {
        const size_t n_words = (size_t)3 * P.n_oracles * P.coset;
        const uint32_t cmask = P.coset - 1;
        auto word = [&](size_t w) -> uint64_t {
            const size_t e = w / 3;                    /* element index in the slice: j + k*coset */
            const uint32_t limb = (uint32_t)(w - 3 * e);
            const uint32_t j = (uint32_t)e & cmask;
            const size_t k = e >> P.coset_log;
            const size_t pos = (P.kind == B200_DOMAIN_MULTIPLICATIVE_COSET) ? (i + (size_t)j * P.leaves) : (i * P.coset + j);
            const uint64_t *base = P.oracle_table ? P.oracle_table[k] : P.oracles[k];
            return __ldg(base + pos * 3 + limb);
        };
        blake2b_words(h, P.digest_len, n_words, word);
        if (P.salts) {                                  /* zk: H(H(leaf) || salt)  (blake2b.tcc:127-136) */
            uint64_t inner[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) inner[t] = h[t];
            const uint64_t *salt = P.salts + i * P.salt_words;
            const uint32_t sw = P.salt_words;
            auto word2 = [&](size_t w) -> uint64_t {
                uint64_t v = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) if (w == (size_t)t) v = inner[t];
                if (w >= dw) v = salt[w - dw];
                return v;
            };
            blake2b_words(h, P.digest_len, (size_t)dw + sw, word2);
        }
        uint64_t *dst = P.nodes + (P.leaves - 1 + i) * dw;
#pragma unroll
        for (int t = 0; t < 8; ++t) if ((uint32_t)t < dw) { dst[t] = h[t]; sdig[threadIdx.x * 8 + t] = h[t]; }
    }
    /* reduce the CTA's subtree in shared memory: level l has MERKLE_CTA >> l nodes */
    uint32_t width = MERKLE_CTA;
    size_t level_nodes = P.leaves;              /* nodes of the current level in the whole tree */
    for (uint32_t l = 0; l < P.fused_levels; ++l) {
        __syncthreads();
        width >>= 1; level_nodes >>= 1;
        uint64_t r[8];
        const bool active = threadIdx.x < width && (blockIdx.x * (size_t)width + threadIdx.x) < level_nodes;
        if (active) blake2b_two_to_one(r, P.digest_len, &sdig[(2 * threadIdx.x) * 8], &sdig[(2 * threadIdx.x + 1) * 8], dw);
        __syncthreads();
        if (active) {
            const size_t node = (level_nodes - 1) + blockIdx.x * (size_t)width + threadIdx.x;
            uint64_t *dst = P.nodes + node * dw;
#pragma unroll
            for (int t = 0; t < 8; ++t) if ((uint32_t)t < dw) { dst[t] = r[t]; sdig[threadIdx.x * 8 + t] = r[t]; }
        }
    }
}

/* one level of inner nodes: node j in [first, first+count) = H(node 2j+1 || node 2j+2) (merkle_tree.tcc:209-218) */
__global__ void merkle_level_kernel(uint64_t *nodes, size_t first, size_t count, uint32_t dw, uint32_t digest_len)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    const size_t j = first + t;
    uint64_t l[8], r[8], h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        l[q] = ((uint32_t)q < dw) ? nodes[(2 * j + 1) * dw + q] : 0;
        r[q] = ((uint32_t)q < dw) ? nodes[(2 * j + 2) * dw + q] : 0;
    }
    blake2b_two_to_one(h, digest_len, l, r, dw);
#pragma unroll
    for (int q = 0; q < 8; ++q) if ((uint32_t)q < dw) nodes[j * dw + q] = h[q];
}

/* top of the tree (<= 2^10 nodes wide) in a single CTA */
__global__ void __launch_bounds__(1024) merkle_top_kernel(uint64_t *nodes, size_t width, uint32_t dw, uint32_t digest_len)
{
    /* width = number of nodes of the highest level already computed (power of two <= 2048) */
    for (size_t w = width >> 1; w >= 1; w >>= 1) {
        if (threadIdx.x < w) {
            const size_t j = (w - 1) + threadIdx.x;
            uint64_t l[8], r[8], h[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                l[q] = ((uint32_t)q < dw) ? nodes[(2 * j + 1) * dw + q] : 0;
                r[q] = ((uint32_t)q < dw) ? nodes[(2 * j + 2) * dw + q] : 0;
            }
            blake2b_two_to_one(h, digest_len, l, r, dw);
#pragma unroll
            for (int q = 0; q < 8; ++q) if ((uint32_t)q < dw) nodes[j * dw + q] = h[q];
        }
        __threadfence_block();
        __syncthreads();
    }
}


/* ---------------------------------------------------------------------------------------------- */
/* fast path (32-byte digests): one thread owns 8 consecutive leaves / nodes and the 3 levels above
   them (15 / 7 compressions, all lanes busy, no shared memory).                                  */
__device__ __forceinline__ void two_to_one_32(uint64_t out[4], const uint64_t l[4], const uint64_t r[4])
{
    uint64_t h[8];
    blake2b_init(h, 32);
    uint64_t m[16];
#pragma unroll
    for (int t = 0; t < 4; ++t) { m[t] = l[t]; m[4 + t] = r[t]; }
#pragma unroll
    for (int t = 8; t < 16; ++t) m[t] = 0;
    blake2b_compress(h, m, 64, true);
#pragma unroll
    for (int t = 0; t < 4; ++t) out[t] = h[t];
}

__device__ __forceinline__ void store_digest32(uint64_t *nodes, size_t heap_index, const uint64_t d[4])
{
    ulonglong2 *p = reinterpret_cast<ulonglong2 *>(nodes + heap_index * 4);
    p[0] = make_ulonglong2(d[0], d[1]);
    p[1] = make_ulonglong2(d[2], d[3]);
}

/* reduce 8 digests held in registers through three levels; `width` = nodes of the level of d[],
   `first` = index of d[0] within that level */
__device__ __forceinline__ void reduce8_32(uint64_t *nodes, uint64_t d[8][4], size_t width, size_t first)
{
    uint64_t e[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { two_to_one_32(e[q], d[2 * q], d[2 * q + 1]); store_digest32(nodes, (width >> 1) - 1 + (first >> 1) + q, e[q]); }
    uint64_t f[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q) { two_to_one_32(f[q], e[2 * q], e[2 * q + 1]); store_digest32(nodes, (width >> 2) - 1 + (first >> 2) + q, f[q]); }
    uint64_t g[4];
    two_to_one_32(g, f[0], f[1]);
    store_digest32(nodes, (width >> 3) - 1 + (first >> 3), g);
}

__global__ void __launch_bounds__(128) merkle_leaf8_kernel(const LeafParams P)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t * 8 >= P.leaves) return;
    uint64_t d[8][4];
    const size_t n_words = (size_t)3 * P.n_oracles * P.coset;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const size_t i = t * 8 + q;
        uint64_t h[8];
        blake2b_init(h, 32);
        /* stream the slice: for oracle k, for j < coset, 3 limbs (uniform control flow across the warp) */
        uint32_t k = 0, j = 0, limb = 0;
        const uint64_t *base = P.oracle_table ? P.oracle_table[0] : P.oracles[0];
        size_t pos = (P.kind == B200_DOMAIN_MULTIPLICATIVE_COSET) ? i : i * P.coset;
        const size_t jstep = (P.kind == B200_DOMAIN_MULTIPLICATIVE_COSET) ? P.leaves : 1;
        size_t done = 0;
        for (;;) {
            const size_t remaining = n_words - done;
            const bool last = remaining <= 16;
            uint64_t m[16];
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                uint64_t v = 0;
                if ((size_t)w < remaining) {
                    v = __ldg(base + (pos + (size_t)j * jstep) * 3 + limb);
                    if (++limb == 3) {
                        limb = 0;
                        if (++j == P.coset) { j = 0; ++k; if (k < P.n_oracles) base = P.oracle_table ? P.oracle_table[k] : P.oracles[k]; }
                    }
                }
                m[w] = v;
            }
            done += last ? remaining : 16;
            blake2b_compress(h, m, (uint64_t)done * 8, last);
            if (last) break;
        }
        if (P.salts) {                                  /* zk: H(H(leaf) || salt)  (blake2b.tcc:127-136) */
            const uint64_t *salt = P.salts + i * P.salt_words;
            uint64_t m[16];
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                uint64_t v = 0;
                if (w < 4) v = h[w];
                else if ((uint32_t)(w - 4) < P.salt_words) v = salt[w - 4];
                m[w] = v;
            }
            blake2b_init(h, 32);
            blake2b_compress(h, m, (uint64_t)(4 + P.salt_words) * 8, true);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) d[q][w] = h[w];
        store_digest32(P.nodes, P.leaves - 1 + i, h);
    }
    reduce8_32(P.nodes, d, P.leaves, t * 8);
}

/* ---------------------------------------------------------------------------------------------- */
/* wide leaves (Ligero: one column of several hundred interleaved rows = ~12 KB = 96 compressions per leaf,
   merkle_tree.tcc:123-145, and only 2^13 leaves): WARP-COOPERATIVE BLAKE2b.  Four lanes ("a quad") share one hash
   state, lane q holding column q of the 4x4 state matrix (a, b, c, d = v[q], v[4+q], v[8+q], v[12+q]).  The column step of
   a round is one G per lane; for the diagonal step b, c, d are rotated by 1, 2, 3 lanes inside the quad with shuffles,
   one G, and rotated back.  The 128-byte message block is fetched by the quad together (lane q loads words q, q+4, q+8,
   q+12 of the block: consecutive leaves of one row are consecutive in memory, so the 8 quads of a warp read 192
   contiguous bytes per row) and parked in shared memory, from which every lane picks the two words each of its G's
   needs.  8 leaves per warp instead of 8 leaves per THREAD: 2^13 leaves give 1024 warps (7 per SM) instead of 32.   */
__constant__ uint8_t c_blake_sigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

#define QUAD_MSG_STRIDE 17   /* u64 words per quad in shared memory (16 + 1 of padding: quads land on different banks) */

__device__ __forceinline__ uint64_t quad_shfl(uint64_t v, int src_lane)
{
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src_lane);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src_lane);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

/* One compression by a quad.  h0 / h1: this lane's two words of the chaining value (h[q], h[4+q]); msg: the quad's block in
   shared memory; sig[r] packs this lane's four message indices of round r (4 bits each: column x, y, diagonal x, y). */
__device__ __forceinline__ void quad_compress(uint64_t &h0, uint64_t &h1, const uint64_t *msg, const uint32_t sig[12], uint32_t q,
                                              uint64_t t, bool last)
{
    BLAKE_ONE_DECL
    uint64_t a = h0, b = h1;
    uint64_t c = q == 0 ? BLAKE_IV0 : q == 1 ? BLAKE_IV1 : q == 2 ? BLAKE_IV2 : BLAKE_IV3;
    uint64_t d = q == 0 ? (BLAKE_IV4 ^ t) : q == 1 ? BLAKE_IV5 : q == 2 ? (last ? ~BLAKE_IV6 : BLAKE_IV6) : BLAKE_IV7;
    const int lane = threadIdx.x & 31, base = lane & ~3;
    const int up1 = base | ((q + 1) & 3), up2 = base | ((q + 2) & 3), up3 = base | ((q + 3) & 3);
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        const uint32_t s = sig[r];
        BLAKE_G(a, b, c, d, msg[s & 15u], msg[(s >> 4) & 15u]);
        /* diagonalise: lane q takes b of column q+1, c of column q+2, d of column q+3 */
        b = quad_shfl(b, up1); c = quad_shfl(c, up2); d = quad_shfl(d, up3);
        BLAKE_G(a, b, c, d, msg[(s >> 8) & 15u], msg[(s >> 12) & 15u]);
        b = quad_shfl(b, up3); c = quad_shfl(c, up2); d = quad_shfl(d, up1);
    }
    h0 ^= a ^ c;
    h1 ^= b ^ d;
}

__global__ void __launch_bounds__(128) merkle_leaf_quad_kernel(const LeafParams P)
{
    __shared__ uint64_t smsg[(128 / 4) * QUAD_MSG_STRIDE];
    const uint32_t q = threadIdx.x & 3;
    const size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 2;     /* leaf of this quad */
    const bool active = i < P.leaves;                                          /* whole quads only: leaves is a power of two >= 8 */
    uint64_t *msg = smsg + (threadIdx.x >> 2) * QUAD_MSG_STRIDE;
    uint32_t sig[12];
#pragma unroll
    for (int r = 0; r < 12; ++r)
        sig[r] = (uint32_t)c_blake_sigma[r][2 * q] | ((uint32_t)c_blake_sigma[r][2 * q + 1] << 4) |
                 ((uint32_t)c_blake_sigma[r][8 + 2 * q] << 8) | ((uint32_t)c_blake_sigma[r][8 + 2 * q + 1] << 12);

    const size_t n_words = (size_t)3 * P.n_oracles * P.coset;
    const uint32_t cmask = P.coset - 1;
    /* chaining value: lane q holds h[q] and h[4+q] */
    uint64_t h0 = (q == 0 ? (BLAKE_IV0 ^ 0x01010000ULL ^ 32ull) : q == 1 ? BLAKE_IV1 : q == 2 ? BLAKE_IV2 : BLAKE_IV3);
    uint64_t h1 = (q == 0 ? BLAKE_IV4 : q == 1 ? BLAKE_IV5 : q == 2 ? BLAKE_IV6 : BLAKE_IV7);
    size_t done = 0;
    for (;;) {                                                                 /* uniform trip count across the grid */
        const size_t remaining = n_words - done;
        const bool last = remaining <= 16;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t w = done + q + 4 * u;                                 /* word of the leaf this lane fetches */
            uint64_t v = 0;
            if (active && w < n_words) {
                const size_t e = w / 3;
                const uint32_t limb = (uint32_t)(w - 3 * e);
                const uint32_t j = (uint32_t)e & cmask;
                const size_t k = e >> P.coset_log;
                const size_t pos = (P.kind == B200_DOMAIN_MULTIPLICATIVE_COSET) ? (i + (size_t)j * P.leaves) : (i * P.coset + j);
                const uint64_t *src = P.oracle_table ? P.oracle_table[k] : P.oracles[k];
                v = __ldg(src + pos * 3 + limb);
            }
            msg[q + 4 * u] = v;
        }
        __syncwarp();
        done += last ? remaining : 16;
        quad_compress(h0, h1, msg, sig, q, (uint64_t)done * 8, last);
        __syncwarp();                                                          /* the block is rewritten next iteration */
        if (last) break;
    }
    if (P.salts) {                                  /* zk: H(H(leaf) || salt)  (blake2b.tcc:127-136): one more block */
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t w = q + 4 * u;
            uint64_t v = 0;
            if (u == 0) v = h0;                                               /* words 0..3 = the inner digest (h[0..3]) */
            else if (active && w - 4 < P.salt_words) v = P.salts[i * P.salt_words + (w - 4)];
            msg[w] = v;
        }
        __syncwarp();
        h0 = (q == 0 ? (BLAKE_IV0 ^ 0x01010000ULL ^ 32ull) : q == 1 ? BLAKE_IV1 : q == 2 ? BLAKE_IV2 : BLAKE_IV3);
        h1 = (q == 0 ? BLAKE_IV4 : q == 1 ? BLAKE_IV5 : q == 2 ? BLAKE_IV6 : BLAKE_IV7);
        quad_compress(h0, h1, msg, sig, q, (uint64_t)(4 + P.salt_words) * 8, true);
    }
    if (active) P.nodes[(P.leaves - 1 + i) * 4 + q] = h0;                      /* 32-byte digest = h[0..3], one word per lane */
}

/* [round 2] One hash per thread at EVERY level.  Fusing three levels into the leaf thread (merkle_leaf8_kernel above) saves
   writing and re-reading the digests — 2 x 32 B per node, a few per cent of a pass that is ALU-bound anyway — but costs a
   64-register digest array, 8-leaf slices that are 8 x slice bytes apart between lanes, and 1/8 of the threads.  Measured on
   the 4 x 2^23 tree: 2.73 ms fused, 1.81 ms with one launch per level.  Leaf i is hashed by thread i, so for every word index
   the lanes of a warp read addresses one leaf (48-192 B) apart; each level is a launch of width/2 threads reading 64
   contiguous bytes each. */
#ifndef B200_LEAF1_MINBLOCKS
#define B200_LEAF1_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(256, B200_LEAF1_MINBLOCKS) merkle_leaf1_kernel(const LeafParams P)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= P.leaves) return;
    const size_t n_words = (size_t)3 * P.n_oracles * P.coset;
    uint64_t h[8];
    blake2b_init(h, 32);
    uint32_t k = 0, j = 0, limb = 0;
    const uint64_t *base = P.oracle_table ? P.oracle_table[0] : P.oracles[0];
    const size_t pos = (P.kind == B200_DOMAIN_MULTIPLICATIVE_COSET) ? i : i * P.coset;
    const size_t jstep = (P.kind == B200_DOMAIN_MULTIPLICATIVE_COSET) ? P.leaves : 1;
    size_t done = 0;
    for (;;) {
        const size_t remaining = n_words - done;
        const bool last = remaining <= 16;
        uint64_t m[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            uint64_t v = 0;
            if ((size_t)w < remaining) {
                v = __ldg(base + (pos + (size_t)j * jstep) * 3 + limb);
                if (++limb == 3) {
                    limb = 0;
                    if (++j == P.coset) { j = 0; ++k; if (k < P.n_oracles) base = P.oracle_table ? P.oracle_table[k] : P.oracles[k]; }
                }
            }
            m[w] = v;
        }
        done += last ? remaining : 16;
        blake2b_compress(h, m, (uint64_t)done * 8, last);
        if (last) break;
    }
    if (P.salts) {                                  /* zk: H(H(leaf) || salt)  (blake2b.tcc:127-136) */
        const uint64_t *salt = P.salts + i * P.salt_words;
        uint64_t m[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            uint64_t v = 0;
            if (w < 4) v = h[w];
            else if ((uint32_t)(w - 4) < P.salt_words) v = salt[w - 4];
            m[w] = v;
        }
        blake2b_init(h, 32);
        blake2b_compress(h, m, (uint64_t)(4 + P.salt_words) * 8, true);
    }
    store_digest32(P.nodes, P.leaves - 1 + i, h);
}

/* level of `count` nodes starting at heap index `first`: node j = H(node 2j+1 || node 2j+2), 32-byte digests */
__global__ void __launch_bounds__(256) merkle_level32_kernel(uint64_t *nodes, size_t first, size_t count)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    const size_t j = first + t;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(nodes + (2 * j + 1) * 4);     /* children are adjacent: 64 bytes */
    const ulonglong2 a = src[0], b = src[1], c2 = src[2], d = src[3];
    const uint64_t l[4] = {a.x, a.y, b.x, b.y}, r[4] = {c2.x, c2.y, d.x, d.y};
    uint64_t out[4];
    two_to_one_32(out, l, r);
    store_digest32(nodes, j, out);
}

__global__ void __launch_bounds__(128) merkle_reduce8_kernel(uint64_t *nodes, size_t width)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t * 8 >= width) return;
    uint64_t d[8][4];
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(nodes + (width - 1 + t * 8) * 4);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const ulonglong2 a = src[2 * q], b = src[2 * q + 1];
        d[q][0] = a.x; d[q][1] = a.y; d[q][2] = b.x; d[q][3] = b.y;
    }
    reduce8_32(nodes, d, width, t * 8);
}

int merkle_build(b200_ctx *c, const b200_buf *oracles, size_t n_oracles, size_t n_elems, size_t coset,
                 int kind, const uint8_t *salts, size_t salt_len, size_t digest_len, b200_bytes nodes)
{
    if (!oracles || n_oracles == 0) return set_err(c, B200_ERR_INVALID, "no oracles");
    if (coset == 0 || (coset & (coset - 1)) || n_elems % coset)
        return set_err(c, B200_ERR_LOGIC, "Attempting to construct a Merkle tree with a constituent vector of wrong size");
    const size_t leaves = n_elems / coset;
    if (leaves < 2 || (leaves & (leaves - 1)))   /* merkle_tree.tcc:27-31 */
        return set_err(c, B200_ERR_INVALID, "Merkle tree size must be a power of two, and at least 2.");
    if (digest_len == 0 || digest_len > 64 || digest_len % 8)
        return set_err(c, B200_ERR_INVALID, "digest length must be a multiple of 8 bytes, at most 64");
    if (salts && (salt_len % 8 || salt_len == 0 || digest_len + salt_len > 128))
        return set_err(c, B200_ERR_INVALID, "unsupported zk salt length");
    if (kind != B200_DOMAIN_AFFINE_SUBSPACE && kind != B200_DOMAIN_MULTIPLICATIVE_COSET)
        return set_err(c, B200_ERR_INVALID, "unknown domain kind");

    LeafParams P; memset(&P, 0, sizeof P);
    P.n_oracles = (uint32_t)n_oracles; P.coset = (uint32_t)coset; P.coset_log = ceil_log2(coset);
    P.kind = kind; P.leaves = leaves; P.digest_len = (uint32_t)digest_len; P.digest_words = (uint32_t)(digest_len / 8);
    P.nodes = (uint64_t *)nodes;
    if (n_oracles <= 64) {
        for (size_t k = 0; k < n_oracles; ++k) P.oracles[k] = (const uint64_t *)oracles[k];
    } else {                                       /* Ligero: hundreds of rows per leaf */
        const uint64_t **tab = (const uint64_t **)scratch_get(c, 3, n_oracles * sizeof(uint64_t *));
        if (!tab) return set_err(c, B200_ERR_NOMEM, "cannot allocate oracle table");
        B200_CUDA(c, cudaMemcpyAsync(tab, oracles, n_oracles * sizeof(uint64_t *), cudaMemcpyHostToDevice, c->stream));
        P.oracle_table = tab;
    }
    if (salts) {
        uint8_t *ds = (uint8_t *)scratch_get(c, 2, leaves * salt_len);
        if (!ds) return set_err(c, B200_ERR_NOMEM, "cannot allocate salt buffer");
        B200_CUDA(c, cudaMemcpyAsync(ds, salts, leaves * salt_len, cudaMemcpyHostToDevice, c->stream));
        P.salts = (const uint64_t *)ds; P.salt_words = (uint32_t)(salt_len / 8);
    }
    size_t width;   /* nodes in the highest level already computed */
    static const bool no_quad = getenv("B200_NO_QUAD_LEAF") != nullptr;           /* ablation switch */
    if (!no_quad && digest_len == 32 && leaves >= 8 && n_oracles * coset * 24 >= 1024 && leaves * 4 <= ((size_t)1 << 22) &&
        (!salts || salt_len <= 96)) {
        /* wide leaves, few of them: four lanes per leaf (merkle_leaf_quad_kernel), then the usual 8-to-1 reductions */
        merkle_leaf_quad_kernel<<<(unsigned)((leaves * 4 + 127) / 128), 128, 0, c->stream>>>(P);
        B200_LAUNCH_CHECK(c);
        width = leaves;
        while (width > 2048) {
            const size_t th = width / 8;
            merkle_reduce8_kernel<<<(unsigned)((th + 127) / 128), 128, 0, c->stream>>>(P.nodes, width);
            B200_LAUNCH_CHECK(c);
            width >>= 3;
        }
    } else if (!getenv("B200_MERKLE_FLAT") && !getenv("B200_MERKLE_FUSED") && digest_len == 32 && leaves >= 4096 && (!salts || salt_len <= 96)) {
        /* one hash per thread at every level (see merkle_leaf1_kernel) */
        merkle_leaf1_kernel<<<(unsigned)((leaves + 255) / 256), 256, 0, c->stream>>>(P);
        B200_LAUNCH_CHECK(c);
        width = leaves;
        while (width > 2048) {
            const size_t count = width >> 1, first = count - 1;
            merkle_level32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(P.nodes, first, count);
            B200_LAUNCH_CHECK(c);
            width = count;
        }
    } else if (!getenv("B200_MERKLE_FLAT") && digest_len == 32 && leaves >= 4096 && (!salts || salt_len <= 96)) {
        /* round-1 path (B200_MERKLE_FUSED=1, kept for comparison): 8 leaves + 3 levels per thread */
        const size_t threads = leaves / 8;
        merkle_leaf8_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, c->stream>>>(P);
        B200_LAUNCH_CHECK(c);
        width = leaves >> 3;
        while (width > 2048) {
            const size_t th = width / 8;
            merkle_reduce8_kernel<<<(unsigned)((th + 127) / 128), 128, 0, c->stream>>>(P.nodes, width);
            B200_LAUNCH_CHECK(c);
            width >>= 3;
        }
    } else {
        /* generic path: a CTA of 256 leaves owns the 8 levels above them */
        uint32_t fused = (leaves >= MERKLE_CTA && !getenv("B200_MERKLE_FLAT")) ? 8 : 0;   /* B200_MERKLE_FLAT: generic kernels, one launch per level */
        P.fused_levels = fused;
        const unsigned blocks = (unsigned)((leaves + MERKLE_CTA - 1) / MERKLE_CTA);
        merkle_leaf_kernel<<<blocks, MERKLE_CTA, 0, c->stream>>>(P);
        B200_LAUNCH_CHECK(c);
        width = leaves >> fused;
        while (width > 2048) {
            const size_t count = width >> 1, first = count - 1;
            merkle_level_kernel<<<(unsigned)((count + 127) / 128), 128, 0, c->stream>>>(P.nodes, first, count, P.digest_words, P.digest_len);
            B200_LAUNCH_CHECK(c);
            width = count;
        }
    }
    if (width > 1) {
        merkle_top_kernel<<<1, 1024, 0, c->stream>>>(P.nodes, width, P.digest_words, P.digest_len);
        B200_LAUNCH_CHECK(c);
    }
    return 0;
}


/* ---------------------------------------------------------------------------------------------- */
/* proof-of-work grind: pow::solve_pow for binary digests (libiop/bcs/pow.tcc:87-102),
   acceptance test verify_pow_internal (pow.tcc:143-158).  Candidates are tried in the reference's order —
   the challenge itself, then the challenge with its last 8 bytes replaced by 0, 1, 2, ... — and the FIRST
   hit in that order is returned (atomicMin over candidate indices, batches of 2^22 hashes).             */
struct PowParams {
    uint64_t chal[4];
    uint64_t mask, upper;
    uint64_t base;           /* first candidate counter of this batch */
    unsigned long long *best; /* device: smallest accepted candidate index so far (index 0 = challenge itself) */
};

__global__ void __launch_bounds__(256) pow_grind_kernel(const PowParams P, uint64_t count)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint64_t idx = P.base + t;        /* candidate index; 0 = unmodified challenge, k >= 1 = last word k-1 */
    uint64_t m[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = P.chal[i]; m[4 + i] = P.chal[i]; }
    if (idx) m[7] = idx - 1;
#pragma unroll
    for (int i = 8; i < 16; ++i) m[i] = 0;
    uint64_t h[8];
    blake2b_init(h, 32);
    blake2b_compress(h, m, 64, true);
    if ((h[3] & P.mask) <= P.upper) atomicMin(P.best, (unsigned long long)idx);
}

int pow_solve(b200_ctx *c, const uint8_t *challenge, size_t digest_len, uint32_t pow_bitlen, uint64_t pow_upperbound, uint8_t *pow_out)
{
    if (digest_len != 32) return set_err(c, B200_ERR_INVALID, "proof of work is implemented for 32-byte BLAKE2b digests");
    if (pow_bitlen >= 31) return set_err(c, B200_ERR_INVALID, "pow_bitlen must be < 31 (pow.tcc:149 uses an int shift)");
    PowParams P;
    memcpy(P.chal, challenge, 32);
    P.mask = ((uint64_t)1 << pow_bitlen) - 1; P.upper = pow_upperbound;
    unsigned long long *best = (unsigned long long *)scratch_get(c, 3, 64);
    if (!best) return set_err(c, B200_ERR_NOMEM, "cannot allocate PoW result");
    P.best = best;
    const unsigned long long none = ~0ull;
    B200_CUDA(c, cudaMemcpyAsync(best, &none, sizeof none, cudaMemcpyHostToDevice, c->stream));
    const uint64_t batch = (uint64_t)1 << 22;
    unsigned long long found = none;
    for (uint64_t base = 0; found == none; base += batch) {
        if (base > ((uint64_t)1 << 40)) return set_err(c, B200_ERR_LOGIC, "proof of work not found after 2^40 hashes");
        P.base = base;
        pow_grind_kernel<<<(unsigned)(batch / 256), 256, 0, c->stream>>>(P, batch);
        B200_LAUNCH_CHECK(c);
        B200_CUDA(c, cudaMemcpyAsync(&found, best, sizeof found, cudaMemcpyDeviceToHost, c->stream));
        B200_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    memcpy(pow_out, challenge, 32);
    if (found) { const uint64_t k = found - 1; memcpy(pow_out + 24, &k, 8); }
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* verifier random messages: blake2b_FieldT_randomness_extractor (libiop/bcs/hashing/blake2b.tcc:231-256).  Element i is
   the keyed hash BLAKE2b(outlen = 24, key = u64 i, message = root || u64 squeeze index) (blake2b.tcc:162-183); for a
   prime field the bits above the modulus' top bit are cleared and a value >= modulus is re-drawn with key += n
   (rejection sampling, blake2b.tcc:185-229).  Ligero squeezes 4 * N elements after its first round.        */
struct SqueezeParams {
    uint64_t msg[16];        /* root || index, zero padded */
    uint32_t msg_len;        /* bytes, <= 128, multiple of 8 */
    uint64_t n;
    uint64_t *out;           /* n elements of 3 x u64 */
    int prime_field;
    uint64_t modulus[3], top_mask;   /* prime field only: limb 2 keeps (top_mask) */
};

__global__ void __launch_bounds__(256) squeeze_kernel(const SqueezeParams P)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    uint64_t key = i;
    uint64_t h[8], m[16];
    for (;;) {
        h[0] = BLAKE_IV0 ^ 0x01010000ULL ^ (8ull << 8) ^ 24ull;      /* digest 24 bytes, key 8 bytes */
        h[1] = BLAKE_IV1; h[2] = BLAKE_IV2; h[3] = BLAKE_IV3; h[4] = BLAKE_IV4; h[5] = BLAKE_IV5; h[6] = BLAKE_IV6; h[7] = BLAKE_IV7;
        m[0] = key;
#pragma unroll
        for (int t = 1; t < 16; ++t) m[t] = 0;
        blake2b_compress(h, m, 128, false);                          /* the key block */
#pragma unroll
        for (int t = 0; t < 16; ++t) m[t] = P.msg[t];
        blake2b_compress(h, m, 128 + (uint64_t)P.msg_len, true);
        if (!P.prime_field) break;
        h[2] &= P.top_mask;
        const bool below = h[2] != P.modulus[2] ? h[2] < P.modulus[2]
                         : h[1] != P.modulus[1] ? h[1] < P.modulus[1] : h[0] < P.modulus[0];
        if (below) break;
        key += P.n;
    }
    uint64_t *o = P.out + 3 * i;
    o[0] = h[0]; o[1] = h[1]; o[2] = h[2];
}

int squeeze_field_elements(b200_ctx *c, int field, const b200_fp_params *fp, const uint8_t *root_plus_index, size_t msg_len,
                           size_t n, b200_buf out)
{
    if (msg_len == 0 || msg_len > 128 || msg_len % 8)
        return set_err(c, B200_ERR_INVALID, "squeeze: root || index must be 8..128 bytes, a multiple of 8");
    if (field == B200_FIELD_FP && !fp) return set_err(c, B200_ERR_INVALID, "squeeze: prime field without parameters");
    if (!n) return 0;
    SqueezeParams P;
    memset(&P, 0, sizeof P);
    memcpy(P.msg, root_plus_index, msg_len);
    P.msg_len = (uint32_t)msg_len; P.n = n; P.out = (uint64_t *)out;
    P.prime_field = field == B200_FIELD_FP;
    if (P.prime_field) {
        for (int k = 0; k < 3; ++k) P.modulus[k] = fp->modulus[k];
        if (!P.modulus[2]) return set_err(c, B200_ERR_INVALID, "squeeze: the modulus must use its top limb (3-limb Fp_model)");
        uint64_t mask = ~0ull;
        while ((mask >> 1) >= P.modulus[2]) mask >>= 1;              /* smallest 2^k - 1 >= the top limb */
        P.top_mask = mask;
    }
    squeeze_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(P);
    B200_LAUNCH_CHECK(c);
    return 0;
}

struct BatchParams { const uint64_t *in; uint64_t *out; size_t msg_words, n; uint32_t digest_len, dw; };
__global__ void blake2b_batch_kernel(const BatchParams P)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint64_t *msg = P.in + i * P.msg_words;
    uint64_t h[8];
    blake2b_words(h, P.digest_len, P.msg_words, [&](size_t w) { return msg[w]; });
    for (uint32_t t = 0; t < P.dw; ++t) P.out[i * P.dw + t] = h[t];
}

int blake2b_batch(b200_ctx *c, b200_bytes in, size_t msg_len, size_t n_msgs, size_t digest_len, b200_bytes out)
{
    if (msg_len % 8 || digest_len % 8 || digest_len == 0 || digest_len > 64)
        return set_err(c, B200_ERR_INVALID, "message and digest lengths must be multiples of 8 bytes");
    if (!n_msgs) return 0;
    BatchParams P{(const uint64_t *)in, (uint64_t *)out, msg_len / 8, n_msgs, (uint32_t)digest_len, (uint32_t)(digest_len / 8)};
    blake2b_batch_kernel<<<(unsigned)((n_msgs + 127) / 128), 128, 0, c->stream>>>(P);
    B200_LAUNCH_CHECK(c);
    return 0;
}

} // namespace b200
