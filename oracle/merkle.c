/*
 * oracle/merkle.c — BCS Merkle tree build + hashchain, CPU restatement
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * construct_with_leaves_serialized_by_cosets   libiop/bcs/merkle_tree.tcc:93-151
 * compute_inner_nodes                          libiop/bcs/merkle_tree.tcc:201-229
 * leaf positions: additive  pos = i*coset + j  libiop/algebra/field_subset/subspace.tcc:87-91
 *                 multiplic. pos = i + j*leaves libiop/algebra/field_subset/subgroup.tcc:192-197
 * blake2b_field_element_hash                   libiop/bcs/hashing/blake2b.tcc:141-160
 * blake2b_leafhash::zk_hash = H(H(leaf)||salt) libiop/bcs/hashing/blake2b.tcc:127-136
 * blake2b_two_to_one_hash                      libiop/bcs/hashing/blake2b.cpp:28-48
 * blake2b_hashchain                            libiop/bcs/hashing/blake2b.tcc:10-110
 * randomness extractors                        libiop/bcs/hashing/blake2b.tcc:163-257, blake2b.cpp:50-74
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

void orc_merkle_build(uint8_t *nodes, const f192 *const *oracles, size_t n_oracles,
                      size_t n_elems, size_t coset_size, int domain_kind,
                      const uint8_t *salts, size_t salt_len, size_t digest_len)
{
    const size_t leaves = n_elems / coset_size;
    const size_t slice_elems = n_oracles * coset_size;
    f192 *slice = (f192 *)malloc(slice_elems * sizeof(f192));
    uint8_t tmp[128];
    for (size_t i = 0; i < leaves; ++i) {
        for (size_t j = 0; j < coset_size; ++j) {
            const size_t pos = (domain_kind == 2) ? (i + j * leaves) : (i * coset_size + j);
            for (size_t k = 0; k < n_oracles; ++k)
                slice[j + k * coset_size] = oracles[k][pos];      /* merkle_tree.tcc:127-134 */
        }
        uint8_t *dst = nodes + (leaves - 1 + i) * digest_len;
        if (salts) {
            orc_blake2b(tmp, digest_len, (const uint8_t *)slice, slice_elems * sizeof(f192), NULL, 0);
            memcpy(tmp + digest_len, salts + i * salt_len, salt_len);
            orc_blake2b(dst, digest_len, tmp, digest_len + salt_len, NULL, 0);
        } else {
            orc_blake2b(dst, digest_len, (const uint8_t *)slice, slice_elems * sizeof(f192), NULL, 0);
        }
    }
    /* inner nodes, heap order (merkle_tree.tcc:201-229) */
    for (size_t j = leaves - 1; j-- > 0; ) {
        memcpy(tmp, nodes + (2 * j + 1) * digest_len, digest_len);
        memcpy(tmp + digest_len, nodes + (2 * j + 2) * digest_len, digest_len);
        orc_blake2b(nodes + j * digest_len, digest_len, tmp, 2 * digest_len, NULL, 0);
    }
    free(slice);
}

/* ---- hashchain (blake2b.tcc:10-110) ---- */
void orc_hashchain_init(orc_hashchain *hc, size_t security_parameter)
{
    hc->digest_len = (2 * security_parameter + 7) / 8;          /* blake2b.tcc:15-16 */
    memset(hc->state, ' ', hc->digest_len);                     /* blake2b.tcc:18 */
    hc->squeeze_index = 0;
}

/* absorb_hash_digest hashes only the first digest_len bytes of state||input, i.e. the old state
   (blake2b.tcc:51-66, the length argument at line 60). */
void orc_hashchain_absorb(orc_hashchain *hc)
{
    uint8_t out[64];
    orc_blake2b(out, hc->digest_len, hc->state, hc->digest_len, NULL, 0);
    memcpy(hc->state, out, hc->digest_len);
}

void orc_hashchain_squeeze_gf192(orc_hashchain *hc, f192 *out, size_t n)
{
    hc->squeeze_index++;                                        /* blake2b.tcc:79 */
    uint8_t buf[72];
    memcpy(buf, hc->state, hc->digest_len);
    memcpy(buf + hc->digest_len, &hc->squeeze_index, 8);        /* blake2b.tcc:237-240 */
    for (size_t i = 0; i < n; ++i) {
        const uint64_t key = i;                                 /* blake2b.tcc:165-181 */
        orc_blake2b((uint8_t *)&out[i], 24, buf, hc->digest_len + 8, (const uint8_t *)&key, 8);
    }
}

void orc_hashchain_squeeze_fp(orc_hashchain *hc, const orc_fp_params *P, f192 *out, size_t n)
{
    hc->squeeze_index++;
    uint8_t buf[72];
    memcpy(buf, hc->state, hc->digest_len);
    memcpy(buf + hc->digest_len, &hc->squeeze_index, 8);
    for (size_t i = 0; i < n; ++i) {
        uint64_t key = i;                                       /* blake2b.tcc:185-229 */
        for (;;) {
            f192 el;
            orc_blake2b((uint8_t *)&el, 24, buf, hc->digest_len + 8, (const uint8_t *)&key, 8);
            for (int bit = 191; bit >= (int)P->num_bits; --bit)
                el.w[bit / 64] &= ~(1ull << (bit % 64));
            int lt = 0;
            for (int l = 2; l >= 0; --l) {
                if (el.w[l] < P->mod[l]) { lt = 1; break; }
                if (el.w[l] > P->mod[l]) { lt = 0; break; }
            }
            key += n;
            if (lt) { out[i] = el; break; }
        }
    }
}

uint64_t orc_hashchain_squeeze_position(orc_hashchain *hc, uint64_t range)
{
    hc->squeeze_index++;                                        /* blake2b.tcc:93-101 */
    uint64_t r = 0;
    orc_blake2b((uint8_t *)&r, 8, hc->state, hc->digest_len,
                (const uint8_t *)&hc->squeeze_index, 8);        /* blake2b.cpp:50-74 */
    return r % range;
}

/* pow::solve_pow_internal for binary digests (libiop/bcs/pow.tcc:87-102) with verify_pow_internal (pow.tcc:143-158).
   Returns the number of hashes tried. */
uint64_t orc_pow_solve(uint8_t *pow_out, const uint8_t *challenge, size_t digest_len, uint32_t pow_bitlen, uint64_t pow_upperbound)
{
    uint8_t buf[128], hash[64];
    memcpy(pow_out, challenge, digest_len);
    const size_t num_words = digest_len / 8;
    uint64_t pow_int = 0, tries = 0;
    for (;;) {
        memcpy(buf, challenge, digest_len);
        memcpy(buf + digest_len, pow_out, digest_len);
        orc_blake2b(hash, digest_len, buf, 2 * digest_len, NULL, 0);     /* node_hasher(challenge, pow) */
        ++tries;
        uint64_t lsw;
        memcpy(&lsw, hash + (num_words - 1) * 8, 8);
        if ((lsw & (uint64_t)((1 << pow_bitlen) - 1)) <= pow_upperbound) return tries;
        memcpy(pow_out + (num_words - 1) * 8, &pow_int, 8);
        pow_int += 1;
    }
}
