/*
 * oracle.h — CPU restatement of the libiop prover hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This directory is the parity oracle for the B200 kernels.  Nothing in the
 * product (libiop_b200/, include/) may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function restates the algorithm of a reference function (cited
 * file:line relative to /root/reference).  The oracle is pinned against
 *   - the ten GF(2^192) multiplication KATs of libff
 *     (depends/libff/libff/algebra/fields/tests/test_binary_fields.cpp:124-135),
 *   - SURVEY.md Appendix C vectors C1..C6 (reference-generated),
 *   - golden fixtures under tests/golden/ produced by oracle/_ref (the unmodified
 *     reference compiled from /root/reference by oracle/Makefile).
 *
 * Element layout (both fields): 3 x uint64 little-endian limbs = 24 bytes.
 *   gf192  : GF(2)[x]/(x^192+x^7+x^2+x+1), polynomial basis     (gf192.hpp:19-20,98-99)
 *   Fp     : Montgomery representation, R = 2^192                (fp.hpp:42)
 */
#ifndef B200_ORACLE_H
#define B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t w[3]; } f192;

/* ---------------- GF(2^192) ---------------- (gf192.cpp:50-59, 72-184, 245-288) */
void orc_gf192_mul(f192 *r, const f192 *a, const f192 *b);
void orc_gf192_mul_bitwise(f192 *r, const f192 *a, const f192 *b); /* gf192.cpp:143-177 */
void orc_gf192_inv(f192 *r, const f192 *a);
void orc_gf192_mul_vec(f192 *r, const f192 *a, const f192 *b, size_t n);
void orc_gf192_inv_vec(f192 *r, const f192 *a, size_t n);

/* ---------------- Fp (3 limbs, Montgomery) ---------------- (fp.tcc:25-194,323-532,670-722) */
typedef struct {
    uint64_t mod[3];
    uint64_t r2[3];    /* R^2 mod p */
    uint64_t one[3];   /* R mod p   */
    uint64_t inv;      /* -p^{-1} mod 2^64 */
    uint32_t two_adicity;
    uint32_t num_bits;
    uint64_t generator[3]; /* multiplicative generator, Montgomery */
} orc_fp_params;

/* fills params from a decimal modulus string and small generator; returns 0 on success */
int  orc_fp_params_init(orc_fp_params *P, const uint64_t mod[3], uint64_t generator_small);
void orc_fp_params_edwards(orc_fp_params *P); /* edwards_fields.cpp:24-46 */
void orc_fp_mul(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b);
void orc_fp_add(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b);
void orc_fp_sub(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b);
void orc_fp_inv(const orc_fp_params *P, f192 *r, const f192 *a);
void orc_fp_pow_u64(const orc_fp_params *P, f192 *r, const f192 *a, uint64_t e);
void orc_fp_from_u64(const orc_fp_params *P, f192 *r, uint64_t v);
void orc_fp_to_plain(const orc_fp_params *P, f192 *r, const f192 *a); /* out of Montgomery */
/* generator of the order-2^logn subgroup: gen^((p-1)/2^logn)   (subgroup.tcc:56-59) */
void orc_fp_subgroup_generator(const orc_fp_params *P, f192 *g, uint32_t logn);
void orc_fp_mul_vec(const orc_fp_params *P, f192 *r, const f192 *a, const f192 *b, size_t n);

/* ---------------- FFTs ---------------- */
/* additive_FFT (libiop/algebra/fft.tcc:39-124). in: n_coeffs <= 2^m coefficients; out: 2^m evals */
void orc_additive_fft(f192 *out, const f192 *coeffs, size_t n_coeffs,
                      uint32_t m, const f192 *basis, const f192 *shift);
/* additive_IFFT (fft.tcc:126-204) */
void orc_additive_ifft(f192 *out, const f192 *evals,
                       uint32_t m, const f192 *basis, const f192 *shift);
/* naive_FFT (fft.tcc:14-37) Horner at every point; field: 0 = gf192, 1 = Fp */
void orc_naive_fft_gf192(f192 *out, const f192 *coeffs, size_t n_coeffs,
                         uint32_t m, const f192 *basis, const f192 *shift);
void orc_naive_fft_fp(const orc_fp_params *P, f192 *out, const f192 *coeffs, size_t n_coeffs,
                      uint32_t logn, const f192 *g, const f192 *shift);
/* multiplicative_FFT_degree_aware (fft.tcc:236-317) */
void orc_mult_fft(const orc_fp_params *P, f192 *out, const f192 *coeffs, size_t n_coeffs,
                  uint32_t logn, const f192 *g, const f192 *shift);
/* multiplicative_IFFT_internal -> libfqfft iFFT / icosetFFT
   (fft.tcc:343-361, basic_radix2_domain.tcc:49-74, basic_radix2_domain_aux.tcc:45-79,172-180) */
void orc_mult_ifft(const orc_fp_params *P, f192 *out, const f192 *evals,
                   uint32_t logn, const f192 *g, const f192 *shift);

/* ---------------- FRI fold ---------------- */
/* additive_evaluate_next_f_i_over_entire_domain (fri_aux.tcc:37-103) */
void orc_fri_fold_additive(f192 *out, const f192 *f, uint32_t m, const f192 *basis,
                           const f192 *shift, uint32_t eta, const f192 *x);
/* multiplicative_evaluate_next_f_i_over_entire_domain (fri_aux.tcc:107-249) */
void orc_fri_fold_multiplicative(const orc_fp_params *P, f192 *out, const f192 *f, uint32_t logn,
                                 const f192 *g, const f192 *shift, uint32_t eta, const f192 *x);

/* ---------------- BLAKE2b (RFC 7693; libsodium crypto_generichash_blake2b) ---------------- */
int orc_blake2b(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen,
                const uint8_t *key, size_t keylen);

/* ---------------- Merkle tree (merkle_tree.tcc:93-151, 201-229) ---------------- */
/* domain_kind: 1 = affine subspace (pos = i*coset + j), 2 = multiplicative coset (pos = i + j*leaves)
   salts: NULL or num_leaves*salt_len bytes (zk: H(H(leaf)||salt), blake2b.tcc:127-136)
   nodes_out: (2*leaves-1)*digest_len bytes, heap order, root first. */
void orc_merkle_build(uint8_t *nodes_out, const f192 *const *oracles, size_t n_oracles,
                      size_t n_elems, size_t coset_size, int domain_kind,
                      const uint8_t *salts, size_t salt_len, size_t digest_len);

/* ---------------- hashchain / PoW (blake2b.tcc:10-110,163-257; pow.tcc:87-158) ---------------- */
typedef struct { uint8_t state[64]; size_t digest_len; uint64_t squeeze_index; } orc_hashchain;
void   orc_hashchain_init(orc_hashchain *hc, size_t security_parameter);
void   orc_hashchain_absorb(orc_hashchain *hc);                     /* input is ignored, see SURVEY §0.4 */
void   orc_hashchain_squeeze_gf192(orc_hashchain *hc, f192 *out, size_t n);
void   orc_hashchain_squeeze_fp(orc_hashchain *hc, const orc_fp_params *P, f192 *out, size_t n);
uint64_t orc_hashchain_squeeze_position(orc_hashchain *hc, uint64_t range);
/* pow::solve_pow (pow.tcc:87-102,143-158); returns hashes tried */
uint64_t orc_pow_solve(uint8_t *pow_out, const uint8_t *challenge, size_t digest_len, uint32_t pow_bitlen, uint64_t pow_upperbound);

#ifdef __cplusplus
}
#endif
#endif
