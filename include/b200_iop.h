/*
 * b200_iop.h — C ABI of the B200-native libiop prover hot path.
 *
 * This is the drop-in boundary.  The reference (scipr-lab/libiop) has no FFI layer:
 * callers invoke C++ templates on std::vector<FieldT>.  Each entry point below replaces one of
 * those templates; the C++ adapter in libiop_b200/host/ keeps the reference's names and
 * signatures and forwards here.  Reference citations are relative to /root/reference.
 *
 * Conventions
 *   - field element = 3 x uint64 little-endian limbs = 24 bytes, for both
 *       gf192       GF(2)[x]/(x^192+x^7+x^2+x+1)          depends/libff/.../binary/gf192.hpp:19-20,98-99
 *       Fp (3 limb) Montgomery form, R = 2^192            depends/libff/.../prime_base/fp.hpp:42
 *   - all sizes are in elements unless the name says bytes
 *   - every call returns 0 or a negative b200_status; b200_last_error(ctx) holds the message
 *   - a context is bound to one device and one stream; calls are stream-ordered and asynchronous,
 *     results become visible to the host through download/gather (which synchronise)
 *   - a context is NOT thread-safe (neither is the reference: SURVEY.md §8b "Threading")
 *   - there is no CPU fallback: without a CUDA device b200_ctx_create fails
 */
#ifndef B200_IOP_H
#define B200_IOP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;
typedef uint64_t b200_buf;   /* device array of 24-byte field elements (a device address) */
typedef uint64_t b200_bytes; /* device byte array (digests, salts)          (a device address) */

typedef enum {
    B200_OK = 0,
    B200_ERR_INVALID = -1,   /* std::invalid_argument in the reference (e.g. fft.tcc:333,368) */
    B200_ERR_LOGIC = -2,     /* std::logic_error in the reference (e.g. merkle_tree.tcc:98-108) */
    B200_ERR_CUDA = -3,
    B200_ERR_NOMEM = -4
} b200_status;

enum { B200_DOMAIN_AFFINE_SUBSPACE = 1, B200_DOMAIN_MULTIPLICATIVE_COSET = 2 };
enum { B200_FIELD_GF192 = 1, B200_FIELD_FP = 2 };

/* Prime-field description: Fp_model<3, modulus> statics (fp.hpp:40-60; edwards_fields.cpp:24-46). */
typedef struct {
    uint64_t modulus[3];
    uint64_t r2[3];        /* Rsquared */
    uint64_t one_mont[3];  /* R mod p  */
    uint64_t inv;          /* -modulus^-1 mod 2^64 */
    uint32_t two_adicity;  /* s */
    uint32_t num_bits;
} b200_fp_params;

/* field_subset<FieldT> (libiop/algebra/field_subset/field_subset.hpp:25-77) flattened.
 *   kind 1: affine_subspace  — element idx = shift + sum_k bit_k(idx)*basis[k]      (subspace.tcc:54-71,161-170)
 *   kind 2: multiplicative_coset — element idx = shift * generator^idx, Montgomery  (subgroup.tcc:280-304) */
typedef struct {
    int kind;
    uint32_t dim;              /* log2(#elements) */
    const uint64_t *basis;     /* kind 1: dim*3 limbs (host pointer) */
    uint64_t shift[3];
    uint64_t generator[3];     /* kind 2 */
} b200_domain;

/* ---- context / memory --------------------------------------------------------------------- */
b200_ctx *b200_ctx_create(int device);
void b200_ctx_destroy(b200_ctx *);
const char *b200_last_error(b200_ctx *);
int b200_sync(b200_ctx *);
void *b200_stream(b200_ctx *);                       /* cudaStream_t the context launches on */
int b200_device_info(b200_ctx *, int *sm_count, size_t *total_mem, int *cc_major, int *cc_minor);

/* Device buffers come from a caching allocator: b200_free keeps the block for the next b200_alloc of the same size
 * (stream-ordered reuse, no synchronisation); b200_trim returns all cached blocks to the driver. */
b200_buf b200_alloc(b200_ctx *, size_t n_elems);     /* 0 on failure */
void b200_free(b200_ctx *, b200_buf);
int b200_trim(b200_ctx *);
b200_bytes b200_alloc_bytes(b200_ctx *, size_t n_bytes);
void b200_free_bytes(b200_ctx *, b200_bytes);
void *b200_host_alloc(b200_ctx *, size_t n_bytes);   /* pinned host memory for upload/download */
void b200_host_free(b200_ctx *, void *);

int b200_upload(b200_ctx *, b200_buf dst, size_t dst_offset, const uint64_t *src, size_t n_elems);
int b200_download(b200_ctx *, uint64_t *dst, b200_buf src, size_t src_offset, size_t n_elems);
int b200_upload_bytes(b200_ctx *, b200_bytes dst, const uint8_t *src, size_t n_bytes);
int b200_download_bytes(b200_ctx *, uint8_t *dst, b200_bytes src, size_t offset, size_t n_bytes);
/* blocks until every b200_upload / b200_upload_bytes issued so far has left the host buffer (kernels queued behind them keep
 * running): the point after which the caller may release or rewrite its host vectors */
int b200_wait_uploads(b200_ctx *);
int b200_copy(b200_ctx *, b200_buf dst, size_t dst_offset, b200_buf src, size_t src_offset, size_t n_elems);
/* dst[i * dst_stride] = src[src_offset + i * src_stride], i < n_elems: the evaluations over a sub-coset of a multiplicative
 * domain (IFFT_of_known_degree_over_field_subset, libiop/algebra/fft.tcc:446-455) taken from a device-resident codeword, and
 * the placement of a sub-coset's values inside the larger coset (field_subset::reindex_by_subset, subgroup.tcc:150-173) */
int b200_copy_strided(b200_ctx *, b200_buf dst, size_t dst_stride, b200_buf src, size_t src_offset, size_t src_stride, size_t n_elems);
int b200_fill_zero(b200_ctx *, b200_buf dst, size_t dst_offset, size_t n_elems);
/* iop_protocol::get_oracle_evaluation_at_point (libiop/iop/iop.tcc:670-688): query answers */
int b200_gather(b200_ctx *, uint64_t *dst, b200_buf src, const size_t *idx, size_t k);

/* the same for many device arrays at one list of positions, one launch and one transfer: dst[s * k + j] = srcs[s][idx[j]],
 * s < n_srcs, j < k (bcs_prover::get_transcript, bcs/bcs_prover.tcc:197-207: every oracle of a Merkle tree at every position
 * queried in that tree; Ligero commits ~10^3 oracles per tree) */
int b200_gather_multi(b200_ctx *, uint64_t *dst, const b200_buf *srcs, size_t n_srcs, const size_t *idx, size_t k);
/* device timing on the context's stream (CUDA events) */
int b200_timer_start(b200_ctx *);
int b200_timer_stop(b200_ctx *, float *ms);          /* synchronises */

/* ---- GF(2^192): additive FFT ---------------------------------------------------------------- */
/* FFT_over_field_subset<gf192>(coeffs, L) -> additive_FFT      libiop/algebra/fft.tcc:39-124, 414-419
 * coeffs: n_coeffs <= 2^L->dim monomial coefficients (zero-padded, fft.tcc:43-44); evals_out: 2^dim. */
int b200_gf192_fft(b200_ctx *, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf evals_out);
/* IFFT_over_field_subset<gf192>(evals, H) -> additive_IFFT     libiop/algebra/fft.tcc:126-204, 428-433 */
int b200_gf192_ifft(b200_ctx *, const b200_domain *H, b200_buf evals, b200_buf coeffs_out);
/* `batch` polynomials over one domain in one call — the row loops of Ligero's interleaved protocols, which run
 * IFFT_over_field_subset(row, systematic_domain) then FFT_over_field_subset(coefficients, codeword_domain) once per row
 * (protocols/encoded/ligero/ligero.tcc:293-331, interleaved_lincheck_ot.tcc:283-330, interleaved_rowcheck.tcc).
 * Row r reads coeffs + r * n_coeffs (evals + r * 2^dim) and writes evals_out + r * 2^dim (coeffs_out + r * 2^dim). */
int b200_gf192_fft_batch(b200_ctx *, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, size_t batch, b200_buf evals_out);
int b200_gf192_ifft_batch(b200_ctx *, const b200_domain *H, b200_buf evals, size_t batch, b200_buf coeffs_out);
/* Fused low-degree extension: the evaluations over H (2^H->dim) of a polynomial of degree < |H| -> its evaluations over L,
 * i.e. IFFT_over_field_subset(evals, H) followed by FFT_over_field_subset(coefficients, L) as the prover does for every
 * witness-derived oracle (protocols/encoded/r1cs_rs_iop/r1cs_rs_iop.tcc:449-479; Ligero's row encodings, ligero.tcc:293-331)
 * without forming the coefficients: the basis-only halves of the two transforms cancel.  Requires L->basis[0 .. H->dim) ==
 * H->basis (B200_ERR_INVALID otherwise: use the two calls).  Row r of a batch reads evals + r * 2^H->dim, writes out + r * 2^L->dim. */
int b200_gf192_lde(b200_ctx *, const b200_domain *H, const b200_domain *L, b200_buf evals, b200_buf evals_out);
int b200_gf192_lde_batch(b200_ctx *, const b200_domain *H, const b200_domain *L, b200_buf evals, size_t batch, b200_buf evals_out);
/* The two halves of the fused extension as separate calls.  `interpolate`: evaluations over H -> the 2^H->dim coefficients of
 * the polynomial in the basis-dependent form the reference holds between additive_IFFT's inverse butterflies and its bit
 * reversal (fft.tcc:139-170); `evaluate`: such coefficients w.r.t. L->basis[0 .. coeff_dim) -> evaluations over L.  The form is
 * LINEAR in the polynomial and splits along cosets: for H' = H + L->basis[dim H], a polynomial f of degree < 2|H| satisfies
 * f = A + (Z_H / Z_H(basis[dim H])) * B with coeffs(B) = interpolate_H(f|H) + interpolate_{H'}(f|H'), so the quotient of the
 * sumcheck (sumcheck.tcc:344-388: IFFT, division by Z_H, FFT) is two interpolations, a sum, one scaling and one evaluation.
 * Neither half runs in place. */
int b200_gf192_lde_interpolate(b200_ctx *, const b200_domain *H, b200_buf evals, b200_buf coeffs_out);
int b200_gf192_lde_evaluate(b200_ctx *, const b200_domain *L, uint32_t coeff_dim, b200_buf coeffs, b200_buf evals_out);
/* element-wise helpers (gf192::operator*=, inverse: gf192.cpp:72-184, 245-288) */
int b200_gf192_mul(b200_ctx *, b200_buf out, b200_buf a, b200_buf b, size_t n);
int b200_gf192_inv(b200_ctx *, b200_buf out, b200_buf a, size_t n);

/* ---- Fp: radix-2 NTT ------------------------------------------------------------------------ */
/* FFT_over_field_subset<Fp>(coeffs, L) -> multiplicative_FFT_degree_aware   fft.tcc:236-317, 407-412 */
int b200_fp_ntt(b200_ctx *, const b200_fp_params *, const b200_domain *L, b200_buf coeffs, size_t n_coeffs,
                b200_buf evals_out);
/* IFFT_over_field_subset<Fp>(evals, H) -> libfqfft iFFT / icosetFFT        fft.tcc:343-361 */
int b200_fp_intt(b200_ctx *, const b200_fp_params *, const b200_domain *H, b200_buf evals, b200_buf coeffs_out);
/* element-wise helpers (Fp_model::mul_reduce/+=/-=/invert: fp.tcc:25-194,323-532,670-722) */
int b200_fp_mul(b200_ctx *, const b200_fp_params *, b200_buf out, b200_buf a, b200_buf b, size_t n);
int b200_fp_add(b200_ctx *, const b200_fp_params *, b200_buf out, b200_buf a, b200_buf b, size_t n);
int b200_fp_sub(b200_ctx *, const b200_fp_params *, b200_buf out, b200_buf a, b200_buf b, size_t n);
int b200_fp_inv(b200_ctx *, const b200_fp_params *, b200_buf out, b200_buf a, size_t n);

/* ---- FRI commit-phase fold ------------------------------------------------------------------- */
/* evaluate_next_f_i_over_entire_domain(f_i, L_i, 2^eta, x_i)   libiop/protocols/ldt/fri/fri_aux.tcc:6-249
 * field = B200_FIELD_GF192 (L_i affine subspace) or B200_FIELD_FP (L_i multiplicative coset; fp != NULL).
 * f_i: 2^dim elements; f_next_out: 2^(dim-eta). */
int b200_fri_fold(b200_ctx *, int field, const b200_fp_params *fp, const b200_domain *L_i, uint32_t eta,
                  const uint64_t x_i[3], b200_buf f_i, b200_buf f_next_out);

/* ---- BCS Merkle tree ------------------------------------------------------------------------- */
/* merkle_tree::construct_with_leaves_serialized_by_cosets + compute_inner_nodes
 *                                                              libiop/bcs/merkle_tree.tcc:93-151, 201-229
 * leaf i = BLAKE2b( for oracle k, for j < coset_size: oracle_k[pos(i,j)] ), pos per domain_kind
 * (subspace.tcc:87-91 / subgroup.tcc:192-197); zk: H(H(leaf) || salt_i)  (blake2b.tcc:127-136).
 * nodes_out: (2*leaves-1)*digest_len bytes in heap order, root first (merkle_tree.tcc:145,201-229). */
int b200_merkle_build(b200_ctx *, const b200_buf *oracles, size_t n_oracles, size_t n_elems, size_t coset_size,
                      int domain_kind, const uint8_t *zk_salts_host /* NULL or leaves*salt_len */, size_t salt_len,
                      size_t digest_len, b200_bytes nodes_out);
int b200_merkle_root(b200_ctx *, b200_bytes nodes, uint8_t *root_out, size_t digest_len); /* get_root(), merkle_tree.tcc:232-240 */
/* node fetch for get_set_membership_proof (merkle_tree.tcc:243-336) */
int b200_merkle_nodes(b200_ctx *, b200_bytes nodes, const size_t *node_idx, size_t k, size_t digest_len, uint8_t *out);
/* blake2b_two_to_one_hash / blake2b_field_element_hash over a batch (blake2b.cpp:28-48, blake2b.tcc:141-160):
 * out[i] = BLAKE2b-digest_len( in[i*msg_len .. (i+1)*msg_len) ) */
int b200_blake2b_batch(b200_ctx *, b200_bytes in, size_t msg_len, size_t n_msgs, size_t digest_len, b200_bytes out);

/* ---- element-wise passes between the LDE, Merkle and FRI calls and point evaluation (N4) --------
 * field = B200_FIELD_GF192 or B200_FIELD_FP (fp != NULL); every buffer holds n elements unless stated. */
/* out[i] = sum_k coeffs[k] * vecs[k][i]: random linear combinations — multi_lincheck's f_combined_Mz
 * (lincheck/basic_lincheck_aux.tcc:117-123), combined_LDT_virtual_oracle (ldt_reducer_aux.tcc:52-66). coeffs: k*3 host limbs. */
int b200_vec_lincomb(b200_ctx *, int field, const b200_fp_params *fp, b200_buf out, size_t n, size_t k,
                     const uint64_t *coeffs, const b200_buf *vecs);
/* out[i] = a[i]*b[i] (+|-) c[i]*d[i]; b == 0: a[i]; c == 0: no second term; d == 0: c[i].  Covers fz = fw*Z + f_1v
 * (r1cs_rs_iop.tcc:216-221), lincheck's combined*p1 - fz*p2 (basic_lincheck_aux.tcc:131-137), Az*Bz - Cz (rowcheck.tcc:61-82),
 * f - (c*x^(|H|-1) + Z_H*h) (sumcheck.tcc:95-99). */
int b200_vec_mul_add(b200_ctx *, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, b200_buf b, b200_buf c,
                     b200_buf d, int subtract, size_t n);
/* out[a] = sum_{r < split} x_r[a]*y_r[a] - sum_{split <= r < rows} x_r[a]*y_r[a]; x, y: rows*n elements, row-major.  The
 * evaluations of the response polynomial of Ligero's interleaved lincheck: sum over target rows of r_i * f_i minus sum over
 * input rows of s_i * f_i (protocols/encoded/ligero/interleaved_lincheck_ot.tcc:283-330). */
int b200_vec_rows_dot(b200_ctx *, int field, const b200_fp_params *fp, b200_buf out, b200_buf x, b200_buf y, size_t n, size_t rows,
                      size_t split);
/* out[i] = a[i] * mul + add with host scalars mul / add (NULL = absent): f - claimed_eval (boundary_constraint.tcc:56-60),
 * x - y_i (lagrange_polynomial.tcc:76-85), k * v (batch_inverse_and_mul, algebra/utils.tcc:38-69). */
int b200_vec_affine(b200_ctx *, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, const uint64_t *mul, const uint64_t *add, size_t n);
/* first index whose element is zero, n if there is none (the y = x position in lagrange_polynomial.tcc:86-104) */
int b200_vec_find_zero(b200_ctx *, b200_buf a, size_t n, size_t *index_out);
/* out[i] = a[i] * table[(i >> shift) & mask] (a == 0: the table value itself): factors that repeat over cosets of H in L,
 * e.g. Z_H^-1 in rowcheck (rowcheck.tcc:36-84: additive shift = log|H|, multiplicative shift = 0, mask = |L|/|H| - 1). */
int b200_vec_mul_table(b200_ctx *, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, b200_buf table,
                       uint32_t shift, size_t mask, size_t n);
/* GF(2^192): out[i] = prod_{k < n_factors} ( T_k[0] + sum_{j: bit j of i} T_k[1+j] ), i < 2^dim; tables = n_factors x (dim+1)
 * host elements.  n_factors = 1 is all_subset_sums (algebra/utils.tcc:8-30) = any F2-affine function over an affine subspace
 * (its elements, a vanishing / linearized polynomial, x^(2^e)); more factors give subspace_element_powers
 * (algebra/exponentiation.tcc:22-54), one factor per set bit of the exponent. */
int b200_gf192_subset_sum_products(b200_ctx *, b200_buf out, uint32_t dim, uint32_t n_factors, const uint64_t *tables);
/* Fp: out[i] = a * r^i + add (add may be NULL): coset elements and their powers (exponentiation.tcc:56-72), X^|H| - h^|H| over
 * a coset (vanishing_polynomial.tcc:115-165). */
int b200_fp_geometric(b200_ctx *, const b200_fp_params *fp, b200_buf out, size_t n, const uint64_t a[3], const uint64_t r[3],
                      const uint64_t add[3]);
/* polynomial<FieldT>::evaluation_at_point (algebra/polynomials/polynomial.tcc:98-109): out = sum_i coeffs[i] * x^i (host result) */
int b200_poly_eval(b200_ctx *, int field, const b200_fp_params *fp, b200_buf coeffs, size_t n, const uint64_t x[3], uint64_t out[3]);

/* Sparse matrix-vector product over CSR rows: y[i] = sum_{k in [row_ptr[i], row_ptr[i+1])} coeff[k] * x[col[k]].
 * r1cs_constraint_system::create_Az_Bz_Cz_from_variable_assignment (relations/r1cs.tcc:236-266), one call per matrix.
 * row_ptr (n_rows+1), col (nnz), coeff (nnz*3 limbs) are host arrays; x (n_x elements) and y_out (n_rows) device buffers. */
int b200_spmv_csr(b200_ctx *, int field, const b200_fp_params *fp, size_t n_rows, const uint64_t *row_ptr, const uint32_t *col,
                  const uint64_t *coeff, b200_buf x, size_t n_x, b200_buf y_out);

/* The same product with the matrix already on the device (uploaded once per constraint system and used for A z, B z, C z and,
 * transposed, for multi_lincheck's p_alpha_ABC): row_ptr (n_rows+1 x u64), col (nnz x u32) are device byte arrays, coeff a device
 * element array.  Indices are trusted (the host-array entry point above validates them). */
int b200_spmv_csr_dev(b200_ctx *, int field, const b200_fp_params *fp, size_t n_rows, b200_bytes row_ptr, b200_bytes col, b200_buf coeff,
                      b200_buf x, b200_buf y_out);

/* Index-driven copies of 24-byte elements between device arrays (field-agnostic), idx / map device arrays of u32:
 *   b200_gather_dev   out[k] = table[map ? map[idx[k]] : idx[k]], k < n  — the row / column elements and the per-column factors of
 *                     every non-zero entry in matrix_indexer::compute_oracles_over_K (r1cs_rs_iop/fractal_indexer.tcc:52-123), read
 *                     from per-domain tables through the flattened entry list (map = field_subset::reindex_by_subset as a table)
 *   b200_scatter_dev  out[map[c]] = x[c], c < n (map injective) — compute_p_alpha_M's placement of column c at its position in
 *                     the summation domain (lincheck/common.tcc:22-27)
 *   b200_fill         dst[dst_offset + i] = value, i < n — the padding entries of the index oracles (fractal_indexer.tcc:104-121) */
int b200_gather_dev(b200_ctx *, b200_buf out, b200_buf table, b200_bytes idx, b200_bytes map, size_t n);
int b200_scatter_dev(b200_ctx *, b200_buf out, b200_buf x, b200_bytes map, size_t n);
int b200_fill(b200_ctx *, b200_buf dst, size_t dst_offset, const uint64_t value[3], size_t n);

/* row_out[k] = i for k in [row_ptr[i], row_ptr[i+1]), i < n_rows: the explicit row index of every entry of a CSR matrix whose
 * row pointers (n_rows + 1 x u64) are on the device — what the entry-list consumers above and below address entries by, made
 * on the device instead of being written and uploaded by the host (4 bytes per entry each way). */
int b200_csr_row_indices_dev(b200_ctx *, size_t n_rows, b200_bytes row_ptr, b200_bytes row_out);

/* Transpose of a device-resident sparse matrix given as an entry list (row[k], col[k], coeff[k]), k < nnz (u32 device arrays and a
 * device element array, as left by the flattening of an R1CS matrix): the entries regrouped by column as CSR over the columns —
 * col_ptr_out (n_cols + 1 x u64), src_out[k'] = the row of the entry now at k', coeff_out[k'].  The order of the entries inside one
 * column is unspecified: the only use is the sum of b200_spmv_csr_dev, which is exact and commutative in either field.  This is
 * compute_p_alpha_M's transposed product (lincheck/common.tcc:3-36) for fields without an atomic addition on the device. */
int b200_csr_transpose_dev(b200_ctx *, size_t n_cols, size_t nnz, b200_bytes col, b200_bytes row, b200_buf coeff,
                           b200_bytes col_ptr_out, b200_bytes src_out, b200_buf coeff_out);

/* GF(2^192) scatter-accumulate over a list of matrix entries: y[dst[k]] += scale * coeff[k] * x[src[k]], k < nnz.
 * multi_lincheck_virtual_oracle::set_challenge's p_alpha_ABC (protocols/encoded/lincheck/basic_lincheck_aux.tcc:63-84): the
 * transposed product (sum_m r_m M_m)^T (1, alpha, alpha^2, ...) — one call per matrix with scale = r_m, x = the powers of alpha,
 * dst = the column's position in the summation domain, src = the row.  dst, src (nnz) and coeff (nnz*3 limbs) are host arrays;
 * y (n_y elements, accumulated into) and x (n_x) are device buffers.  Additions are XOR, so entries may hit one y[.] in any order. */
int b200_gf192_scatter_mul_add(b200_ctx *, b200_buf y, size_t n_y, size_t nnz, const uint32_t *dst, const uint32_t *src,
                               const uint64_t *coeff, b200_buf x, size_t n_x, const uint64_t scale[3]);

/* the same with the entry list on the device: dst, src device byte arrays of nnz x u32, coeff a device element array (trusted) */
int b200_gf192_scatter_mul_add_dev(b200_ctx *, b200_buf y, size_t nnz, b200_bytes dst, b200_bytes src, b200_buf coeff, b200_buf x,
                                   const uint64_t scale[3]);

/* polynomial_over_vanishing_polynomial (algebra/polynomials/vanishing_polynomial.tcc:312-371,
 * linearized_polynomial.tcc:238-291): P = Q*Z + R for the sparse divisor Z = lead*y^deg + sum_j coefs[j]*y^pows[j]
 * (pows[j] < deg; additive: the linearized vanishing polynomial of a subspace, multiplicative: y^deg - c).
 * p: n_p > deg coefficients; q_out: n_p - deg; r_out: deg coefficients (0 = not wanted); lead_inv = lead^-1 (host). */
int b200_poly_divide_sparse(b200_ctx *, int field, const b200_fp_params *fp, b200_buf p, size_t n_p, size_t deg,
                            const uint64_t lead_inv[3], size_t n_terms, const uint64_t *coefs, const size_t *pows,
                            b200_buf q_out, b200_buf r_out);

/* ---- proof of work ---------------------------------------------------------- */
/* pow<FieldT, binary_hash_digest>::solve_pow (libiop/bcs/pow.tcc:87-102, acceptance pow.tcc:143-158), called at
 * bcs_prover.tcc:52-59: first candidate, in the reference's order, whose BLAKE2b(challenge || candidate) has its last
 * little-endian word & (2^pow_bitlen - 1) <= pow_upperbound.  challenge, pow_out: digest_len (= 32) host bytes. */
int b200_pow_solve(b200_ctx *, const uint8_t *challenge, size_t digest_len, uint32_t pow_bitlen, uint64_t pow_upperbound,
                   uint8_t *pow_out);

/* ---- verifier random messages of the BCS hash chain ------------------------------------------------ */
/* blake2b_FieldT_randomness_extractor (libiop/bcs/hashing/blake2b.tcc:231-256), reached from blake2b_hashchain::squeeze
 * (blake2b.tcc:76-85) in bcs_protocol::squeeze_verifier_random_messages (bcs_common.tcc:599-615): out[i] = keyed
 * BLAKE2b(digest 24 bytes, key = u64 i, message = root_plus_index) read as a field element; for B200_FIELD_FP with the
 * bits above the modulus' top bit cleared and values >= modulus re-drawn with key += n (blake2b.tcc:185-229).
 * root_plus_index: msg_len host bytes (root || u64 squeeze index; 8..128, multiple of 8). */
int b200_squeeze_field_elements(b200_ctx *, int field, const b200_fp_params *fp, const uint8_t *root_plus_index,
                                size_t msg_len, size_t n, b200_buf out);

/* ---- instrumentation (bench.py): kernel-launch counter and per-launch CUDA-event timing of the FFT tile passes ---- */
unsigned long long b200_launch_count(b200_ctx *);
void b200_transfer_bytes(b200_ctx *, unsigned long long *h2d, unsigned long long *d2h);   /* running totals of upload / download bytes */
int b200_profile_begin(b200_ctx *);
int b200_profile_end(b200_ctx *, char *buf, size_t cap);  /* "label\tms\n" per recorded launch; returns #records */

/* library identification */
const char *b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_IOP_H */
