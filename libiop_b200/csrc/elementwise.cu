/*
 * elementwise.cu — the element-wise-over-L passes that sit between the LDE, Merkle and FRI calls of the prover
 *  and polynomial evaluation at a point (N4), sm_100a.
 *
 * The reference evaluates its virtual oracles with scalar host loops over std::vector<FieldT>:
 *   multi_lincheck_virtual_oracle::evaluated_contents   protocols/encoded/lincheck/basic_lincheck_aux.tcc:102-144
 *   fz_virtual_oracle::evaluated_contents               protocols/encoded/r1cs_rs_iop/r1cs_rs_iop.tcc:181-226
 *   sumcheck_g_oracle::evaluated_contents               protocols/encoded/sumcheck/sumcheck.tcc:58-121 (+ sumcheck_aux.tcc:4-32)
 *   rowcheck_ABC_virtual_oracle::evaluated_contents     protocols/encoded/common/rowcheck.tcc:18-87
 *   combined_LDT_virtual_oracle::evaluated_contents     protocols/ldt/ldt_reducer_aux.tcc:40-129
 *   all_subset_sums / subspace_element_powers           algebra/utils.tcc:8-30, algebra/exponentiation.tcc:3-54
 *   polynomial::evaluation_at_point                     algebra/polynomials/polynomial.tcc:98-109
 * Every one of them is a short arithmetic expression per codeword position.  They are expressed here as five
 * field-generic primitives (GF(2^192) and the 3-limb Montgomery prime field); field arithmetic is exact, so any
 * evaluation order gives the reference's bits.  All are one pass over HBM: 24 B read per input element, 24 B written.
 */
#include <algorithm>
#include <cstring>
#include <vector>

#include "ctx.h"
#include "fp192.cuh"
#include "gf192.cuh"

namespace b200 {

/* ---- field policies ------------------------------------------------------------------------------------------- */
struct GfF {
    typedef gf192 T;
    __device__ __forceinline__ T add(const T &a, const T &b) const { return gf_add(a, b); }
    __device__ __forceinline__ T sub(const T &a, const T &b) const { return gf_add(a, b); }
    __device__ __forceinline__ T mul(const T &a, const T &b) const { return gf_mul(a, b); }
    __device__ __forceinline__ T zero() const { return gf_zero(); }
    __host__ T hmul(const T &a, const T &b) const { return gf_mul(a, b); }
    __host__ T hone() const { return gf_one(); }
};
struct FpF {
    typedef fp192 T;
    fp_dev_params P;
    __device__ __forceinline__ T add(const T &a, const T &b) const { return fp_add(a, b, P); }
    __device__ __forceinline__ T sub(const T &a, const T &b) const { return fp_sub(a, b, P); }
    __device__ __forceinline__ T mul(const T &a, const T &b) const { return fp_mul(a, b, P); }
    __device__ __forceinline__ T zero() const { return fp_zero(); }
    __host__ T hmul(const T &a, const T &b) const { return fp_mul(a, b, P); }
    __host__ T hone() const { return fp_one(P); }
};

template<typename T> __device__ __forceinline__ T ld_el(const T *p, size_t i)
{
    const uint64_t *q = reinterpret_cast<const uint64_t *>(p + i);
    const uint64_t q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    T r;
    r.w[0] = (uint32_t)q0; r.w[1] = (uint32_t)(q0 >> 32);
    r.w[2] = (uint32_t)q1; r.w[3] = (uint32_t)(q1 >> 32);
    r.w[4] = (uint32_t)q2; r.w[5] = (uint32_t)(q2 >> 32);
    return r;
}
template<typename T> __device__ __forceinline__ void st_el(T *p, size_t i, const T &v)
{
    uint64_t *q = reinterpret_cast<uint64_t *>(p + i);
    q[0] = (uint64_t)v.w[0] | ((uint64_t)v.w[1] << 32);
    q[1] = (uint64_t)v.w[2] | ((uint64_t)v.w[3] << 32);
    q[2] = (uint64_t)v.w[4] | ((uint64_t)v.w[5] << 32);
}
template<typename T> static T host_el(const uint64_t *p)
{
    T r;
    for (int i = 0; i < 3; ++i) { r.w[2 * i] = (uint32_t)p[i]; r.w[2 * i + 1] = (uint32_t)(p[i] >> 32); }
    return r;
}

static inline unsigned ew_grid(b200_ctx *c, size_t n, unsigned threads)
{
    const size_t want = (n + threads - 1) / threads;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count * 8));
}

/* ---- 1. random linear combination: out[i] = sum_k coeff[k] * v_k[i] ------------------------------------------ */
#define B200_LINCOMB_MAX 12
template<typename F> struct LinCombParams {
    typename F::T coeff[B200_LINCOMB_MAX];
    const typename F::T *v[B200_LINCOMB_MAX];
    typename F::T *out;
    size_t n;
    int k;
    int accumulate;   /* out[i] += ... */
};
template<typename F> __global__ void __launch_bounds__(256) lincomb_kernel(const LinCombParams<F> P, const F f)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < P.n; i += (size_t)gridDim.x * blockDim.x) {
        typename F::T acc = P.accumulate ? ld_el(P.out, i) : f.zero();
#pragma unroll 1
        for (int k = 0; k < P.k; ++k) acc = f.add(acc, f.mul(P.coeff[k], ld_el(P.v[k], i)));
        st_el(P.out, i, acc);
    }
}

/* ---- 2. out[i] = a[i]*b[i] (+|-) c[i]*d[i]; b == NULL -> a[i]; c == NULL -> no second term; d == NULL -> c[i] --- */
template<typename F> __global__ void __launch_bounds__(256)
mul_add_kernel(typename F::T *out, const typename F::T *a, const typename F::T *b, const typename F::T *c, const typename F::T *d,
               int subtract, size_t n, const F f)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        typename F::T r = ld_el(a, i);
        if (b) r = f.mul(r, ld_el(b, i));
        if (c) {
            typename F::T s = ld_el(c, i);
            if (d) s = f.mul(s, ld_el(d, i));
            r = subtract ? f.sub(r, s) : f.add(r, s);
        }
        st_el(out, i, r);
    }
}

/* ---- 2b. out[i] = a[i] * mul + add with scalar mul / add (either may be absent) -------------------------------- */
template<typename F> __global__ void __launch_bounds__(256)
affine_kernel(typename F::T *out, const typename F::T *a, const typename F::T mul, const typename F::T add, int has_mul, int has_add,
              size_t n, const F f)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        typename F::T r = ld_el(a, i);
        if (has_mul) r = f.mul(r, mul);
        if (has_add) r = f.add(r, add);
        st_el(out, i, r);
    }
}

/* ---- 2c. first index holding the zero element (n if none): the x = y position of a Lagrange denominator vector ---- */
__global__ void __launch_bounds__(256) find_zero_kernel(const uint64_t *a, size_t n, unsigned long long *first)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if ((__ldg(a + 3 * i) | __ldg(a + 3 * i + 1) | __ldg(a + 3 * i + 2)) == 0) atomicMin(first, (unsigned long long)i);
}

/* ---- 3. out[i] = a[i] * table[(i >> shift) & mask]   (values that repeat over cosets, e.g. Z_H^-1: rowcheck.tcc:36-84) */
template<typename F> __global__ void __launch_bounds__(256)
mul_table_kernel(typename F::T *out, const typename F::T *a, const typename F::T *table, uint32_t shift, size_t mask, size_t n, const F f)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const typename F::T t = ld_el(table, (i >> shift) & mask);
        st_el(out, i, a ? f.mul(ld_el(a, i), t) : t);
    }
}

/* ---- 4. GF(2^192): out[i] = prod_k ( off_k + sum_{j: bit j of i} B_k[j] ) -------------------------------------
   K = 1 is all_subset_sums (utils.tcc:8-30): the elements of an affine subspace, or any F2-affine function of them
   (a linearized / vanishing polynomial, x^(2^e)); K > 1 is subspace_element_powers (exponentiation.tcc:22-54). */
__global__ void __launch_bounds__(256)
subset_sum_products_kernel(gf192 *out, const gf192 *tables /* K x (dim+1): offset, basis[0..dim) */, uint32_t dim, uint32_t K, size_t n)
{
    extern __shared__ uint32_t sh_raw[];
    gf192 *sh = reinterpret_cast<gf192 *>(sh_raw);
    for (uint32_t t = threadIdx.x; t < K * (dim + 1); t += blockDim.x) sh[t] = tables[t];
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        gf192 acc;
#pragma unroll 1
        for (uint32_t k = 0; k < K; ++k) {
            const gf192 *tb = sh + k * (dim + 1);
            gf192 v = tb[0];
#pragma unroll 1
            for (uint32_t j = 0; j < dim; ++j)
                if ((i >> j) & 1u) gf_addto(v, tb[1 + j]);
            acc = k ? gf_mul(acc, v) : v;
        }
        st_el(out, i, acc);
    }
}

/* ---- 5. Fp: out[i] = a * r^i + c  (coset elements, their powers, X^|H| - h^|H| over a coset) -------------------- */
#define B200_GEOM_RUN 16
struct GeomParams {
    fp192 pw[40];     /* r^(RUN * 2^j) */
    fp192 a, r, c;
    fp192 *out;
    size_t n;
};
__global__ void __launch_bounds__(256) geometric_kernel(const GeomParams P, const FpF f)
{
    const size_t runs = (P.n + B200_GEOM_RUN - 1) / B200_GEOM_RUN;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < runs; t += (size_t)gridDim.x * blockDim.x) {
        fp192 cur = P.a;
#pragma unroll 1
        for (int j = 0; j < 40; ++j)
            if ((t >> j) & 1u) cur = f.mul(cur, P.pw[j]);
        const size_t base = t * B200_GEOM_RUN;
#pragma unroll 1
        for (int e = 0; e < B200_GEOM_RUN && base + e < P.n; ++e) {
            st_el(P.out, base + e, f.add(cur, P.c));
            cur = f.mul(cur, P.r);
        }
    }
}

/* ---- 6. polynomial evaluation at one point --------------------------------------------------------------------
   Each thread runs Horner over RUN consecutive coefficients; the 256 partial values of a block are combined by a
   tree whose level l multiplies by x^(RUN*2^l); one value per block is written and the same kernel is applied to
   those values with x^(RUN*256) until one is left. */
#define B200_PE_RUN 8
#define B200_PE_THREADS 256
template<typename F> struct PolyEvalParams {
    typename F::T lvl[8];     /* y^(RUN * 2^l), l = 0..7 */
    typename F::T y;
    const typename F::T *in;
    typename F::T *out;
    size_t n;
};
template<typename F> __global__ void __launch_bounds__(B200_PE_THREADS) poly_eval_kernel(const PolyEvalParams<F> P, const F f)
{
    __shared__ typename F::T sh[B200_PE_THREADS];
    const size_t base = ((size_t)blockIdx.x * B200_PE_THREADS + threadIdx.x) * B200_PE_RUN;
    typename F::T acc = f.zero();
#pragma unroll 1
    for (int e = B200_PE_RUN - 1; e >= 0; --e) {
        acc = f.mul(acc, P.y);
        if (base + e < P.n) acc = f.add(acc, ld_el(P.in, base + e));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < 8; ++l) {
        const int stride = 1 << l;
        if ((threadIdx.x & (2 * stride - 1)) == 0)
            sh[threadIdx.x] = f.add(sh[threadIdx.x], f.mul(sh[threadIdx.x + stride], P.lvl[l]));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_el(P.out, blockIdx.x, sh[0]);
}

/* ---- host side ------------------------------------------------------------------------------------------------- */
template<typename F>
static int lincomb_run(b200_ctx *c, const F &f, b200_buf out, size_t n, size_t k, const uint64_t *coeffs, const b200_buf *vecs)
{
    if (!n) return 0;
    if (!k) { B200_CUDA(c, cudaMemsetAsync((void *)(uintptr_t)out, 0, n * 24, c->stream)); return 0; }
    for (size_t done = 0; done < k; done += B200_LINCOMB_MAX) {
        LinCombParams<F> P;
        P.out = (typename F::T *)(uintptr_t)out;
        P.n = n;
        P.k = (int)std::min<size_t>(B200_LINCOMB_MAX, k - done);
        P.accumulate = done ? 1 : 0;
        for (int j = 0; j < P.k; ++j) {
            P.coeff[j] = host_el<typename F::T>(coeffs + 3 * (done + j));
            P.v[j] = (const typename F::T *)(uintptr_t)vecs[done + j];
        }
        lincomb_kernel<F><<<ew_grid(c, n, 256), 256, 0, c->stream>>>(P, f);
        B200_LAUNCH_CHECK(c);
    }
    return 0;
}

template<typename F>
static int poly_eval_run(b200_ctx *c, const F &f, b200_buf coeffs, size_t n, const uint64_t *x, uint64_t *out)
{
    typedef typename F::T T;
    if (!n) { std::memset(out, 0, 24); return 0; }
    const size_t per_block = (size_t)B200_PE_RUN * B200_PE_THREADS;
    const size_t blocks0 = (n + per_block - 1) / per_block;
    T *ping = (T *)scratch_get(c, 3, 2 * (blocks0 + 1) * sizeof(T));
    if (!ping) return set_err(c, B200_ERR_NOMEM, "cannot allocate polynomial evaluation scratch");
    T *pong = ping + blocks0 + 1;
    T y = host_el<T>(x);
    const T *in = (const T *)(uintptr_t)coeffs;
    size_t len = n;
    T *dst = ping;
    for (;;) {
        PolyEvalParams<F> P;
        P.y = y; P.in = in; P.out = dst; P.n = len;
        T p = y;                                          /* y^RUN, RUN = 2^3 */
        for (int s = 0; s < 3; ++s) p = f.hmul(p, p);
        for (int l = 0; l < 8; ++l) { P.lvl[l] = p; p = f.hmul(p, p); }
        const size_t blocks = (len + per_block - 1) / per_block;
        poly_eval_kernel<F><<<(unsigned)blocks, B200_PE_THREADS, 0, c->stream>>>(P, f);
        B200_LAUNCH_CHECK(c);
        if (blocks == 1) break;
        y = p;                                            /* y^(RUN*256): the step between block values */
        in = dst; len = blocks;
        dst = (dst == ping) ? pong : ping;
    }
    B200_CUDA(c, cudaMemcpyAsync(out, dst, 24, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}


/* ---- 7. division of a dense polynomial by a sparse monic-up-to-a-constant one ------------------
   P = Q*Z + R with Z = lead*y^deg + sum_j coef_j*y^(pow_j), pow_j <= deg/2: the vanishing polynomial of an affine
   subspace (a linearized polynomial, linearized_polynomial.tcc:238-291) or of a multiplicative coset (y^deg - c,
   vanishing_polynomial.tcc:312-358).  The reference clears the coefficients of P one at a time from the top; written
   as a recurrence on the quotient,
       q[i] = lead^-1 * ( P[deg + i] - sum_j coef_j * q[i + deg - pow_j] )          (terms with index >= |q| vanish)
       r[t] = P[t] - sum_j coef_j * q[t - pow_j]                                   (0 <= t - pow_j < |q|)
   q[i] only needs q at distance >= deg - max_j pow_j, so blocks of that many coefficients are independent: one launch
   per block from the top, in place. */
#define B200_DIV_MAX_TERMS 48
template<typename F> struct DivParams {
    typename F::T coef[B200_DIV_MAX_TERMS];
    size_t dist[B200_DIV_MAX_TERMS];          /* quotient pass: deg - pow_j; remainder pass: pow_j */
    typename F::T lead_inv;
    const typename F::T *p;                   /* quotient pass: P + deg; remainder pass: P */
    typename F::T *q;
    typename F::T *r;
    size_t qn, lo, hi;                        /* range handled by this launch */
    int n_terms;
};
template<typename F> __global__ void __launch_bounds__(256) divide_quotient_kernel(const DivParams<F> P, const F f)
{
    for (size_t i = P.lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < P.hi; i += (size_t)gridDim.x * blockDim.x) {
        typename F::T acc = ld_el(P.p, i);
#pragma unroll 1
        for (int j = 0; j < P.n_terms; ++j) {
            const size_t src = i + P.dist[j];
            if (src < P.qn) acc = f.sub(acc, f.mul(P.coef[j], ld_el(P.q, src)));
        }
        st_el(P.q, i, f.mul(acc, P.lead_inv));
    }
}
template<typename F> __global__ void __launch_bounds__(256) divide_remainder_kernel(const DivParams<F> P, const F f)
{
    for (size_t t = P.lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < P.hi; t += (size_t)gridDim.x * blockDim.x) {
        typename F::T acc = ld_el(P.p, t);
#pragma unroll 1
        for (int j = 0; j < P.n_terms; ++j) {
            if (t < P.dist[j]) continue;
            const size_t src = t - P.dist[j];
            if (src < P.qn) acc = f.sub(acc, f.mul(P.coef[j], ld_el(P.q, src)));
        }
        st_el(P.r, t, acc);
    }
}

template<typename F>
static int divide_run(b200_ctx *c, const F &f, b200_buf p, size_t n_p, size_t deg, const uint64_t *lead_inv, size_t n_terms,
                      const uint64_t *coefs, const size_t *pows, b200_buf q_out, b200_buf r_out)
{
    typedef typename F::T T;
    if (n_p <= deg) return set_err(c, B200_ERR_INVALID, "polynomial division: dividend degree below divisor degree");
    if (n_terms > B200_DIV_MAX_TERMS) return set_err(c, B200_ERR_INVALID, "polynomial division: too many divisor terms");
    size_t max_pow = 0;
    for (size_t j = 0; j < n_terms; ++j) {
        if (pows[j] >= deg) return set_err(c, B200_ERR_INVALID, "polynomial division: term degree not below the leading degree");
        max_pow = std::max(max_pow, pows[j]);
    }
    const size_t qn = n_p - deg, block = deg - max_pow;
    if ((qn + block - 1) / block > 4096) return set_err(c, B200_ERR_INVALID, "polynomial division: divisor not sparse enough (too many dependent blocks)");
    DivParams<F> P;
    P.n_terms = (int)n_terms;
    P.lead_inv = host_el<T>(lead_inv);
    for (size_t j = 0; j < n_terms; ++j) { P.coef[j] = host_el<T>(coefs + 3 * j); P.dist[j] = deg - pows[j]; }
    P.p = (const T *)(uintptr_t)p + deg; P.q = (T *)(uintptr_t)q_out; P.r = nullptr; P.qn = qn;
    for (size_t hi = qn; hi > 0;) {
        const size_t lo = hi > block ? hi - block : 0;
        P.lo = lo; P.hi = hi;
        divide_quotient_kernel<F><<<ew_grid(c, hi - lo, 256), 256, 0, c->stream>>>(P, f);
        B200_LAUNCH_CHECK(c);
        hi = lo;
    }
    if (r_out) {
        for (size_t j = 0; j < n_terms; ++j) P.dist[j] = pows[j];
        P.p = (const T *)(uintptr_t)p; P.r = (T *)(uintptr_t)r_out; P.lo = 0; P.hi = deg;
        divide_remainder_kernel<F><<<ew_grid(c, deg, 256), 256, 0, c->stream>>>(P, f);
        B200_LAUNCH_CHECK(c);
    }
    return 0;
}

/* ---- 8. sparse matrix-vector product, CSR rows: y[i] = sum_k coeff[k] * x[col[k]] --------------
   r1cs_constraint_system::create_Az_Bz_Cz_from_variable_assignment (relations/r1cs.tcc:236-266): R1CS rows hold a handful of
   terms each, so one thread per row; the gather from x is the only irregular access and x (24 MiB at N = 2^20) sits in L2. */
template<typename F> __global__ void __launch_bounds__(256)
spmv_csr_kernel(typename F::T *y, const uint64_t *row_ptr, const uint32_t *col, const typename F::T *coeff, const typename F::T *x,
                size_t n_rows, const F f)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_rows; i += (size_t)gridDim.x * blockDim.x) {
        typename F::T acc = f.zero();
        for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) acc = f.add(acc, f.mul(ld_el(coeff, k), ld_el(x, col[k])));
        st_el(y, i, acc);
    }
}

/* ---- 9. GF(2^192) scatter-accumulate over matrix entries: y[dst[k]] ^= scale * coeff[k] * x[src[k]] ------------------------
   multi_lincheck's p_alpha_ABC (basic_lincheck_aux.tcc:63-84): addition is XOR, so colliding destinations are resolved with
   three 64-bit atomic XORs and the result does not depend on the order of the entries. */
__global__ void __launch_bounds__(256)
scatter_mul_add_gf192_kernel(gf192 *y, const uint32_t *dst, const uint32_t *src, const gf192 *coeff, const gf192 *x, const gf192 scale,
                             size_t nnz)
{
    const gf192_prepared sp = gf_prepare(scale);
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nnz; k += (size_t)gridDim.x * blockDim.x) {
        const gf192 v = gf_mul(gf_mul(ld_el(coeff, k), sp), ld_el(x, src[k]));
        unsigned long long *q = reinterpret_cast<unsigned long long *>(y + dst[k]);
        atomicXor(q, (unsigned long long)v.w[0] | ((unsigned long long)v.w[1] << 32));
        atomicXor(q + 1, (unsigned long long)v.w[2] | ((unsigned long long)v.w[3] << 32));
        atomicXor(q + 2, (unsigned long long)v.w[4] | ((unsigned long long)v.w[5] << 32));
    }
}

static bool field_ok(b200_ctx *c, int field, const b200_fp_params *fp)
{
    if (field == B200_FIELD_GF192) return true;
    if (field == B200_FIELD_FP && fp) return true;
    set_err(c, B200_ERR_INVALID, "unknown field id %d (or missing prime-field parameters)", field);
    return false;
}
static FpF make_fpf(const b200_fp_params *fp)
{
    FpF f;
    f.P = make_fp_dev_params(fp->modulus, fp->r2, fp->one_mont, fp->inv);
    return f;
}

} // namespace b200

using namespace b200;

/* ---- row-wise inner product of two interleaved codes: out[a] = sum_{r < split} x_r[a] y_r[a] - sum_{r >= split} x_r[a] y_r[a]
   (x, y: rows x n, row-major).  The response polynomial of Ligero's interleaved lincheck, interleaved_lincheck_ot.tcc:283-330. */
template<typename F> __global__ void __launch_bounds__(128)
rows_dot_kernel(typename F::T *out, const typename F::T *x, const typename F::T *y, size_t n, size_t rows, size_t split, const F f)
{
    const size_t a = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (a >= n) return;
    typename F::T plus = f.zero(), minus = f.zero();
    for (size_t r = 0; r < split; ++r) plus = f.add(plus, f.mul(ld_el(x, r * n + a), ld_el(y, r * n + a)));
    for (size_t r = split; r < rows; ++r) minus = f.add(minus, f.mul(ld_el(x, r * n + a), ld_el(y, r * n + a)));
    st_el(out, a, f.sub(plus, minus));
}

extern "C" {

int b200_vec_lincomb(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, size_t n, size_t k,
                     const uint64_t *coeffs, const b200_buf *vecs)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "vec_lincomb", n * k);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (field == B200_FIELD_GF192) return lincomb_run(c, GfF(), out, n, k, coeffs, vecs);
    return lincomb_run(c, make_fpf(fp), out, n, k, coeffs, vecs);
}

int b200_vec_mul_add(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, b200_buf b, b200_buf cc,
                     b200_buf d, int subtract, size_t n)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "vec_mul_add", n);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (!n) return 0;
    if (!a) return set_err(c, B200_ERR_INVALID, "b200_vec_mul_add: operand a is required");
    if (field == B200_FIELD_GF192)
        mul_add_kernel<GfF><<<ew_grid(c, n, 256), 256, 0, c->stream>>>((gf192 *)out, (const gf192 *)a, (const gf192 *)b, (const gf192 *)cc,
                                                                      (const gf192 *)d, subtract, n, GfF());
    else
        mul_add_kernel<FpF><<<ew_grid(c, n, 256), 256, 0, c->stream>>>((fp192 *)out, (const fp192 *)a, (const fp192 *)b, (const fp192 *)cc,
                                                                      (const fp192 *)d, subtract, n, make_fpf(fp));
    B200_LAUNCH_CHECK(c);
    return 0;
}

int b200_vec_rows_dot(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf x, b200_buf y, size_t n, size_t rows,
                      size_t split)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "vec_rows_dot", n * rows);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (split > rows) return set_err(c, B200_ERR_INVALID, "b200_vec_rows_dot: split beyond the number of rows");
    if (!n) return 0;
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (field == B200_FIELD_GF192)
        rows_dot_kernel<GfF><<<grid, 128, 0, c->stream>>>((gf192 *)out, (const gf192 *)x, (const gf192 *)y, n, rows, split, GfF());
    else
        rows_dot_kernel<FpF><<<grid, 128, 0, c->stream>>>((fp192 *)out, (const fp192 *)x, (const fp192 *)y, n, rows, split, make_fpf(fp));
    B200_LAUNCH_CHECK(c);
    return 0;
}

int b200_vec_affine(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, const uint64_t *mul, const uint64_t *add, size_t n)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "vec_affine", n);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (!n) return 0;
    if (field == B200_FIELD_GF192)
        affine_kernel<GfF><<<ew_grid(c, n, 256), 256, 0, c->stream>>>((gf192 *)out, (const gf192 *)a, mul ? host_el<gf192>(mul) : gf_zero(),
                                                                     add ? host_el<gf192>(add) : gf_zero(), mul != nullptr, add != nullptr, n, GfF());
    else
        affine_kernel<FpF><<<ew_grid(c, n, 256), 256, 0, c->stream>>>((fp192 *)out, (const fp192 *)a, mul ? host_el<fp192>(mul) : fp_zero(),
                                                                     add ? host_el<fp192>(add) : fp_zero(), mul != nullptr, add != nullptr, n, make_fpf(fp));
    B200_LAUNCH_CHECK(c);
    return 0;
}

int b200_vec_find_zero(b200_ctx *c, b200_buf a, size_t n, size_t *index_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "vec_find_zero", n);
    unsigned long long *dev = (unsigned long long *)scratch_get(c, 3, 64);
    if (!dev) return set_err(c, B200_ERR_NOMEM, "cannot allocate scratch");
    unsigned long long first = (unsigned long long)n;
    B200_CUDA(c, cudaMemcpyAsync(dev, &first, 8, cudaMemcpyHostToDevice, c->stream));
    if (n) {
        find_zero_kernel<<<ew_grid(c, n, 256), 256, 0, c->stream>>>((const uint64_t *)(uintptr_t)a, n, dev);
        B200_LAUNCH_CHECK(c);
    }
    B200_CUDA(c, cudaMemcpyAsync(&first, dev, 8, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    *index_out = (size_t)first;
    return 0;
}

int b200_vec_mul_table(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf out, b200_buf a, b200_buf table,
                       uint32_t shift, size_t mask, size_t n)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "vec_mul_table", n);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (!n) return 0;
    if (field == B200_FIELD_GF192)
        mul_table_kernel<GfF><<<ew_grid(c, n, 256), 256, 0, c->stream>>>((gf192 *)out, (const gf192 *)a, (const gf192 *)table, shift, mask, n, GfF());
    else
        mul_table_kernel<FpF><<<ew_grid(c, n, 256), 256, 0, c->stream>>>((fp192 *)out, (const fp192 *)a, (const fp192 *)table, shift, mask, n, make_fpf(fp));
    B200_LAUNCH_CHECK(c);
    return 0;
}

int b200_gf192_subset_sum_products(b200_ctx *c, b200_buf out, uint32_t dim, uint32_t n_factors, const uint64_t *tables)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "subset_sum_products", ((size_t)n_factors << 32) | dim);
    if (dim > 40 || n_factors == 0 || n_factors > 64) return set_err(c, B200_ERR_INVALID, "subset sums: dim <= 40 and 1..64 factors");
    const size_t n = (size_t)1 << dim;
    const size_t tb = (size_t)n_factors * (dim + 1) * 24;
    void *dev = scratch_get(c, 3, tb);
    if (!dev) return set_err(c, B200_ERR_NOMEM, "cannot allocate subset-sum tables");
    void *stage = staging_get(c, tb);
    if (!stage) return set_err(c, B200_ERR_NOMEM, "cannot allocate staging");
    B200_CUDA(c, cudaStreamSynchronize(c->stream));           /* staging may still feed an earlier copy */
    std::memcpy(stage, tables, tb);
    B200_CUDA(c, cudaMemcpyAsync(dev, stage, tb, cudaMemcpyHostToDevice, c->stream));
    if (tb > 48 * 1024)
        B200_CUDA(c, cudaFuncSetAttribute(subset_sum_products_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb));
    subset_sum_products_kernel<<<ew_grid(c, n, 256), 256, tb, c->stream>>>((gf192 *)out, (const gf192 *)dev, dim, n_factors, n);
    B200_LAUNCH_CHECK(c);
    return 0;
}

int b200_fp_geometric(b200_ctx *c, const b200_fp_params *fp, b200_buf out, size_t n, const uint64_t a[3], const uint64_t r[3],
                      const uint64_t add[3])
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "fp_geometric", n);
    if (!fp) return set_err(c, B200_ERR_INVALID, "missing prime-field parameters");
    if (!n) return 0;
    const FpF f = make_fpf(fp);
    GeomParams P;
    P.a = host_el<fp192>(a); P.r = host_el<fp192>(r);
    P.c = add ? host_el<fp192>(add) : fp_zero();
    P.out = (fp192 *)(uintptr_t)out; P.n = n;
    fp192 p = P.r;
    for (int s = 0; s < 4; ++s) p = f.hmul(p, p);          /* r^RUN, RUN = 2^4 */
    for (int j = 0; j < 40; ++j) { P.pw[j] = p; p = f.hmul(p, p); }
    const size_t runs = (n + B200_GEOM_RUN - 1) / B200_GEOM_RUN;
    geometric_kernel<<<ew_grid(c, runs, 256), 256, 0, c->stream>>>(P, f);
    B200_LAUNCH_CHECK(c);
    return 0;
}

int b200_poly_eval(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf coeffs, size_t n, const uint64_t x[3], uint64_t out[3])
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "poly_eval", n);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (field == B200_FIELD_GF192) return poly_eval_run(c, GfF(), coeffs, n, x, out);
    return poly_eval_run(c, make_fpf(fp), coeffs, n, x, out);
}

int b200_spmv_csr(b200_ctx *c, int field, const b200_fp_params *fp, size_t n_rows, const uint64_t *row_ptr, const uint32_t *col,
                  const uint64_t *coeff, b200_buf x, size_t n_x, b200_buf y_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "spmv_csr", n_rows);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (!n_rows) return 0;
    const size_t nnz = (size_t)row_ptr[n_rows];
    for (size_t k = 0; k < nnz; ++k)
        if (col[k] >= n_x) return set_err(c, B200_ERR_INVALID, "spmv: column index %u outside the vector (%zu)", col[k], n_x);
    /* matrix goes up with the call (the prover builds a fresh constraint system per proof): row_ptr | coeff | col */
    const size_t b_rp = (n_rows + 1) * sizeof(uint64_t), b_cf = nnz * 24, b_cl = nnz * sizeof(uint32_t);
    char *dev = (char *)scratch_get(c, 2, b_rp + b_cf + b_cl + 64);
    if (!dev) return set_err(c, B200_ERR_NOMEM, "cannot allocate %zu bytes for the sparse matrix", b_rp + b_cf + b_cl);
    B200_CUDA(c, cudaMemcpyAsync(dev, row_ptr, b_rp, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaMemcpyAsync(dev + b_rp, coeff, b_cf, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaMemcpyAsync(dev + b_rp + b_cf, col, b_cl, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += b_rp + b_cf + b_cl;
    const uint64_t *d_rp = (const uint64_t *)dev;
    const uint32_t *d_cl = (const uint32_t *)(dev + b_rp + b_cf);
    if (field == B200_FIELD_GF192)
        spmv_csr_kernel<GfF><<<ew_grid(c, n_rows, 256), 256, 0, c->stream>>>((gf192 *)y_out, d_rp, d_cl, (const gf192 *)(dev + b_rp),
                                                                            (const gf192 *)x, n_rows, GfF());
    else
        spmv_csr_kernel<FpF><<<ew_grid(c, n_rows, 256), 256, 0, c->stream>>>((fp192 *)y_out, d_rp, d_cl, (const fp192 *)(dev + b_rp),
                                                                            (const fp192 *)x, n_rows, make_fpf(fp));
    B200_LAUNCH_CHECK(c);
    B200_CUDA(c, cudaStreamSynchronize(c->stream));      /* the host arrays may be released by the caller */
    return 0;
}

int b200_spmv_csr_dev(b200_ctx *c, int field, const b200_fp_params *fp, size_t n_rows, b200_bytes row_ptr, b200_bytes col, b200_buf coeff,
                      b200_buf x, b200_buf y_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "spmv_csr_dev", n_rows);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (!n_rows) return 0;
    if (field == B200_FIELD_GF192)
        spmv_csr_kernel<GfF><<<ew_grid(c, n_rows, 256), 256, 0, c->stream>>>((gf192 *)y_out, (const uint64_t *)row_ptr, (const uint32_t *)col,
                                                                            (const gf192 *)coeff, (const gf192 *)x, n_rows, GfF());
    else
        spmv_csr_kernel<FpF><<<ew_grid(c, n_rows, 256), 256, 0, c->stream>>>((fp192 *)y_out, (const uint64_t *)row_ptr, (const uint32_t *)col,
                                                                            (const fp192 *)coeff, (const fp192 *)x, n_rows, make_fpf(fp));
    B200_LAUNCH_CHECK(c);
    return 0;
}
int b200_gf192_scatter_mul_add_dev(b200_ctx *c, b200_buf y, size_t nnz, b200_bytes dst, b200_bytes src, b200_buf coeff, b200_buf x,
                                   const uint64_t scale[3])
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "scatter_mul_add_dev", nnz);
    if (!nnz) return 0;
    scatter_mul_add_gf192_kernel<<<ew_grid(c, nnz, 256), 256, 0, c->stream>>>((gf192 *)y, (const uint32_t *)dst, (const uint32_t *)src,
                                                                             (const gf192 *)coeff, (const gf192 *)x, host_el<gf192>(scale), nnz);
    B200_LAUNCH_CHECK(c);
    return 0;
}
int b200_gf192_scatter_mul_add(b200_ctx *c, b200_buf y, size_t n_y, size_t nnz, const uint32_t *dst, const uint32_t *src,
                               const uint64_t *coeff, b200_buf x, size_t n_x, const uint64_t scale[3])
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "scatter_mul_add", nnz);
    if (!nnz) return 0;
    for (size_t k = 0; k < nnz; ++k)
        if (dst[k] >= n_y || src[k] >= n_x)
            return set_err(c, B200_ERR_INVALID, "scatter: entry %zu (%u <- %u) outside the vectors (%zu, %zu)", k, dst[k], src[k], n_y, n_x);
    /* the entry list goes up with the call: coeff | dst | src */
    const size_t b_cf = nnz * 24, b_ix = nnz * sizeof(uint32_t);
    char *dev = (char *)scratch_get(c, 2, b_cf + 2 * b_ix + 64);
    if (!dev) return set_err(c, B200_ERR_NOMEM, "cannot allocate %zu bytes for the entry list", b_cf + 2 * b_ix);
    B200_CUDA(c, cudaMemcpyAsync(dev, coeff, b_cf, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaMemcpyAsync(dev + b_cf, dst, b_ix, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(c, cudaMemcpyAsync(dev + b_cf + b_ix, src, b_ix, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += b_cf + 2 * b_ix;
    scatter_mul_add_gf192_kernel<<<ew_grid(c, nnz, 256), 256, 0, c->stream>>>((gf192 *)y, (const uint32_t *)(dev + b_cf),
                                                                             (const uint32_t *)(dev + b_cf + b_ix), (const gf192 *)dev,
                                                                             (const gf192 *)x, host_el<gf192>(scale), nnz);
    B200_LAUNCH_CHECK(c);
    B200_CUDA(c, cudaStreamSynchronize(c->stream));      /* the host arrays may be released by the caller */
    return 0;
}
int b200_poly_divide_sparse(b200_ctx *c, int field, const b200_fp_params *fp, b200_buf p, size_t n_p, size_t deg,
                            const uint64_t lead_inv[3], size_t n_terms, const uint64_t *coefs, const size_t *pows,
                            b200_buf q_out, b200_buf r_out)
{
    B200_ENTER(c);
    B200_TRACE_SCOPE(c, "poly_divide_sparse", n_p);
    if (!field_ok(c, field, fp)) return B200_ERR_INVALID;
    if (field == B200_FIELD_GF192) return divide_run(c, GfF(), p, n_p, deg, lead_inv, n_terms, coefs, pows, q_out, r_out);
    return divide_run(c, make_fpf(fp), p, n_p, deg, lead_inv, n_terms, coefs, pows, q_out, r_out);
}

} /* extern "C" */
