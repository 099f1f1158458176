/*
 * ntt.cu — radix-2 NTT / inverse NTT over a 3-limb Montgomery prime field on multiplicative cosets.
 *
 * Replaces multiplicative_FFT_degree_aware (libiop/algebra/fft.tcc:236-317) and
 * multiplicative_IFFT_internal (fft.tcc:343-361 -> libfqfft basic_radix2_domain::iFFT / icosetFFT,
 * depends/libfqfft/libfqfft/evaluation_domain/domains/basic_radix2_domain.tcc:49-74,
 * basic_radix2_domain_aux.tcc:45-79,172-180).  Output idx is p(shift * g^idx) (subgroup.tcc:280-304),
 * values in Montgomery form, fully reduced — unique, hence bit-identical to the reference.
 *
 * B200 design: the array stays in HBM/L2; each launch stages a tile of 2^TB elements in shared
 * memory and runs every butterfly level whose partner bit lies in the tile (same tile-pass engine
 * as additive_fft.cu).  Like the reference's "degree aware" variant, a polynomial with 2^d
 * coefficients on a 2^n domain skips the first n-d levels: the bit-reversed, replicated start
 * vector is produced by the gather of the first pass.  Twiddles g^k are cached per domain.
 */
#include <algorithm>
#include <cstring>

#include "ctx.h"
#include "fp192.cuh"

namespace b200 {

struct NttPlan {
    uint32_t logn = 0;
    fp192 g, ginv;
    /* Twiddles, one contiguous table per butterfly level: level s (half-size 2^s) uses w_s^j, j < 2^s, with w_s the
       primitive 2^(s+1)-th root g^(2^(logn-1-s)); table s starts at entry 2^s - 1 (2^logn - 1 entries in all).  A single
       table g^k read with stride 2^(logn-1-s) costs a separate 32-byte sector per butterfly at every level but the last
       (ncu, round 2: 589 MB of DRAM reads for a pass whose data is 192 MB); per-level tables are read in runs. */
    fp192 *w = nullptr;
    fp192 *winv = nullptr;
    ~NttPlan() { cudaFree(w); cudaFree(winv); }
};

struct NttPassParams {
    const fp192 *src;
    fp192 *dst;
    const fp192 *w;
    fp_dev_params fp;
    uint32_t logn, lo, hi, c;
    uint32_t first_stage, n_stages;  /* stages first_stage .. first_stage+n_stages-1 (bit positions) */
    uint32_t gather_d;               /* != 0: src holds 2^gather_d coefficients, load src[bitrev_logn(i) & (2^d-1)] */
    uint32_t n_src;                  /* number of valid source coefficients when gathering */
};

__device__ __forceinline__ fp192 ldg_fp(const fp192 *p)
{
    const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
    const uint64_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    fp192 r;
    r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
    r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
    r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
    return r;
}

struct FpTile {
    uint64_t *p0, *p1, *p2;
    __device__ __forceinline__ fp192 ld(uint32_t L) const
    {
        const uint64_t a = p0[L], b = p1[L], c = p2[L];
        fp192 r;
        r.w[0] = (uint32_t)a; r.w[1] = (uint32_t)(a >> 32);
        r.w[2] = (uint32_t)b; r.w[3] = (uint32_t)(b >> 32);
        r.w[4] = (uint32_t)c; r.w[5] = (uint32_t)(c >> 32);
        return r;
    }
    __device__ __forceinline__ void st(uint32_t L, const fp192 &x) const
    {
        p0[L] = (uint64_t)x.w[0] | ((uint64_t)x.w[1] << 32);
        p1[L] = (uint64_t)x.w[2] | ((uint64_t)x.w[3] << 32);
        p2[L] = (uint64_t)x.w[4] | ((uint64_t)x.w[5] << 32);
    }
};

#define B200_NTT_PAD 4

__global__ void __launch_bounds__(256) ntt_tile_pass(const NttPassParams P)
{
    extern __shared__ uint64_t smem[];
    const uint32_t mainb = P.hi - P.lo;
    const uint32_t tb = mainb + P.c;
    const uint32_t T = 1u << tb;
    FpTile t;
    t.p0 = smem; t.p1 = smem + T + B200_NTT_PAD; t.p2 = smem + 2 * (T + B200_NTT_PAD);
    const uint32_t midbits = P.lo - P.c;
    const uint32_t mid = blockIdx.x & ((1u << midbits) - 1u);
    const uint32_t top = blockIdx.x >> midbits;

    auto global_index = [&](uint32_t L) -> uint32_t {
        const uint32_t w = L & ((1u << P.c) - 1u);
        const uint32_t u = L >> P.c;
        return w | (mid << P.c) | (u << P.lo) | (top << P.hi);
    };

    for (uint32_t L = threadIdx.x; L < T; L += blockDim.x) {
        const uint32_t i = global_index(L);
        if (P.gather_d) {
            const uint32_t k = (__brev(i) >> (32 - P.logn)) & ((1u << P.gather_d) - 1u);
            t.st(L, k < P.n_src ? ldg_fp(P.src + k) : fp_zero());
        } else {
            t.st(L, ldg_fp(P.src + i));
        }
    }
    for (uint32_t s = P.first_stage; s < P.first_stage + P.n_stages; ++s) {
        __syncthreads();
        /* a[k+j+m] = a[k+j] - w^j a[k+j+m]; a[k+j] += w^j a[k+j+m]   (fft.tcc:302-312, aux.tcc:65-75) */
        const uint32_t pos = s - P.lo + P.c;
        for (uint32_t k = threadIdx.x; k < (T >> 1); k += blockDim.x) {
            const uint32_t low = k & ((1u << pos) - 1u);
            const uint32_t L0 = ((k >> pos) << (pos + 1)) | low, L1 = L0 | (1u << pos);
            const uint32_t j = global_index(L0) & ((1u << s) - 1u);
            const fp192 a = t.ld(L0);
            fp192 b = t.ld(L1);
            if (j != 0) b = fp_mul(b, ldg_fp(P.w + (((size_t)1 << s) - 1) + j), P.fp);
            t.st(L0, fp_add(a, b, P.fp));
            t.st(L1, fp_sub(a, b, P.fp));
        }
    }
    __syncthreads();
    uint64_t *out = reinterpret_cast<uint64_t *>(P.dst);
    for (uint32_t L = threadIdx.x; L < T; L += blockDim.x) {
        uint64_t *q = out + (size_t)global_index(L) * 3;
        q[0] = t.p0[L]; q[1] = t.p1[L]; q[2] = t.p2[L];
    }
}

/* tab[len + e] = tab[e] * factor */
__global__ void fp_power_table_step(fp192 *tab, fp192 factor, uint32_t len, fp_dev_params P)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < len) tab[len + e] = fp_mul(tab[e], factor, P);
}

/* x[i] *= c * base^i  (coset scaling: aux.tcc:172-180; iFFT 1/n: basic_radix2_domain.tcc:55-59).
   base^i is assembled from two 2^10-entry tables (i = hi*1024 + lo). */
__global__ void fp_scale_by_powers(fp192 *x, size_t n, const fp192 *pow_lo, const fp192 *pow_hi, fp_dev_params P)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fp192 f = ldg_fp(pow_lo + (i & 1023));
        if (i >> 10) f = fp_mul(f, ldg_fp(pow_hi + (i >> 10)), P);
        x[i] = fp_mul(x[i], f, P);
    }
}

__global__ void fp_fill_kernel(fp192 *dst, const fp192 *src_elem, size_t n)
{
    const fp192 v = src_elem ? *src_elem : fp_zero();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

static fp_dev_params dev_params(const b200_fp_params *P)
{
    return make_fp_dev_params(P->modulus, P->r2, P->one_mont, P->inv);
}
static fp192 from_limbs(const uint64_t v[3])
{
    fp192 r;
    for (int i = 0; i < 3; ++i) { r.w[2 * i] = (uint32_t)v[i]; r.w[2 * i + 1] = (uint32_t)(v[i] >> 32); }
    return r;
}

static int build_fp_power_table(b200_ctx *c, fp192 *tab, const fp192 &first, fp192 base, uint32_t logn_entries, const fp_dev_params &fp)
{
    B200_CUDA(c, cudaMemcpyAsync(tab, &first, sizeof first, cudaMemcpyHostToDevice, c->stream));
    for (uint32_t b = 0; b < logn_entries; ++b) {
        const uint32_t len = 1u << b;
        fp_power_table_step<<<(len + 255) / 256, 256, 0, c->stream>>>(tab, base, len, fp);
        base = fp_mul(base, base, fp);
    }
    B200_LAUNCH_CHECK(c);
    return 0;
}

static int get_ntt_plan(b200_ctx *c, const b200_fp_params *P, const b200_domain *D, bool inverse, std::shared_ptr<NttPlan> *out)
{
    const fp_dev_params fp = dev_params(P);
    std::string key((const char *)P->modulus, 24);
    key.append((const char *)D->generator, 24);
    key.push_back((char)D->dim);
    std::shared_ptr<NttPlan> pl;
    auto it = c->ntt_plans.find(key);
    if (it != c->ntt_plans.end()) pl = it->second;
    else {
        pl = std::make_shared<NttPlan>();
        pl->logn = D->dim;
        pl->g = from_limbs(D->generator);
        pl->ginv = fp_inv(pl->g, fp);
        c->ntt_plans[key] = pl;
    }
    const uint32_t logn = D->dim;
    fp192 **slot = inverse ? &pl->winv : &pl->w;
    if (!*slot) {
        fp192 *tab = nullptr;
        const size_t entries = ((size_t)1 << logn);           /* sum over levels of 2^s, rounded up */
        if (cudaMalloc(&tab, entries * sizeof(fp192)) != cudaSuccess) {
            cudaGetLastError();
            return set_err(c, B200_ERR_NOMEM, "cannot allocate NTT twiddle tables");
        }
        /* level s: powers of w_s = g^(2^(logn-1-s)); computed from the top level down by squaring the root */
        fp192 root = inverse ? pl->ginv : pl->g;               /* w_(logn-1) = g */
        for (uint32_t s = logn; s-- > 0;) {
            int rc = build_fp_power_table(c, tab + (((size_t)1 << s) - 1), fp_one(fp), root, s, fp);
            if (rc) { cudaFree(tab); return rc; }
            root = fp_mul(root, root, fp);
        }
        B200_CUDA(c, cudaStreamSynchronize(c->stream));
        *slot = tab;
    }
    *out = pl;
    return 0;
}

#ifndef B200_NTT_TILE_BITS
#define B200_NTT_TILE_BITS 11   /* 48 KiB tiles, 4 CTAs per SM: 2^20 -> 2^23 in 1.92 ms vs 2.14 ms with 2^12 (profiles/r2_variants.md) */
#endif
static const uint32_t kNttTileBits = B200_NTT_TILE_BITS;
static const uint32_t kNttCoalesceBits = 4;

static int run_stages(b200_ctx *c, NttPassParams P, uint32_t logn, uint32_t first, const fp192 *src, fp192 *dst,
                      uint32_t gather_d, uint32_t n_src)
{
    bool &attr = c->attr_ntt;
    if (!attr) {
        B200_CUDA(c, cudaFuncSetAttribute(ntt_tile_pass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(3 * (((size_t)1 << 13) + B200_NTT_PAD) * sizeof(uint64_t))));
        attr = true;
    }
    const uint32_t tb = std::min(kNttTileBits, logn);
    uint32_t s = first;
    bool first_pass = true;
    while (s < logn || first_pass) {
        uint32_t lo, hi, cb;
        if (s < tb) { lo = 0; cb = 0; hi = tb; }                      /* contiguous tile covers stages s..tb-1 */
        else {
            cb = kNttCoalesceBits; lo = s; hi = std::min(logn, s + tb - cb);
            while (hi - lo + cb < tb && lo > cb) --lo;
            if (lo == cb) { lo = 0; cb = 0; }
        }
        P.lo = lo; P.hi = hi; P.c = cb; P.logn = logn;
        P.first_stage = s; P.n_stages = hi > s ? hi - s : 0;
        P.src = first_pass ? src : dst; P.dst = dst;
        P.gather_d = first_pass ? gather_d : 0; P.n_src = n_src;
        const uint32_t tbits = (hi - lo) + cb;
        const size_t smem = 3 * (((size_t)1 << tbits) + B200_NTT_PAD) * sizeof(uint64_t);
        ntt_tile_pass<<<1u << (logn - tbits), 256, smem, c->stream>>>(P);
        B200_LAUNCH_CHECK(c);
        s = hi;
        first_pass = false;
    }
    return 0;
}

static bool is_one(const uint64_t v[3], const b200_fp_params *P)
{
    return v[0] == P->one_mont[0] && v[1] == P->one_mont[1] && v[2] == P->one_mont[2];
}

/* scale x[i] by c0 * base^i using two small power tables built on the fly */
static int scale_by_powers(b200_ctx *c, fp192 *x, size_t n, const fp192 &c0, const fp192 &base, const fp_dev_params &fp)
{
    const uint32_t logn = ceil_log2(n);
    const uint32_t hi_bits = logn > 10 ? logn - 10 : 0;
    fp192 *tabs = (fp192 *)scratch_get(c, 2, (1024 + ((size_t)1 << hi_bits)) * sizeof(fp192));
    if (!tabs) return set_err(c, B200_ERR_NOMEM, "cannot allocate power tables");
    fp192 *lo = tabs, *hi = tabs + 1024;
    int rc = build_fp_power_table(c, lo, c0, base, 10, fp);
    if (rc) return rc;
    rc = build_fp_power_table(c, hi, fp_one(fp), fp_pow_u64(base, 1024, fp), hi_bits, fp);
    if (rc) return rc;
    fp_scale_by_powers<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, c->stream>>>(x, n, lo, hi, fp);
    B200_LAUNCH_CHECK(c);
    B200_CUDA(c, cudaStreamSynchronize(c->stream)); /* c0 / base temporaries were copied from the stack */
    return 0;
}

int fp_ntt(b200_ctx *c, const b200_fp_params *Pm, const b200_domain *L, b200_buf coeffs, size_t n_coeffs, b200_buf out)
{
    if (!Pm || !L || L->kind != B200_DOMAIN_MULTIPLICATIVE_COSET)
        return set_err(c, B200_ERR_INVALID, "multiplicative FFT needs a multiplicative coset domain and field parameters");
    const uint32_t logn = L->dim;
    if (logn > Pm->two_adicity || logn > 31) return set_err(c, B200_ERR_INVALID, "domain size 2^%u exceeds the field's 2-adicity", logn);
    const size_t n = (size_t)1 << logn;
    if (n_coeffs > n) return set_err(c, B200_ERR_INVALID, "more coefficients than domain elements");
    const fp_dev_params fp = dev_params(Pm);
    fp192 *o = (fp192 *)out;
    if (n_coeffs <= 1 || logn == 0) {
        fp_fill_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1184), 256, 0, c->stream>>>(o, n_coeffs ? (const fp192 *)coeffs : nullptr, n);
        B200_LAUNCH_CHECK(c);
        return 0;
    }
    std::shared_ptr<NttPlan> pl;
    int rc = get_ntt_plan(c, Pm, L, false, &pl);
    if (rc) return rc;
    const uint32_t d = ceil_log2(n_coeffs);
    const fp192 *src = (const fp192 *)coeffs;
    if (!is_one(L->shift, Pm)) {           /* coefficient i times shift^i (fft.tcc:246-249) */
        fp192 *X = (fp192 *)scratch_get(c, 0, n_coeffs * sizeof(fp192));
        if (!X) return set_err(c, B200_ERR_NOMEM, "cannot allocate NTT work array");
        B200_CUDA(c, cudaMemcpyAsync(X, src, n_coeffs * sizeof(fp192), cudaMemcpyDeviceToDevice, c->stream));
        rc = scale_by_powers(c, X, n_coeffs, fp_one(fp), from_limbs(L->shift), fp);
        if (rc) return rc;
        src = X;
    }
    NttPassParams P; memset(&P, 0, sizeof P);
    P.fp = fp; P.w = pl->w;
    /* bit-reverse + replicate (fft.tcc:267-289) fused into the first pass; levels logn-d .. logn-1 (fft.tcc:295-315) */
    return run_stages(c, P, logn, logn - d, src, o, d, (uint32_t)n_coeffs);
}

int fp_intt(b200_ctx *c, const b200_fp_params *Pm, const b200_domain *H, b200_buf evals, b200_buf out)
{
    if (!Pm || !H || H->kind != B200_DOMAIN_MULTIPLICATIVE_COSET)
        return set_err(c, B200_ERR_INVALID, "multiplicative IFFT needs a multiplicative coset domain and field parameters");
    const uint32_t logn = H->dim;
    if (logn > Pm->two_adicity || logn > 31) return set_err(c, B200_ERR_INVALID, "domain size 2^%u exceeds the field's 2-adicity", logn);
    const size_t n = (size_t)1 << logn;
    if (logn == 0) {                       /* fft.tcc:397-401 */
        B200_CUDA(c, cudaMemcpyAsync((void *)out, (const void *)evals, sizeof(fp192), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    const fp_dev_params fp = dev_params(Pm);
    std::shared_ptr<NttPlan> pl;
    int rc = get_ntt_plan(c, Pm, H, true, &pl);
    if (rc) return rc;
    const fp192 *src = (const fp192 *)evals;
    fp192 *o = (fp192 *)out;
    if ((const void *)src == (void *)o) {  /* the gather of the first pass is not in-place safe */
        fp192 *X = (fp192 *)scratch_get(c, 0, n * sizeof(fp192));
        if (!X) return set_err(c, B200_ERR_NOMEM, "cannot allocate NTT work array");
        B200_CUDA(c, cudaMemcpyAsync(X, src, n * sizeof(fp192), cudaMemcpyDeviceToDevice, c->stream));
        src = X;
    }
    NttPassParams P; memset(&P, 0, sizeof P);
    P.fp = fp; P.w = pl->winv;
    rc = run_stages(c, P, logn, 0, src, o, logn, (uint32_t)n);   /* FFT with omega^-1 (basic_radix2_domain.tcc:53) */
    if (rc) return rc;
    /* times 1/n (basic_radix2_domain.tcc:55-59) and shift^-i for cosets (basic_radix2_domain.tcc:70-74) */
    const fp192 ninv = fp_inv(fp_from_u64((uint64_t)n, fp), fp);
    const fp192 sinv = is_one(H->shift, Pm) ? fp_one(fp) : fp_inv(from_limbs(H->shift), fp);
    return scale_by_powers(c, o, n, ninv, sinv, fp);
}

/* ---- element-wise helpers ---- */
__global__ void fp_binop_kernel(fp192 *out, const fp192 *a, const fp192 *b, size_t n, int op, fp_dev_params P)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fp192 r;
        if (op == 0) r = fp_mul(a[i], b[i], P);
        else if (op == 1) r = fp_add(a[i], b[i], P);
        else r = fp_sub(a[i], b[i], P);
        out[i] = r;
    }
}

/* Element-wise inversion, Montgomery's trick per thread (the reference's batch_inverse, algebra/utils.tcc:38-69, runs it over
   the whole vector sequentially): RUN consecutive elements share one Fermat inversion for 3 multiplications each.
   Zero maps to zero (the reference's has_zeroes path, utils.tcc:73-103). */
#define B200_FP_INV_RUN 16
__global__ void __launch_bounds__(128) fp_inv_kernel(fp192 *out, const fp192 *a, size_t n, fp_dev_params P)
{
    const size_t runs = (n + B200_FP_INV_RUN - 1) / B200_FP_INV_RUN;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < runs; t += (size_t)gridDim.x * blockDim.x) {
        const size_t base = t * B200_FP_INV_RUN;
        const int cnt = (int)((n - base < B200_FP_INV_RUN) ? n - base : B200_FP_INV_RUN);
        fp192 x[B200_FP_INV_RUN], pre[B200_FP_INV_RUN];
        fp192 acc = fp_one(P);
#pragma unroll 1
        for (int e = 0; e < cnt; ++e) {
            x[e] = a[base + e];
            pre[e] = acc;
            if (!fp_is_zero(x[e])) acc = fp_mul(acc, x[e], P);
        }
        fp192 inv = fp_inv(acc, P);
#pragma unroll 1
        for (int e = cnt - 1; e >= 0; --e) {
            if (fp_is_zero(x[e])) { out[base + e] = fp_zero(); continue; }
            out[base + e] = fp_mul(inv, pre[e], P);
            inv = fp_mul(inv, x[e], P);
        }
    }
}

int fp_binop_vec(b200_ctx *c, const b200_fp_params *Pm, int op, b200_buf out, b200_buf a, b200_buf b, size_t n)
{
    if (!Pm) return set_err(c, B200_ERR_INVALID, "field parameters required");
    if (!n) return 0;
    if (op == 3) {
        const size_t runs = (n + B200_FP_INV_RUN - 1) / B200_FP_INV_RUN;
        fp_inv_kernel<<<(unsigned)std::min<size_t>((runs + 127) / 128, 148 * 16), 128, 0, c->stream>>>((fp192 *)out, (const fp192 *)a, n, dev_params(Pm));
    } else {
        fp_binop_kernel<<<(unsigned)std::min<size_t>((n + 127) / 128, 148 * 8), 128, 0, c->stream>>>(
            (fp192 *)out, (const fp192 *)a, (const fp192 *)(b ? b : a), n, op, dev_params(Pm));
    }
    B200_LAUNCH_CHECK(c);
    return 0;
}

} // namespace b200
