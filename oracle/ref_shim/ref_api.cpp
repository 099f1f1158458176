/*
 * oracle/ref_shim/ref_api.cpp — C entry points onto the UNMODIFIED reference (TEST INFRASTRUCTURE).
 *
 * Compiled by oracle/Makefile against the sources where they lie under /root/reference
 * (libiop + vendored libff/libfqfft); output goes to oracle/_ref/libiopref.so.  No reference
 * source is copied into this repository: this file only *calls* the reference templates
 *   FFT_over_field_subset / IFFT_over_field_subset      libiop/algebra/fft.hpp:62-88
 *   evaluate_next_f_i_over_entire_domain                libiop/protocols/ldt/fri/fri_aux.hpp:23-28
 *   merkle_tree::construct_with_leaves_serialized_by_cosets  libiop/bcs/merkle_tree.hpp:88-90
 *   blake2b_hashchain                                   libiop/bcs/hashing/blake2b.hpp
 *   aurora_snark_prover / aurora_snark_verifier         libiop/snark/aurora_snark.hpp:75-92
 * It is used (a) to pin oracle/*.c, (b) to generate tests/golden/ fixtures, and (c) as the
 * `cpu_baseline.kind = "reference"` arm of bench.py.
 */
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include <libff/algebra/fields/binary/gf192.hpp>
#include <libff/algebra/curves/edwards/edwards_pp.hpp>

#include "libiop/algebra/fft.hpp"
#include "libiop/algebra/field_subset/field_subset.hpp"
#include "libiop/algebra/field_subset/subspace.hpp"
#include "libiop/algebra/field_subset/subgroup.hpp"
#include "libiop/protocols/ldt/fri/fri_aux.hpp"
#include "libiop/bcs/merkle_tree.hpp"
#include "libiop/bcs/hashing/blake2b.hpp"
#include "libiop/bcs/hashing/hash_enum.hpp"
#include "libiop/bcs/pow.hpp"
#include "libiop/snark/aurora_snark.hpp"
#include "libiop/snark/fractal_snark.hpp"
#include "libiop/snark/ligero_snark.hpp"
#include "libiop/relations/examples/r1cs_examples.hpp"

#define REFN(name) ref_##name

using namespace libiop;
typedef libff::gf192 G;
typedef libff::edwards_Fr Fr;

extern "C" void b200_shim_randombytes_seed(uint64_t seed);

static_assert(sizeof(G) == 24, "gf192 must be 24 bytes");
static_assert(sizeof(Fr) == 24, "edwards_Fr must be 24 bytes");

template<typename F> static std::vector<F> vec_from(const uint64_t *p, size_t n)
{
    std::vector<F> v(n);
    std::memcpy((void *)v.data(), p, n * 24);
    return v;
}
template<typename F> static void vec_to(uint64_t *p, const std::vector<F> &v)
{
    std::memcpy(p, (const void *)v.data(), v.size() * 24);
}
template<typename F> static F el_from(const uint64_t *p)
{
    F e; std::memcpy((void *)&e, p, 24); return e;
}

template<typename F, typename D>
class exposed_merkle_tree : public merkle_tree<F, D> {
public:
    using merkle_tree<F, D>::merkle_tree;
    const std::vector<D> &nodes() const { return this->inner_nodes_; }
    const std::vector<zk_salt_type> &salts() const { return this->zk_leaf_randomness_elements_; }
};

extern "C" {

void REFN(init)(void)
{
    static bool done = false;
    if (!done) { libff::edwards_pp::init_public_params(); done = true; }
    const bool prof = std::getenv("B200_REF_PROFILE") != nullptr;   /* libff::enter_block timers */
    libff::inhibit_profiling_info = !prof;
    libff::inhibit_profiling_counters = !prof;
}

void REFN(seed)(uint64_t seed) { b200_shim_randombytes_seed(seed); }

/* ---- field ops ---- */
void REFN(gf192_mul)(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        G r = el_from<G>(a + 3 * i) * el_from<G>(b + 3 * i);
        std::memcpy(out + 3 * i, &r, 24);
    }
}
void REFN(gf192_inv)(uint64_t *out, const uint64_t *a, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        G r = el_from<G>(a + 3 * i).inverse();
        std::memcpy(out + 3 * i, &r, 24);
    }
}
void REFN(fp_mul)(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        Fr r = el_from<Fr>(a + 3 * i) * el_from<Fr>(b + 3 * i);
        std::memcpy(out + 3 * i, &r, 24);
    }
}
void REFN(fp_add)(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        Fr r = el_from<Fr>(a + 3 * i) + el_from<Fr>(b + 3 * i);
        std::memcpy(out + 3 * i, &r, 24);
    }
}
void REFN(fp_sub)(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        Fr r = el_from<Fr>(a + 3 * i) - el_from<Fr>(b + 3 * i);
        std::memcpy(out + 3 * i, &r, 24);
    }
}
void REFN(fp_inv)(uint64_t *out, const uint64_t *a, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        Fr r = el_from<Fr>(a + 3 * i).inverse();
        std::memcpy(out + 3 * i, &r, 24);
    }
}
/* Montgomery limbs of the small integer v, of the multiplicative generator, of the subgroup generator */
void REFN(fp_from_u64)(uint64_t *out, uint64_t v) { Fr r(v); std::memcpy(out, &r, 24); }
void REFN(fp_generator)(uint64_t *out) { Fr r = Fr::multiplicative_generator; std::memcpy(out, &r, 24); }
void REFN(fp_subgroup_generator)(uint64_t *out, uint32_t logn)
{
    multiplicative_subgroup<Fr> s((size_t)1 << logn);
    Fr r = s.generator(); std::memcpy(out, &r, 24);
}
void REFN(fp_params)(uint64_t *mod, uint64_t *r2, uint64_t *inv, uint32_t *s)
{
    std::memcpy(mod, Fr::mod.data, 24);
    std::memcpy(r2, Fr::Rsquared.data, 24);
    *inv = Fr::inv; *s = (uint32_t)Fr::s;
}

/* ---- FFTs ---- */
static field_subset<G> make_subspace(uint32_t m, const uint64_t *basis, const uint64_t *shift)
{
    std::vector<G> b = vec_from<G>(basis, m);
    return field_subset<G>(affine_subspace<G>(b, el_from<G>(shift)));
}

void REFN(additive_fft)(uint64_t *out, const uint64_t *coeffs, size_t n_coeffs,
                      uint32_t m, const uint64_t *basis, const uint64_t *shift)
{
    vec_to(out, FFT_over_field_subset<G>(vec_from<G>(coeffs, n_coeffs), make_subspace(m, basis, shift)));
}
void REFN(additive_ifft)(uint64_t *out, const uint64_t *evals,
                       uint32_t m, const uint64_t *basis, const uint64_t *shift)
{
    vec_to(out, IFFT_over_field_subset<G>(vec_from<G>(evals, (size_t)1 << m), make_subspace(m, basis, shift)));
}
void REFN(naive_fft_gf192)(uint64_t *out, const uint64_t *coeffs, size_t n_coeffs,
                         uint32_t m, const uint64_t *basis, const uint64_t *shift)
{
    vec_to(out, naive_FFT<G>(vec_from<G>(coeffs, n_coeffs), make_subspace(m, basis, shift)));
}
void REFN(mult_fft)(uint64_t *out, const uint64_t *coeffs, size_t n_coeffs, uint32_t logn, const uint64_t *shift)
{
    field_subset<Fr> dom((size_t)1 << logn, el_from<Fr>(shift));
    vec_to(out, FFT_over_field_subset<Fr>(vec_from<Fr>(coeffs, n_coeffs), dom));
}
void REFN(mult_ifft)(uint64_t *out, const uint64_t *evals, uint32_t logn, const uint64_t *shift)
{
    field_subset<Fr> dom((size_t)1 << logn, el_from<Fr>(shift));
    vec_to(out, IFFT_over_field_subset<Fr>(vec_from<Fr>(evals, (size_t)1 << logn), dom));
}

/* ---- FRI fold ---- */
void REFN(fri_fold_additive)(uint64_t *out, const uint64_t *f, uint32_t m, const uint64_t *basis,
                           const uint64_t *shift, uint32_t eta, const uint64_t *x)
{
    auto fi = std::make_shared<std::vector<G>>(vec_from<G>(f, (size_t)1 << m));
    auto r = evaluate_next_f_i_over_entire_domain<G>(fi, make_subspace(m, basis, shift),
                                                     (size_t)1 << eta, el_from<G>(x));
    vec_to(out, *r);
}
void REFN(fri_fold_mult)(uint64_t *out, const uint64_t *f, uint32_t logn,
                       const uint64_t *shift, uint32_t eta, const uint64_t *x)
{
    auto fi = std::make_shared<std::vector<Fr>>(vec_from<Fr>(f, (size_t)1 << logn));
    field_subset<Fr> dom((size_t)1 << logn, el_from<Fr>(shift));
    auto r = evaluate_next_f_i_over_entire_domain<Fr>(fi, dom, (size_t)1 << eta, el_from<Fr>(x));
    vec_to(out, *r);
}

} /* extern C */
/* ---- Merkle ---- */
template<typename F>
static void merkle_build(uint8_t *nodes_out, uint8_t *salts_out, const uint64_t *const *oracles,
                         size_t n_oracles, size_t n_elems, size_t coset, int zk, size_t security)
{
    const size_t leaves = n_elems / coset;
    const size_t dlen = (2 * security + 7) / 8;
    exposed_merkle_tree<F, binary_hash_digest> mt(
        leaves, std::make_shared<blake2b_leafhash<F>>(security), blake2b_two_to_one_hash,
        dlen, zk != 0, security);
    std::vector<std::shared_ptr<std::vector<F>>> cols;
    for (size_t k = 0; k < n_oracles; ++k)
        cols.emplace_back(std::make_shared<std::vector<F>>(vec_from<F>(oracles[k], n_elems)));
    mt.construct_with_leaves_serialized_by_cosets(cols, coset);
    const auto &nodes = mt.nodes();
    for (size_t i = 0; i < nodes.size(); ++i) std::memcpy(nodes_out + i * dlen, nodes[i].data(), dlen);
    if (zk && salts_out) {
        const auto &s = mt.salts();
        for (size_t i = 0; i < s.size(); ++i) std::memcpy(salts_out + i * dlen, s[i].data(), dlen);
    }
}
extern "C" {
void REFN(merkle_build)(uint8_t *nodes_out, uint8_t *salts_out, const uint64_t *const *oracles,
                      size_t n_oracles, size_t n_elems, size_t coset, int field_kind, int zk, size_t security)
{
    if (field_kind == 2) merkle_build<Fr>(nodes_out, salts_out, oracles, n_oracles, n_elems, coset, zk, security);
    else merkle_build<G>(nodes_out, salts_out, oracles, n_oracles, n_elems, coset, zk, security);
}

/* ---- proof of work (pow.tcc:87-158) ---- */
void REFN(pow_solve)(uint8_t *pow_out, const uint8_t *challenge, size_t work_parameter, size_t cost_per_hash,
                     uint32_t *bitlen_out, uint64_t *upper_out)
{
    const pow_parameters params(work_parameter, cost_per_hash);
    const libiop::pow<G, binary_hash_digest> p(params, 32);
    const binary_hash_digest sol = p.solve_pow(blake2b_two_to_one_hash, binary_hash_digest((const char *)challenge, 32));
    std::memcpy(pow_out, sol.data(), 32);
    if (bitlen_out) *bitlen_out = (uint32_t)params.pow_bitlen();
    if (upper_out) *upper_out = params.pow_upperbound();
}

/* ---- verifier random messages: blake2b_FieldT_randomness_extractor (bcs/hashing/blake2b.tcc:231-256) ---- */
void REFN(squeeze_gf192)(uint64_t *out, const uint8_t *root, size_t digest_len, size_t index, size_t n)
{
    const std::vector<G> v = blake2b_FieldT_randomness_extractor<G>(binary_hash_digest((const char *)root, digest_len), index, n);
    if (n) std::memcpy(out, v.data(), n * 24);
}
void REFN(squeeze_edwards)(uint64_t *out, const uint8_t *root, size_t digest_len, size_t index, size_t n)
{
    REFN(init)();
    const std::vector<libff::edwards_Fr> v =
        blake2b_FieldT_randomness_extractor<libff::edwards_Fr>(binary_hash_digest((const char *)root, digest_len), index, n);
    for (size_t i = 0; i < n; ++i) std::memcpy(out + 3 * i, v[i].mont_repr.data, 24);
}

/* ---- hashchain ---- */
void REFN(hashchain_demo)(uint64_t *squeezed /*2 elems*/, uint64_t *positions /*3*/, const uint8_t *absorbed32)
{
    blake2b_hashchain<G, binary_hash_digest> hc(128);
    hc.absorb(binary_hash_digest((const char *)absorbed32, 32));
    vec_to(squeezed, hc.squeeze(2));
    std::vector<size_t> q = hc.squeeze_query_positions(3, (size_t)1 << 20);
    for (int i = 0; i < 3; ++i) positions[i] = q[i];
}

} /* extern C */
/* ---- whole Aurora prover: canonical transcript dump ---- */
template<typename F>
static void dump_transcript(std::vector<uint8_t> &o, const bcs_transformation_transcript<F, binary_hash_digest> &t)
{
    auto u64 = [&](uint64_t v) { const uint8_t *p = (const uint8_t *)&v; o.insert(o.end(), p, p + 8); };
    auto bytes = [&](const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; o.insert(o.end(), q, q + n); };
    u64(t.prover_messages_.size());
    for (auto &m : t.prover_messages_) { u64(m.size()); if (!m.empty()) bytes(m.data(), m.size() * 24); }
    u64(t.MT_roots_.size());
    for (auto &r : t.MT_roots_) { u64(r.size()); bytes(r.data(), r.size()); }
    u64(t.query_positions_.size());
    for (auto &q : t.query_positions_) { u64(q.size()); for (size_t p : q) u64(p); }
    u64(t.query_responses_.size());
    for (auto &qr : t.query_responses_) {
        u64(qr.size());
        for (auto &v : qr) { u64(v.size()); if (!v.empty()) bytes(v.data(), v.size() * 24); }
    }
    u64(t.MT_leaf_positions_.size());
    for (auto &q : t.MT_leaf_positions_) { u64(q.size()); for (size_t p : q) u64(p); }
    u64(t.MT_set_membership_proofs_.size());
    for (auto &pf : t.MT_set_membership_proofs_) {
        u64(pf.auxiliary_hashes.size());
        for (auto &h : pf.auxiliary_hashes) { u64(h.size()); bytes(h.data(), h.size()); }
        u64(pf.randomness_hashes.size());
        for (auto &h : pf.randomness_hashes) { u64(h.size()); bytes(h.data(), h.size()); }
    }
    u64(t.proof_of_work_.size()); bytes(t.proof_of_work_.data(), t.proof_of_work_.size());
}

extern "C" {
/* Runs generate_r1cs_example<gf192>(2^log_n, k, 2^log_n - 1) with the deterministic RNG seeded by
   `seed`, then aurora_snark_prover (+verifier if verify != 0).  Returns prover seconds; writes the
   canonical dump into buf (up to cap bytes) and its length into *len; *accepted = verifier bit. */
double REFN(aurora_gf192)(uint32_t log_n, uint32_t k_inputs, int make_zk, uint32_t rs_extra, uint32_t fri_loc,
                        uint64_t seed, int verify, uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    REFN(init)();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = generate_r1cs_example<G>(n, k_inputs, n - 1);
    aurora_snark_parameters<G, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic,
        blake2b_type, (size_t)fri_loc, (size_t)rs_extra, make_zk != 0, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    const auto t0 = std::chrono::steady_clock::now();
    const aurora_snark_argument<G, binary_hash_digest> proof = aurora_snark_prover<G, binary_hash_digest>(
        ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<uint8_t> o;
    dump_transcript(o, proof);
    if (len) *len = o.size();
    if (buf && o.size() <= cap) std::memcpy(buf, o.data(), o.size());
    if (accepted) {
        *accepted = -1;
        if (verify)
            *accepted = aurora_snark_verifier<G, binary_hash_digest>(
                ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

/* the same with the number of variables chosen freely: generate_r1cs_example<gf192>(2^log_n, k, num_variables) — variable
   counts that are not a power of two minus one, variable domains larger or smaller than the constraint domain */
double REFN(aurora_gf192_shape)(uint32_t log_n, uint32_t k_inputs, uint64_t num_variables, int make_zk, uint32_t rs_extra, uint32_t fri_loc,
                              uint64_t seed, int verify, uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    REFN(init)();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = generate_r1cs_example<G>(n, k_inputs, (size_t)num_variables);
    aurora_snark_parameters<G, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic,
        blake2b_type, (size_t)fri_loc, (size_t)rs_extra, make_zk != 0, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    const auto t0 = std::chrono::steady_clock::now();
    const aurora_snark_argument<G, binary_hash_digest> proof = aurora_snark_prover<G, binary_hash_digest>(
        ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<uint8_t> o;
    dump_transcript(o, proof);
    if (len) *len = o.size();
    if (buf && o.size() <= cap) std::memcpy(buf, o.data(), o.size());
    if (accepted) {
        *accepted = -1;
        if (verify)
            *accepted = aurora_snark_verifier<G, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

/* the same prover call `repeats` times on one generated instance (throughput measurements, bench/e2e_throughput.py) */
void REFN(aurora_gf192_repeat)(uint32_t log_n, uint32_t k_inputs, int make_zk, uint32_t rs_extra, uint32_t fri_loc,
                               uint64_t seed, uint32_t repeats, double *secs_out)
{
    REFN(init)();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = generate_r1cs_example<G>(n, k_inputs, n - 1);
    aurora_snark_parameters<G, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic,
        blake2b_type, (size_t)fri_loc, (size_t)rs_extra, make_zk != 0, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    for (uint32_t r = 0; r < repeats; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        const aurora_snark_argument<G, binary_hash_digest> proof = aurora_snark_prover<G, binary_hash_digest>(
            ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
        secs_out[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
}

} /* extern "C" */

/* ---- deterministic R1CS instances for both field kinds ----------------------------------------
   Same shape as generate_r1cs_example (libiop/relations/examples/r1cs_examples.tcc:23-78): constraint i is
   z[i mod m] * z[(i+7) mod m] = c_i * z[(2i+1) mod m]; the assignment comes from a seeded xorshift64 instead of
   FieldT::random_element() because Fp_model draws from std::random_device (bigint.tcc:194-206), which cannot be
   seeded — the CPU and the GPU-integrated builds must see the same instance. */
static uint64_t xs64(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static void det_elem(G &e, uint64_t &s) { e = G(xs64(s), xs64(s), xs64(s)); }
static void det_elem(Fr &e, uint64_t &s)
{
    e = Fr((long)(xs64(s) >> 1), true) * Fr((long)(xs64(s) >> 1), true) + Fr((long)(xs64(s) >> 1), true);
    e = e * e * Fr((long)(xs64(s) >> 1), true) + e;
}
template<typename F>
static r1cs_example<F> det_example(size_t num_constraints, size_t num_inputs, size_t num_variables, uint64_t seed)
{
    uint64_t st = seed ? seed : 0x9E3779B97F4A7C15ull;
    r1cs_constraint_system<F> cs;
    cs.primary_input_size_ = num_inputs;
    cs.auxiliary_input_size_ = num_variables - num_inputs;
    std::vector<F> z(num_variables);
    for (auto &e : z) det_elem(e, st);
    for (size_t i = 0; i < num_constraints; ++i) {
        linear_combination<F> A, B, C;
        const size_t a = i % num_variables, b = (i + 7) % num_variables, c = (2 * i + 1) % num_variables;
        A.add_term(a + 1, F::one());
        B.add_term(b + 1, F::one());
        const F ab = z[a] * z[b];
        if (z[c].is_zero()) C.add_term(0, ab); else C.add_term(c + 1, ab * z[c].inverse());
        cs.add_constraint(r1cs_constraint<F>(A, B, C));
    }
    r1cs_primary_input<F> primary(z.begin(), z.begin() + num_inputs);
    r1cs_auxiliary_input<F> aux(z.begin() + num_inputs, z.end());
    return r1cs_example<F>(std::move(cs), std::move(primary), std::move(aux));
}

template<typename T>
static void finish_dump(const T &proof, uint8_t *buf, size_t cap, size_t *len)
{
    std::vector<uint8_t> o;
    dump_transcript(o, proof);
    if (len) *len = o.size();
    if (buf && o.size() <= cap) std::memcpy(buf, o.data(), o.size());
}

extern "C" {

/* Aurora over the multiplicative edwards_Fr domains (README.md:160-164 `--is_multiplicative 1 --field_size=181`), non-zk */
double REFN(aurora_edwards)(uint32_t log_n, uint32_t k_inputs, uint32_t rs_extra, uint32_t fri_loc, uint64_t seed,
                            int verify, uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    REFN(init)();
    const size_t n = (size_t)1 << log_n;
    r1cs_example<Fr> ex = det_example<Fr>(n, k_inputs, n - 1, seed);
    aurora_snark_parameters<Fr, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic, blake2b_type,
        (size_t)fri_loc, (size_t)rs_extra, false, multiplicative_coset_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    const auto t0 = std::chrono::steady_clock::now();
    const auto proof = aurora_snark_prover<Fr, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    finish_dump(proof, buf, cap, len);
    if (accepted) *accepted = verify ? (aurora_snark_verifier<Fr, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return std::chrono::duration<double>(t1 - t0).count();
}

/* Fractal over edwards_Fr (BASELINE config 3 shape; instrument_fractal_snark.cpp:86-180: rho = 3, localization 2, k = 0) */
double REFN(fractal_edwards)(uint32_t log_n, uint32_t rs_extra, uint32_t fri_loc, uint64_t seed, int verify,
                             uint8_t *buf, size_t cap, size_t *len, int *accepted, double *indexer_secs)
{
    REFN(init)();
    const size_t n = (size_t)1 << log_n;
    r1cs_example<Fr> ex = det_example<Fr>(n, 0, n - 1, seed);
    fractal_snark_parameters<Fr, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic, blake2b_type,
        (size_t)fri_loc, (size_t)rs_extra, false, multiplicative_coset_type,
        std::make_shared<r1cs_constraint_system<Fr>>(ex.constraint_system_));
    const auto ti0 = std::chrono::steady_clock::now();
    std::pair<bcs_prover_index<Fr, binary_hash_digest>, bcs_verifier_index<Fr, binary_hash_digest>> index =
        fractal_snark_indexer(params);
    const auto ti1 = std::chrono::steady_clock::now();
    if (indexer_secs) *indexer_secs = std::chrono::duration<double>(ti1 - ti0).count();
    const auto t0 = std::chrono::steady_clock::now();
    const auto proof = fractal_snark_prover<Fr, binary_hash_digest>(index.first, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    finish_dump(proof, buf, cap, len);
    /* the index roots are part of what the verifier trusts: append them so parity covers the indexer too */
    if (buf && len) {
        size_t off = *len;
        for (auto &r : index.second.index_MT_roots_) { if (off + r.size() <= cap) std::memcpy(buf + off, r.data(), r.size()); off += r.size(); }
        *len = off;
    }
    if (accepted) *accepted = verify ? (fractal_snark_verifier<Fr, binary_hash_digest>(index.second, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return std::chrono::duration<double>(t1 - t0).count();
}

/* Ligero over GF(2^192) (BASELINE config 4 shape; instrument_ligero_snark.cpp:66-117: ratio 0.1, rho = 2) */
double REFN(ligero_gf192)(uint32_t log_n, double height_width_ratio, uint32_t rs_extra, int make_zk, uint64_t seed, int verify,
                          uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    REFN(init)();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = det_example<G>(n, 15, n - 1, seed);
    ligero_snark_parameters<G, binary_hash_digest> params;
    params.security_level_ = 128;
    params.LDT_reducer_soundness_type_ = LDT_reducer_soundness_type::optimistic_heuristic;
    params.height_width_ratio_ = (float)height_width_ratio;
    params.RS_extra_dimensions_ = rs_extra;
    params.make_zk_ = make_zk != 0;
    params.domain_type_ = affine_subspace_type;
    params.bcs_params_ = default_bcs_params<G, binary_hash_digest>(blake2b_type, 128, log_n);
    const auto t0 = std::chrono::steady_clock::now();
    const auto proof = ligero_snark_prover<G, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    finish_dump(proof, buf, cap, len);
    if (accepted) *accepted = verify ? (ligero_snark_verifier<G, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return std::chrono::duration<double>(t1 - t0).count();
}

} /* extern "C" */
