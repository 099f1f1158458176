/*
 * prover_api.cpp — the reference's own SNARK entry points with the B200 library dropped in (libiop_b200_dropin.so).
 *
 * This translation unit includes the UNMODIFIED libiop headers from the reference checkout (B200_REF_ROOT, never
 * copied into this repository) with libiop_b200/dropin/ first on the include path, so that the shadow .tcc files there
 * specialise the hot-path templates for libff::gf192 / libff::edwards_Fr.  Host code — the IOP layer, the protocols, the
 * BCS transformation, the hash chain, the verifier — is the reference's; the data-parallel work runs on the device:
 *     aurora_snark_prover / aurora_snark_verifier      libiop/snark/aurora_snark.hpp:75-92
 *     fractal_snark_indexer / prover / verifier        libiop/snark/fractal_snark.hpp
 *     ligero_snark_prover / verifier                   libiop/snark/ligero_snark.hpp
 * The extern "C" functions below run them on generated R1CS instances and return the canonical transcript bytes
 * (libiop_b200/host/b200_transcript.hpp); libiop_b200/prover.py binds them for tests and bench.py.
 * A C++ application links the same way and calls the reference API directly.
 */
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include <libff/algebra/fields/binary/gf192.hpp>
#include <libff/algebra/curves/edwards/edwards_pp.hpp>

#include "libiop/snark/aurora_snark.hpp"
#include "libiop/snark/fractal_snark.hpp"
#include "libiop/snark/ligero_snark.hpp"
#include "libiop/relations/examples/r1cs_examples.hpp"

#include "libiop_b200/host/b200_device.hpp"
#include "libiop_b200/host/b200_residency.hpp"
#include "libiop_b200/host/b200_transcript.hpp"

/* the process's large blocks (the prover's oracles) come from page-locked memory: DMA in place, no page faults */
B200_DEFINE_PINNED_OPERATOR_NEW

using namespace libiop;
typedef libff::gf192 G;
typedef libff::edwards_Fr Fr;

extern "C" void b200_shim_randombytes_seed(uint64_t seed);
extern "C" void b200_shim_randombytes_secure(void);

namespace {

/* The instance is always generated from the seeded stream (reproducible benchmarks and parity runs).  The PROOF's own
   randomness — zero-knowledge masks and Merkle salts, all drawn through randombytes_buf — comes from the unpredictable
   generator unless the parity harness asks for the seeded stream with B200_DETERMINISTIC_RNG=1 (tests/conftest.py sets it
   so that zk transcripts can be compared bit for bit with the CPU reference's). */
void arm_proof_randomness(bool make_zk)
{
    const char *e = std::getenv("B200_DETERMINISTIC_RNG");
    if (make_zk && !(e && e[0] == '1')) b200_shim_randombytes_secure();
}

void init_once()
{
    static bool done = false;
    if (!done) { libff::edwards_pp::init_public_params(); done = true; }
    const bool prof = std::getenv("B200_REF_PROFILE") != nullptr;     /* libff::enter_block timers of the reference */
    libff::inhibit_profiling_info = !prof;
    libff::inhibit_profiling_counters = !prof;
}

/* Deterministic instances of the shape of generate_r1cs_example (relations/examples/r1cs_examples.tcc:23-78):
   constraint i is z[i mod m] * z[(i+7) mod m] = c_i * z[(2i+1) mod m].  Fp_model::random_element draws from
   std::random_device (bigint.tcc:194-206) and cannot be seeded, so the assignment comes from a seeded xorshift64;
   the parity harness builds the CPU reference's instance with the same generator. */
uint64_t xorshift(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
void next_element(G &e, uint64_t &s) { e = G(xorshift(s), xorshift(s), xorshift(s)); }
void next_element(Fr &e, uint64_t &s)
{
    e = Fr((long)(xorshift(s) >> 1), true) * Fr((long)(xorshift(s) >> 1), true) + Fr((long)(xorshift(s) >> 1), true);
    e = e * e * Fr((long)(xorshift(s) >> 1), true) + e;
}
template<typename F>
r1cs_example<F> seeded_example(size_t num_constraints, size_t num_inputs, size_t num_variables, uint64_t seed)
{
    uint64_t st = seed ? seed : 0x9E3779B97F4A7C15ull;
    r1cs_constraint_system<F> cs;
    cs.primary_input_size_ = num_inputs;
    cs.auxiliary_input_size_ = num_variables - num_inputs;
    std::vector<F> z(num_variables);
    for (auto &e : z) next_element(e, st);
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

template<typename ProofT>
void emit(const ProofT &proof, uint8_t *buf, size_t cap, size_t *len)
{
    const std::vector<uint8_t> bytes = libiop_b200::canonical_transcript_bytes(proof);
    if (len) *len = bytes.size();
    if (buf && bytes.size() <= cap) std::memcpy(buf, bytes.data(), bytes.size());
}

double seconds(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b)
{
    return std::chrono::duration<double>(b - a).count();
}

} // namespace

extern "C" {

void b200_prover_seed(uint64_t seed) { b200_shim_randombytes_seed(seed); }

/* Aurora over GF(2^192), affine-subspace domains (BASELINE configs 1, 2, 5): generate_r1cs_example<gf192>(2^log_n, k, 2^log_n - 1)
   under the seeded randombytes stream, security 128, optimistic-heuristic LDT reducer, heuristic FRI, BLAKE2b
   (SURVEY.md §8d).  Returns prover seconds; *accepted = reference verifier verdict (or -1 if verify == 0). */
double b200_prover_aurora_gf192(uint32_t log_n, uint32_t k_inputs, int make_zk, uint32_t rs_extra, uint32_t fri_loc,
                                uint64_t seed, int verify, uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    init_once();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = generate_r1cs_example<G>(n, k_inputs, n - 1);
    aurora_snark_parameters<G, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic,
        blake2b_type, (size_t)fri_loc, (size_t)rs_extra, make_zk != 0, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    arm_proof_randomness(make_zk != 0);
    const auto t0 = std::chrono::steady_clock::now();
    const aurora_snark_argument<G, binary_hash_digest> proof = aurora_snark_prover<G, binary_hash_digest>(
        ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    emit(proof, buf, cap, len);
    if (accepted)
        *accepted = verify ? (aurora_snark_verifier<G, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return seconds(t0, t1);
}

/* the same with the number of variables chosen freely (generate_r1cs_example<gf192>(2^log_n, k, num_variables)): variable
   counts that are not a power of two minus one, variable domains larger or smaller than the constraint domain */
double b200_prover_aurora_gf192_shape(uint32_t log_n, uint32_t k_inputs, uint64_t num_variables, int make_zk, uint32_t rs_extra,
                                      uint32_t fri_loc, uint64_t seed, int verify, uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    init_once();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = generate_r1cs_example<G>(n, k_inputs, (size_t)num_variables);
    aurora_snark_parameters<G, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic,
        blake2b_type, (size_t)fri_loc, (size_t)rs_extra, make_zk != 0, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    arm_proof_randomness(make_zk != 0);
    const auto t0 = std::chrono::steady_clock::now();
    const aurora_snark_argument<G, binary_hash_digest> proof = aurora_snark_prover<G, binary_hash_digest>(
        ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    emit(proof, buf, cap, len);
    if (accepted)
        *accepted = verify ? (aurora_snark_verifier<G, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return seconds(t0, t1);
}

/* The same prover call `repeats` times on ONE generated instance (instance generation costs more than the proof at
   N = 2^20); secs_out[repeats].  Used for throughput measurements with several prover processes per GPU. */
void b200_prover_aurora_gf192_repeat(uint32_t log_n, uint32_t k_inputs, int make_zk, uint32_t rs_extra, uint32_t fri_loc,
                                     uint64_t seed, uint32_t repeats, double *secs_out)
{
    init_once();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = generate_r1cs_example<G>(n, k_inputs, n - 1);
    aurora_snark_parameters<G, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic,
        blake2b_type, (size_t)fri_loc, (size_t)rs_extra, make_zk != 0, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    arm_proof_randomness(make_zk != 0);
    for (uint32_t r = 0; r < repeats; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        const aurora_snark_argument<G, binary_hash_digest> proof = aurora_snark_prover<G, binary_hash_digest>(
            ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
        secs_out[r] = seconds(t0, std::chrono::steady_clock::now());
    }
}

/* The steps of aurora_snark_prover (snark/aurora_snark.tcc:120-146) timed one by one on one generated instance:
   secs_out[0] protocol construction + registrations, [1] produce_proof, [2] get_transcript, [3] destruction of the
   protocol objects and the prover's oracles, [4] destruction of the argument.  A diagnostic for host-side work that the
   reference's own enter_block timers do not cover (bench/dropin_phases.py); the product path is the call above. */
void b200_prover_aurora_gf192_phases(uint32_t log_n, uint32_t k_inputs, int make_zk, uint32_t rs_extra, uint32_t fri_loc,
                                     uint64_t seed, uint32_t repeats, double *secs_out)
{
    init_once();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = generate_r1cs_example<G>(n, k_inputs, n - 1);
    aurora_snark_parameters<G, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic,
        blake2b_type, (size_t)fri_loc, (size_t)rs_extra, make_zk != 0, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    arm_proof_randomness(make_zk != 0);
    for (uint32_t r = 0; r < repeats; ++r) {
        double *out = secs_out + 5 * r;
        auto t0 = std::chrono::steady_clock::now();
        {
            std::unique_ptr<aurora_snark_argument<G, binary_hash_digest>> transcript;
            {
                bcs_prover<G, binary_hash_digest> IOP(params.bcs_params_);
                aurora_iop<G> full_protocol(IOP, ex.constraint_system_, params.iop_params_);
                full_protocol.register_interactions();
                IOP.seal_interaction_registrations();
                full_protocol.register_queries();
                IOP.seal_query_registrations();
                auto t1 = std::chrono::steady_clock::now(); out[0] = seconds(t0, t1); t0 = t1;
                full_protocol.produce_proof(ex.primary_input_, ex.auxiliary_input_);
                t1 = std::chrono::steady_clock::now(); out[1] = seconds(t0, t1); t0 = t1;
                transcript.reset(new aurora_snark_argument<G, binary_hash_digest>(IOP.get_transcript()));
                t1 = std::chrono::steady_clock::now(); out[2] = seconds(t0, t1); t0 = t1;
            }
            auto t1 = std::chrono::steady_clock::now(); out[3] = seconds(t0, t1); t0 = t1;
        }
        out[4] = seconds(t0, std::chrono::steady_clock::now());
    }
}

/* Aurora over edwards_Fr, multiplicative cosets (README.md:160-164 `--is_multiplicative 1 --field_size=181`), non-zk */
double b200_prover_aurora_edwards(uint32_t log_n, uint32_t k_inputs, uint32_t rs_extra, uint32_t fri_loc, uint64_t seed,
                                  int verify, uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    init_once();
    const size_t n = (size_t)1 << log_n;
    r1cs_example<Fr> ex = seeded_example<Fr>(n, k_inputs, n - 1, seed);
    aurora_snark_parameters<Fr, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic, blake2b_type,
        (size_t)fri_loc, (size_t)rs_extra, false, multiplicative_coset_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
    const auto t0 = std::chrono::steady_clock::now();
    const auto proof = aurora_snark_prover<Fr, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    emit(proof, buf, cap, len);
    if (accepted)
        *accepted = verify ? (aurora_snark_verifier<Fr, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return seconds(t0, t1);
}

/* Fractal over edwards_Fr (BASELINE config 3; instrument_fractal_snark.cpp:86-180: rho = 3, localization 2, no inputs).
   The index Merkle roots are appended to the transcript bytes so that parity covers the indexer as well. */
double b200_prover_fractal_edwards(uint32_t log_n, uint32_t rs_extra, uint32_t fri_loc, uint64_t seed, int verify,
                                   uint8_t *buf, size_t cap, size_t *len, int *accepted, double *indexer_secs)
{
    init_once();
    const size_t n = (size_t)1 << log_n;
    r1cs_example<Fr> ex = seeded_example<Fr>(n, 0, n - 1, seed);
    fractal_snark_parameters<Fr, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic, blake2b_type,
        (size_t)fri_loc, (size_t)rs_extra, false, multiplicative_coset_type,
        std::make_shared<r1cs_constraint_system<Fr>>(ex.constraint_system_));
    const auto i0 = std::chrono::steady_clock::now();
    std::pair<bcs_prover_index<Fr, binary_hash_digest>, bcs_verifier_index<Fr, binary_hash_digest>> index = fractal_snark_indexer(params);
    const auto i1 = std::chrono::steady_clock::now();
    if (indexer_secs) *indexer_secs = seconds(i0, i1);
    const auto t0 = std::chrono::steady_clock::now();
    const auto proof = fractal_snark_prover<Fr, binary_hash_digest>(index.first, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    emit(proof, buf, cap, len);
    if (buf && len) {
        size_t off = *len;
        for (auto &r : index.second.index_MT_roots_) { if (off + r.size() <= cap) std::memcpy(buf + off, r.data(), r.size()); off += r.size(); }
        *len = off;
    }
    if (accepted)
        *accepted = verify ? (fractal_snark_verifier<Fr, binary_hash_digest>(index.second, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return seconds(t0, t1);
}

/* Ligero over GF(2^192) (BASELINE config 4; instrument_ligero_snark.cpp:66-117: height/width ratio 0.1, rho = 2) */
double b200_prover_ligero_gf192(uint32_t log_n, double height_width_ratio, uint32_t rs_extra, int make_zk, uint64_t seed, int verify,
                                uint8_t *buf, size_t cap, size_t *len, int *accepted)
{
    init_once();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = seeded_example<G>(n, 15, n - 1, seed);
    ligero_snark_parameters<G, binary_hash_digest> params;
    params.security_level_ = 128;
    params.LDT_reducer_soundness_type_ = LDT_reducer_soundness_type::optimistic_heuristic;
    params.height_width_ratio_ = (float)height_width_ratio;
    params.RS_extra_dimensions_ = rs_extra;
    params.make_zk_ = make_zk != 0;
    params.domain_type_ = affine_subspace_type;
    params.bcs_params_ = default_bcs_params<G, binary_hash_digest>(blake2b_type, 128, log_n);
    arm_proof_randomness(make_zk != 0);
    const auto t0 = std::chrono::steady_clock::now();
    const auto proof = ligero_snark_prover<G, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
    const auto t1 = std::chrono::steady_clock::now();
    emit(proof, buf, cap, len);
    if (accepted)
        *accepted = verify ? (ligero_snark_verifier<G, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params) ? 1 : 0) : -1;
    return seconds(t0, t1);
}

/* The steps of ligero_snark_prover (snark/ligero_snark.tcc:40-72) timed one by one on one generated instance, like
   b200_prover_aurora_gf192_phases: secs_out[0] ligero_iop construction, [1] register_interactions, [2] register_queries (+ seals),
   [3] produce_proof, [4] get_transcript, [5] destruction of the protocol objects and oracles.  Diagnostic only. */
void b200_prover_ligero_gf192_phases(uint32_t log_n, double height_width_ratio, uint32_t rs_extra, uint64_t seed, uint32_t repeats,
                                     double *secs_out)
{
    init_once();
    b200_shim_randombytes_seed(seed);
    const size_t n = (size_t)1 << log_n;
    r1cs_example<G> ex = seeded_example<G>(n, 15, n - 1, seed);
    ligero_snark_parameters<G, binary_hash_digest> params;
    params.security_level_ = 128;
    params.LDT_reducer_soundness_type_ = LDT_reducer_soundness_type::optimistic_heuristic;
    params.height_width_ratio_ = (float)height_width_ratio;
    params.RS_extra_dimensions_ = rs_extra;
    params.make_zk_ = false;
    params.domain_type_ = affine_subspace_type;
    params.bcs_params_ = default_bcs_params<G, binary_hash_digest>(blake2b_type, 128, log_n);
    for (uint32_t r = 0; r < repeats; ++r) {
        double *out = secs_out + 6 * r;
        auto t0 = std::chrono::steady_clock::now();
        const auto lap = [&t0]() { const auto t1 = std::chrono::steady_clock::now(); const double s = seconds(t0, t1); t0 = t1; return s; };
        {
            const ligero_iop_parameters<G> iop_params = obtain_iop_parameters_from_ligero_snark_params<G>(
                params, ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());
            bcs_prover<G, binary_hash_digest> IOP(params.bcs_params_);
            ligero_iop<G> full_protocol(IOP, ex.constraint_system_, iop_params);
            out[0] = lap();
            full_protocol.register_interactions();
            IOP.seal_interaction_registrations();
            out[1] = lap();
            full_protocol.register_queries();
            IOP.seal_query_registrations();
            out[2] = lap();
            full_protocol.produce_proof(ex.primary_input_, ex.auxiliary_input_);
            out[3] = lap();
            const ligero_snark_argument<G, binary_hash_digest> transcript = IOP.get_transcript();
            out[4] = lap();
        }
        out[5] = lap();
    }
}

/* device-side counters of the adapter's context: kernels launched, bytes uploaded / downloaded so far */
void b200_prover_device_stats(uint64_t *launches, uint64_t *h2d_bytes, uint64_t *d2h_bytes)
{
    unsigned long long h = 0, d = 0;
    b200_transfer_bytes(libiop_b200::context(), &h, &d);
    *launches = b200_launch_count(libiop_b200::context());
    *h2d_bytes = h; *d2h_bytes = d;
}

/* host-runtime counters: pinned arena and residency cache */
void b200_prover_arena_stats(uint64_t *slabs, uint64_t *reserved_bytes, uint64_t *free_bytes)
{
    libiop_b200::pinned_arena *a = libiop_b200::pinned_new_detail::arena();
    *slabs = a ? a->slab_count() : 0;
    *reserved_bytes = a ? a->reserved_bytes() : 0;
    *free_bytes = a ? a->free_bytes() : 0;
}
/* stale = lazy vectors the host touched (each one a whole-vector download), see b200_prover_lazy_stats */
void b200_prover_residency_stats(uint64_t *hits, uint64_t *misses, uint64_t *stale, uint64_t *resident_bytes)
{
    libiop_b200::residency &r = libiop_b200::residency::instance();
    *hits = r.hits(); *misses = r.misses(); *stale = r.faults(); *resident_bytes = r.resident_bytes();
}

/* device-resident oracles: host accesses that forced a whole-vector download (count, bytes), gathers of the query phase */
void b200_prover_lazy_stats(uint64_t *faults, uint64_t *fault_bytes, uint64_t *gathers)
{
    libiop_b200::residency &r = libiop_b200::residency::instance();
    *faults = r.faults(); *fault_bytes = r.fault_bytes(); *gathers = r.gathers();
}

} /* extern "C" */
