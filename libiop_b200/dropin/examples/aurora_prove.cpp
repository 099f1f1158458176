/*
 * aurora_prove.cpp — a plain C++ program on the reference API with the B200 library dropped in: what a libiop user's
 * binary looks like after the integration (cf. the reference's libiop/profiling/instrument_aurora_snark.cpp:86-200,
 * whose option parsing needs Boost, absent from this image).  Nothing here is B200-specific except the one macro
 * below; the build line (-I libiop_b200/dropin first) does the rest.
 *
 *   aurora_prove [log_n = 16] [make_zk = 0] [repeats = 3]
 * prints the prover and verifier times of each repetition and a checksum of the transcript bytes.
 */
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include <libff/algebra/fields/binary/gf192.hpp>
#include <libff/algebra/curves/edwards/edwards_pp.hpp>

#include "libiop/snark/aurora_snark.hpp"
#include "libiop/relations/examples/r1cs_examples.hpp"

#include "libiop_b200/host/b200_device.hpp"
#include "libiop_b200/host/b200_transcript.hpp"

B200_DEFINE_PINNED_OPERATOR_NEW          /* the process's large blocks (the prover's oracles) are page-locked */

int main(int argc, char **argv)
{
    using namespace libiop;
    typedef libff::gf192 FieldT;
    const size_t log_n = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 16;
    const bool make_zk = argc > 2 && std::atoi(argv[2]) != 0;
    const int repeats = argc > 3 ? std::atoi(argv[3]) : 3;

    libff::inhibit_profiling_info = true;
    libff::inhibit_profiling_counters = true;
    const size_t n = size_t(1) << log_n;
    r1cs_example<FieldT> ex = generate_r1cs_example<FieldT>(n, 31, n - 1);
    aurora_snark_parameters<FieldT, binary_hash_digest> params(
        128, LDT_reducer_soundness_type::optimistic_heuristic, FRI_soundness_type::heuristic, blake2b_type,
        2 /* FRI localization */, 3 /* RS extra dimensions */, make_zk, affine_subspace_type,
        ex.constraint_system_.num_constraints(), ex.constraint_system_.num_variables());

    for (int r = 0; r < repeats; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        const aurora_snark_argument<FieldT, binary_hash_digest> proof =
            aurora_snark_prover<FieldT, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, ex.auxiliary_input_, params);
        const auto t1 = std::chrono::steady_clock::now();
        const bool ok = aurora_snark_verifier<FieldT, binary_hash_digest>(ex.constraint_system_, ex.primary_input_, proof, params);
        const auto t2 = std::chrono::steady_clock::now();
        const std::vector<uint8_t> bytes = libiop_b200::canonical_transcript_bytes(proof);
        uint64_t h = 1469598103934665603ull;                       /* FNV-1a over the canonical transcript bytes */
        for (uint8_t b : bytes) { h ^= b; h *= 1099511628211ull; }
        std::fprintf(stderr, "aurora_prove N=2^%zu zk=%d run %d: prover %.3f s, verifier %.3f s (%s), transcript %zu bytes, fnv1a %016llx\n",
                     log_n, (int)make_zk, r, std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(),
                     ok ? "accept" : "REJECT", bytes.size(), (unsigned long long)h);
        if (!ok) return 1;
    }
    return 0;
}
