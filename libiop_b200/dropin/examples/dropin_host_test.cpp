/*
 * dropin_host_test.cpp — CPU-only checks of the host logic inside the drop-in shadows (no device call is made):
 *   1. sample_salts_flat draws the zk salts with the same randombytes_buf call pattern as the reference's
 *      merkle_tree::sample_leaf_randomness (merkle_tree.tcc:36-72): same bytes, leaf by leaf, for batch-sized and small leaf counts;
 *   2. membership_proof_nodes lists exactly the sibling nodes, in the order, that the reference's
 *      get_set_membership_proof (merkle_tree.tcc:242-336) emits, checked against the reference on a CPU-built tree.
 */
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <vector>

#include <libff/algebra/fields/binary/gf64.hpp>
#include <libff/algebra/fields/binary/gf192.hpp>

#include "libiop/bcs/merkle_tree.hpp"
#include "libiop/bcs/hashing/blake2b.hpp"

extern "C" void b200_shim_randombytes_seed(uint64_t seed);
extern "C" void b200_shim_randombytes_secure(void);

using namespace libiop;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

/* gf64 is not a field the drop-in serves, so this tree runs the reference's own code paths throughout */
typedef libff::gf64 RefField;
class exposed_tree : public merkle_tree<RefField, binary_hash_digest> {
public:
    using merkle_tree<RefField, binary_hash_digest>::merkle_tree;
    void sample() { this->sample_leaf_randomness(); }
    const std::vector<zk_salt_type> &salts() const { return this->zk_leaf_randomness_elements_; }
    const std::vector<binary_hash_digest> &nodes() const { return this->inner_nodes_; }
};

int main()
{
    libff::inhibit_profiling_info = true;
    libff::inhibit_profiling_counters = true;
    /* 1. salts */
    for (std::size_t leaves : {std::size_t(2), std::size_t(8), std::size_t(1024), std::size_t(1) << 15, std::size_t(1) << 16, std::size_t(1) << 18}) {   /* tree sizes are powers of two >= 2 */
        exposed_tree t(leaves, std::make_shared<blake2b_leafhash<RefField>>(128), blake2b_two_to_one_hash, 32, true, 128);
        b200_shim_randombytes_seed(1234 + leaves);
        t.sample();
        b200_shim_randombytes_seed(1234 + leaves);
        const std::size_t zk_bytes = t.salts()[0].size();
        const std::string flat = b200_glue::sample_salts_flat(leaves, zk_bytes);
        CHECK(t.salts().size() == leaves && flat.size() == leaves * zk_bytes);
        bool same = true;
        for (std::size_t i = 0; i < leaves && same; ++i) same = (flat.compare(i * zk_bytes, zk_bytes, t.salts()[i]) == 0);
        CHECK(same);
        CHECK(b200_glue::salt_at(t.salts(), leaves, leaves / 2) == t.salts()[leaves / 2]);
        CHECK(b200_glue::salt_at({flat}, leaves, leaves / 2) == t.salts()[leaves / 2]);
    }
    /* 1b. a salt length that does not divide the 1 MiB batch (security 100 -> 25 bytes): leafs_per_batch * 25 exceeds the batch;
           every salt byte must still come from randombytes_buf (no read past the batch buffer) */
    {
        b200_shim_randombytes_seed(99);
        const std::size_t leaves = std::size_t(1) << 17, zk_bytes = 25;
        const std::string flat = b200_glue::sample_salts_flat(leaves, zk_bytes);
        CHECK(flat.size() == leaves * zk_bytes);
        std::size_t zeros = 0;
        for (std::size_t i = flat.size() - 65536; i < flat.size(); ++i) zeros += flat[i] == 0;
        CHECK(zeros < 1024);                      /* ~256 expected from a uniform stream */
    }
    /* 1c. without the parity switch randombytes_buf is the unpredictable generator: two draws differ, and differ from the seeded stream */
    {
        unsigned char a[32], b[32], c[32];
        b200_shim_randombytes_secure();
        randombytes_buf(a, 32); randombytes_buf(b, 32);
        b200_shim_randombytes_seed(0);
        randombytes_buf(c, 32);
        CHECK(std::memcmp(a, b, 32) != 0 && std::memcmp(a, c, 32) != 0 && std::memcmp(b, c, 32) != 0);
    }
    /* 2. pruned authentication paths: node list vs the reference's proof on a CPU-built tree */
    std::mt19937_64 rng(7);
    for (std::size_t log_leaves : {1, 2, 5, 10}) {
        const std::size_t leaves = std::size_t(1) << log_leaves;
        exposed_tree t(leaves, std::make_shared<blake2b_leafhash<RefField>>(128), blake2b_two_to_one_hash, 32, false, 128);
        std::vector<RefField> column(leaves);
        for (auto &e : column) e = RefField(rng());
        t.construct({column});
        for (int trial = 0; trial < 20; ++trial) {
            std::vector<std::size_t> positions;
            const std::size_t k = 1 + rng() % (leaves < 12 ? leaves : 12);
            for (std::size_t i = 0; i < k; ++i) positions.push_back(rng() % leaves);
            const merkle_tree_set_membership_proof<binary_hash_digest> ref = t.get_set_membership_proof(positions);
            std::vector<std::size_t> S(positions);
            std::sort(S.begin(), S.end());
            S.erase(std::unique(S.begin(), S.end()), S.end());
            const std::vector<std::size_t> wanted = b200_glue::membership_proof_nodes(leaves, S);
            CHECK(wanted.size() == ref.auxiliary_hashes.size());
            bool same = wanted.size() == ref.auxiliary_hashes.size();
            for (std::size_t i = 0; i < wanted.size() && same; ++i) same = (t.nodes()[wanted[i]] == ref.auxiliary_hashes[i]);
            CHECK(same);
        }
    }
    /* 3. short verifier messages stay on the host in the gf192 override of blake2b_FieldT_randomness_extractor:
          element i is the keyed hash of root || index (blake2b.tcc:162-183) */
    {
        const binary_hash_digest root(32, 'r');
        const std::size_t index = 5, n = 9;
        const std::vector<libff::gf192> got = blake2b_FieldT_randomness_extractor<libff::gf192>(root, index, n);
        CHECK(got.size() == n);
        unsigned char msg[40];
        std::memcpy(msg, root.data(), 32);
        std::memcpy(msg + 32, &index, 8);
        for (std::size_t i = 0; i < n && i < got.size(); ++i) {
            libff::gf192 want;
            crypto_generichash_blake2b((unsigned char *)&want, sizeof(want), msg, sizeof(msg), (const unsigned char *)&i, sizeof(i));
            CHECK(got[i] == want);
        }
    }
    std::printf("%d failed\n", failures);
    return failures ? 1 : 0;
}
