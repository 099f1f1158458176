// CPU unit test of the device field arithmetic (libiop_b200/csrc/gf192.cuh, fp192.cuh are
// __host__ __device__): compares against the oracle (oracle/liboracle.so) on random and extreme inputs.
// Built and run by tests/test_host_field.py; no GPU needed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../libiop_b200/csrc/gf192.cuh"
#include "../../libiop_b200/csrc/fp192.cuh"
#include "../../oracle/oracle.h"

using namespace b200;

static gf192 to_gf(const f192 &a) { gf192 r; memcpy(&r, &a, 24); return r; }
static f192 from_gf(const gf192 &a) { f192 r; memcpy(&r, &a, 24); return r; }
static bool same(const gf192 &a, const f192 &b) { return memcmp(&a, &b, 24) == 0; }

int main()
{
    std::mt19937_64 rng(12345);
    int fails = 0;
    // clmul32 worst cases: all-ones exercises the maximum per-position carry counts of both spacings
    {
        const uint32_t vals[] = {0xffffffffu, 0x80000001u, 0x42108421u, 0x84210842u, 0x11111111u, 0xdeadbeefu, 1u, 0u};
        for (uint32_t a : vals) for (uint32_t b : vals) {
            uint64_t want = 0;
            for (int i = 0; i < 32; ++i) if ((b >> i) & 1) want ^= (uint64_t)a << i;
            const uint64_t g4 = clmul32_s4(split4(a), split4(b));
            const uint64_t g5 = clmul32_s5(split5(a), split5(b));
            if (g4 != want || g5 != want) { printf("clmul32 mismatch a=%08x b=%08x s4=%d s5=%d\n", a, b, g4 == want, g5 == want); ++fails; }
        }
        for (int it = 0; it < 200000; ++it) {
            const uint32_t a = (uint32_t)rng(), b = (uint32_t)rng();
            uint64_t want = 0;
            for (int i = 0; i < 32; ++i) if ((b >> i) & 1) want ^= (uint64_t)a << i;
            if (clmul32_s4(split4(a), split4(b)) != want || clmul32_s5(split5(a), split5(b)) != want) { ++fails; break; }
        }
    }
    // gf192 mul / sqr / inv vs oracle
    for (int it = 0; it < 20000; ++it) {
        f192 a, b, r;
        for (int i = 0; i < 3; ++i) { a.w[i] = rng(); b.w[i] = rng(); }
        if (it < 8) { for (int i = 0; i < 3; ++i) a.w[i] = (it & 1) ? ~0ull : 0x8000000000000001ull; }
        if (it < 4) { for (int i = 0; i < 3; ++i) b.w[i] = ~0ull; }
        orc_gf192_mul(&r, &a, &b);
        if (!same(gf_mul(to_gf(a), to_gf(b)), r)) { if (fails < 5) printf("gf_mul mismatch at %d\n", it); ++fails; }
        orc_gf192_mul(&r, &a, &a);
        if (!same(gf_sqr(to_gf(a)), r)) { if (fails < 5) printf("gf_sqr mismatch at %d\n", it); ++fails; }
        if (it < 200 && (a.w[0] | a.w[1] | a.w[2])) {
            orc_gf192_inv(&r, &a);
            if (!same(gf_inv(to_gf(a)), r)) { if (fails < 5) printf("gf_inv mismatch at %d\n", it); ++fails; }
        }
    }
    // Fp (edwards_Fr) vs oracle
    orc_fp_params OP; orc_fp_params_edwards(&OP);
    const fp_dev_params P = make_fp_dev_params(OP.mod, OP.r2, OP.one, OP.inv);
    for (int it = 0; it < 20000; ++it) {
        f192 a, b, r;
        for (int i = 0; i < 3; ++i) { a.w[i] = rng(); b.w[i] = rng(); }
        a.w[2] &= (1ull << 52) - 1; b.w[2] &= (1ull << 52) - 1;   // < 2^180 < p
        if (it == 0) { memcpy(a.w, OP.mod, 24); a.w[0] -= 1; memcpy(b.w, a.w, 24); }   // p-1
        fp192 da, db; memcpy(&da, &a, 24); memcpy(&db, &b, 24);
        fp192 g;
        orc_fp_mul(&OP, &r, &a, &b); g = fp_mul(da, db, P);
        if (memcmp(&g, &r, 24)) { if (fails < 5) printf("fp_mul mismatch at %d\n", it); ++fails; }
        orc_fp_add(&OP, &r, &a, &b); g = fp_add(da, db, P);
        if (memcmp(&g, &r, 24)) { if (fails < 5) printf("fp_add mismatch at %d\n", it); ++fails; }
        orc_fp_sub(&OP, &r, &a, &b); g = fp_sub(da, db, P);
        if (memcmp(&g, &r, 24)) { if (fails < 5) printf("fp_sub mismatch at %d\n", it); ++fails; }
        if (it < 50) {
            orc_fp_inv(&OP, &r, &a); g = fp_inv(da, P);
            if (memcmp(&g, &r, 24)) { if (fails < 5) printf("fp_inv mismatch at %d\n", it); ++fails; }
        }
    }
    printf(fails ? "FAIL %d\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
