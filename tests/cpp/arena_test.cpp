/* Host-only unit test of libiop_b200/host/b200_pinned_arena.hpp: the free-list allocator behind the pinned
   operator new, driven with malloc'd slabs (no GPU), plus make_uninitialized_vector. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../libiop_b200/host/b200_pinned_arena.hpp"

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static size_t slab_calls = 0;
static void *slab(std::size_t n) { ++slab_calls; void *p = nullptr; return posix_memalign(&p, 4096, n) == 0 ? p : nullptr; }   /* page-aligned like cudaHostAlloc */
static void *no_slab(std::size_t) { return nullptr; }

int main()
{
    using libiop_b200::pinned_arena;
    {   /* random allocate / release: blocks never overlap, contents survive, everything coalesces back */
        pinned_arena a(&slab, std::size_t(8) << 20);
        std::mt19937_64 rng(1);
        struct blk { unsigned char *p; size_t n; unsigned char tag; };
        std::vector<blk> live;
        for (int it = 0; it < 4000; ++it) {
            if (live.size() < 8 || (rng() % 2 && live.size() < 200)) {   /* at most ~300 MiB live: 38 of the 256 slab slots */
                const size_t n = 1 + rng() % (std::size_t(3) << 20);
                unsigned char *p = (unsigned char *)a.allocate(n);
                CHECK(p != nullptr);
                CHECK(((uintptr_t)p & 4095) == 0);
                const unsigned char tag = (unsigned char)(rng() & 0xff);
                std::memset(p, tag, n);
                live.push_back({p, n, tag});
            } else {
                const size_t k = rng() % live.size();
                blk b = live[k];
                live[k] = live.back(); live.pop_back();
                bool ok = true;
                for (size_t i = 0; i < b.n; i += 997) ok &= (b.p[i] == b.tag);
                CHECK(ok && b.p[b.n - 1] == b.tag);
                CHECK(a.release(b.p));
            }
        }
        for (auto &b : live) {
            bool ok = true;
            for (size_t i = 0; i < b.n; i += 997) ok &= (b.p[i] == b.tag);
            CHECK(ok);
            CHECK(a.release(b.p));
        }
        CHECK(a.free_bytes() == a.reserved_bytes());
        /* after full coalescing a slab-sized request is served without growing */
        const size_t before = a.slab_count();
        void *big = a.allocate(std::size_t(8) << 20);
        CHECK(big != nullptr && a.slab_count() == before);
        CHECK(a.release(big));
        int local;
        CHECK(!a.release(&local));                       /* foreign pointers are declined */
        void *huge = a.allocate(std::size_t(40) << 20);  /* larger than a slab: gets its own */
        CHECK(huge != nullptr);
        CHECK(a.release(huge));
    }
    {   /* page-locking unavailable: the arena declines and the caller falls back to malloc */
        pinned_arena a(&no_slab, std::size_t(1) << 20);
        CHECK(a.allocate(1 << 20) == nullptr);
        CHECK(a.slab_count() == 0);
    }
    {   /* make_uninitialized_vector: right size, writable, destructible, growable */
        struct el { unsigned long long w[3]; el() : w{7, 7, 7} {} };
        std::vector<el> v = libiop_b200::make_uninitialized_vector<el>(100000);
        CHECK(v.size() == 100000 && v.capacity() >= 100000);
        for (auto &e : v) e.w[0] = 1;
        v.push_back(el());
        CHECK(v.size() == 100001 && v[100000].w[1] == 7 && v[5].w[0] == 1);
        std::vector<el> e0 = libiop_b200::make_uninitialized_vector<el>(0);
        CHECK(e0.empty());
    }
    std::printf("%d failed\n", failures);
    return failures ? 1 : 0;
}
