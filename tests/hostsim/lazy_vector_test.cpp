/* Unit test of the device-resident ("lazy") host vectors of libiop_b200/host/b200_residency.hpp + b200_pinned_arena.hpp,
   run against the host simulator of the C ABI (no GPU): registration by ownership, use of the device copy by adapter calls,
   materialisation on ANY host access (read, write, by-value copy), release on destruction, device-to-device clones, the
   query-phase gathers.  Also built against the real library by tests/cpp (same source) for the -m gpu run. */
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../libiop_b200/host/b200_device.hpp"
#include "../../libiop_b200/host/b200_residency.hpp"

B200_DEFINE_PINNED_OPERATOR_NEW

using namespace libiop_b200;
struct el { uint64_t w[3]; bool operator==(const el &o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2]; } };

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static unsigned long long d2h() { unsigned long long h, d; b200_transfer_bytes(context(), &h, &d); return d; }
static unsigned long long h2d() { unsigned long long h, d; b200_transfer_bytes(context(), &h, &d); return h; }

static device_buffer device_pattern(std::size_t n, uint64_t seed)
{
    std::vector<el> h(n);
    for (std::size_t i = 0; i < n; ++i) h[i] = el{{seed + i, seed * 3 + i, ~i}};
    device_buffer b(n);
    throw_status(context(), b200_upload(context(), b.get(), 0, reinterpret_cast<const uint64_t *>(h.data()), n));
    throw_status(context(), b200_sync(context()));
    return b;
}
static std::vector<el> read_back(const device_buffer &b, std::size_t n)
{
    std::vector<el> h(n);
    throw_status(context(), b200_download(context(), reinterpret_cast<uint64_t *>(h.data()), b.get(), 0, n));
    return h;
}

int main()
{
    residency &res = residency::instance();
    const std::size_t n = 100000;
    {   /* 1. a lazy vector costs no transfer and adapter calls find its device copy */
        const unsigned long long d0 = d2h(), u0 = h2d();
        std::vector<el> v = make_lazy_vector<el>(device_pattern(n, 7), n);
        const unsigned long long u1 = h2d();
        CHECK(v.size() == n && lazy_region::instance().contains(v.data()) && res.is_resident(v.data()));
        CHECK(d2h() == d0);
        { device_view view(v); CHECK(!view.buf().owns()); CHECK(read_back(view.buf(), n)[5] == (el{{7 + 5, 21 + 5, ~5ull}})); }
        { device_view part(v.data() + 100, 50); CHECK(!part.buf().owns()); CHECK(read_back(part.buf(), 50)[1] == (el{{7 + 101, 21 + 101, ~101ull}})); }
        CHECK(h2d() == u1 && u1 - u0 == n * 24);             /* only device_pattern's own upload */
        /* moving the vector keeps the registration (same storage) */
        std::vector<el> w = std::move(v);
        CHECK(res.is_resident(w.data()));
        /* 2. query-phase access: single elements and batches without materialising */
        CHECK(resident_element(w, 4242) == (el{{7 + 4242, 21 + 4242, ~4242ull}}));
        std::vector<std::size_t> pos = {0, 99999, 31337};
        el got[3];
        CHECK(res.gather(w.data(), pos, reinterpret_cast<uint64_t *>(got)) && got[1] == (el{{7 + 99999, 21 + 99999, ~99999ull}}));
        CHECK(res.is_resident(w.data()) && res.faults() == 0);
        /* 3. a plain host read materialises the whole vector, correctly, and ends the registration */
        const unsigned long long d1 = d2h();
        CHECK(w[31337] == (el{{7 + 31337, 21 + 31337, ~31337ull}}));
        CHECK(res.faults() == 1 && !res.is_resident(w.data()) && d2h() - d1 == n * 24);
        bool all = true;
        for (std::size_t i = 0; i < n; ++i) all &= (w[i] == el{{7 + i, 21 + i, ~(uint64_t)i}});
        CHECK(all);
        /* 4. the host may now change it: a later adapter call uploads the CURRENT contents (no stale device copy) */
        w[3] = el{{1, 2, 3}};
        const unsigned long long u2 = h2d();
        { device_view view(w); CHECK(view.buf().owns()); CHECK(read_back(view.buf(), n)[3] == (el{{1, 2, 3}})); }
        CHECK(h2d() - u2 == n * 24);
    }
    CHECK(res.resident_bytes() == 0);
    {   /* 5. destruction drops the device copy; the address is reused by the next lazy vector without confusion */
        const void *first;
        { std::vector<el> a = make_lazy_vector<el>(device_pattern(n, 1), n); first = a.data(); CHECK(res.resident_bytes() == n * 24); }
        CHECK(res.resident_bytes() == 0 && !res.is_resident(first));
        std::vector<el> b = make_lazy_vector<el>(device_pattern(n, 2), n);
        CHECK(b.data() == first);                               /* first fit: same block again */
        CHECK(resident_element(b, 10) == (el{{2 + 10, 6 + 10, ~10ull}}));
    }
    {   /* 6. a write faults too (in-place mutation cannot leave a stale device copy behind) */
        std::vector<el> v = make_lazy_vector<el>(device_pattern(n, 9), n);
        const std::size_t f0 = res.faults();
        v[n - 1] = el{{5, 5, 5}};
        CHECK(res.faults() == f0 + 1 && !res.is_resident(v.data()));
        CHECK(v[0] == (el{{9, 27, ~0ull}}) && v[n - 1] == (el{{5, 5, 5}}));
    }
    {   /* 7. a by-value copy made by code that knows nothing about lazy vectors is a correct copy */
        std::vector<el> v = make_lazy_vector<el>(device_pattern(n, 11), n);
        std::vector<el> copy(v);
        CHECK(copy.size() == n && copy[77] == (el{{11 + 77, 33 + 77, ~77ull}}) && copy[n - 1] == v[n - 1]);
    }
    {   /* 8. clone_vector keeps a device-resident vector on the device; the clone outlives the original */
        std::unique_ptr<std::vector<el>> v(new std::vector<el>(make_lazy_vector<el>(device_pattern(n, 13), n)));
        const unsigned long long d0 = d2h();
        std::vector<el> c = clone_vector(*v);
        CHECK(res.is_resident(c.data()) && res.is_resident(v->data()) && c.data() != v->data() && d2h() == d0);
        v.reset();
        CHECK(resident_element(c, 500) == (el{{13 + 500, 39 + 500, ~500ull}}));
        /* host_readable: one plain transfer into ordinary memory, the vector itself stays lazy */
        const host_readable<el> h(c);
        CHECK(&h.get() != &c && h.get()[500] == (el{{13 + 500, 39 + 500, ~500ull}}) && res.is_resident(c.data()));
        std::vector<el> plain(10, el{{1, 1, 1}});
        CHECK(&host_readable<el>(plain).get() == &plain);
        CHECK(clone_vector(plain) == plain);
    }
    {   /* 9. small results are simply downloaded */
        std::vector<el> s = make_lazy_vector<el>(device_pattern(100, 3), 100);
        CHECK(!lazy_region::instance().contains(s.data()) && s[99] == (el{{3 + 99, 9 + 99, ~99ull}}));
    }
    {   /* 10. deferred vectors: the contents are computed on first use only; inputs stay attached */
        int runs = 0;
        auto copy_input = [&runs](device_buffer &out, std::map<std::string, device_buffer> &in) {
            ++runs;
            throw_status(context(), b200_copy(context(), out.get(), 0, in.at("evals").get(), 0, out.size()));
        };
        {
            std::map<std::string, device_buffer> in; in.emplace("evals", device_pattern(n, 17));
            std::vector<el> d = make_deferred_vector<el>(n, std::move(in), copy_input);
            CHECK(res.is_resident(d.data()) && res.is_pending(d.data()) && runs == 0);
            CHECK(res.attachment(d.data(), "evals", n) != 0 && runs == 0);            /* reading an attachment does not trigger it */
            { device_view view(d); CHECK(runs == 1 && !view.buf().owns()); CHECK(read_back(view.buf(), n)[9] == (el{{17 + 9, 51 + 9, ~9ull}})); }
            CHECK(!res.is_pending(d.data()));
            { device_view again(d); CHECK(runs == 1); }
        }
        {   /* never used: never computed */
            std::map<std::string, device_buffer> in; in.emplace("evals", device_pattern(n, 19));
            std::vector<el> d = make_deferred_vector<el>(n, std::move(in), copy_input);
            CHECK(runs == 1);
        }
        CHECK(runs == 1 && res.resident_bytes() == 0);
        {   /* a host touch computes, then downloads */
            std::map<std::string, device_buffer> in; in.emplace("evals", device_pattern(n, 23));
            std::vector<el> d = make_deferred_vector<el>(n, std::move(in), copy_input);
            CHECK(d[n - 1] == (el{{23 + n - 1, 69 + n - 1, ~(uint64_t)(n - 1)}}) && runs == 2 && !res.is_resident(d.data()));
        }
        {   /* small: computed at once */
            std::map<std::string, device_buffer> in; in.emplace("evals", device_pattern(64, 29));
            std::vector<el> d = make_deferred_vector<el>(64, std::move(in), copy_input);
            CHECK(runs == 3 && d[5] == (el{{29 + 5, 87 + 5, ~5ull}}));
        }
    }
    CHECK(res.resident_bytes() == 0);
    std::printf("%d failed\n", failures);
    return failures ? 1 : 0;
}
