/*
 * b200_fields.hpp — host-side field element types of the adapter (used to describe domains and
 * challenges; bulk arithmetic happens on the device).
 *   gf192      same encoding and semantics as libff::gf192 (gf192.hpp:19-99, gf192.cpp:50-59,72-184,245-288)
 *   edwards_Fr same encoding as libff::edwards_Fr = Fp_model<3, r> in Montgomery form
 *              (fp.hpp:40-60, edwards_fields.cpp:24-46)
 * and field_traits<FieldT>, which tells the generic hot-path templates which kernels a field maps to.
 * A maintainer plugging the adapter under the reference's own types specialises field_traits for
 * libff::gf192 / libff::edwards_Fr.
 */
#ifndef LIBIOP_B200_FIELDS_HPP_
#define LIBIOP_B200_FIELDS_HPP_

#include <cstdint>
#include <cstring>

#include "../../include/b200_iop.h"
#include "../csrc/fp192.cuh"
#include "../csrc/gf192.cuh"

namespace libiop_b200 {

class gf192 {
    uint64_t value_[3]; /* little-endian, gf192.hpp:98-99 */
    static b200::gf192 dev(const gf192 &a) { b200::gf192 r; std::memcpy(&r, a.value_, 24); return r; }
    static gf192 host(const b200::gf192 &a) { gf192 r; std::memcpy(r.value_, &a, 24); return r; }
public:
    gf192() : value_{0, 0, 0} {}
    explicit gf192(uint64_t low) : value_{low, 0, 0} {}
    gf192(uint64_t high, uint64_t mid, uint64_t low) : value_{low, mid, high} {} /* gf192.cpp:32-35 */
    gf192 operator+(const gf192 &o) const { return gf192(value_[2] ^ o.value_[2], value_[1] ^ o.value_[1], value_[0] ^ o.value_[0]); }
    gf192 operator-(const gf192 &o) const { return *this + o; }
    gf192 &operator+=(const gf192 &o) { *this = *this + o; return *this; }
    gf192 operator*(const gf192 &o) const { return host(b200::gf_mul(dev(*this), dev(o))); }
    gf192 &operator*=(const gf192 &o) { *this = *this * o; return *this; }
    gf192 squared() const { return host(b200::gf_sqr(dev(*this))); }
    gf192 inverse() const { return host(b200::gf_inv(dev(*this))); }
    bool operator==(const gf192 &o) const { return value_[0] == o.value_[0] && value_[1] == o.value_[1] && value_[2] == o.value_[2]; }
    bool operator!=(const gf192 &o) const { return !(*this == o); }
    bool is_zero() const { return !(value_[0] | value_[1] | value_[2]); }
    static gf192 zero() { return gf192(); }
    static gf192 one() { return gf192(1); }
    const uint64_t *words() const { return value_; }
    static constexpr std::size_t extension_degree() { return 192; }
};
static_assert(sizeof(gf192) == 24, "gf192 must be 3 x u64");

/* edwards_Fr: r = 1552511030102430251236801561344621993261920897571225601, s = 31, generator 19 */
struct edwards_Fr_params {
    static const b200_fp_params &get()
    {
        static const b200_fp_params P = make();
        return P;
    }
private:
    static b200_fp_params make()
    {
        b200_fp_params P;
        std::memset(&P, 0, sizeof P);
        const char *dec = "1552511030102430251236801561344621993261920897571225601";
        for (const char *c = dec; *c; ++c) {
            unsigned __int128 carry = (unsigned)(*c - '0');
            for (int i = 0; i < 3; ++i) {
                const unsigned __int128 v = (unsigned __int128)P.modulus[i] * 10 + carry;
                P.modulus[i] = (uint64_t)v; carry = v >> 64;
            }
        }
        uint64_t x = 1;
        for (int i = 0; i < 6; ++i) x *= 2 - P.modulus[0] * x;
        P.inv = (uint64_t)0 - x;
        uint64_t t[3] = {1, 0, 0};
        for (int i = 0; i < 384; ++i) {       /* R and R^2 mod p by doubling */
            uint64_t c = 0, u[3];
            for (int k = 0; k < 3; ++k) { const unsigned __int128 s = (unsigned __int128)t[k] * 2 + c; t[k] = (uint64_t)s; c = (uint64_t)(s >> 64); }
            uint64_t b = 0;
            for (int k = 0; k < 3; ++k) { const unsigned __int128 dd = (unsigned __int128)t[k] - P.modulus[k] - b; u[k] = (uint64_t)dd; b = (uint64_t)(dd >> 64) & 1; }
            if (c || !b) std::memcpy(t, u, sizeof t);
            if (i == 191) std::memcpy(P.one_mont, t, sizeof t);
        }
        std::memcpy(P.r2, t, sizeof t);
        P.two_adicity = 31; P.num_bits = 181;
        return P;
    }
};

template<typename Params>
class Fp192 {
    uint64_t mont_repr_[3];
    static b200::fp_dev_params dp() { const b200_fp_params &P = Params::get(); return b200::make_fp_dev_params(P.modulus, P.r2, P.one_mont, P.inv); }
    static b200::fp192 dev(const Fp192 &a) { b200::fp192 r; std::memcpy(&r, a.mont_repr_, 24); return r; }
    static Fp192 host(const b200::fp192 &a) { Fp192 r; std::memcpy(r.mont_repr_, &a, 24); return r; }
public:
    Fp192() : mont_repr_{0, 0, 0} {}
    explicit Fp192(uint64_t v) { *this = host(b200::fp_from_u64(v, dp())); }
    Fp192 operator+(const Fp192 &o) const { return host(b200::fp_add(dev(*this), dev(o), dp())); }
    Fp192 operator-(const Fp192 &o) const { return host(b200::fp_sub(dev(*this), dev(o), dp())); }
    Fp192 operator*(const Fp192 &o) const { return host(b200::fp_mul(dev(*this), dev(o), dp())); }
    Fp192 &operator*=(const Fp192 &o) { *this = *this * o; return *this; }
    Fp192 squared() const { return *this * *this; }
    Fp192 inverse() const { return host(b200::fp_inv(dev(*this), dp())); }
    Fp192 pow(uint64_t e) const { return host(b200::fp_pow_u64(dev(*this), e, dp())); }
    bool operator==(const Fp192 &o) const { return !std::memcmp(mont_repr_, o.mont_repr_, 24); }
    bool operator!=(const Fp192 &o) const { return !(*this == o); }
    static Fp192 zero() { return Fp192(); }
    static Fp192 one() { return Fp192(1); }
    static Fp192 multiplicative_generator() { return Fp192(19); } /* edwards_fields.cpp:43 */
    /* generator of the order-2^logn subgroup: generator^((p-1)/2^logn)  (subgroup.tcc:56-59) */
    static Fp192 subgroup_generator(uint32_t logn)
    {
        const b200_fp_params &P = Params::get();
        uint64_t e[3] = {P.modulus[0] - 1, P.modulus[1], P.modulus[2]};
        for (uint32_t i = 0; i < logn; ++i) { e[0] = (e[0] >> 1) | (e[1] << 63); e[1] = (e[1] >> 1) | (e[2] << 63); e[2] >>= 1; }
        Fp192 acc = one(), base = multiplicative_generator();
        for (int i = 0; i < 192; ++i) { if ((e[i / 64] >> (i % 64)) & 1) acc = acc * base; base = base * base; }
        return acc;
    }
    const uint64_t *words() const { return mont_repr_; }
};
typedef Fp192<edwards_Fr_params> edwards_Fr;
static_assert(sizeof(edwards_Fr) == 24, "Fp elements must be 3 x u64");

/* field_traits: kind = B200_FIELD_GF192 / B200_FIELD_FP, fp() = parameters (nullptr for gf192) */
template<typename FieldT> struct field_traits;
template<> struct field_traits<gf192> {
    static constexpr int kind = B200_FIELD_GF192;
    static const b200_fp_params *fp() { return nullptr; }
};
template<typename Params> struct field_traits<Fp192<Params>> {
    static constexpr int kind = B200_FIELD_FP;
    static const b200_fp_params *fp() { return &Params::get(); }
};

} // namespace libiop_b200
#endif
