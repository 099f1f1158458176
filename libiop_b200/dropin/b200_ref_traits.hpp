/*
 * b200_ref_traits.hpp — binds the C++ adapter (libiop_b200/host) to the reference's own field types
 * (part of the drop-in build, libiop_b200/dropin/Makefile).
 * This is the code a libiop maintainer would add.
 */
#ifndef B200_REF_TRAITS_HPP_
#define B200_REF_TRAITS_HPP_
#include <cstring>
#include <type_traits>

#include <libff/algebra/fields/binary/gf192.hpp>
#include <libff/algebra/curves/edwards/edwards_fields.hpp>

#include "libiop_b200/host/b200_device.hpp"
#include "libiop_b200/host/b200_fields.hpp"

#define B200_STR2(x) #x
#define B200_STR(x) B200_STR2(x)
#define B200_REF_FILE(rel) B200_STR(B200_REF_ROOT/rel)

namespace libiop_b200 {
template<> struct field_traits<libff::gf192> {
    static constexpr int kind = B200_FIELD_GF192;
    static const b200_fp_params *fp() { return nullptr; }
};
template<> struct field_traits<libff::edwards_Fr> {
    static constexpr int kind = B200_FIELD_FP;
    static const b200_fp_params *fp()
    {
        static b200_fp_params P;
        static bool init = false;
        if (!init) {   /* Fp_model statics set by edwards_pp::init_public_params() (edwards_fields.cpp:24-46) */
            std::memset(&P, 0, sizeof P);
            std::memcpy(P.modulus, libff::edwards_Fr::mod.data, 24);
            std::memcpy(P.r2, libff::edwards_Fr::Rsquared.data, 24);
            const libff::edwards_Fr one = libff::edwards_Fr::one();
            std::memcpy(P.one_mont, one.mont_repr.data, 24);
            P.inv = libff::edwards_Fr::inv;
            P.two_adicity = (uint32_t)libff::edwards_Fr::s;
            P.num_bits = (uint32_t)libff::edwards_Fr::num_bits;
            init = true;
        }
        return &P;
    }
};
template<typename FieldT> struct b200_enabled : std::false_type {};
template<> struct b200_enabled<libff::gf192> : std::true_type {};
template<> struct b200_enabled<libff::edwards_Fr> : std::true_type {};
} // namespace libiop_b200
#endif
