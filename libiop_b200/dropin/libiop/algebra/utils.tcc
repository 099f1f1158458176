/* Shadow of libiop/algebra/utils.tcc for the drop-in build.
   batch_inverse_and_mul (utils.tcc:38-103) — Montgomery's batch inversion, a loop-carried chain of 3n multiplications on
   the host — runs on the device for long vectors (b200_gf192_inv / b200_fp_inv: one inversion per run of 32 / 16 elements,
   runs in parallel), followed by the scalar multiplication.  A zero element yields zero, which is what the reference's
   has_zeroes path returns (utils.tcc:73-103); without zeroes both give the unique inverses. */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_elementwise.hpp"
#include "libiop_b200/host/b200_residency.hpp"

namespace libiop {
template<typename FieldT> std::vector<FieldT> ref_cpu_batch_inverse_and_mul(const std::vector<FieldT> &vec, const FieldT &k, const bool has_zeroes);
}
#define batch_inverse ref_cpu_batch_inverse
#define batch_inverse_and_mul ref_cpu_batch_inverse_and_mul
#include B200_REF_FILE(libiop/algebra/utils.tcc)
#undef batch_inverse_and_mul
#undef batch_inverse

namespace libiop {
namespace b200_glue {
template<typename FieldT>
std::vector<FieldT> inverse_and_mul(const std::vector<FieldT> &vec, const FieldT &k, const bool has_zeroes, std::true_type)
{
    using namespace libiop_b200;
    /* short host vectors stay on the host; a device-resident vector is inverted where it is, whatever its length */
    if (vec.size() < 16384 && !residency::instance().is_resident(vec.data())) return ref_cpu_batch_inverse_and_mul<FieldT>(vec, k, has_zeroes);
    const device_view in(vec);
    device_buffer out(vec.size());
    vec_inverse<FieldT>(out, in.buf(), vec.size());
    if (k != FieldT::one()) vec_affine<FieldT>(out, out, &k, nullptr, vec.size());
    return make_lazy_vector<FieldT>(std::move(out), vec.size());
}
template<typename FieldT>
std::vector<FieldT> inverse_and_mul(const std::vector<FieldT> &vec, const FieldT &k, const bool has_zeroes, std::false_type)
{ return ref_cpu_batch_inverse_and_mul<FieldT>(vec, k, has_zeroes); }
}
template<typename FieldT>
std::vector<FieldT> batch_inverse_and_mul(const std::vector<FieldT> &vec, const FieldT &k, const bool has_zeroes)
{ return b200_glue::inverse_and_mul<FieldT>(vec, k, has_zeroes, libiop_b200::b200_enabled<FieldT>()); }
template<typename FieldT>
std::vector<FieldT> batch_inverse(const std::vector<FieldT> &vec, const bool has_zeroes)
{ return batch_inverse_and_mul<FieldT>(vec, FieldT::one(), has_zeroes); }
} // namespace libiop
