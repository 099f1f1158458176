/*
 * b200_dropin_common.hpp — shared helpers of the drop-in overrides (libiop_b200/dropin/libiop/...).
 *
 * Each override file shadows one reference .tcc on the include path: it includes the reference's own file unmodified
 * from where it lies and then explicitly specialises individual member functions for the field types the B200 library
 * serves (libff::gf192, libff::edwards_Fr).  No reference source is copied; every other field keeps the CPU code.
 */
#ifndef B200_DROPIN_COMMON_HPP_
#define B200_DROPIN_COMMON_HPP_
#include <memory>
#include <vector>

#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_elementwise.hpp"
#include "libiop_b200/host/b200_hotpath.hpp"
#include "libiop_b200/host/b200_residency.hpp"

namespace libiop {
namespace b200_glue {

/* device view of a host vector: its resident copy when the adapter produced the vector, else an asynchronous upload
   (the vector outlives the next synchronising call in every caller) */
template<typename FieldT>
libiop_b200::device_view to_device(const std::vector<FieldT> &v)
{
    return libiop_b200::device_view(v);
}

/* result of a device pass as the shared host vector the IOP layer expects: a lazy vector, i.e. the contents stay in the
   device buffer until (unless) the host reads them (libiop_b200/host/b200_residency.hpp) */
template<typename FieldT>
std::shared_ptr<std::vector<FieldT>> to_host(libiop_b200::device_buffer &b, std::size_t n)
{
    return std::make_shared<std::vector<FieldT>>(libiop_b200::make_lazy_vector<FieldT>(std::move(b), n));
}

} // namespace b200_glue
} // namespace libiop

/* stamps `MACRO(FIELD)` for every field the B200 library serves */
#define B200_FOR_EACH_FIELD(MACRO) MACRO(libff::gf192) MACRO(libff::edwards_Fr)
#endif
