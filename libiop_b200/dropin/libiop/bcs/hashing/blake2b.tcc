/* Shadow of libiop/bcs/hashing/blake2b.tcc for the drop-in build: the verifier's random messages
   (blake2b_FieldT_randomness_extractor, blake2b.tcc:231-256, one keyed BLAKE2b per field element, reached from
   blake2b_hashchain::squeeze at blake2b.tcc:76-85) are drawn on the B200 (b200_squeeze_field_elements) once a message
   is long enough to pay for a launch — Ligero's lincheck / rowcheck challenges are N elements each.  Short messages
   keep the reference's per-element sampler on the host. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/bcs/hashing/blake2b.tcc)

namespace libiop {
namespace b200_glue {
constexpr std::size_t squeeze_on_device_from = 2048;   /* elements; below this the host loop is faster than launch + copy */
} // namespace b200_glue

#define B200_SQUEEZE_SPECIALISE(FIELD)                                                                                     \
template<> inline std::vector<FIELD> blake2b_FieldT_randomness_extractor<FIELD>(                                           \
    const binary_hash_digest &root, const std::size_t index, const std::size_t num_elements)                               \
{                                                                                                                          \
    std::vector<unsigned char> root_plus_index(root.size() + sizeof(index));                                               \
    std::memcpy(root_plus_index.data(), root.data(), root.size());                                                         \
    std::memcpy(root_plus_index.data() + root.size(), &index, sizeof(index));                                              \
    if (num_elements >= b200_glue::squeeze_on_device_from && root_plus_index.size() % 8 == 0 && root_plus_index.size() <= 128) \
        return libiop_b200::squeeze_field_elements<FIELD>(root_plus_index.data(), root_plus_index.size(), num_elements);   \
    std::vector<FIELD> result;                                                                                             \
    result.reserve(num_elements);                                                                                          \
    for (std::size_t i = 0; i < num_elements; ++i)                                                                         \
        result.emplace_back(blake2b_FieldT_rejection_sample<FIELD>(FIELD::zero(), root_plus_index.data(),                  \
                                                                   root_plus_index.size(), i, num_elements));              \
    return result;                                                                                                         \
}
B200_FOR_EACH_FIELD(B200_SQUEEZE_SPECIALISE)
#undef B200_SQUEEZE_SPECIALISE
} // namespace libiop
