/* Shadow of libiop/protocols/ldt/fri/fri_aux.tcc for the drop-in build:
   evaluate_next_f_i_over_entire_domain (fri_aux.hpp:23-28) forwards to the B200 adapter. */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_hotpath.hpp"

#define evaluate_next_f_i_over_entire_domain ref_cpu_evaluate_next_f_i_over_entire_domain
#include B200_REF_FILE(libiop/protocols/ldt/fri/fri_aux.tcc)
#undef evaluate_next_f_i_over_entire_domain

namespace libiop {
namespace b200_glue {
template<typename FieldT> std::shared_ptr<std::vector<FieldT>> fold(const std::shared_ptr<std::vector<FieldT>> &f, const field_subset<FieldT> &d,
                                                                    size_t cs, const FieldT x, std::true_type)
{ return libiop_b200::evaluate_next_f_i_over_entire_domain<FieldT>(f, d, cs, x); }
template<typename FieldT> std::shared_ptr<std::vector<FieldT>> fold(const std::shared_ptr<std::vector<FieldT>> &f, const field_subset<FieldT> &d,
                                                                    size_t cs, const FieldT x, std::false_type)
{ return ref_cpu_evaluate_next_f_i_over_entire_domain<FieldT>(f, d, cs, x); }
}
template<typename FieldT>
std::shared_ptr<std::vector<FieldT>> evaluate_next_f_i_over_entire_domain(
    const std::shared_ptr<std::vector<FieldT>> &f_i_evals, const field_subset<FieldT> &f_i_domain, const size_t coset_size, const FieldT x_i)
{ return b200_glue::fold<FieldT>(f_i_evals, f_i_domain, coset_size, x_i, libiop_b200::b200_enabled<FieldT>()); }
} // namespace libiop
