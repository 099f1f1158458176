/* Shadow of libiop/protocols/ligero_iop.tcc for the drop-in build.
   ligero_iop's constructor (ligero_iop.tcc:303-334) copies the caller's constraint system into a value member
   (ligero_iop.hpp:78) that nothing reads afterwards — interleaved_r1cs_protocol receives the caller's reference and keeps
   its own copy (ligero.hpp:43).  For the served fields the member keeps the sizes only: one deep copy (and its destruction)
   of 2^18 heap-allocated constraints less per proof.  -DB200_DROPIN_COPY_CONSTRAINT_SYSTEM restores the reference's copy. */
#include "b200_dropin_common.hpp"
#include "libiop/protocols/aurora_iop.hpp"          /* b200_glue::constraint_system_sizes (aurora_iop.tcc shadow) */
#include B200_REF_FILE(libiop/protocols/ligero_iop.tcc)

namespace libiop {
#ifndef B200_DROPIN_COPY_CONSTRAINT_SYSTEM
#define B200_LIGERO_IOP_CTOR(FIELD)                                                                                        \
template<> inline ligero_iop<FIELD>::ligero_iop(iop_protocol<FIELD> &IOP, const r1cs_constraint_system<FIELD> &constraint_system, \
                                                const ligero_iop_parameters<FIELD> &parameters) :                          \
    IOP_(IOP), constraint_system_(b200_glue::constraint_system_sizes(constraint_system)), parameters_(parameters)          \
{                                                                                                                          \
    const std::size_t systematic_size = 1ull << parameters.systematic_domain_dim();                                        \
    const std::size_t codeword_dim = parameters.systematic_domain_dim() + this->parameters_.RS_extra_dimensions();         \
    this->codeword_domain_ = field_subset<FIELD>(1ull << codeword_dim);                                                    \
    /* the systematic domains are shifted out of the codeword domain (ligero_iop.tcc:314-318) */                           \
    const FIELD shift = this->codeword_domain_.element_outside_of_subset();                                                \
    const domain_handle codeword_handle = this->IOP_.register_domain(this->codeword_domain_);                              \
    const domain_handle systematic_handle = this->IOP_.register_domain(field_subset<FIELD>(systematic_size, shift));       \
    const domain_handle extended_handle = this->IOP_.register_domain(field_subset<FIELD>(systematic_size << 1, shift));    \
    this->protocol_ = std::make_shared<interleaved_r1cs_protocol<FIELD>>(                                                  \
        this->IOP_, codeword_handle, systematic_handle, extended_handle, constraint_system, this->parameters_.encoded_ligero_params_); \
    this->LDT_reducer_ = std::make_shared<LDT_instance_reducer<FIELD, direct_LDT_protocol<FIELD>>>(                        \
        this->IOP_, codeword_handle, this->parameters_.ldt_reducer_params_);                                               \
}
B200_FOR_EACH_FIELD(B200_LIGERO_IOP_CTOR)
#undef B200_LIGERO_IOP_CTOR
#endif
} // namespace libiop
