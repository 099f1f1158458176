/* Shadow of libiop/protocols/aurora_iop.tcc for the drop-in build.
   aurora_iop's constructor (aurora_iop.tcc:262-322) deep-copies the caller's constraint system TWICE per proof — once
   into the value member constraint_system_ (aurora_iop.hpp:78; read only for num_inputs() in the constructor) and once
   into the shared_ptr the encoded protocol and its three r1cs_sparse_matrix views keep (aurora_iop.tcc:296-297).  At
   N = 2^20 that is 2 x 3 x 2^20 heap-allocated linear combinations built and torn down again: ~75 ms of a 0.22 s proof
   once the data-parallel work runs on the device (profiles/r2_dropin_phases_N20.txt).
   For the served fields the constructor keeps the sizes only in the value member and hands the protocol a non-owning
   shared_ptr to the CALLER's constraint system, which aurora_snark_prover / aurora_snark_verifier (snark/aurora_snark.tcc:
   120-146, 149-179) keep alive for longer than the aurora_iop object.  -DB200_DROPIN_COPY_CONSTRAINT_SYSTEM restores the
   reference's copies for callers that cannot promise that. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/aurora_iop.tcc)

namespace libiop {
namespace b200_glue {
template<typename FieldT>
r1cs_constraint_system<FieldT> constraint_system_sizes(const r1cs_constraint_system<FieldT> &cs)
{
    r1cs_constraint_system<FieldT> s;
    s.primary_input_size_ = cs.primary_input_size_;
    s.auxiliary_input_size_ = cs.auxiliary_input_size_;
    return s;
}
template<typename FieldT>
std::shared_ptr<r1cs_constraint_system<FieldT>> borrow_constraint_system(const r1cs_constraint_system<FieldT> &cs)
{
#ifdef B200_DROPIN_COPY_CONSTRAINT_SYSTEM
    return std::make_shared<r1cs_constraint_system<FieldT>>(cs);
#else
    /* a control block of its own (so that weak_ptr observers — b200_constraint_matrices.hpp — see it die), no deletion */
    return std::shared_ptr<r1cs_constraint_system<FieldT>>(const_cast<r1cs_constraint_system<FieldT> *>(&cs),
                                                           [](r1cs_constraint_system<FieldT> *) {});
#endif
}
} // namespace b200_glue

#ifndef B200_DROPIN_COPY_CONSTRAINT_SYSTEM
#define B200_AURORA_IOP_CTOR(FIELD)                                                                                        \
template<> inline aurora_iop<FIELD>::aurora_iop(iop_protocol<FIELD> &IOP, const r1cs_constraint_system<FIELD> &constraint_system, \
                                                const aurora_iop_parameters<FIELD> &parameters) :                          \
    IOP_(IOP), constraint_system_(b200_glue::constraint_system_sizes(constraint_system)), parameters_(parameters)          \
{                                                                                                                          \
    if (!libff::is_power_of_2(constraint_system.num_inputs() + 1))                                                         \
        throw std::invalid_argument("number of inputs in the constraint system must be one less than a power of two.");    \
    /* the codeword domain's shift: an element outside the unshifted domain of that size (aurora_iop.tcc:278-285) */       \
    const FIELD shift = field_subset<FIELD>(1ull << parameters.codeword_domain_dim()).element_outside_of_subset();         \
    const domain_handle constraint_handle = IOP.register_domain(field_subset<FIELD>(1ull << parameters.constraint_domain_dim())); \
    const domain_handle variable_handle = IOP.register_domain(field_subset<FIELD>(1ull << parameters.variable_domain_dim())); \
    const field_subset<FIELD> codeword_domain(1ull << parameters.codeword_domain_dim(), shift);                            \
    this->codeword_domain_handle_ = IOP.register_domain(codeword_domain);                                                  \
    this->protocol_ = std::make_shared<encoded_aurora_protocol<FIELD>>(                                                    \
        this->IOP_, constraint_handle, variable_handle, this->codeword_domain_handle_,                                     \
        b200_glue::borrow_constraint_system(constraint_system), parameters.encoded_aurora_params_);                        \
    this->LDT_reducer_ = std::make_shared<LDT_instance_reducer<FIELD, FRI_protocol<FIELD>>>(                               \
        this->IOP_, this->codeword_domain_handle_, this->parameters_.LDT_reducer_params_);                                 \
    this->IOP_.set_round_parameters(round_parameters<FIELD>(this->parameters_.FRI_params_.quotient_map_domain(codeword_domain))); \
}
B200_FOR_EACH_FIELD(B200_AURORA_IOP_CTOR)
#undef B200_AURORA_IOP_CTOR
#endif
} // namespace libiop
