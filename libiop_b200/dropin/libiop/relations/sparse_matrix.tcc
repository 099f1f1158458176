/* Shadow of libiop/relations/sparse_matrix.tcc for the drop-in build.
   r1cs_sparse_matrix's constructor (sparse_matrix.tcc:5-12) is where a constraint system is first seen together with the
   shared_ptr that owns it; for the served fields it also registers that pointer with the device-resident copy of the
   matrices (b200_constraint_matrices.hpp), which create_Az_Bz_Cz_from_variable_assignment (r1cs.tcc shadow) and
   multi_lincheck's set_challenge (basic_lincheck_aux.tcc shadow) then share.  Everything else is the reference's file. */
#include "b200_constraint_matrices.hpp"
#include B200_REF_FILE(libiop/relations/sparse_matrix.tcc)

namespace libiop {
#define B200_SPARSE_MATRIX_SPECIALISE(FIELD)                                                                               \
template<> inline r1cs_sparse_matrix<FIELD>::r1cs_sparse_matrix(std::shared_ptr<r1cs_constraint_system<FIELD>> constraint_system, \
                                                                const r1cs_sparse_matrix_type matrix_type) :               \
    constraint_system_(constraint_system), matrix_type_(matrix_type)                                                       \
{                                                                                                                          \
    b200_glue::constraint_matrices<FIELD>::instance().remember(constraint_system);                                         \
}
B200_SPARSE_MATRIX_SPECIALISE(libff::gf192)
B200_SPARSE_MATRIX_SPECIALISE(libff::edwards_Fr)
#undef B200_SPARSE_MATRIX_SPECIALISE
} // namespace libiop
