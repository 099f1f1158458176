/* Shadow of libiop/protocols/encoded/r1cs_rs_iop/r1cs_rs_iop.tcc for the drop-in build.
   fz_virtual_oracle::evaluated_contents (r1cs_rs_iop.tcc:181-226): f_z = f_w * Z_{1,v} + f_{1,v} over the codeword
   domain.  On the B200: Z_{1,v} over L is generated on the device from dim(L)+1 host values, f_{1,v} is LDE'd on the
   device from its |1,v| coefficients, and one fused pass forms the result; only f_w goes up and f_z comes down. */
#include "b200_dropin_common.hpp"
#include "b200_matrix_views.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/r1cs_rs_iop/r1cs_rs_iop.tcc)

namespace libiop {
/* create_fw_prime_evals (r1cs_rs_iop.tcc:423-446) reads f_{1,v} over the whole variable domain element by element; that
   vector is an FFT result, i.e. device-resident.  These overloads (preferred over the reference's template for the served
   fields) hand the reference's loop a host copy fetched with one transfer. */
#define B200_FW_PRIME_OVERLOAD(FIELD)                                                                                      \
inline std::vector<FIELD> create_fw_prime_evals(const std::vector<FIELD> &f_1v_over_variable_domain, const std::size_t primary_input_size, \
                                                const r1cs_auxiliary_input<FIELD> &auxiliary_input, const field_subset<FIELD> &variable_domain) \
{                                                                                                                          \
    const libiop_b200::host_readable<FIELD> f_1v(f_1v_over_variable_domain);                                               \
    return create_fw_prime_evals<FIELD>(f_1v.get(), primary_input_size, auxiliary_input, variable_domain);                 \
}
B200_FOR_EACH_FIELD(B200_FW_PRIME_OVERLOAD)
#undef B200_FW_PRIME_OVERLOAD

/* encoded_aurora_protocol::compute_fprime_ABCz_over_codeword_domain (r1cs_rs_iop.tcc:449-479): Az, Bz, Cz over the constraint
   domain -> f_Az, f_Bz, f_Cz over the codeword domain.  The reference interpolates each (IFFT over H), adds Z_H * R in the
   zero-knowledge case, and evaluates over L (FFT).  Without zero knowledge nothing happens between the two transforms, and
   the B200 library extends evaluations over H straight to L (b200_gf192_lde: the basis-only halves of the two transforms
   cancel) — 3 of the proof's 8 codeword-size LDEs, with no coefficient vector crossing PCIe.  The zero-knowledge case keeps
   the reference's three steps. */
#define B200_ABCZ_SPECIALISE(FIELD)                                                                                        \
template<> inline void encoded_aurora_protocol<FIELD>::compute_fprime_ABCz_over_codeword_domain(                           \
    std::vector<FIELD> &Az, std::vector<FIELD> &Bz, std::vector<FIELD> &Cz)                                                \
{                                                                                                                          \
    if (!this->params_.make_zk()) {                                                                                        \
        this->fprime_Az_over_codeword_domain_ = libiop_b200::LDE_over_field_subsets<FIELD>(Az, this->constraint_domain_, this->codeword_domain_); \
        this->fprime_Bz_over_codeword_domain_ = libiop_b200::LDE_over_field_subsets<FIELD>(Bz, this->constraint_domain_, this->codeword_domain_); \
        this->fprime_Cz_over_codeword_domain_ = libiop_b200::LDE_over_field_subsets<FIELD>(Cz, this->constraint_domain_, this->codeword_domain_); \
        return;                                                                                                            \
    }                                                                                                                      \
    polynomial<FIELD> f_Az(IFFT_over_field_subset<FIELD>(Az, this->constraint_domain_));                                   \
    polynomial<FIELD> f_Bz(IFFT_over_field_subset<FIELD>(Bz, this->constraint_domain_));                                   \
    polynomial<FIELD> f_Cz(IFFT_over_field_subset<FIELD>(Cz, this->constraint_domain_));                                   \
    const vanishing_polynomial<FIELD> constraint_vp(this->constraint_domain_);                                             \
    f_Az += constraint_vp * this->R_Az_;                                                                                   \
    f_Bz += constraint_vp * this->R_Bz_;                                                                                   \
    f_Cz += constraint_vp * this->R_Cz_;                                                                                   \
    this->fprime_Az_over_codeword_domain_ = FFT_over_field_subset<FIELD>(f_Az.coefficients(), this->codeword_domain_);     \
    this->fprime_Bz_over_codeword_domain_ = FFT_over_field_subset<FIELD>(f_Bz.coefficients(), this->codeword_domain_);     \
    this->fprime_Cz_over_codeword_domain_ = FFT_over_field_subset<FIELD>(f_Cz.coefficients(), this->codeword_domain_);     \
}
B200_FOR_EACH_FIELD(B200_ABCZ_SPECIALISE)
#undef B200_ABCZ_SPECIALISE

namespace b200_glue {
/* f_Az, f_Bz, f_Cz over the codeword domain, one matrix at a time: flatten M (host), then M z and its extension to L
   (device, asynchronous) — so the device extends A z while the host flattens B, and B z while it flattens C
   (create_Az_Bz_Cz_from_variable_assignment + compute_fprime_ABCz_over_codeword_domain, r1cs_rs_iop.tcc:449-479, 571-598,
   without zero knowledge: nothing happens between the products and their extensions).  false = this constraint system
   has no device-resident matrices (too small, or not registered); nothing is done then. */
template<typename FieldT>
bool extend_constraint_products(const r1cs_constraint_system<FieldT> &cs, const std::vector<FieldT> &variable_assignment,
                                const field_subset<FieldT> &constraint_domain, const field_subset<FieldT> &codeword_domain,
                                std::vector<FieldT> *const out[3])
{
    using namespace libiop_b200;
    const std::size_t n_H = constraint_domain.num_elements(), n_L = codeword_domain.num_elements();
    if (cs.constraints_.size() != n_H) return false;
    for (int m = 0; m < 3; ++m) {
        const flat_matrix *M = constraint_matrices<FieldT>::instance().get_matrix(&cs, m);
        if (!M) { if (m) throw std::logic_error("B200 drop-in: constraint matrices went away mid-proof"); return false; }
        const device_view d_z(variable_assignment);
        device_buffer d_y(n_H), d_out(n_L);
        throw_status(context(), b200_spmv_csr_dev(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), M->n_rows,
                                                  M->row_ptr->get(), M->col->get(), M->coeff.get(), d_z.get(), d_y.get()));
        if (!LDE_over_field_subsets_device<FieldT>(d_y, constraint_domain, codeword_domain, d_out)) {
            device_buffer coeffs(n_H);
            IFFT_over_field_subset_device<FieldT>(d_y, constraint_domain, coeffs);
            FFT_over_field_subset_device<FieldT>(coeffs, n_H, codeword_domain, d_out);
        }
        *out[m] = make_lazy_vector<FieldT>(std::move(d_out), n_L);
    }
    return true;
}

/* f_w over the codeword domain WITHOUT a coefficient vector (submit_witness_oracles, r1cs_rs_iop.tcc:497-569).
   The reference interpolates \hat f_w = (witness - f_{1,v}) over the variable domain (IFFT), divides the coefficient vector by
   the vanishing polynomial of the input positions, and evaluates the quotient over L (FFT).  For subspace domains over
   GF(2^192) the same codeword comes out of evaluations alone:
       z            = (1, primary, auxiliary) on the device (the vector create_Az_Bz_Cz multiplies next — it goes up once),
       \hat f_w|V   = z + f_{1,v}|V         (zero at the input positions by construction: f_{1,v} interpolates them),
       \hat f_w|L   = fused LDE V -> L      (+ Z_V|L * mask|L in the zero-knowledge case),
       f_w|L        = \hat f_w|L / Z_in|L,  Z_in|L constant on the 2^dim(in) elements that differ only in the input basis
                                            directions: |L| / 2^dim(in) inversions and one table-lookup multiplication.
   Field arithmetic is exact, so the values are those of the reference's quotient polynomial; 32 ms -> 8 ms of an N = 2^20
   proof (profiles/r2_host_gaps_N20_v1.txt vs _v2).  Returns false — nothing done — when the domains do not qualify. */
template<typename FieldT>
bool witness_codeword_on_device(const r1cs_primary_input<FieldT> &primary_input, const r1cs_auxiliary_input<FieldT> &auxiliary_input,
                                const std::vector<FieldT> &f_1v_coefficients, const field_subset<FieldT> &input_variable_domain,
                                const field_subset<FieldT> &variable_domain, const field_subset<FieldT> &codeword_domain,
                                const polynomial<FieldT> *fw_mask, std::vector<FieldT> &fw_over_codeword_domain,
                                std::vector<FieldT> *variable_assignment /* null: the caller's products run on the host */)
{
    using namespace libiop_b200;
    const field_subset_type type = variable_domain.type();
    if (codeword_domain.type() != type || input_variable_domain.type() != type) return false;
    if ((type == affine_subspace_type) != (field_traits<FieldT>::kind == B200_FIELD_GF192)) return false;
    const std::size_t n_V = variable_domain.num_elements(), n_L = codeword_domain.num_elements();
    const std::size_t n_z = 1 + primary_input.size() + auxiliary_input.size();
    if (n_V < 4096 || n_z > n_V || n_V >= (std::size_t(1) << 32) || primary_input.size() + 1 != input_variable_domain.num_elements()) return false;

    /* z on the device */
    device_buffer d_z(n_z);
    std::vector<FieldT> head({ FieldT::one() });
    head.insert(head.end(), primary_input.begin(), primary_input.end());
    d_z.upload_async(head);
    d_z.upload_async(auxiliary_input, head.size());
    throw_status(context(), b200_wait_uploads(context()));                /* `head` dies with this scope */

    /* \hat f_w over V: variable c sits at position c of a subspace, at reindex_by_subset(dim(in), c) of a coset */
    device_buffer d_f1v(n_V), d_fw(n_V);
    { const device_view c(f_1v_coefficients); FFT_over_field_subset_device<FieldT>(c.buf(), f_1v_coefficients.size(), variable_domain, d_f1v); }
    if (type == affine_subspace_type) {
        vec_mul_add<FieldT>(d_fw, d_z, nullptr, &d_f1v, nullptr, true, n_z);
        if (n_z < n_V) throw_status(context(), b200_fill_zero(context(), d_fw.get(), n_z, n_V - n_z));
    } else {
        const b200_bytes position = reindex_table<FieldT>(variable_domain, input_variable_domain.dimension(), n_z);
        device_buffer d_at(n_z);
        throw_status(context(), b200_gather_dev(context(), d_at.get(), d_f1v.get(), position, 0, n_z));
        vec_mul_add<FieldT>(d_at, d_z, nullptr, &d_at, nullptr, true, n_z);
        throw_status(context(), b200_fill_zero(context(), d_fw.get(), 0, n_V));
        throw_status(context(), b200_scatter_dev(context(), d_fw.get(), d_at.get(), position, n_z));
    }

    /* over L */
    device_buffer d_L(n_L);
    if (!LDE_over_field_subsets_device<FieldT>(d_fw, variable_domain, codeword_domain, d_L)) {
        IFFT_over_field_subset_device<FieldT>(d_fw, variable_domain, d_f1v);
        FFT_over_field_subset_device<FieldT>(d_f1v, n_V, codeword_domain, d_L);
    }
    if (fw_mask) {
        device_buffer d_mask(n_L), d_Zv(n_L);
        const device_view c(fw_mask->coefficients());
        FFT_over_field_subset_device<FieldT>(c.buf(), fw_mask->coefficients().size(), codeword_domain, d_mask);
        vanishing_polynomial_over_domain<FieldT>(d_Zv, vanishing_polynomial<FieldT>(variable_domain), codeword_domain);
        vec_mul_add<FieldT>(d_L, d_L, nullptr, &d_Zv, &d_mask, false, n_L);
    }

    /* / Z_in: the table of Z_in over L collapses — along the basis directions Z_in kills (subspaces), or to its period
       |L| / deg Z_in in the position (cosets: Z_in = x^deg - c and x_i^deg = s^deg (g^deg)^i) */
    const vanishing_polynomial<FieldT> input_vp(input_variable_domain);
    std::size_t n_q = 0;
    uint32_t shift_bits = 0;
    device_buffer d_q, d_qinv;
    if (type == affine_subspace_type) {
        std::vector<FieldT> t = affine_function_table<FieldT>(codeword_domain, [&input_vp](const FieldT &x) { return input_vp.evaluation_at_point(x); },
                                                              FieldT::one());
        std::size_t lead = 0;
        while (lead < codeword_domain.dimension() && t[1 + lead].is_zero()) ++lead;
        t.erase(t.begin() + 1, t.begin() + 1 + lead);
        const std::size_t q_dim = codeword_domain.dimension() - lead;
        n_q = std::size_t(1) << q_dim; shift_bits = (uint32_t)lead;
        d_q = device_buffer(n_q); d_qinv = device_buffer(n_q);
        subspace_function_product<FieldT>(d_q, q_dim, { t });
    } else {
        const std::size_t deg = input_vp.degree();
        if (deg == 0 || n_L % deg) return false;
        n_q = n_L / deg;
        d_q = device_buffer(n_q); d_qinv = device_buffer(n_q);
        const FieldT a = field_pow(codeword_domain.shift(), deg);
        geometric_sequence<FieldT>(d_q, n_q, a, field_pow(codeword_domain.generator(), deg), input_vp.evaluation_at_point(codeword_domain.shift()) - a);
    }
    vec_inverse<FieldT>(d_qinv, d_q, n_q);
    throw_status(context(), b200_vec_mul_table(context(), field_traits<FieldT>::kind, field_traits<FieldT>::fp(), d_L.get(), d_L.get(),
                                               d_qinv.get(), shift_bits, n_q - 1, n_L));
    /* no synchronisation: the temporaries go back to the context's stream-ordered block cache */

    fw_over_codeword_domain = make_lazy_vector<FieldT>(std::move(d_L), n_L);
    if (variable_assignment) *variable_assignment = make_lazy_vector<FieldT>(std::move(d_z), n_z);
    return true;
}
} // namespace b200_glue

/* encoded_aurora_protocol::submit_witness_oracles (r1cs_rs_iop.tcc:481-614) with the witness codeword formed on the device
   (above) and the variable assignment handed to create_Az_Bz_Cz_from_variable_assignment as the device-resident vector it
   already is.  Domains that do not qualify run the reference's sequence of calls.  The order of the random draws
   (R_Az, R_Bz, R_Cz, then the f_w mask) is the reference's. */
#define B200_SUBMIT_WITNESS_SPECIALISE(FIELD)                                                                              \
template<> inline void encoded_aurora_protocol<FIELD>::submit_witness_oracles(                                             \
    const r1cs_primary_input<FIELD> &primary_input, const r1cs_auxiliary_input<FIELD> &auxiliary_input)                    \
{                                                                                                                          \
    this->fz_oracle_->set_primary_input(primary_input);                                                                    \
    libff::enter_block("Submit witness oracles");                                                                          \
    if (this->params_.holographic()) this->holographic_multi_lincheck_->submit_sumcheck_masking_polynomials();             \
    else this->multi_lincheck_->submit_sumcheck_masking_polynomials();                                                     \
    libff::enter_block("Compute randomized f_w");                                                                          \
    if (this->params_.make_zk()) {                                                                                         \
        this->R_Az_ = polynomial<FIELD>::random_polynomial(this->params_.query_bound());                                   \
        this->R_Bz_ = polynomial<FIELD>::random_polynomial(this->params_.query_bound());                                   \
        this->R_Cz_ = polynomial<FIELD>::random_polynomial(this->params_.query_bound());                                   \
    }                                                                                                                      \
    std::vector<FIELD> f_1v_evaluations({ FIELD::one() });                                                                 \
    f_1v_evaluations.insert(f_1v_evaluations.end(), primary_input.begin(), primary_input.end());                           \
    this->f_1v_coefficients_ = IFFT_over_field_subset<FIELD>(f_1v_evaluations, this->input_variable_domain_);              \
    if (this->params_.make_zk()) this->fw_mask_ = polynomial<FIELD>::random_polynomial(this->fw_mask_degree_);             \
    std::vector<FIELD> variable_assignment;                                                                                \
    /* z stays on the device only if the products with A, B, C run there (a host loop would fault it back in) */           \
    const bool device_products = b200_glue::constraint_matrices<FIELD>::instance().serves(this->constraint_system_.get()); \
    const bool witness_on_device = b200_glue::witness_codeword_on_device<FIELD>(primary_input, auxiliary_input,            \
            this->f_1v_coefficients_, this->input_variable_domain_, this->variable_domain_, this->codeword_domain_,        \
            this->params_.make_zk() ? &this->fw_mask_ : nullptr, this->fw_over_codeword_domain_,                           \
            device_products ? &variable_assignment : nullptr);                                                             \
    if (!witness_on_device) {                                                                                              \
        /* the reference's route: interpolate, add the mask, divide the coefficients, evaluate */                          \
        const std::vector<FIELD> f_1v_over_variable_domain = FFT_over_field_subset<FIELD>(this->f_1v_coefficients_, this->variable_domain_); \
        polynomial<FIELD> fw_prime(IFFT_over_field_subset<FIELD>(                                                          \
            create_fw_prime_evals(f_1v_over_variable_domain, primary_input.size(), auxiliary_input, this->variable_domain_), \
            this->variable_domain_));                                                                                      \
        if (this->params_.make_zk()) fw_prime += vanishing_polynomial<FIELD>(this->variable_domain_) * this->fw_mask_;     \
        fw_prime = polynomial_over_vanishing_polynomial(fw_prime, vanishing_polynomial<FIELD>(this->input_variable_domain_)).first; \
        this->fw_over_codeword_domain_ = FFT_over_field_subset<FIELD>(fw_prime.coefficients(), this->codeword_domain_);    \
    }                                                                                                                      \
    if (!witness_on_device || !device_products) {                                                                          \
        variable_assignment.assign(1, FIELD::one());                                                                       \
        variable_assignment.insert(variable_assignment.end(), primary_input.begin(), primary_input.end());                 \
        variable_assignment.insert(variable_assignment.end(), auxiliary_input.begin(), auxiliary_input.end());             \
    }                                                                                                                      \
    libff::leave_block("Compute randomized f_w");                                                                          \
    libff::enter_block("Compute A/B/Cz");                                                                                  \
    std::vector<FIELD> *const extended[3] = { &this->fprime_Az_over_codeword_domain_, &this->fprime_Bz_over_codeword_domain_, \
                                              &this->fprime_Cz_over_codeword_domain_ };                                    \
    const bool pipelined = !this->params_.make_zk() && b200_glue::extend_constraint_products<FIELD>(                       \
        *this->constraint_system_, variable_assignment, this->constraint_domain_, this->codeword_domain_, extended);       \
    std::vector<FIELD> Az, Bz, Cz;                                                                                         \
    if (!pipelined) this->constraint_system_->create_Az_Bz_Cz_from_variable_assignment(variable_assignment, Az, Bz, Cz);   \
    libff::leave_block("Compute A/B/Cz");                                                                                  \
    libff::enter_block("Compute f_{A/B/Cz} over codeword domain");                                                         \
    if (!pipelined) this->compute_fprime_ABCz_over_codeword_domain(Az, Bz, Cz);                                            \
    libff::leave_block("Compute f_{A/B/Cz} over codeword domain");                                                         \
    libff::enter_block("Call IOP oracle submission routines");                                                             \
    this->IOP_.submit_oracle(this->fw_handle_, std::move(this->fw_over_codeword_domain_));                                 \
    this->IOP_.submit_oracle(this->fAz_handle_, std::move(this->fprime_Az_over_codeword_domain_));                         \
    this->IOP_.submit_oracle(this->fBz_handle_, std::move(this->fprime_Bz_over_codeword_domain_));                         \
    this->IOP_.submit_oracle(this->fCz_handle_, std::move(this->fprime_Cz_over_codeword_domain_));                         \
    std::vector<FIELD>().swap(this->fw_over_codeword_domain_);                                                             \
    std::vector<FIELD>().swap(this->fprime_Az_over_codeword_domain_);                                                      \
    std::vector<FIELD>().swap(this->fprime_Bz_over_codeword_domain_);                                                      \
    std::vector<FIELD>().swap(this->fprime_Cz_over_codeword_domain_);                                                      \
    libff::leave_block("Call IOP oracle submission routines");                                                             \
    libff::leave_block("Submit witness oracles");                                                                          \
}
B200_FOR_EACH_FIELD(B200_SUBMIT_WITNESS_SPECIALISE)
#undef B200_SUBMIT_WITNESS_SPECIALISE

#define B200_FZ_SPECIALISE(FIELD)                                                                                          \
template<> inline std::shared_ptr<std::vector<FIELD>> fz_virtual_oracle<FIELD>::evaluated_contents(                        \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    libff::enter_block("fz evaluated contents");                                                                           \
    if (constituent_oracle_evaluations.size() != 1)                                                                        \
        throw std::invalid_argument("fz_virtual_oracle has one constituent oracle.");                                      \
    if (this->primary_input_.size() != this->primary_input_size_)                                                          \
        throw std::logic_error("Evaluation requested before primary_input is set.");                                       \
    const std::shared_ptr<std::vector<FIELD>> &fw = constituent_oracle_evaluations[0];                                     \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    if (fw->size() != n)                                                                                                   \
        throw std::invalid_argument("Provided fw evaluations don't match the declared codeword domain size.");             \
    using namespace libiop_b200;                                                                                           \
    device_view d_fw = b200_glue::to_device(*fw); device_buffer d_Z(n), d_f1v(n), d_out(n);                                            \
    const vanishing_polynomial<FIELD> input_vp(this->input_variable_domain_);                                              \
    vanishing_polynomial_over_domain<FIELD>(d_Z, input_vp, this->codeword_domain_);                                        \
    std::vector<FIELD> f_1v_evaluations({ FIELD::one() });                                                                 \
    f_1v_evaluations.insert(f_1v_evaluations.end(), this->primary_input_.begin(), this->primary_input_.end());             \
    const std::vector<FIELD> f_1v_coefficients = IFFT_over_field_subset<FIELD>(f_1v_evaluations, this->input_variable_domain_); \
    device_view d_c = b200_glue::to_device(f_1v_coefficients);                                                             \
    FFT_over_field_subset_device<FIELD>(d_c.buf(), f_1v_coefficients.size(), this->codeword_domain_, d_f1v);                     \
    vec_mul_add<FIELD>(d_out, d_fw.buf(), &d_Z, &d_f1v, nullptr, false, n);                                                      \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    libff::leave_block("fz evaluated contents");                                                                           \
    return result;                                                                                                         \
}
B200_FOR_EACH_FIELD(B200_FZ_SPECIALISE)
#undef B200_FZ_SPECIALISE
} // namespace libiop
