/* Shadow of libiop/protocols/encoded/sumcheck/sumcheck.tcc for the drop-in build.
   sumcheck_g_oracle::evaluated_contents (sumcheck.tcc:58-121):
     additive        g = f - ( eps^-1 mu * x^(|H|-1) + Z_H(x) * h )            x^(|H|-1) = x^|H| * x^-1
     multiplicative  g = ( f - |H|^-1 mu - Z_H(x) * h ) * x^-1
   On the B200: Z_H over L and mu' * x^|H| are F2-affine in the position (generated on the device from dim(L)+1 host
   values each), x^-1 is an element-wise inversion (additive) or a geometric sequence (multiplicative), and the
   expression is two or three fused passes.  The reference's zero handling (batch_inverse_and_mul with has_zeroes,
   algebra/utils.tcc:73-103: the inverse of 0 is taken as 0) is what the device inversion x^(2^192-2) gives. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/sumcheck/sumcheck.tcc)

namespace libiop {
#define B200_SUMCHECK_G_SPECIALISE(FIELD)                                                                                  \
template<> inline std::shared_ptr<std::vector<FIELD>> sumcheck_g_oracle<FIELD>::evaluated_contents(                        \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != 2)                                                                        \
        throw std::invalid_argument("sumcheck_g_oracle has two constituent oracles");                                      \
    libff::enter_block("Sumcheck: g evaluated contents");                                                                  \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    device_view d_f = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                          \
    device_view d_h = b200_glue::to_device(*constituent_oracle_evaluations[1]);                                          \
    device_buffer d_Z(n), d_t(n), d_u(n), d_out(n);                                                                        \
    vanishing_polynomial_over_domain<FIELD>(d_Z, this->Z_, this->codeword_domain_);                                        \
    if (this->field_subset_type_ == affine_subspace_type) {                                                                \
        /* x^-1 over L: the domain's table (kept on the device across proofs); d_u = eps^-1 mu * x^|H| */                  \
        const device_buffer &d_xinv = subspace_element_inverses<FIELD>(this->codeword_domain_);                            \
        domain_element_powers<FIELD>(d_u, this->codeword_domain_, this->summation_domain_.num_elements(),                  \
                                     this->eps_inv_times_claimed_sum_);                                                    \
        vec_mul_add<FIELD>(d_u, d_u, &d_xinv, &d_Z, &d_h.buf(), false, n);       /* mu' x^(|H|-1) + Z_H h */                     \
        vec_mul_add<FIELD>(d_out, d_f.buf(), nullptr, &d_u, nullptr, true, n);   /* f - (...) */                                 \
    } else {                                                                                                               \
        /* d_t = f - Z_H h; d_u = constant |H|^-1 mu ... folded: (f - Z_H h - c) * x^-1 */                                 \
        vec_mul_add<FIELD>(d_t, d_f.buf(), nullptr, &d_Z, &d_h.buf(), true, n);                                                        \
        geometric_sequence<FIELD>(d_u, n, this->codeword_domain_.shift().inverse(), this->codeword_domain_.generator().inverse(), \
                                  FIELD::zero());                                                                          \
        /* (t - c) * xinv = t*xinv - c*xinv: second term via the same sequence scaled by c */                              \
        device_buffer d_v(n);                                                                                              \
        geometric_sequence<FIELD>(d_v, n, this->codeword_domain_.shift().inverse() * this->order_H_inv_times_claimed_sum_, \
                                  this->codeword_domain_.generator().inverse(), FIELD::zero());                            \
        vec_mul_add<FIELD>(d_out, d_t, &d_u, &d_v, nullptr, true, n);                                                      \
    }                                                                                                                      \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    libff::leave_block("Sumcheck: g evaluated contents");                                                                  \
    return result;                                                                                                         \
}
B200_FOR_EACH_FIELD(B200_SUMCHECK_G_SPECIALISE)
#undef B200_SUMCHECK_G_SPECIALISE

namespace b200_glue {
/* floor(P / Z_H) with P's n_p coefficients resident on the device (polynomial_over_vanishing_polynomial,
   vanishing_polynomial.tcc:360-371, quotient only).  false = outside the range of the device recurrence (the conditions of
   the shadows of vanishing_polynomial.tcc / linearized_polynomial.tcc); nothing is computed then. */
template<typename FieldT>
bool quotient_by_vanishing_on_device(const libiop_b200::device_buffer &d_p, std::size_t n_p, const vanishing_polynomial<FieldT> &Z,
                                     libiop_b200::device_buffer &d_q)
{
    const std::size_t deg = Z.degree();
    if (n_p <= deg || n_p - deg < 4096) return false;
    std::vector<FieldT> coefs;
    std::vector<std::size_t> pows;
    FieldT lead = FieldT::one();
    if (Z.type() == affine_subspace_type) {
        const linearized_polynomial<FieldT> lin = Z.get_linearized_polynomial();
        if (lin.num_terms() < 2 || lin.degree() != deg || (n_p - deg) / (deg - deg / 2) > 64) return false;
        std::size_t p_pow = deg / 2;                       /* lin[j] multiplies y^(2^(j-1)) for j >= 1, lin[0] is the constant */
        for (std::size_t j = lin.num_terms() - 1; j--; ) { coefs.push_back(lin[j]); pows.push_back(p_pow); p_pow /= 2; }
        lead = lin[lin.num_terms() - 1];
    } else {
        if ((n_p - deg) / deg > 64) return false;
        coefs.push_back(Z.constant_coefficient()); pows.push_back(0);
    }
    d_q = libiop_b200::device_buffer(n_p - deg);
    libiop_b200::poly_divide_sparse_device<FieldT>(d_p, n_p, deg, lead, coefs, pows, d_q, nullptr);
    return true;
}
/* h = floor(f / Z_H) over the codeword domain L straight from f|L, subspace domains over GF(2^192), deg f < 2|H|, L's first
   dim(H) basis vectors spanning H's direction space.  With b = L.basis[dim H], C_0 = shift(L) + span and C_1 = C_0 + b (the
   first 2|H| positions of L), the fused extension's intermediate form splits as f = A + (Z_H / w) B, w = Z_H(b) - Z_H(0), with
   coeffs(B) = interpolate_{C_0}(f|C_0) + interpolate_{C_1}(f|C_1) (include/b200_iop.h) — so h = B / w and
       h|L = evaluate_L( (interpolate(f|C_0) + interpolate(f|C_1)) / w ):
   2 x 20 + 8 x 20 butterfly stages of 2^19 instead of the reference's IFFT(2^21) + division + FFT(2^23) (12.8 -> 5.7 ms at
   N = 2^20).  Exact arithmetic: the same field elements.  false = the domains do not qualify, nothing is computed. */
template<typename FieldT>
bool sumcheck_quotient_over_codeword_domain(const std::vector<FieldT> &f_over_L, std::size_t degree_bound, const field_subset<FieldT> &H,
                                            const field_subset<FieldT> &L, libiop_b200::device_buffer &h_over_L)
{
    using namespace libiop_b200;
    if (field_traits<FieldT>::kind != B200_FIELD_GF192 || H.type() != affine_subspace_type || L.type() != affine_subspace_type) return false;
    const std::size_t d = H.dimension(), n_H = H.num_elements(), n_L = L.num_elements();
    if (n_H < residency::min_elems || L.dimension() < d + 1 || degree_bound > 2 * n_H || degree_bound <= n_H || f_over_L.size() != n_L) return false;
    const std::vector<FieldT> &bh = H.basis(), &bl = L.basis();
    for (std::size_t k = 0; k < d; ++k) if (!(bh[k] == bl[k])) return false;
    const vanishing_polynomial<FieldT> Z(H);
    const FieldT w = Z.evaluation_at_point(bl[d]) - Z.evaluation_at_point(FieldT::zero());
    if (w.is_zero()) return false;
    const FieldT w_inv = w.inverse();
    /* the two cosets of span(H) the first 2|H| evaluations lie on */
    std::vector<FieldT> span(bl.begin(), bl.begin() + d);
    const field_subset<FieldT> C0(affine_subspace<FieldT>(span, L.shift())), C1(affine_subspace<FieldT>(span, L.shift() + bl[d]));
    const device_view f(f_over_L.data(), 2 * n_H);
    const device_buffer f0 = device_buffer::alias(f.get(), n_H), f1 = device_buffer::alias(f.get() + 24 * n_H, n_H);
    device_buffer x0(n_H), x1(n_H), B(n_H);
    std::vector<uint64_t> store0, store1, storeL;
    const b200_domain d0 = make_domain<FieldT>(C0, store0), d1 = make_domain<FieldT>(C1, store1), dL = make_domain<FieldT>(L, storeL);
    throw_status(context(), b200_gf192_lde_interpolate(context(), &d0, f0.get(), x0.get()));
    throw_status(context(), b200_gf192_lde_interpolate(context(), &d1, f1.get(), x1.get()));
    vec_lincomb<FieldT>(B, n_H, { w_inv, w_inv }, { &x0, &x1 });
    h_over_L = device_buffer(n_L);
    throw_status(context(), b200_gf192_lde_evaluate(context(), &dL, (uint32_t)d, B.get(), h_over_L.get()));
    return true;
}
} // namespace b200_glue

/* batch_sumcheck_protocol::calculate_and_submit_proof (sumcheck.tcc:344-388): interpolate the combined f from its codeword,
   divide by Z_H, submit h = the quotient over the codeword domain.  The reference moves a coefficient vector through three
   host objects on the way (the IFFT result, the dividend, a copy of the quotient); here the chain
       known-degree IFFT -> sparse division -> FFT
   runs from device buffer to device buffer and h is submitted as the device-resident codeword it already is: three 24-50 MB
   transfers each way and a 25 MB host copy less per proof at N = 2^20 (profiles/r2_host_gaps_N20_v1.txt vs _v2). */
#define B200_SUMCHECK_PROVE_SPECIALISE(FIELD)                                                                              \
template<> inline void batch_sumcheck_protocol<FIELD>::calculate_and_submit_proof()                                        \
{                                                                                                                          \
    const std::vector<FIELD> challenge = this->IOP_.obtain_verifier_random_message(this->challenge_handle_);               \
    this->combined_f_oracle_->set_random_coefficients(challenge);                                                          \
    const std::shared_ptr<std::vector<FIELD>> f_evaluations =                                                              \
        this->IOP_.get_oracle_evaluations(std::make_shared<virtual_oracle_handle>(this->combined_f_oracle_handle_));       \
    this->g_oracle_->set_claimed_sum(this->get_combined_claimed_sum(challenge));                                           \
    using namespace libiop_b200;                                                                                           \
    {                                                                                                                      \
        device_buffer h_over_L;                                                                                            \
        if (b200_glue::sumcheck_quotient_over_codeword_domain<FIELD>(*f_evaluations, this->degree_bound_, this->summation_domain_, \
                                                                     this->codeword_domain_, h_over_L)) {                  \
            const std::size_t n = this->codeword_domain_.num_elements();                                                   \
            this->IOP_.submit_oracle(this->h_handle_, oracle<FIELD>(make_lazy_vector<FIELD>(std::move(h_over_L), n)));     \
            return;                                                                                                        \
        }                                                                                                                  \
    }                                                                                                                      \
    const vanishing_polynomial<FIELD> Z(this->summation_domain_);                                                          \
    const device_buffer d_f = IFFT_of_known_degree_over_field_subset_to_device<FIELD>(*f_evaluations, this->degree_bound_, this->codeword_domain_); \
    device_buffer d_h;                                                                                                     \
    if (b200_glue::quotient_by_vanishing_on_device<FIELD>(d_f, this->degree_bound_, Z, d_h)) {                             \
        const std::size_t n = this->codeword_domain_.num_elements();                                                       \
        device_buffer d_out(n);                                                                                            \
        FFT_over_field_subset_device<FIELD>(d_h, this->degree_bound_ - Z.degree(), this->codeword_domain_, d_out);         \
        this->IOP_.submit_oracle(this->h_handle_, oracle<FIELD>(make_lazy_vector<FIELD>(std::move(d_out), n)));            \
        return;                                                                                                            \
    }                                                                                                                      \
    std::vector<FIELD> coefficients = d_f.template download<FIELD>(d_f.size());                                            \
    coefficients.resize(this->degree_bound_);                                                                              \
    const std::pair<polynomial<FIELD>, polynomial<FIELD>> h_and_g =                                                        \
        polynomial_over_vanishing_polynomial<FIELD>(polynomial<FIELD>(std::move(coefficients)), Z);                        \
    this->IOP_.submit_oracle(this->h_handle_, oracle<FIELD>(FFT_over_field_subset<FIELD>(h_and_g.first.coefficients(), this->codeword_domain_))); \
}
B200_FOR_EACH_FIELD(B200_SUMCHECK_PROVE_SPECIALISE)
#undef B200_SUMCHECK_PROVE_SPECIALISE
} // namespace libiop
