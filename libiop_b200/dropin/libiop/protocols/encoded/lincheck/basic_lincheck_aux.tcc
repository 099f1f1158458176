/* Shadow of libiop/protocols/encoded/lincheck/basic_lincheck_aux.tcc for the drop-in build.
   multi_lincheck_virtual_oracle::evaluated_contents (basic_lincheck_aux.tcc:102-144):
       result = (sum_m r_m * f_Mz_m) * p_alpha_prime|L  -  f_z * p_alpha_ABC|L.
   On the B200 the two LDEs stay on the device, the random linear combination and the final expression are two passes,
   and only the constituent oracles go up / the result comes down.
   multi_lincheck_virtual_oracle::evaluation_at_point (:147-181) evaluates both degree-|H| polynomials by Horner for every
   query position (2 x 2^20 multiplications per position at N = 2^20: 2 s of bcs_prover::get_transcript); here the
   coefficient vectors are cached on the device and evaluated by b200_poly_eval — and, when the point is an element of the
   codeword domain (every prover-side query), the value is simply read out of the low-degree extension that
   evaluated_contents left on the device. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/lincheck/basic_lincheck_aux.tcc)

namespace libiop {
namespace b200_glue {
template<typename FieldT> libiop_b200::device_vector_cache<FieldT> &polynomial_cache()
{
    static libiop_b200::device_vector_cache<FieldT> cache(4);
    return cache;
}
/* The low-degree extensions of p_alpha' and p_alpha_ABC over the codeword domain, left on the device by
   evaluated_contents.  The prover's query phase asks for both polynomials at ~100 points that all lie in the codeword
   domain (iop.tcc:670-724 passes the position along with the point): those values are already in these arrays, so a query
   costs one 24-byte gather instead of a degree-2^20 Horner evaluation.  An entry is tied to its coefficient vector by
   address, length and a content sample, and to the domain by value. */
template<typename FieldT>
class codeword_lde_cache {
    struct entry { const FieldT *ptr; std::size_t n; std::vector<FieldT> sample; field_subset<FieldT> domain; libiop_b200::device_buffer lde; };
    std::list<entry> entries_;
    static std::vector<FieldT> sample_of(const std::vector<FieldT> &v)
    {
        std::vector<FieldT> s;
        const std::size_t n = v.size(), k = n < 32 ? n : 32;
        for (std::size_t i = 0; i < k; ++i) s.push_back(v[(k > 1) ? (i * (n - 1)) / (k - 1) : 0]);
        return s;
    }
public:
    void put(const std::vector<FieldT> &coefficients, const field_subset<FieldT> &domain, libiop_b200::device_buffer &&lde)
    {
        for (auto it = entries_.begin(); it != entries_.end();)
            if (it->ptr == coefficients.data()) it = entries_.erase(it); else ++it;
        if (entries_.size() >= 4) entries_.pop_back();
        entries_.push_front(entry{coefficients.data(), coefficients.size(), sample_of(coefficients), domain, std::move(lde)});
    }
    const libiop_b200::device_buffer *find(const std::vector<FieldT> &coefficients, const field_subset<FieldT> &domain)
    {
        for (auto &e : entries_)
            if (e.ptr == coefficients.data() && e.n == coefficients.size() && e.sample == sample_of(coefficients) && e.domain == domain)
                return &e.lde;
        return nullptr;
    }
};
template<typename FieldT> codeword_lde_cache<FieldT> &lde_cache()
{
    static codeword_lde_cache<FieldT> cache;
    return cache;
}
/* value of the polynomial at `x`: from its resident low-degree extension when x is element `position` of the domain */
template<typename FieldT>
bool evaluate_from_lde(const std::vector<FieldT> &coefficients, const field_subset<FieldT> &domain, std::size_t position, const FieldT &x,
                       FieldT &value)
{
    const libiop_b200::device_buffer *lde = lde_cache<FieldT>().find(coefficients, domain);
    if (!lde || position >= domain.num_elements() || domain.element_by_index(position) != x) return false;
    libiop_b200::throw_status(libiop_b200::context(), b200_gather(libiop_b200::context(), reinterpret_cast<uint64_t *>(&value), lde->get(), &position, 1));
    return true;
}

template<typename FieldT> FieldT evaluate_polynomial(const std::vector<FieldT> &coefficients, const FieldT &x)
{
    if (coefficients.size() < 4096) {           /* not worth a launch: same Horner recurrence as polynomial.tcc:98-109 */
        FieldT r = FieldT::zero();
        for (auto it = coefficients.rbegin(); it != coefficients.rend(); ++it) { r *= x; r += *it; }
        return r;
    }
    return libiop_b200::poly_eval<FieldT>(polynomial_cache<FieldT>().get(coefficients), coefficients.size(), x);
}
}

#define B200_LINCHECK_SPECIALISE(FIELD)                                                                                    \
template<> inline std::shared_ptr<std::vector<FIELD>> multi_lincheck_virtual_oracle<FIELD>::evaluated_contents(            \
    const std::vector<std::shared_ptr<std::vector<FIELD>>> &constituent_oracle_evaluations) const                          \
{                                                                                                                          \
    libff::enter_block("multi_lincheck evaluated contents");                                                               \
    if (constituent_oracle_evaluations.size() != this->matrices_.size() + 1)                                               \
        throw std::invalid_argument("multi_lincheck uses more constituent oracles than what was provided.");               \
    using namespace libiop_b200;                                                                                           \
    const std::size_t n = this->codeword_domain_.num_elements();                                                           \
    device_buffer d_p1(n), d_p2(n), d_comb(n), d_out(n);                                                                   \
    {                                                                                                                      \
        device_view c1 = b200_glue::to_device(this->p_alpha_prime_.coefficients());                                            \
        FFT_over_field_subset_device<FIELD>(c1.buf(), this->p_alpha_prime_.coefficients().size(), this->codeword_domain_, d_p1); \
        device_view c2 = b200_glue::to_device(this->p_alpha_ABC_.coefficients());                                                \
        FFT_over_field_subset_device<FIELD>(c2.buf(), this->p_alpha_ABC_.coefficients().size(), this->codeword_domain_, d_p2);   \
        throw_status(context(), b200_sync(context()));                                                                     \
    }                                                                                                                      \
    device_view d_fz = b200_glue::to_device(*constituent_oracle_evaluations[0]);                                           \
    std::vector<device_view> d_Mz;                                                                                         \
    std::vector<const device_buffer *> Mz_ptrs;                                                                            \
    d_Mz.reserve(this->matrices_.size());                                                                                  \
    for (std::size_t m = 0; m < this->matrices_.size(); ++m) {                                                             \
        if (constituent_oracle_evaluations[m + 1]->size() != n) throw std::invalid_argument("Vectors of mismatched size."); \
        d_Mz.emplace_back(b200_glue::to_device(*constituent_oracle_evaluations[m + 1]));                                   \
    }                                                                                                                      \
    for (auto &b : d_Mz) Mz_ptrs.push_back(&b.buf());                                                                            \
    vec_lincomb<FIELD>(d_comb, n, this->r_Mz_, Mz_ptrs);                                                                   \
    vec_mul_add<FIELD>(d_out, d_comb, &d_p1, &d_fz.buf(), &d_p2, true, n);                                                       \
    std::shared_ptr<std::vector<FIELD>> result = b200_glue::to_host<FIELD>(d_out, n);                                      \
    b200_glue::lde_cache<FIELD>().put(this->p_alpha_prime_.coefficients(), this->codeword_domain_, std::move(d_p1));       \
    b200_glue::lde_cache<FIELD>().put(this->p_alpha_ABC_.coefficients(), this->codeword_domain_, std::move(d_p2));         \
    libff::leave_block("multi_lincheck evaluated contents");                                                               \
    return result;                                                                                                         \
}                                                                                                                          \
template<> inline FIELD multi_lincheck_virtual_oracle<FIELD>::evaluation_at_point(                                         \
    const std::size_t evaluation_position, const FIELD evaluation_point,                                                   \
    const std::vector<FIELD> &constituent_oracle_evaluations) const                                                        \
{                                                                                                                          \
    if (constituent_oracle_evaluations.size() != this->matrices_.size() + 1)                                               \
        throw std::invalid_argument("multi_lincheck uses more constituent oracles than what was provided.");               \
    if (this->use_lagrange_) throw std::logic_error("B200 drop-in: the Lagrange variant of multi_lincheck is not wired"); \
    FIELD p_alpha_prime_X, p_alpha_ABC_X;                                                                                  \
    if (!b200_glue::evaluate_from_lde<FIELD>(this->p_alpha_prime_.coefficients(), this->codeword_domain_, evaluation_position,  \
                                             evaluation_point, p_alpha_prime_X))                                           \
        p_alpha_prime_X = b200_glue::evaluate_polynomial<FIELD>(this->p_alpha_prime_.coefficients(), evaluation_point);    \
    if (!b200_glue::evaluate_from_lde<FIELD>(this->p_alpha_ABC_.coefficients(), this->codeword_domain_, evaluation_position,    \
                                             evaluation_point, p_alpha_ABC_X))                                             \
        p_alpha_ABC_X = b200_glue::evaluate_polynomial<FIELD>(this->p_alpha_ABC_.coefficients(), evaluation_point);        \
    FIELD f_combined_Mz_x = FIELD::zero();                                                                                 \
    for (std::size_t i = 0; i < this->r_Mz_.size(); i++) f_combined_Mz_x += this->r_Mz_[i] * constituent_oracle_evaluations[i + 1]; \
    return f_combined_Mz_x * p_alpha_prime_X - constituent_oracle_evaluations[0] * p_alpha_ABC_X;                          \
}
B200_FOR_EACH_FIELD(B200_LINCHECK_SPECIALISE)
#undef B200_LINCHECK_SPECIALISE
} // namespace libiop
