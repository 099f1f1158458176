/* Shadow of libiop/protocols/encoded/ligero/interleaved_lincheck_ot.tcc for the drop-in build.
   interleaved_lincheck_ot_protocol::calculate_and_submit_responses (interleaved_lincheck_ot.tcc:194-346) encodes, per
   interaction, one row per target oracle (a slice of the verifier's random vector r) and one per input oracle (a slice of
   r^T M), multiplies each codeword with the oracle of the same row on the host and sums.  For GF(2^192) the rows are
   encoded as one batch on the B200 and the sum of products is one pass (b200_vec_rows_dot); r^T M stays a host loop over
   the sparse matrix.  Other fields keep the reference's code. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/ligero/interleaved_lincheck_ot.tcc)

namespace libiop {

template<> inline void interleaved_lincheck_ot_protocol<libff::gf192>::calculate_and_submit_responses(
    const std::vector<libff::gf192> &supplementary_input, const std::size_t supplementary_input_size,
    const std::vector<libff::gf192> &supplementary_target, const std::size_t supplementary_target_size,
    std::vector<std::vector<libff::gf192>> random_linear_combinations)
{
    typedef libff::gf192 F;
    using libiop_b200::device_buffer;
    using libiop_b200::device_view;
    b200_ctx *ctx = libiop_b200::context();
    const std::size_t row_len = this->systematic_domain_size_, code_len = this->codeword_domain_size_;
    const std::size_t n_target = this->num_oracles_target_, n_input = this->num_oracles_input_, n_rows = n_target + n_input;
    const auto row_of = [code_len](const device_buffer &b, std::size_t r) { return b.get() + r * code_len * sizeof(F); };

    /* codewords of the public parts (interleaved_lincheck_ot.tcc:210-247): only the leading rows are non-zero */
    libff::enter_block("Input consistency");
    const std::size_t n_public_input = std::min(n_input, (supplementary_input_size + row_len - 1) / row_len);
    const std::size_t n_public_target = std::min(n_target, (supplementary_target_size + row_len - 1) / row_len);
    if (supplementary_input.size() < n_public_input * row_len || supplementary_target.size() < n_public_target * row_len)
        throw std::invalid_argument("interleaved lincheck: supplementary vector shorter than its declared size");
    const device_buffer public_input =
        libiop_b200::encode_rows_device<F>(supplementary_input.data(), n_public_input, this->systematic_domain_, this->codeword_domain_);
    const device_buffer public_target =
        libiop_b200::encode_rows_device<F>(supplementary_target.data(), n_public_target, this->systematic_domain_, this->codeword_domain_);
    libff::leave_block("Input consistency");

    if (random_linear_combinations.size() == 0) {
        random_linear_combinations.resize(this->num_interactions_);
        for (std::size_t h = 0; h < this->num_interactions_; ++h)
            random_linear_combinations[h] = this->IOP_.obtain_verifier_random_message(this->random_linear_combination_handles_[h]);
    }

    /* the committed codewords, rows [targets | inputs], each plus its public part; the same for every interaction */
    device_buffer committed(n_rows * code_len);
    const auto place = [&](std::size_t r, const oracle_handle_ptr &handle, const device_buffer &pub, std::size_t n_pub, std::size_t k) {
        const std::shared_ptr<std::vector<F>> evaluations = this->IOP_.get_oracle_evaluations(handle);
        if (evaluations->size() != code_len) throw std::invalid_argument("interleaved lincheck: oracle of the wrong size");
        const device_view resident(*evaluations);
        libiop_b200::throw_status(ctx, b200_copy(ctx, committed.get(), r * code_len, resident.get(), 0, code_len));
        if (k < n_pub)
            libiop_b200::throw_status(ctx, b200_vec_mul_add(ctx, B200_FIELD_GF192, nullptr, row_of(committed, r), row_of(committed, r), 0,
                                                            row_of(pub, k), 0, 0, code_len));
        libiop_b200::throw_status(ctx, b200_sync(ctx));       /* `evaluations` may be a temporary of a virtual oracle */
    };
    for (std::size_t i = 0; i < n_target; ++i) place(i, this->target_vector_row_oracle_handles_[i], public_target, n_public_target, i);
    for (std::size_t i = 0; i < n_input; ++i) place(n_target + i, this->input_vector_row_oracle_handles_[i], public_input, n_public_input, i);

    for (std::size_t h = 0; h < this->num_interactions_; ++h) {
        const std::vector<F> &r = random_linear_combinations[h];
        /* rows to encode: the target rows are r itself, the input rows are r^T M (interleaved_lincheck_ot.tcc:266-281) */
        std::vector<F> combination(n_rows * row_len, F::zero());
        std::copy(r.begin(), r.begin() + std::min(r.size(), n_target * row_len), combination.begin());
        F *from_matrix = combination.data() + n_target * row_len;
        for (std::size_t j = 0; j < this->constraint_matrix_.size(); ++j)
            for (const auto &term : this->constraint_matrix_[j]) from_matrix[term.first] += r[j] * term.second;

        const device_buffer encoded =
            libiop_b200::encode_rows_device<F>(combination.data(), n_rows, this->systematic_domain_, this->codeword_domain_);
        /* sum over targets of r_i f_i  -  sum over inputs of s_i f_i  (+ the blinding row) */
        device_buffer response(code_len), coefficients(code_len);
        libiop_b200::throw_status(ctx, b200_vec_rows_dot(ctx, B200_FIELD_GF192, nullptr, response.get(), encoded.get(), committed.get(),
                                                         code_len, n_rows, n_target));
        if (this->make_zk_) {
            const std::shared_ptr<std::vector<F>> blinding = this->IOP_.get_oracle_evaluations(this->blinding_vector_row_oracle_handles_[h]);
            const device_view resident(*blinding);
            libiop_b200::throw_status(ctx, b200_vec_mul_add(ctx, B200_FIELD_GF192, nullptr, response.get(), response.get(), 0,
                                                            resident.get(), 0, 0, code_len));
            libiop_b200::throw_status(ctx, b200_sync(ctx));
        }
        libiop_b200::IFFT_over_field_subset_device<F>(response, this->codeword_domain_, coefficients);
        this->IOP_.submit_prover_message(this->response_handles_[h], coefficients.download<F>(this->response_size_));
    }
}

/* The constructor (interleaved_lincheck_ot.tcc:12-44) receives its constraint matrix — a std::vector<std::map> with one heap
   node per non-zero entry — by value and copies it once more into the member; the specialisations move the parameter in
   (2^18 map nodes less to build and to destroy per lincheck and proof at N = 2^18). */
#define B200_LINCHECK_OT_CTOR(FIELD)                                                                                       \
template<> inline interleaved_lincheck_ot_protocol<FIELD>::interleaved_lincheck_ot_protocol(                               \
    iop_protocol<FIELD> &IOP, const domain_handle &codeword_domain_handle, const domain_handle &systematic_domain_handle,  \
    const domain_handle &extended_systematic_domain_handle, const std::size_t num_oracles_input,                           \
    const std::size_t num_oracles_target, const std::size_t num_queries, const std::size_t num_interactions, const bool make_zk, \
    const field_subset_type domain_type, naive_sparse_matrix<FIELD> constraint_matrix) :                                   \
    IOP_(IOP), codeword_domain_handle_(codeword_domain_handle), systematic_domain_handle_(systematic_domain_handle),       \
    num_oracles_input_(num_oracles_input), num_oracles_target_(num_oracles_target), num_queries_(num_queries),             \
    num_interactions_(num_interactions), make_zk_(make_zk), field_subset_type_(domain_type),                               \
    constraint_matrix_(std::move(constraint_matrix))                                                                       \
{                                                                                                                          \
    this->codeword_domain_ = this->IOP_.get_domain(this->codeword_domain_handle_);                                         \
    this->codeword_domain_size_ = this->codeword_domain_.num_elements();                                                   \
    this->systematic_domain_ = this->IOP_.get_domain(this->systematic_domain_handle_);                                     \
    this->systematic_domain_size_ = this->systematic_domain_.num_elements();                                               \
    this->extended_systematic_domain_ = this->IOP_.get_domain(extended_systematic_domain_handle);                          \
    this->response_size_ = 2 * this->systematic_domain_size_;                                                              \
}
B200_FOR_EACH_FIELD(B200_LINCHECK_OT_CTOR)
#undef B200_LINCHECK_OT_CTOR
} // namespace libiop
