/* Shadow of libiop/protocols/encoded/ligero/interleaved_rowcheck.tcc for the drop-in build.
   interleaved_rowcheck_protocol::calculate_and_submit_responses (interleaved_rowcheck.tcc:151-206) sums
   r_i * (x_i * y_i - z_i) over the oracle rows on the host, one codeword position at a time.  For GF(2^192) the rows are
   read from their resident device copies: one b200_vec_mul_add per row for x_i * y_i - z_i, one b200_vec_lincomb for the
   weighted sum, then the reference's IFFT over the codeword domain.  Other fields keep the reference's code. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/ligero/interleaved_rowcheck.tcc)

namespace libiop {

template<> inline void interleaved_rowcheck_protocol<libff::gf192>::calculate_and_submit_responses()
{
    typedef libff::gf192 F;
    using libiop_b200::device_buffer;
    using libiop_b200::device_view;
    b200_ctx *ctx = libiop_b200::context();
    const std::size_t code_len = this->codeword_domain_size_, n_rows = this->num_oracles_;

    /* x_i * y_i - z_i for every row; the same for every interaction */
    device_buffer products(std::max<std::size_t>(n_rows, 1) * code_len);
    std::vector<b200_buf> row_ptrs(n_rows);
    for (std::size_t i = 0; i < n_rows; ++i) {
        const std::shared_ptr<std::vector<F>> x = this->IOP_.get_oracle_evaluations(this->x_vector_row_oracle_handles_[i]);
        const std::shared_ptr<std::vector<F>> y = this->IOP_.get_oracle_evaluations(this->y_vector_row_oracle_handles_[i]);
        const std::shared_ptr<std::vector<F>> z = this->IOP_.get_oracle_evaluations(this->z_vector_row_oracle_handles_[i]);
        if (x->size() != code_len || y->size() != code_len || z->size() != code_len)
            throw std::invalid_argument("interleaved rowcheck: oracle of the wrong size");
        const device_view dx(*x), dy(*y), dz(*z);
        row_ptrs[i] = products.get() + i * code_len * sizeof(F);
        libiop_b200::throw_status(ctx, b200_vec_mul_add(ctx, B200_FIELD_GF192, nullptr, row_ptrs[i], dx.get(), dy.get(), dz.get(), 0, 1, code_len));
        libiop_b200::throw_status(ctx, b200_sync(ctx));       /* x, y, z may be temporaries of virtual oracles */
    }

    for (std::size_t h = 0; h < this->num_interactions_; ++h) {
        const std::vector<F> r = this->IOP_.obtain_verifier_random_message(this->random_linear_combination_handles_[h]);
        if (r.size() < n_rows) throw std::invalid_argument("interleaved rowcheck: random message shorter than the number of rows");
        device_buffer response(code_len), coefficients(code_len);
        libiop_b200::throw_status(ctx, b200_vec_lincomb(ctx, B200_FIELD_GF192, nullptr, response.get(), code_len, n_rows,
                                                        reinterpret_cast<const uint64_t *>(r.data()), row_ptrs.data()));
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

} // namespace libiop
