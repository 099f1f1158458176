/* Shadow of libiop/protocols/encoded/ligero/ligero.tcc for the drop-in build.
   interleaved_r1cs_protocol::submit_witness_oracles (ligero.tcc:217-333) Reed-Solomon encodes the witness and the three
   products A z, B z, C z one 2^11-point row at a time (IFFT over the systematic domain, FFT over the codeword domain:
   512 + 512 latency-bound calls at BASELINE config 4).  For GF(2^192) all rows go to the B200 as ONE interleaved batch
   (b200_gf192_ifft_batch + b200_gf192_fft_batch); the host part — building z and the sparse products — is unchanged
   in meaning.  Other fields keep the reference's code. */
#include "b200_dropin_common.hpp"
#include B200_REF_FILE(libiop/protocols/encoded/ligero/ligero.tcc)

namespace libiop {

template<> inline void interleaved_r1cs_protocol<libff::gf192>::submit_witness_oracles(
    const r1cs_primary_input<libff::gf192> &primary_input, const r1cs_auxiliary_input<libff::gf192> &auxiliary_input)
{
    typedef libff::gf192 F;
    libff::enter_block("Submit witness oracles");
    const std::size_t row_len = this->systematic_domain_size_;
    const std::size_t n_w = this->num_oracles_input_, n_abc = this->num_oracles_vectors_;

    libff::enter_block("Generate extended witness and auxiliary witness");
    /* z = (1, v, w, 0...) over the matrix width */
    std::vector<F> z;
    z.reserve(std::max(this->matrix_width_, 1 + primary_input.size() + auxiliary_input.size()));
    z.emplace_back(F::one());
    z.insert(z.end(), primary_input.begin(), primary_input.end());
    z.insert(z.end(), auxiliary_input.begin(), auxiliary_input.end());
    if (z.size() < this->matrix_width_) z.resize(this->matrix_width_, F::zero());
    /* every row that is encoded, row-major: [w rows | a rows | b rows | c rows]; the w rows are z with (1, v) zeroed */
    std::vector<F> rows((n_w + 3 * n_abc) * row_len, F::zero());
    const std::size_t public_len = 1 + primary_input.size();
    for (std::size_t i = public_len; i < public_len + auxiliary_input.size() && i < n_w * row_len; ++i) rows[i] = z[i];
    libff::leave_block("Generate extended witness and auxiliary witness");

    libff::enter_block("Perform matrix multiplications");
    const auto times_z = [&](const naive_sparse_matrix<F> &M, F *product) {
        const std::size_t height = std::min(this->matrix_height_, n_abc * row_len);
        for (std::size_t i = 0; i < height; ++i) {
            F sum = F::zero();
            for (const auto &term : M[i]) sum += term.second * z[term.first];
            product[i] = sum;
        }
    };
    times_z(this->A_matrix_, rows.data() + n_w * row_len);
    times_z(this->B_matrix_, rows.data() + (n_w + n_abc) * row_len);
    times_z(this->C_matrix_, rows.data() + (n_w + 2 * n_abc) * row_len);
    libff::leave_block("Perform matrix multiplications");

    libff::enter_block("Submit input oracles");
    std::vector<std::vector<F>> codewords =
        libiop_b200::encode_rows<F>(rows.data(), n_w + 3 * n_abc, this->systematic_domain_, this->codeword_domain_);
    for (std::size_t i = 0; i < n_w; ++i)
        this->IOP_.submit_oracle(this->w_vector_handles_[i], oracle<F>(std::move(codewords[i])));
    libff::leave_block("Submit input oracles");

    libff::enter_block("Submit vector oracles");
    for (std::size_t i = 0; i < n_abc; ++i) {
        this->IOP_.submit_oracle(this->a_vector_handles_[i], oracle<F>(std::move(codewords[n_w + i])));
        this->IOP_.submit_oracle(this->b_vector_handles_[i], oracle<F>(std::move(codewords[n_w + n_abc + i])));
        this->IOP_.submit_oracle(this->c_vector_handles_[i], oracle<F>(std::move(codewords[n_w + 2 * n_abc + i])));
    }
    libff::leave_block("Submit vector oracles");
    libff::leave_block("Submit witness oracles");
}

} // namespace libiop
