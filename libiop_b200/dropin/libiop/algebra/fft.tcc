/*
 * Shadow of libiop/algebra/fft.tcc for the drop-in build (reached through the fft.hpp shadow next to this file).
 * The reference's own file is included unmodified from where it lies, with its three dispatchers renamed to ref_cpu_*;
 * the dispatchers are then defined here under their own names — taking const references, see fft.hpp — and forward to the
 * B200 adapter for libff::gf192 / libff::edwards_Fr (every other field keeps the CPU path).
 */
#include "b200_ref_traits.hpp"
#include "libiop_b200/host/b200_hotpath.hpp"

#ifndef FFT_over_field_subset      /* normally still renamed by the fft.hpp shadow that is including us */
#define FFT_over_field_subset ref_cpu_FFT_over_field_subset
#define IFFT_over_field_subset ref_cpu_IFFT_over_field_subset
#define IFFT_of_known_degree_over_field_subset ref_cpu_IFFT_of_known_degree_over_field_subset
#endif
#include B200_REF_FILE(libiop/algebra/fft.tcc)
#undef FFT_over_field_subset
#undef IFFT_over_field_subset
#undef IFFT_of_known_degree_over_field_subset

namespace libiop {
namespace b200_glue {
template<typename FieldT> std::vector<FieldT> fft(const std::vector<FieldT> &c, const field_subset<FieldT> &d, std::true_type)
{ return libiop_b200::FFT_over_field_subset_cref<FieldT>(c, d); }
template<typename FieldT> std::vector<FieldT> fft(const std::vector<FieldT> &c, const field_subset<FieldT> &d, std::false_type)
{ return ref_cpu_FFT_over_field_subset<FieldT>(c, d); }
template<typename FieldT> std::vector<FieldT> ifft(const std::vector<FieldT> &e, const field_subset<FieldT> &d, std::true_type)
{ return libiop_b200::IFFT_over_field_subset_cref<FieldT>(e, d); }
template<typename FieldT> std::vector<FieldT> ifft(const std::vector<FieldT> &e, const field_subset<FieldT> &d, std::false_type)
{ return ref_cpu_IFFT_over_field_subset<FieldT>(e, d); }
template<typename FieldT> std::vector<FieldT> ifft_kd(const std::vector<FieldT> &e, size_t deg, const field_subset<FieldT> &d, std::true_type)
{ return libiop_b200::IFFT_of_known_degree_over_field_subset_cref<FieldT>(e, deg, d); }
template<typename FieldT> std::vector<FieldT> ifft_kd(const std::vector<FieldT> &e, size_t deg, const field_subset<FieldT> &d, std::false_type)
{ return ref_cpu_IFFT_of_known_degree_over_field_subset<FieldT>(e, deg, d); }
}

template<typename FieldT>
std::vector<FieldT> FFT_over_field_subset(const std::vector<FieldT> &coeffs, const field_subset<FieldT> &domain)
{ return b200_glue::fft<FieldT>(coeffs, domain, libiop_b200::b200_enabled<FieldT>()); }
template<typename FieldT>
std::vector<FieldT> IFFT_over_field_subset(const std::vector<FieldT> &evals, const field_subset<FieldT> &domain)
{ return b200_glue::ifft<FieldT>(evals, domain, libiop_b200::b200_enabled<FieldT>()); }
template<typename FieldT>
std::vector<FieldT> IFFT_of_known_degree_over_field_subset(const std::vector<FieldT> &evals, size_t degree, const field_subset<FieldT> &domain)
{ return b200_glue::ifft_kd<FieldT>(evals, degree, domain, libiop_b200::b200_enabled<FieldT>()); }
} // namespace libiop
