/* Shadow of libiop/algebra/fft.hpp for the drop-in build.
   The reference declares its three dispatchers with the vectors BY VALUE (fft.hpp:62-88), so every caller copies its
   argument before the call: 24 MiB per coefficient vector and 192 MiB for the codeword that batch_sumcheck hands to
   IFFT_of_known_degree_over_field_subset at N = 2^20 (sumcheck.tcc:352-355) — and a by-value copy READS the vector on the
   host, which a device-resident (lazy) codeword must not be (libiop_b200/host/b200_residency.hpp).
   Here the reference's header is included unmodified with the three names renamed (they remain available as the CPU
   implementations ref_cpu_*), and the dispatchers are declared — and defined in the fft.tcc shadow next to this file — under
   their own names taking const references.  Call sites are unchanged: FFT_over_field_subset<FieldT>(coefficients, domain). */
#ifndef B200_DROPIN_ALGEBRA_FFT_HPP_
#define B200_DROPIN_ALGEBRA_FFT_HPP_

#define FFT_over_field_subset ref_cpu_FFT_over_field_subset
#define IFFT_over_field_subset ref_cpu_IFFT_over_field_subset
#define IFFT_of_known_degree_over_field_subset ref_cpu_IFFT_of_known_degree_over_field_subset
#include B200_REF_FILE(libiop/algebra/fft.hpp)       /* ends by including "libiop/algebra/fft.tcc": the shadow, which also lifts the renames */
#undef FFT_over_field_subset
#undef IFFT_over_field_subset
#undef IFFT_of_known_degree_over_field_subset

#endif
