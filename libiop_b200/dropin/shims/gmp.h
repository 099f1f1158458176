/* Minimal <gmp.h> shim: only libgmp.so.10 (runtime) is installed in this image, not the header.
   Declares exactly the GMP entry points libff/libiop use; links against the
   real /usr/lib/x86_64-linux-gnu/libgmp.so.10. */
#ifndef B200_SHIM_GMP_H
#define B200_SHIM_GMP_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef long mp_size_t;
typedef unsigned long mp_bitcnt_t;
typedef mp_limb_t *mp_ptr;
typedef const mp_limb_t *mp_srcptr;
#define GMP_NUMB_BITS 64
#define GMP_LIMB_BITS 64
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#define mpn_add_n __gmpn_add_n
#define mpn_sub_n __gmpn_sub_n
#define mpn_add_1 __gmpn_add_1
#define mpn_sub_1 __gmpn_sub_1
#define mpn_add __gmpn_add
#define mpn_sub __gmpn_sub
#define mpn_addmul_1 __gmpn_addmul_1
#define mpn_mul_n __gmpn_mul_n
#define mpn_mul __gmpn_mul
#define mpn_cmp __gmpn_cmp
#define mpn_copyi __gmpn_copyi
#define mpn_zero __gmpn_zero
#define mpn_set_str __gmpn_set_str
#define mpn_gcdext __gmpn_gcdext
#define mpn_tdiv_qr __gmpn_tdiv_qr
#define mpn_rshift __gmpn_rshift
#define mpn_lshift __gmpn_lshift
#define mpn_divrem_1 __gmpn_divrem_1
#define mpn_get_str __gmpn_get_str
mp_limb_t mpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t mpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t mpn_add(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t mpn_sub(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t mpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
void mpn_mul_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t mpn_mul(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
int mpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
void mpn_copyi(mp_ptr, mp_srcptr, mp_size_t);
void mpn_zero(mp_ptr, mp_size_t);
mp_size_t mpn_set_str(mp_ptr, const unsigned char *, size_t, int);
size_t mpn_get_str(unsigned char *, int, mp_ptr, mp_size_t);
mp_size_t mpn_gcdext(mp_ptr, mp_ptr, mp_size_t *, mp_ptr, mp_size_t, mp_ptr, mp_size_t);
void mpn_tdiv_qr(mp_ptr, mp_ptr, mp_size_t, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t mpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t mpn_lshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t mpn_divrem_1(mp_ptr, mp_size_t, mp_srcptr, mp_size_t, mp_limb_t);

#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_init_set __gmpz_init_set
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_set_ui __gmpz_set_ui
#define mpz_set __gmpz_set
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_mod __gmpz_mod
#define mpz_get_ui __gmpz_get_ui
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_add_ui __gmpz_add_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_tstbit __gmpz_tstbit
#define mpz_cmp_ui __gmpz_cmp_ui
#define gmp_printf __gmp_printf
void mpz_init(mpz_ptr);
void mpz_clear(mpz_ptr);
void mpz_init_set(mpz_ptr, mpz_srcptr);
void mpz_init_set_ui(mpz_ptr, unsigned long);
void mpz_set_ui(mpz_ptr, unsigned long);
void mpz_set(mpz_ptr, mpz_srcptr);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long mpz_get_ui(mpz_srcptr);
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
size_t mpz_sizeinbase(mpz_srcptr, int);
int mpz_tstbit(mpz_srcptr, mp_bitcnt_t);
int mpz_cmp_ui(mpz_srcptr, unsigned long);
int gmp_printf(const char *, ...);
#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : (Z)->_mp_size > 0)
#ifdef __cplusplus
}
#endif
#endif
