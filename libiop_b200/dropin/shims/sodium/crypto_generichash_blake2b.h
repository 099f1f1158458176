/* Shim for <sodium/crypto_generichash_blake2b.h>: one symbol. */
#ifndef B200_SHIM_GENERICHASH_H
#define B200_SHIM_GENERICHASH_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
int crypto_generichash_blake2b(unsigned char *out, size_t outlen,
                               const unsigned char *in, unsigned long long inlen,
                               const unsigned char *key, size_t keylen);
#ifdef __cplusplus
}
#endif
#endif
