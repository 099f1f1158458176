/* Shim for <sodium/randombytes.h>: the reference uses exactly one symbol. */
#ifndef B200_SHIM_RANDOMBYTES_H
#define B200_SHIM_RANDOMBYTES_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void randombytes_buf(void *buf, size_t size);
#ifdef __cplusplus
}
#endif
#endif
