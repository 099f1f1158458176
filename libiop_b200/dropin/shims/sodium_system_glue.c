/*
 * sodium_system_glue.c — built instead of sodium_shim.c when the drop-in links the real libsodium (make SODIUM=system).
 * libsodium then provides crypto_generichash_blake2b and an unpredictable randombytes_buf; this file only adds the parity
 * harness's switch: b200_shim_randombytes_deterministic(seed) installs the seeded xorshift64 stream (the one the CPU
 * reference is run with, oracle/ref_shim/sodium_impl.c) through libsodium's randombytes_set_implementation,
 * b200_shim_randombytes_secure() re-installs libsodium's default generator.
 * libsodium is absent from the build image, so the few declarations needed are restated here (randombytes.h, stable ABI
 * since 1.0); this file is compile-checked but has not been linked against a libsodium in this repository's CI.
 */
#include <stddef.h>
#include <stdint.h>

typedef struct randombytes_implementation {
    const char *(*implementation_name)(void);
    uint32_t (*random)(void);
    void (*stir)(void);
    uint32_t (*uniform)(const uint32_t upper_bound);
    void (*buf)(void *const buf, const size_t size);
    int (*close)(void);
} randombytes_implementation;
extern int randombytes_set_implementation(const randombytes_implementation *impl);
extern randombytes_implementation randombytes_sysrandom_implementation;

static uint64_t det_state = 0x9E3779B97F4A7C15ULL;
static int det_on = 0;

static const char *det_name(void) { return "b200-parity-xorshift64"; }
static void det_buf(void *const buf, const size_t size)
{
    unsigned char *p = (unsigned char *)buf;
    for (size_t i = 0; i < size; ++i) {
        det_state ^= det_state << 13;
        det_state ^= det_state >> 7;
        det_state ^= det_state << 17;
        p[i] = (unsigned char)(det_state >> 32);
    }
}
static uint32_t det_random(void) { uint32_t r; det_buf(&r, sizeof r); return r; }
static void det_stir(void) {}
static int det_close(void) { return 0; }
static randombytes_implementation det_impl = { det_name, det_random, det_stir, NULL, det_buf, det_close };

void b200_shim_randombytes_deterministic(uint64_t seed)
{
    det_state = seed ? seed : 0x9E3779B97F4A7C15ULL;
    if (!det_on) { randombytes_set_implementation(&det_impl); det_on = 1; }
}
void b200_shim_randombytes_secure(void)
{
    if (det_on) { randombytes_set_implementation(&randombytes_sysrandom_implementation); det_on = 0; }
}
int b200_shim_randombytes_is_deterministic(void) { return det_on; }
void b200_shim_randombytes_seed(uint64_t seed) { b200_shim_randombytes_deterministic(seed); }
