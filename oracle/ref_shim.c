/* ref_shim.c -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
 *
 * Linked INTO oracle/_ref/libsimplemoc_ref*.so together with the unmodified
 * reference objects (-Wl,-Bsymbolic so the library binds to these definitions
 * and not to libc's).  It pins the two sources of run-to-run noise in the
 * reference without touching its sources:
 *
 *   rand()   reference src/solver.c:481, src/utils.c:6, src/source.c:188
 *            -> the c-th call returns moc_rand31(seed, c)   (include/moc_rng.h)
 *   srand()  reference src/main.c:20                         -> ignored
 *   time()   reference src/main.c:20, src/solver.c:304       -> constant
 *
 * The OpenMP timing build does NOT link this file (it keeps libc's rand_r).
 */
#include <stdint.h>
#include <time.h>
#include "moc_rng.h"

static uint64_t g_seed = 1;
static uint64_t g_calls = 0;

void ref_shim_reset(uint64_t seed) { g_seed = seed; g_calls = 0; }
uint64_t ref_shim_calls(void) { return g_calls; }
uint64_t ref_shim_seed(void) { return g_seed; }
void ref_shim_set_calls(uint64_t c) { g_calls = c; }

int rand(void) { return (int)moc_rand31(g_seed, g_calls++); }
void srand(unsigned int s) { (void)s; }
time_t time(time_t *t)
{
    if (t) *t = (time_t)1;
    return (time_t)1;
}
