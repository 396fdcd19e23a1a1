/* dropin_glue.c -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
 *
 * Linked into oracle/_ref/SimpleMOC-dropin: the reference's OWN main.c (src/main.c:3-147), init.c, io.c,
 * tracks.c, source.c, utils.c -- all unmodified -- with solver.c and comms.c left out and libmoc_b200.so in
 * their place.  main.c cannot call moc_dropin_configure (it is not edited), so this file exports the hook
 * the library looks up at the first transport_sweep: the seed and position of the host's rand() stream,
 * which here is the pinned one of ref_shim.c.  SMOC_SEED selects the seed (default 1). */
#include <stdint.h>
#include <stdlib.h>

void ref_shim_reset(uint64_t seed);
uint64_t ref_shim_calls(void);
uint64_t ref_shim_seed(void);

void moc_host_rand_state(unsigned long long *seed, unsigned long long *calls)
{
    *seed = ref_shim_seed();
    *calls = ref_shim_calls();
}

__attribute__((constructor)) static void seed_from_environment(void)
{
    const char *s = getenv("SMOC_SEED");
    ref_shim_reset(s ? strtoull(s, NULL, 10) : 1);
}
