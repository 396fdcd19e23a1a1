/* moc_rng.h -- the counter-based random stream that replaces libc rand() on
 * the SimpleMOC hot path.
 *
 * Why it exists.  The reference draws every synthetic input and -- on the hot
 * path -- the source-region id of every 3D segment from the global, sequential
 * libc stream (reference src/solver.c:476-483, src/utils.c:4-7), seeded from
 * time(NULL) (src/main.c:20).  A sequential stream cannot be evaluated by
 * 10^5 CUDA threads at once, and a time() seed cannot be reproduced, so the
 * contract of this repo is:
 *
 *     the c-th call of rand() since the start of the run returns
 *     moc_rand31(seed, c)                       (c = 0, 1, 2, ...)
 *
 * The reference, built unmodified, obeys the same contract when it is linked
 * against oracle/ref_shim.c (which defines rand()/srand()/time()), so the
 * serial CPU code and the parallel GPU code see identical draws.
 *
 * Usable from C99, C++ and CUDA device code.
 */
#ifndef MOC_RNG_H
#define MOC_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MOC_HD __host__ __device__ __forceinline__
#else
#define MOC_HD static inline
#endif

/* glibc's RAND_MAX; (float)MOC_RAND_MAX == 2147483648.0f exactly. */
#define MOC_RAND_MAX 2147483647

/* splitmix64 finaliser over (seed, counter); top 31 bits -> [0, RAND_MAX]. */
MOC_HD uint32_t moc_rand31(uint64_t seed, uint64_t counter)
{
    uint64_t z = seed + (counter + 1ULL) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (uint32_t)(z >> 33);
}

/* urand() of the reference (src/utils.c:4-7): (float)rand() / (float)RAND_MAX */
MOC_HD float moc_urand(uint64_t seed, uint64_t counter)
{
    return (float)moc_rand31(seed, counter) / 2147483648.0f;
}

#endif /* MOC_RNG_H */
