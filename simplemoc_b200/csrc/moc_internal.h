/* moc_internal.h -- shared between the host C part and the CUDA part of
 * libmoc_b200.so.  Not installed; the public interface is include/moc_b200.h. */
#ifndef MOC_INTERNAL_H
#define MOC_INTERNAL_H

#include <stddef.h>
#include <stdint.h>

#include "moc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* printf-style setter for the message returned by moc_last_error() (thread local) */
void moc_set_error(const char *fmt, ...);

/* Host slab allocation.  Pinned (cudaHostAlloc) when a CUDA device is usable so the
 * drop-in entry points can DMA directly; plain calloc otherwise (e.g. unit tests on a
 * CPU-only box).  Memory is zero-filled.  moc_host_free() accepts either kind. */
void *moc_host_alloc(size_t bytes);
void moc_host_free(void *p);

/* stream positions of each block of draws of the synthetic problem construction,
 * in the serial order of the reference */
typedef struct {
    uint64_t az_weight;    /* T2 draws                         tracks.c:11-12   */
    uint64_t n_segments;   /* 2*T2 draws (Box-Muller pairs)    tracks.c:29-33   */
    uint64_t seg_length;   /* S2 draws                         tracks.c:50-58   */
    uint64_t p_weight;     /* T3 draws                         tracks.c:117-148 */
    uint64_t scatter;      /* X*G*G draws                      source.c:45-48   */
    uint64_t xs;           /* X*G*3 draws                      source.c:83-86   */
    uint64_t fine_source;  /* N*fai*G draws                    source.c:157-160 */
    uint64_t sigT;         /* N*G draws                        source.c:170-172 */
    uint64_t regions;      /* 2N-1 draws (index, volume)       source.c:183-198 */
    uint64_t end;          /* first draw of the sweep                            */
} moc_draw_layout;

/* 2D tracks, polar angles, exponential table and the draw layout only (moc_host.c) */
int moc_build_tracks_2d(Input *in, uint64_t seed, Params *out, moc_draw_layout *layout);

#ifdef __cplusplus
}
#endif
#endif
