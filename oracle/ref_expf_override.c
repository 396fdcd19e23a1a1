/* ref_expf_override.c -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
 *
 * Strong definition of interpolateTable() that replaces the reference's own
 * (src/solver.c:1441-1464, weakened with objcopy in oracle/Makefile) in the
 * "exact exponential" oracle build.  Same out-of-range rule as the reference
 * (x > maxVal -> 1), but 1-exp(-x) from libm instead of the (wrong-signed,
 * SURVEY F2) linear table.  It is what the SFU exponential mode of the CUDA
 * kernels is compared with.
 */
#include "SimpleMOC_header.h"

float interpolateTable(Table table, float x)
{
    if (x > table.maxVal)
        return 1.0f;
    return 1.0f - expf(-x);
}
