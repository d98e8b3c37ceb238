/* The drop-in boundary from plain C: include/aps_b200.h must be a valid C header (Julia's ccall,
 * cgo-style bindings and C callers see exactly this), and libaps_b200.so must link and answer
 * argument errors without a GPU. Compiled and run by tests/test_c_abi_from_c.py. */
#include <stdio.h>
#include <string.h>
#include "aps_b200.h"

int main(void) {
    aps_config cfg;
    aps_handle *h = 0;
    double out = 0.0;
    int rc;
    memset(&cfg, 0, sizeof(cfg));
    rc = aps_create(&cfg, &h); /* n_particles = 0: rejected before any CUDA call */
    if (rc != APS_ERR_INVALID || h != 0) return 1;
    if (strstr(aps_last_error(), "n_particles") == 0) return 2;
    if (aps_logsumexp(0, 3, &out) != APS_ERR_INVALID) return 3;          /* null vector */
    if (aps_resample(APS_RESAMPLE_SYSTEMATIC, &out, 0, 1, 0, 0, (int64_t *)&out) != APS_ERR_INVALID) return 4; /* empty weights */
    if (strstr(aps_last_error(), "empty") == 0) return 5;                 /* src/resampling.jl:103,154 */
    printf("%s sizeof(aps_config)=%u sizeof(aps_model)=%u\n", aps_version(), (unsigned)sizeof(aps_config),
           (unsigned)sizeof(aps_model));
    return 0;
}
