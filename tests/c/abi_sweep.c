/* A real particle sweep on the GPU through include/aps_b200.h from plain C -- no Python, no torch:
 * what a cgo / ccall / JNI binding does. LG d=1, SMC + systematic (BASELINE configs[0] shape), then
 * a PGAS conditional sweep on the picked trajectory. Prints every number the Python side
 * (tests/test_c_abi_from_c.py) checks against the oracle.                                        */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "aps_b200.h"

#define N 4096
#define T 8

static void die(const char *what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, aps_last_error());
    exit(10 + rc);
}

static uint64_t fnv(const void *p, size_t n) { /* FNV-1a over the raw bytes */
    const unsigned char *b = (const unsigned char *)p;
    uint64_t h = 1469598103934665603ULL;
    size_t i;
    for (i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ULL;
    return h;
}

int main(void) {
    aps_config cfg;
    aps_handle *h = 0;
    double Y[T], le = 0.0, le2 = 0.0, traj[T];
    static int32_t anc[N];
    static double x[N], w[N];
    int64_t slot = -1;
    int rc, t;
    memset(&cfg, 0, sizeof(cfg));
    cfg.model.obs_kind = APS_OBS_LINEAR_GAUSS; /* test/linear-gaussian.jl:32-42 */
    cfg.model.d = 1;
    cfg.model.dy = 1;
    cfg.model.mu0[0] = 0.0;
    cfg.model.sigma0[0] = 1.0;
    cfg.model.A[0] = 0.5;
    cfg.model.b[0] = 0.2;
    cfg.model.q[0] = 0.1;
    cfg.model.H[0] = 1.0;
    cfg.model.r[0] = 0.1;
    for (t = 1; t < APS_MAX_D; ++t) cfg.model.q[t] = cfg.model.r[t] = 1.0;
    cfg.n_particles = N;
    cfg.n_steps = T;
    cfg.sampler = APS_PGAS;
    cfg.resampler = APS_RESAMPLE_SYSTEMATIC;
    cfg.ess_threshold = 1.0; /* PGAS(n), src/smc.jl:99 */
    cfg.keep_history = 1;
    cfg.world_size = 1;
    for (t = 0; t < T; ++t) Y[t] = 0.35 + 0.01 * t;
    if ((rc = aps_create(&cfg, &h))) die("aps_create", rc);
    if ((rc = aps_set_observations(h, Y, T, 1))) die("aps_set_observations", rc);
    if ((rc = aps_sweep(h, 1234, 0, &le))) die("aps_sweep", rc);
    if ((rc = aps_get_ancestors(h, T + 1, anc))) die("aps_get_ancestors", rc);
    if ((rc = aps_get_states(h, T, x))) die("aps_get_states", rc);
    if ((rc = aps_get_weights(h, w))) die("aps_get_weights", rc);
    if ((rc = aps_pick_trajectory(h, traj, &slot))) die("aps_pick_trajectory", rc);
    if ((rc = aps_sweep(h, 1235, APS_REF_ON_DEVICE, &le2))) die("aps_sweep (conditional)", rc);
    printf("logevidence %.17g\n", le);
    printf("anc_hash %" PRIu64 "\n", fnv(anc, sizeof(anc)));
    printf("x_hash %" PRIu64 "\n", fnv(x, sizeof(x)));
    printf("w_hash %" PRIu64 "\n", fnv(w, sizeof(w)));
    printf("slot %" PRId64 "\n", slot);
    printf("traj");
    for (t = 0; t < T; ++t) printf(" %.17g", traj[t]);
    printf("\nlogevidence2 %.17g\n", le2);
    aps_destroy(h);
    return 0;
}
