/*
 * aps_b200.h -- C ABI of libaps_b200.so: the B200 (sm_100a) particle sweep behind the
 * AdvancedPS.jl sampler surface.
 *
 * The reference (TuringLang/AdvancedPS.jl v0.7.2) has no FFI; its plug-in points are Julia
 * dispatch. Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference). A Julia maintainer binds these with `ccall`
 * inside new methods of AbstractMCMC.sample / AbstractMCMC.step; the
 * Python host mirror in advancedps.jl_b200/ binds the same symbols with ctypes.
 *
 * Conventions
 *   - every function returns an aps_status; aps_last_error() gives the thread-local message
 *     (the Julia wrapper turns non-zero into `error(msg)` to keep ErrorException behaviour,
 *     src/resampling.jl:103,120,154,169; src/container.jl:292-298).
 *   - plain pointers and sizes only. Pointers documented "host or device" are classified with
 *     cudaPointerGetAttributes; everything else is a host pointer that is copied.
 *   - indices crossing the operator boundary are 1-based int64 (Julia Vector{Int}); genealogy
 *     accessors return 0-based int32 as stored on the device.
 *   - calls on one handle must be serialised by the caller; distinct handles are independent.
 */
#ifndef APS_B200_H
#define APS_B200_H

#include <stdint.h>
#include "aps_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum aps_status {
    APS_OK = 0,
    APS_ERR_INVALID = 1,        /* bad argument / empty weights  (src/resampling.jl:103,154)      */
    APS_ERR_WEIGHTS = 2,        /* weights not normalisable, NaN (src/resampling.jl:120,169)      */
    APS_ERR_CUDA = 3,
    APS_ERR_COMM = 4,
    APS_ERR_NOMEM = 5
} aps_status;

typedef enum aps_resampler {      /* src/resampling.jl */
    APS_RESAMPLE_MULTINOMIAL = 0, /* :31-35   */
    APS_RESAMPLE_RESIDUAL = 1,    /* :53-81   */
    APS_RESAMPLE_STRATIFIED = 2,  /* :98-131  */
    APS_RESAMPLE_SYSTEMATIC = 3   /* :149-183 (DEFAULT_RESAMPLER, :185) */
} aps_resampler;

typedef enum aps_sampler { /* src/smc.jl */
    APS_SMC = 0,           /* :1-21   */
    APS_PG = 1,            /* :59-81  */
    APS_PGAS = 2           /* :92-99  */
} aps_sampler;

/* Sampler + model configuration. Mirrors SMC(n, resampler[, threshold]) / PG(...) / PGAS(n)
 * (src/smc.jl:15-21,75-81,99) applied to TracedSSM(model, Y) (src/model.jl:13-22).           */
typedef struct aps_config {
    aps_model model;
    int64_t n_particles;   /* N >= 1 (N >= 2 for PG/PGAS with a reference)                      */
    int64_t n_steps;       /* T = length(Y)                                                     */
    int32_t sampler;       /* aps_sampler                                                       */
    int32_t resampler;     /* aps_resampler                                                     */
    double ess_threshold;  /* ResampleWithESSThreshold (src/resampling.jl:193-204); NaN = bare  */
                           /* resampler function, i.e. resample at every step                   */
    int32_t keep_history;  /* 1: keep all T state/ancestor slabs (needed for trajectories)      */
    int32_t device;        /* CUDA device ordinal                                               */
    int32_t rank;          /* multi-GPU: this process's rank; 0 for single GPU                  */
    int32_t world_size;    /* multi-GPU: number of ranks sharing the particle set; 1 = single   */
} aps_config;

typedef struct aps_handle aps_handle;

/* ---- lifetime: replaces building N Trace objects + ParticleContainer per call
 *      (src/smc.jl:45-51,112-120; src/container.jl:5-27).                                     */
int aps_create(const aps_config *cfg, aps_handle **out);
int aps_destroy(aps_handle *h);

/* Y: T x dy row-major host doubles; copied.  (TracedSSM.Y, src/model.jl:16)                   */
int aps_set_observations(aps_handle *h, const double *Y, int64_t T, int64_t dy);

/* One full (conditional) particle sweep = sweep!(rng, pc, resampler, sampler, ref)
 * (src/container.jl:316-363) including the T+1 resample_propagate! (:171-251) and reweight!
 * (:259-302) rounds. `master_seed` is one rand(rng, UInt64) drawn by the caller from the user's
 * rng (replaces the N+1 draws of seed_from_rng!, src/container.jl:143-159). `ref_traj` is the
 * retained trajectory (T x d row-major host doubles) for PG/PGAS or NULL; pass
 * APS_REF_ON_DEVICE to condition on the trajectory selected by the last aps_pick_trajectory
 * without a host round trip (PGState carried between step calls, src/smc.jl:83-85,112-119).    */
#define APS_REF_ON_DEVICE ((const double *)(uintptr_t)1)
int aps_sweep(aps_handle *h, uint64_t master_seed, const double *ref_traj, double *logevidence);

/* Same sweep, launched kernel by kernel (no CUDA graph) with a CUDA-event pair around every
 * launch on the handle's stream: class_ms[4] / class_launches[4] receive the summed device time
 * and launch count of {propagate, normalise, resample, PGAS-select} kernels. Measurement only.  */
int aps_sweep_profiled(aps_handle *h, uint64_t master_seed, const double *ref_traj, double *logevidence,
                       float *class_ms, int64_t *class_launches);

/* rand(pc.rng, pc) + trajectory extraction (src/container.jl:33-36, src/smc.jl:127).
 * traj_out: T x d host doubles or NULL; index_out: 0-based slot in the final particle set.      */
int aps_pick_trajectory(aps_handle *h, double *traj_out, int64_t *index_out);
/* (sharded handles: a collective call; index_out is the GLOBAL slot, every rank gets the trajectory) */

/* Page-locked host buffers for results the caller wants to OWN (SMCSample.weights is an owned
 * Vector upstream, src/smc.jl:56): a device-to-host copy into such a buffer runs at DMA speed
 * (no staging pass), and unlike aps_get_weights_view the buffer belongs to the caller until
 * aps_host_free. The Python mirror recycles them through a small pool.                          */
int aps_host_alloc(int64_t bytes, void **out);
int aps_host_free(void *p);

/* SMCSample fields (src/smc.jl:23-27,56) materialised lazily.                                  */
int aps_get_weights(aps_handle *h, double *w_out /* N */);                 /* getweights, container.jl:95 */
/* same weights without a staging copy: *w_out points to a pinned host buffer owned by the handle
 * (N doubles), valid until the next call on this handle (Julia: unsafe_wrap; numpy: a view)       */
int aps_get_weights_view(aps_handle *h, const double **w_out);
int aps_get_logweights(aps_handle *h, double *logw_out /* N */);           /* pc.logWs                    */
int aps_get_final_states(aps_handle *h, double *x_out /* N x d */);        /* collect(pc), last state     */
int aps_get_trajectory(aps_handle *h, int64_t slot, double *traj_out /* T x d */);
int aps_get_step_stats(aps_handle *h, double *logz_out /* T */, double *ess_out /* T+1 */,
                       uint8_t *resampled_out /* T+1 */);
/* every trajectory of the final particle set at once: traj_out[t-1][i][k], T x N x d host doubles
 * -- what SMCSample(collect(pc), ...) holds (src/smc.jl:56). One backward pass over the ancestor
 * store on the device, one N x d copy per time step (sharded: this rank's slots).               */
int aps_get_trajectories(aps_handle *h, double *traj_out /* T x N x d */);
/* genealogy slabs, for parity tests: states of time t (1..T) as N x d; ancestors used to
 * build time t (2..T+1; T+1 = final resampled set) as N int32, 0-based.                         */
int aps_get_states(aps_handle *h, int64_t t, double *x_out);
int aps_get_ancestors(aps_handle *h, int64_t t, int32_t *anc_out);
/* Smoothing summary without materialising trajectories on the host: mean_out[t][k] =
 * sum_i W_i X_i[t][k] over the final weighted particle set, X_i = trajectory of final particle i
 * (what examples/gaussian-ssm/script.jl:89-101 computes from collect(pc) on the host). One
 * backward pass over the ancestor store, O(N T) on the device, T x d doubles back. Sharded
 * handles return this rank's partial sums (add the ranks' results on the host).               */
int aps_smoothing_mean(aps_handle *h, double *mean_out /* T x d */);
/* diagnostics: number of "fat" parents (children deferred to the consumer kernel)
 * recorded on this rank at each decision point s = 0..T of the last sweep                        */
int aps_get_fat_counts(aps_handle *h, int32_t *counts_out /* T+1 */);
/* device time (ms, CUDA events on the handle's stream) of the last aps_sweep                    */
int aps_last_sweep_ms(aps_handle *h, float *ms_out);
/* kernels launched by the last aps_sweep (nodes of the replayed CUDA graph count one each)     */
int aps_last_sweep_launches(aps_handle *h, int64_t *n_out);

/* ---- container level (boundary 2/3): the device-resident ParticleContainer driven
 *      call by call, the way the reference's own tests drive theirs (test/container.jl:28-119,
 *      test/pgas.jl:61-91): sweep! is the loop
 *          resample_propagate! -> logZ0 -> reweight! -> logZ1 -> logevidence += logZ1 - logZ0
 *      (src/container.jl:316-363) and each piece is callable. Single-GPU handles only.
 *
 * aps_pc_begin             ParticleContainer(particles, TracedRNG(), rng) + seed_from_rng!
 *                          (src/container.jl:22-27,143-159): N fresh particles, logWs = 0, step
 *                          counter 0; ref_traj as in aps_sweep.
 * aps_pc_resample_propagate resample_propagate!(rng, pc, sampler, resampler, ref) with the handle's
 *                          resampler / ESS threshold (src/container.jl:171-251), incl. update_ref!
 *                          for PGAS (src/pgas.jl:113-128); *resampled_out = 1 if it resampled.
 * aps_pc_reweight          reweight!(pc, ref) (src/container.jl:259-302): every particle advances
 *                          one step and adds its score; *isdone_out = 1 (and nothing changes) when
 *                          all particles had already consumed their T observations.
 * aps_pc_logz              logZ(pc) (src/container.jl:109).
 * aps_set_logweights       pc.logWs = v (a mutable field upstream; test/pgas.jl:82 assigns it).
 * After the loop every accessor above (aps_get_*, aps_pick_trajectory) works as after aps_sweep. */
int aps_pc_begin(aps_handle *h, uint64_t master_seed, const double *ref_traj);
int aps_pc_resample_propagate(aps_handle *h, int32_t *resampled_out);
int aps_pc_reweight(aps_handle *h, int32_t *isdone_out);
int aps_pc_logz(aps_handle *h, double *logz_out);
int aps_set_logweights(aps_handle *h, const double *logw /* N */);

/* ---- operator level (boundary 1): the resampler callable
 *      (rng, w, n) -> Vector{Int} used at src/container.jl:182, and the weight helpers of
 *      src/container.jl:95-119. `w` / outputs may be host or device pointers. The uniform(s)
 *      come from Philox2x64-10 keyed by `key`, counters (i, aps_ctr1(ctr, APS_DOM_RESAMPLE, 0)).  */
int aps_resample(int kind, const double *w, int64_t m, int64_t n, uint64_t key, uint64_t ctr,
                 int64_t *idx_out_1based);
int aps_logsumexp(const double *logw, int64_t n, double *out);             /* logZ, container.jl:109      */
int aps_softmax(const double *logw, int64_t n, double *w_out);             /* getweights, :95             */
int aps_ess(const double *logw, int64_t n, double *out);                   /* effectiveSampleSize, :116   */
int aps_randcat(const double *w, int64_t n, uint64_t key, uint64_t ctr, int64_t *idx_out_1based);
                                                                           /* randcat, resampling.jl:11   */

/* ---- measurement helper: times the dominant resample kernel alone (reads N integer weights,
 *      writes N int32 ancestors) with CUDA events on its stream, L2 flushed between launches.   */
int aps_bench_resample(int kind, int64_t n, int iters, int flush_l2 /* 0 none, 1 write, 2 write + read-back (clean) */, uint64_t seed,
                       float *avg_ms_out, float *min_ms_out);

/* ---- multi-GPU plumbing (one process per GPU; the host exchanges these opaque blobs with
 *      torch.distributed / MPI / Distributed.jl and hands the peers' blobs back).               */
#define APS_IPC_BLOB_BYTES 512
int aps_ipc_export(aps_handle *h, uint8_t *blob_out /* APS_IPC_BLOB_BYTES */);
int aps_ipc_import(aps_handle *h, const uint8_t *blobs /* world_size x APS_IPC_BLOB_BYTES */);

const char *aps_last_error(void);
const char *aps_version(void);

#ifdef __cplusplus
}
#endif
#endif /* APS_B200_H */
