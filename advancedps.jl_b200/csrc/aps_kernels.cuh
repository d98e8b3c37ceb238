// aps_kernels.cuh -- the per-time-step hot path as sm_100a CUDA kernels.
//
//   k_propagate   (K1)  reweight!/advance!           src/container.jl:259-302, src/pgas.jl:53-89
//   k_normalise   (K2)  getweights/logZ/ESS + the ESS decision and evidence accumulation
//                                                     src/container.jl:95-119,233-251,332-359
//   k_resample    (K3)  resample_systematic / resample_stratified + "children grouped by parent"
//                                                     src/resampling.jl:98-183, src/container.jl:182-217
//   k_select_*    (K4/K5) one categorical draw: PGAS ancestor (src/pgas.jl:113-128) and the final
//                        pick (src/container.jl:33-36)
//   k_backtrace         trajectory extraction through the ancestor store (replaces the deep
//                        copies of fork, src/pgas.jl:99-104)
//
// All reductions that feed a comparison are exact integer sums (see aps_math.h), so results do
// not depend on tile size, block scheduling or the number of GPUs.
#pragma once
#include "aps_device.cuh"

struct SweepParams {
    u64 key;      // master seed of this sweep
    int has_ref;  // conditional sweep (PG / PGAS with a retained trajectory)
    int pad;
    u64 epoch;    // sweeps run so far on this handle (sequence numbers of the multi-GPU mailbox)
};

struct DevCtx {
    aps_model_dev md;
    double *x;
    int32_t *anc;
    double *logw;
    u64 *q;
    u64 *tile_sum, *tile_s1, *tile_s2, *tile_prefix;
    StepAcc *acc;
    StepPlan *plan;
    SweepState *st;
    const double *Y;
    const double *ref;
    const SweepParams *sp;
    long long N, T;
    long long NS;  // slab stride: N rounded up to a multiple of 32 (aligned vector access)
    long long x_slabs, anc_slabs;
    long long num_tiles;
    int d, dy;
    int S, Hs;     // weight shift, ESS shift
    int sampler, resampler;
    int bare;      // bare resampler function: resample at every step
    int pad;
    double ess_threshold;
    double logN;
    long long Ng;          // global particle count (== N on one GPU); N is this rank's shard
    long long slot0;       // global index of this rank's first slot
    const PeerTable *peers; // device copy of the peer table, or null on one GPU
    int rank, world;
    int dbg, defer_plan;   // defer_plan: one GPU, few tiles: k_resample derives totals / prefix / plan itself (k_normalise has no last-block phase); dbg: diagnostics only (APS_DEBUG_MULTI): 1 skip waits, 2 no early iterations in the sharded propagate kernel, 4 local gathers, 8 local scatter, 16 in-graph timeline, 32 per-phase times of the fused kernel, 64 maximum posted by the propagate kernel's last block (round-1 scheme), 512 early PDL trigger
    int *fat_cnt;          // [steps] entries in the fat-parent list of each decision point
    FatEntry *fat;         // [steps][APS_FAT_MAX]
    long long fat_steps;   // steps the lists are sized for (T + 2; 1 at the operator level)
    int fat_min, pad3;     // children from which a parent is deferred to the consumer
    // sharded multinomial / residual: receive buffer of routed draws (peer-mapped, behind the fat lists)
    unsigned long long *recv_cnt;  // [fat_steps] draws received at each decision point
    u64 *recv;                     // [recv_cap] their positions, relative to this rank's weight range
    long long recv_cap;
    u64 *rank_woff;                // [fat_steps][APS_MAX_RANKS + 1] exclusive weight prefix of the ranks for the draws of a step
    long long n_override;  // operator level: number of indices to draw (0: N, or N-1 with a reference)
    long long ctr_offset;  // operator level: Philox step counter = plan index + ctr_offset
    // pre-drawn standard normals of the NEXT propagate kernel (k_draw_normals; null: the propagate kernel draws itself):
    // [local pair][2 d] doubles, the pair layout of aps_pair_normals
    double *zbuf;
};

enum { IN_LOGW = 0, IN_W = 1, IN_Q = 2 };

// ---------------------------------------------------------------- init: decision point s = 0
// All log-weights are zero: ESS = N exactly, logZ = log N; particles carry no state yet, so the
// initial resample_propagate! (src/container.jl:325) only decides `resampled[0]`.
__global__ void k_init_sweep(const __grid_constant__ DevCtx c) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        StepPlan p;
        p.M = 0.0;
        p.logZ = c.logN;
        p.ess = (double)c.Ng;
        p.Q = (u64)c.Ng << c.S;
        p.R = 0;
        p.ratio = 0.0;
        p.roff = 0.0;
        p.n = c.Ng - (c.sp->has_ref ? 1 : 0);
        p.resampled = c.bare ? 1 : ((double)c.Ng <= c.ess_threshold * (double)c.Ng ? 1 : 0);
        p.err = 0;
        p.guard = 8;
        p.pad = 0;
        c.plan[0] = p;
        c.st->logev = 0.0;
        c.st->err = 0;
        c.st->picked_slot = -1;
        c.st->spin[0] = c.st->spin[1] = c.st->spin[2] = c.st->spin[3] = 0;
    }
}

// ---------------------------------------------------------------- K0: the state draws of step t, ahead of time
// The standard normals of a step depend on nothing but (key, slot, t) -- half of the propagate kernel's
// instructions (10 Philox rounds, log, sqrt, sincospi per pair) wait for no data. This kernel makes them for
// step t + 1 on a parallel branch of the sweep's graph while the normalise and resample kernels of step t --
// one wave or less each, latency-bound, issue slots mostly idle -- occupy the stream; the propagate kernel then
// only loads them (16 bytes per pair and dimension, L2-resident). Same functions, same values, bit for bit.
// A small persistent grid (a few blocks per SM) on a low-priority stream: it must fill idle slots, not take SMs.
template <int D>
__global__ void __launch_bounds__(512) k_draw_normals(const __grid_constant__ DevCtx c, const long long t) {
    const u64 key = c.sp->key;
    const long long npairs = (c.N + 1) >> 1;
    const long long pair0 = c.slot0 >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // two pairs per iteration: two independent dependency chains per thread (this kernel runs with ONE warp per
    // scheduler, so instruction-level parallelism is all the latency hiding it has)
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += 2 * stride) {
        const long long p1 = p + stride;
        uint64_t w0[2 * D], w1[2 * D];
        aps_pair_words<D>(key, (u64)(pair0 + p), (u64)t, w0);
        aps_pair_words<D>(key, (u64)(pair0 + (p1 < npairs ? p1 : p)), (u64)t, w1);
        double z0[2 * D], z1[2 * D];
        aps_words_to_normals<D>(w0, z0);
        aps_words_to_normals<D>(w1, z1);
        double2 *dst = reinterpret_cast<double2 *>(c.zbuf + p * (2 * D));
#pragma unroll
        for (int j = 0; j < D; ++j) dst[j] = make_double2(z0[2 * j], z0[2 * j + 1]);
        if (p1 < npairs) {
            double2 *dst1 = reinterpret_cast<double2 *>(c.zbuf + p1 * (2 * D));
#pragma unroll
            for (int j = 0; j < D; ++j) dst1[j] = make_double2(z1[2 * j], z1[2 * j + 1]);
        }
    }
}

// ---------------------------------------------------------------- K1: propagate + reweight
// One thread per PAIR of adjacent slots (2p, 2p+1): the pair shares D Philox blocks and their
// Box-Muller normals (aps_pair_normals), ancestors / log-weights / states move as 8- and 16-byte
// vectors. Block maxima of the new log-weights are folded into one atomicMax per block.
#define APS_K1_BOUNDS __launch_bounds__(APS_K1_THREADS, APS_K1_MINBLOCKS)
// PRE: the normals of the step were drawn ahead of time (k_draw_normals) and are loaded from c.zbuf. Without the
// Philox / Box-Muller state the kernel needs fewer registers, and -- 92 instead of 232 instructions per particle --
// it is bound by the latency of its loads (ncu: long scoreboard), so it runs with more resident blocks.
#ifndef APS_K1_MINBLOCKS_PRE
#define APS_K1_MINBLOCKS_PRE 10   // measured at N = 1e6 (ms per sweep): 8 blocks 2.479, 10 blocks 2.438, 12 blocks (40 registers, spills) 2.470
#endif
template <int D, int DY, int OBS, bool MULTI, bool PRE = false>
__global__ void __launch_bounds__(APS_K1_THREADS, (PRE && D == 1) ? APS_K1_MINBLOCKS_PRE : APS_K1_MINBLOCKS) k_propagate(const __grid_constant__ DevCtx c, const long long t,
                                                           double *__restrict__ xt, const double *__restrict__ xp,
                                                           const int32_t *anc) {  // not __restrict__: patched below
    __shared__ u64 red[APS_K1_THREADS / 32];
    if (c.dbg & 512) APS_PDL_TRIGGER();
    const long long N = c.N, NS = c.NS;
    const int has_ref = c.sp->has_ref;   // (sweep parameters: written before the graph is launched)
    const u64 key = c.sp->key;
    const double *__restrict__ y = c.Y + (t - 1) * c.dy;

    u64 bmax = 0;
    unsigned bad = 0;
    double mx = aps_bits2d(0xFFF0000000000000ULL);   // running maximum of this thread's log-weights (-inf)
    bool any = false;
    const long long npairs = (N + 1) >> 1;
    const long long pair0 = c.slot0 >> 1;
    const bool multi = MULTI;
    const u64 seq0 = multi ? c.sp->epoch * (u64)(c.T + 2) : 0ull;
    const long long xoff = xp - c.x;  // slab offset, identical on every rank
    long long p = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x;
    // Children of fat parents (deferred by the resample kernel): every thread first resolves its
    // OWN slots against the list and patches the ancestor store, then runs the unchanged main loop
    // (a separate pre-pass keeps the lookup out of the main loop's register budget).
    auto resolve_fat = [&]() {
        if (t <= 1) return;
        int nfat = MULTI ? __ldcg(&c.fat_cnt[t - 1]) : c.fat_cnt[t - 1];   // (sharded: pushed by the peers, read at L2)
        if (!nfat) return;
        if (nfat > APS_FAT_MAX) nfat = APS_FAT_MAX;
        __shared__ int4 fatl[APS_FAT_MAX];
        fat_stage(fatl, c.fat + (t - 1) * APS_FAT_MAX, nfat);   // (nfat is the same in every thread: uniform)
        int32_t *ancw = const_cast<int32_t *>(anc);
#pragma unroll 1
        for (long long pp = p; pp < npairs; pp += (long long)gridDim.x * APS_K1_THREADS) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const long long i = 2 * pp + h;
                if (i < N) {
                    const int af = fat_lookup(fatl, nfat, (int)(c.slot0 + i));
                    if (af >= 0) ancw[i] = af;
                }
            }
        }
    };
    // the random words do not depend on the ancestors (nor, sharded, on the peers): draw the first pair's
    // before the loads / the wait for the peers
    uint64_t w[2 * D];
    const double *__restrict__ zb = PRE ? c.zbuf : nullptr;   // pre-drawn normals of this step (k_draw_normals)
    if (!PRE && p < npairs) aps_pair_words<D>(key, (u64)(pair0 + p), (u64)t, w);
    APS_PDL_WAIT();   // everything below reads or writes what the previous kernels of the sweep produce
    if (!MULTI) resolve_fat();
    SpanProbe probe(&c.acc[t], 0, c.dbg & 16);
    // log-weights start at zero in every sweep (src/smc.jl:45-51); afterwards they restart from
    // zero only when the previous decision point resampled (reset_logweights!, container.jl:228)
    const bool reset = t == 1 || c.plan[t - 1].resampled != 0;
    // Sharded: the ancestor scatter of step t-1 (peer stores from every rank) must have landed before the
    // slots it touches are read: this rank's resample kernel is complete (stream order), so block 0 tells
    // every rank, and every block waits until all ranks said so. At t = 1 the same exchange makes sure
    // every rank has finished its previous sweep (and trajectory extraction) before any state slab is
    // overwritten.
    // Most slots do not depend on a peer at all: the children of this rank's OWN parents -- global slots
    // [safe_lo, safe_hi), recorded by block 0 of the resample kernel -- have their ancestor entries written
    // by this rank's resample kernel (complete: stream order) and their parents' states in the local slab.
    // Iterations whose 256 slots lie inside that range run BEFORE the wait (phase 0), so the NVLink
    // traversal of the barrier and the skew between the ranks hide behind them; the others follow the
    // wait (phase 1). Not at t = 1, and not when this rank deferred children of fat parents (their
    // ranges cross the safe slots; the list is read after the barrier only).
    __shared__ u64 s_w[APS_MAX_RANKS][4];
    bool early = false;
    int safe_lo = 0, safe_hi = 0;
    if (MULTI) {
        const u64 v0 = 0;
        if (blockIdx.x == 0) mail_post(c.peers, c.rank, c.world, 2, seq0 + (u64)(t - 1) + 1, &v0, 1);
        if (t > 1 && !(c.dbg & 2)) {
            safe_lo = c.acc[t - 1].safe_lo;
            safe_hi = c.acc[t - 1].safe_hi;
            early = safe_hi > safe_lo && __ldcg(&c.fat_cnt[t - 1]) == 0;
        }
    }
    const long long stride = (long long)gridDim.x * APS_K1_THREADS;
    // block-uniform: do all slots of the block's iteration that starts at pair pb descend from this rank's parents?
    auto iteration_is_safe = [&](long long pb) {
        const long long g0 = c.slot0 + 2 * pb, g1 = g0 + 2 * APS_K1_THREADS;
        return early && g0 >= safe_lo && (g1 < c.slot0 + N ? g1 : c.slot0 + N) <= safe_hi;
    };
    bool need_wait = !early;   // a block whose iterations are all safe does not depend on the peers at all
    if (MULTI && early)
        for (long long pb = (long long)blockIdx.x * APS_K1_THREADS; pb < npairs; pb += stride) need_wait = need_wait || !iteration_is_safe(pb);
    long long wp = p;   // the pair whose random words are in w
    // Order inside one iteration: ancestor indices first (they are needed for the only dependent load
    // chain of the loop), then the integer-only Philox rounds while they arrive, then the parent-state
    // gather, then the floating-point half of the draw (log / sqrt / sincospi) while THAT is in flight.
    auto body = [&](const long long p) {
        const long long i0 = 2 * p;
        int2 a2 = make_int2(0, 0);
        if (t > 1) a2 = MULTI ? __ldcg(reinterpret_cast<const int2 *>(anc + i0))   // scattered by the peers: read at L2
                              : *reinterpret_cast<const int2 *>(anc + i0);
        double2 lw2 = make_double2(0.0, 0.0);
        if (!reset) lw2 = *reinterpret_cast<const double2 *>(c.logw + i0);
        double z[2 * D];
        if (PRE) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const double2 v = *reinterpret_cast<const double2 *>(zb + p * (2 * D) + 2 * j);
                z[2 * j] = v.x;
                z[2 * j + 1] = v.y;
            }
        } else {
            if (p != wp) aps_pair_words<D>(key, (u64)(pair0 + p), (u64)t, w);
            if (D > 2) aps_words_to_normals<D>(w, z);   // (d >= 3: registers are the limit -- finish the draw before the gather)
        }
        double xg[2][D];   // parent states of the two slots
        auto gather = [&](int h) {
            long long a = h ? a2.y : a2.x;  // global parent index
            if (i0 + h >= N) a = c.slot0;   // (padding slot of an odd N: its ancestor entry is not written)
            const double *xsrc = xp;
            if (multi && !(c.dbg & 4)) {
                const unsigned al = (unsigned)(a - c.slot0);
                if (al < (unsigned)N) {
                    a = al;  // own shard (the common case)
                } else {     // the parent lives on a peer: read its state over NVLink
                    const int owner = (int)((unsigned)a / (unsigned)N);
                    a -= (long long)owner * N;
                    xsrc = c.peers->x[owner] + xoff;
                }
            }
#pragma unroll
            for (int k = 0; k < D; ++k) xg[h][k] = xsrc[(long long)k * NS + a];
        };
        if (D <= 2 && t > 1) {
            gather(0);
            gather(1);
        }
        if (D <= 2 && !PRE) aps_words_to_normals<D>(w, z);
        double xo[2][D];
        double lwo[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long i = i0 + h;
            double x[D];
            lwo[h] = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = 0.0;
            if (i < N) {
                if (has_ref && c.slot0 + i == c.Ng - 1) {  // the reference keeps the globally last slot
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = c.ref[(t - 1) * D + k];
                } else if (t == 1) {
                    aps_prior_draw<D>(&c.md, z + h * D, x);
                } else {
                    if (D > 2) gather(h);
                    aps_trans_draw<D>(&c.md, xg[h], z + h * D, x);
                }
                const double ll = aps_obs_logpdf<D, DY, OBS>(&c.md, x, y);
                const double lw = (reset ? 0.0 : (h ? lw2.y : lw2.x)) + ll;
                lwo[h] = lw;
                if (lw != lw) bad = 1;
                else {
                    mx = lw > mx ? lw : mx;   // (encoded once per thread after the loop: the encoding is monotone)
                    any = true;
                }
            }
#pragma unroll
            for (int k = 0; k < D; ++k) xo[h][k] = x[k];
        }
#pragma unroll
        for (int k = 0; k < D; ++k)
            *reinterpret_cast<double2 *>(xt + (long long)k * NS + i0) = make_double2(xo[0][k], xo[1][k]);
        *reinterpret_cast<double2 *>(c.logw + i0) = make_double2(lwo[0], lwo[1]);
    };
    if (!MULTI) {
        for (; p < npairs; p += stride) body(p);
    } else {
#pragma unroll 1
        for (int phase = early ? 0 : 1; phase < (need_wait ? 2 : 1); ++phase) {
            if (phase == 1) {
                if (!(c.dbg & 1) && !mail_wait(c.peers, c.rank, c.world, 2, seq0 + (u64)(t - 1) + 1, s_w, 1, c.st->spin, &c.st->err, t == 1 ? 20 : 1))
                    c.st->err = APS_ERR_COMM;
                __syncthreads();
                resolve_fat();  // the peers' pushes into this rank's list are complete now
            }
#pragma unroll 1
            for (long long pb = (long long)blockIdx.x * APS_K1_THREADS; pb < npairs; pb += stride) {
                if (iteration_is_safe(pb) != (phase == 0)) continue;
                if (pb + threadIdx.x < npairs) body(pb + threadIdx.x);
            }
        }
    }
    if (any) bmax = aps_encode_ordered(mx);
    if (!(c.dbg & 512)) APS_PDL_TRIGGER();   // this block's stores are issued: the next kernel may start launching
    bmax = block_max_u64<APS_K1_THREADS / 32>(bmax, red);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (bmax) atomicMax(&c.acc[t].max_enc, bmax);
        if (bad) atomicOr(&c.acc[t].bad, 1u);
    }
    if (multi && (c.dbg & 64)) {   // (64: the round-1 scheme, kept for A/B timing -- block 0 of k_normalise posts otherwise)
        // all-reduce(max), producer side: the last block to finish publishes this shard's maximum
        // to every rank, so it is on its way while this kernel drains and k_normalise launches.
        // Measured (2 x B200): the ticket and, above all, the remote store at the very end of the kernel --
        // the grid is not complete before NVLink has acknowledged it -- cost 1.7 us per step
        __shared__ unsigned s_lastb;
        __shared__ u64 s_pub[2];
        if (threadIdx.x == 0) {
            s_lastb = ticket_release(&c.acc[t].k1_done, 1u) == gridDim.x - 1 ? 1u : 0u;
            if (s_lastb) {
                fence_acq_rel_gpu();
                s_pub[0] = atomicMax(&c.acc[t].max_enc, 0ull);
                s_pub[1] = (u64)atomicOr(&c.acc[t].bad, 0u);
            }
        }
        __syncthreads();
        if (s_lastb) mail_post(c.peers, c.rank, c.world, 0, seq0 + (u64)t + 1, s_pub, 2);
    }
    probe.end();
}

// K1 for even state dimensions >= 4: one thread per SLOT. A pair of slots consumes D Philox blocks,
// slot 2p the normals of blocks 0..D/2-1 and slot 2p+1 those of blocks D/2..D-1 (aps_pair_normals), so
// with an even D each slot can draw its own D/2 blocks -- same draws, no shared block -- and a
// thread carries one particle's state instead of two: about half the registers of the pair kernel
// (d = 4: 128 -> 2 blocks of 256 threads per SM in round 1), twice the resident threads.
#ifndef APS_K1P_MINBLOCKS
#define APS_K1P_MINBLOCKS 6
#endif
#ifndef APS_K1P_MINBLOCKS_PRE
#define APS_K1P_MINBLOCKS_PRE 6
#endif
// PRE: the slot's D normals were drawn ahead of time (k_draw_normals<D>: the pair layout [pair][2 D] IS the slot
// layout [slot][D] -- slot 2p takes the first D normals of its pair, slot 2p+1 the last D)
template <int D, int DY, int OBS, bool MULTI, bool PRE = false>
__global__ void __launch_bounds__(APS_K1_THREADS, PRE ? APS_K1P_MINBLOCKS_PRE : APS_K1P_MINBLOCKS) k_propagate1(const __grid_constant__ DevCtx c, const long long t,
                                                                                 double *__restrict__ xt, const double *__restrict__ xp,
                                                                                 const int32_t *anc) {  // not __restrict__: patched below
    static_assert(D % 2 == 0, "one thread per slot needs an even number of Philox blocks per pair");
    __shared__ u64 red[APS_K1_THREADS / 32];
    const long long N = c.N, NS = c.NS;
    const int has_ref = c.sp->has_ref;
    const u64 key = c.sp->key;
    const double *__restrict__ y = c.Y + (t - 1) * c.dy;
    u64 bmax = 0;
    unsigned bad = 0;
    double mx = aps_bits2d(0xFFF0000000000000ULL);
    bool any = false;
    const bool multi = MULTI;
    const u64 seq0 = multi ? c.sp->epoch * (u64)(c.T + 2) : 0ull;
    const long long xoff = xp - c.x;
    const long long stride = (long long)gridDim.x * APS_K1_THREADS;
    long long i = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x;
    auto resolve_fat = [&]() {
        if (t <= 1) return;
        int nfat = MULTI ? __ldcg(&c.fat_cnt[t - 1]) : c.fat_cnt[t - 1];
        if (!nfat) return;
        if (nfat > APS_FAT_MAX) nfat = APS_FAT_MAX;
        __shared__ int4 fatl[APS_FAT_MAX];
        fat_stage(fatl, c.fat + (t - 1) * APS_FAT_MAX, nfat);
        int32_t *ancw = const_cast<int32_t *>(anc);
#pragma unroll 1
        for (long long ii = i; ii < N; ii += stride) {
            const int af = fat_lookup(fatl, nfat, (int)(c.slot0 + ii));
            if (af >= 0) ancw[ii] = af;
        }
    };
    // this slot's D/2 Philox blocks: counter (pair, ctr1(t, DOM_STATE, half * D/2 + j))
    auto draw_words = [&](long long slot, uint64_t *w) {
        const u64 g = (u64)(c.slot0 + slot);
#pragma unroll
        for (int j = 0; j < D / 2; ++j)
            aps_philox2x64(g >> 1, aps_ctr1((u64)t, APS_DOM_STATE, (uint32_t)((g & 1) * (D / 2) + j)), key, &w[2 * j], &w[2 * j + 1]);
    };
    uint64_t w[D];
    const double *__restrict__ zb = PRE ? c.zbuf : nullptr;
    if (!PRE && i < N) draw_words(i, w);
    if (!MULTI) resolve_fat();
    const bool reset = t == 1 || c.plan[t - 1].resampled != 0;
    if (MULTI) {
        __shared__ u64 s_w[APS_MAX_RANKS][4];
        const u64 v0 = 0;
        if (blockIdx.x == 0) mail_post(c.peers, c.rank, c.world, 2, seq0 + (u64)(t - 1) + 1, &v0, 1);
        if (!(c.dbg & 1) && !mail_wait(c.peers, c.rank, c.world, 2, seq0 + (u64)(t - 1) + 1, s_w, 1, c.st->spin, &c.st->err, t == 1 ? 20 : 1))
            c.st->err = APS_ERR_COMM;
        __syncthreads();
        resolve_fat();
    }
    for (bool first = true; i < N; i += stride, first = false) {
        long long a = 0;
        if (t > 1) a = MULTI ? __ldcg(anc + i) : anc[i];
        double lw_old = 0.0;
        if (!reset) lw_old = c.logw[i];
        double z[D];
        if (PRE) {
#pragma unroll
            for (int j = 0; j < D / 2; ++j) {
                const double2 v = *reinterpret_cast<const double2 *>(zb + i * D + 2 * j);
                z[2 * j] = v.x;
                z[2 * j + 1] = v.y;
            }
        } else {
            if (!first) draw_words(i, w);
#pragma unroll
            for (int j = 0; j < D / 2; ++j) aps_normal_pair(w[2 * j], w[2 * j + 1], &z[2 * j], &z[2 * j + 1]);
        }
        double x[D];
        if (has_ref && c.slot0 + i == c.Ng - 1) {  // the reference keeps the globally last slot
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = c.ref[(t - 1) * D + k];
        } else if (t == 1) {
            aps_prior_draw<D>(&c.md, z, x);
        } else {
            const double *xsrc = xp;
            if (multi && !(c.dbg & 4)) {
                const unsigned al = (unsigned)(a - c.slot0);
                if (al < (unsigned)N) {
                    a = al;
                } else {
                    const int owner = (int)((unsigned)a / (unsigned)N);
                    a -= (long long)owner * N;
                    xsrc = c.peers->x[owner] + xoff;
                }
            }
            double xg[D];
#pragma unroll
            for (int k = 0; k < D; ++k) xg[k] = xsrc[(long long)k * NS + a];
            aps_trans_draw<D>(&c.md, xg, z, x);
        }
        const double ll = aps_obs_logpdf<D, DY, OBS>(&c.md, x, y);
        const double lw = (reset ? 0.0 : lw_old) + ll;
        if (lw != lw) bad = 1;
        else {
            mx = lw > mx ? lw : mx;
            any = true;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) xt[(long long)k * NS + i] = x[k];
        c.logw[i] = lw;
    }
    if (any) bmax = aps_encode_ordered(mx);
    bmax = block_max_u64<APS_K1_THREADS / 32>(bmax, red);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (bmax) atomicMax(&c.acc[t].max_enc, bmax);
        if (bad) atomicOr(&c.acc[t].bad, 1u);
    }
    if (multi && (c.dbg & 64)) {   // (see k_propagate)
        __shared__ unsigned s_lastb;
        __shared__ u64 s_pub[2];
        if (threadIdx.x == 0) {
            s_lastb = ticket_release(&c.acc[t].k1_done, 1u) == gridDim.x - 1 ? 1u : 0u;
            if (s_lastb) {
                fence_acq_rel_gpu();
                s_pub[0] = atomicMax(&c.acc[t].max_enc, 0ull);
                s_pub[1] = (u64)atomicOr(&c.acc[t].bad, 0u);
            }
        }
        __syncthreads();
        if (s_lastb) mail_post(c.peers, c.rank, c.world, 0, seq0 + (u64)t + 1, s_pub, 2);
    }
}

// max of a plain vector (operator-level entry points)
template <int INPUT>
__global__ void __launch_bounds__(APS_K1_THREADS) k_vector_max(const double *__restrict__ in, long long n, StepAcc *acc) {
    __shared__ u64 red[APS_K1_THREADS / 32];
    u64 bmax = 0;
    unsigned bad = 0;
    for (long long i = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; i < n;
         i += (long long)gridDim.x * APS_K1_THREADS) {
        const double v = in[i];
        if (v != v || (INPUT == IN_W && v < 0.0)) bad = 1;
        else {
            const u64 e = aps_encode_ordered(v);
            bmax = e > bmax ? e : bmax;
        }
    }
    bmax = block_max_u64<APS_K1_THREADS / 32>(bmax, red);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (bmax) atomicMax(&acc->max_enc, bmax);
        if (bad) atomicOr(&acc->bad, 1u);
    }
}

// The plan of decision point s from the global integer totals: logZ, ESS, the
// ResampleWithESSThreshold decision (src/container.jl:242-247), the systematic offset R, the
// estimate constants of k_resample, and the evidence accumulation (src/container.jl:341,359).
template <int INPUT>
__device__ __forceinline__ void make_plan(const DevCtx &c, long long s, double M, u64 Q, u64 Q1, u64 Q2, int err,
                                          StepPlan *out) {
    StepPlan p;
    if (INPUT != IN_Q) {
        if (!(M == M) || M == aps_bits2d(0x7FF0000000000000ULL) || M == aps_bits2d(0xFFF0000000000000ULL))
            err = APS_ERR_WEIGHTS;
        if (INPUT == IN_W && !(M > 0.0)) err = APS_ERR_WEIGHTS;
    }
    if (Q == 0 || Q2 == 0) err = APS_ERR_WEIGHTS;
    p.M = M;
    p.Q = Q;
    p.logZ = M + aps_log((double)Q * aps_pow2i(-c.S));
    p.ess = ((double)Q1 * (double)Q1) / (double)Q2;
    p.resampled = c.bare ? 1 : (p.ess <= c.ess_threshold * (double)c.Ng ? 1 : 0);
    p.n = c.n_override > 0 ? c.n_override : c.Ng - (c.sp->has_ref ? 1 : 0);
    uint64_t w0, w1;
    aps_philox2x64(0, aps_ctr1((u64)(s + c.ctr_offset), APS_DOM_RESAMPLE, 0), c.sp->key, &w0, &w1);
    p.R = ceil_uq53(aps_u53(w0), Q);
    p.ratio = 0x1.0p24 * ((double)p.n / (double)Q);
    p.roff = 0x1.0p24 * ((double)p.R / (double)Q);
    p.err = err;
    p.guard = 2 + (int)((p.n + (1LL << 25) - 1) >> 25);
    p.pad = 0;
    *out = p;
}
// the two independent halves of make_plan, for callers that can run them on different warps
// (same arithmetic, disjoint fields): A = weights summary and decision, B = resampling offsets
template <int INPUT>
__device__ __forceinline__ void make_plan_a(const DevCtx &c, long long s, double M, u64 Q, u64 Q1, u64 Q2, int err,
                                            StepPlan *out) {
    if (INPUT != IN_Q) {
        if (!(M == M) || M == aps_bits2d(0x7FF0000000000000ULL) || M == aps_bits2d(0xFFF0000000000000ULL))
            err = APS_ERR_WEIGHTS;
        if (INPUT == IN_W && !(M > 0.0)) err = APS_ERR_WEIGHTS;
    }
    if (Q == 0 || Q2 == 0) err = APS_ERR_WEIGHTS;
    out->M = M;
    out->Q = Q;
    out->logZ = M + aps_log((double)Q * aps_pow2i(-c.S));
    const double ess = ((double)Q1 * (double)Q1) / (double)Q2;
    out->ess = ess;
    out->resampled = c.bare ? 1 : (ess <= c.ess_threshold * (double)c.Ng ? 1 : 0);
    out->err = err;
    out->pad = 0;
}
__device__ __forceinline__ void make_plan_b(const DevCtx &c, long long s, u64 Q, StepPlan *out) {
    const long long n = c.n_override > 0 ? c.n_override : c.Ng - (c.sp->has_ref ? 1 : 0);
    out->n = n;
    uint64_t w0, w1;
    aps_philox2x64(0, aps_ctr1((u64)(s + c.ctr_offset), APS_DOM_RESAMPLE, 0), c.sp->key, &w0, &w1);
    const u64 R = ceil_uq53(aps_u53(w0), Q);
    out->R = R;
    out->ratio = 0x1.0p24 * ((double)n / (double)Q);
    out->roff = 0x1.0p24 * ((double)R / (double)Q);
    out->guard = 2 + (int)((n + (1LL << 25) - 1) >> 25);
}
// evidence bookkeeping, done once per decision point by the thread that records the plan
__device__ __forceinline__ void record_plan(const DevCtx &c, long long s, const StepPlan &p) {
    if (c.st) {
        if (p.err) c.st->err = p.err;
        else if (s >= 1) {
            const StepPlan &pv = c.plan[s - 1];
            const double logZ0 = pv.resampled ? c.logN : pv.logZ;  // logZ(pc) after resample_propagate!
            c.st->logev += p.logZ - logZ0;                         // src/container.jl:341,359
        }
    }
    c.plan[s] = p;
}

// ---------------------------------------------------------------- K2: normalise
// One tile of APS_TILE particles per block: q_i = floor(exp(logw_i - M) 2^S), tile totals of q,
// (q >> Hs) and (q >> Hs)^2. The last block to finish scans the tile totals and writes the plan
// of decision point s: logZ, ESS, the resampling decision, the systematic offset, the evidence.
// K2 is latency-bound at the particle counts of interest (one exp per particle), so it spreads a
// tile over APS_K2_THREADS = 256 threads (8 particles each) instead of the 128 of the resampler;
// all 489 tiles of N = 1e6 are then resident at once (4 blocks per SM).
#ifndef APS_K2_THREADS
#define APS_K2_THREADS 256   // measured standalone at N = 1e6: 256 threads 15.9 us, 512 threads 16.7 us
#endif
#define APS_K2_IPT (APS_TILE / APS_K2_THREADS)
#define APS_K2_WARPS (APS_K2_THREADS / 32)
template <int INPUT>
__global__ void __launch_bounds__(APS_K2_THREADS) k_normalise(const __grid_constant__ DevCtx c, const double *__restrict__ in,
                                                           const long long s) {
    __shared__ u64 red[APS_K2_WARPS];
    __shared__ u64 s_tot[3];
    __shared__ unsigned s_last;
    if (c.dbg & 512) APS_PDL_TRIGGER();
    const long long N = c.N;
    const long long base = (long long)blockIdx.x * APS_TILE;
    StepAcc *acc = &c.acc[s];
    APS_PDL_WAIT();
    SpanProbe probe(acc, 1, APS_TIMELINE && (c.dbg & 16));
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    u64 max_enc = acc->max_enc;
    unsigned bad_in = 0;
    const bool multi = c.world > 1;
    const u64 seq = multi ? c.sp->epoch * (u64)(c.T + 2) + (u64)s + 1 : 0ull;
    double vin[APS_K2_IPT];  // this thread's inputs, loaded before the (sharded) wait for the maxima
    if (INPUT != IN_Q) {
#pragma unroll
        for (int r = 0; r < APS_K2_IPT; ++r) {
            const long long i = base + r * APS_K2_THREADS + threadIdx.x;
            vin[r] = i < N ? in[i] : 0.0;
        }
    }
    if (multi) {
        // all-reduce(max): every block combines the shard maxima published by the propagate kernels
        __shared__ u64 s_m[APS_MAX_RANKS][4];
        __shared__ int s_okm;
        __shared__ u64 s_pubm[2];
        if (threadIdx.x == 0) {
            s_okm = 1;
            s_pubm[0] = max_enc;       // this shard's maximum: the propagate kernel is complete (stream order)
            s_pubm[1] = (u64)acc->bad;
        }
        __syncthreads();
        // producer side: block 0 publishes the shard maximum to every rank at the START of this kernel, so the
        // NVLink traversal overlaps the kernel instead of holding the end of the propagate kernel
        if (blockIdx.x == 0 && !(c.dbg & 64)) mail_post(c.peers, c.rank, c.world, 0, seq, s_pubm, 2);
        if (!(c.dbg & 1) && !mail_wait(c.peers, c.rank, c.world, 0, seq, s_m, 2, c.st ? c.st->spin : nullptr, c.st ? &c.st->err : nullptr)) s_okm = 0;
        if (c.dbg & 1) { if (threadIdx.x < c.world) { s_m[threadIdx.x][0] = acc->max_enc; s_m[threadIdx.x][1] = 0; } }
        __syncthreads();
        max_enc = 0;
        for (int r = 0; r < c.world; ++r) {
            max_enc = s_m[r][0] > max_enc ? s_m[r][0] : max_enc;
            bad_in |= (unsigned)s_m[r][1];
        }
        if (!s_okm && threadIdx.x == 0 && c.st) c.st->err = APS_ERR_COMM;
    }
    const double M = aps_decode_ordered(max_enc);
    const double scale = aps_pow2i(c.S);
    u64 s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int r = 0; r < APS_K2_IPT; ++r) {
        const long long i = base + r * APS_K2_THREADS + threadIdx.x;
        if (i < N) {
            u64 qi;
            if (INPUT == IN_Q) {
                qi = c.q[i];
            } else {
                double e;
                if (INPUT == IN_LOGW) e = aps_exp(vin[r] - M);
                else e = vin[r] / M;
                qi = (e > 0.0) ? (u64)__double2ull_rz(e * scale) : 0ull;
                c.q[i] = qi;
            }
            const u64 qs = qi >> c.Hs;
            s0 += qi;
            s1 += qs;
            s2 += qs * qs;
        }
    }
    if (!(c.dbg & 512)) APS_PDL_TRIGGER();
    // tile totals: warp shuffles, then one shared-memory integer atomic per warp and total
    // (exact integers, so the order is irrelevant; cheaper than three block-wide reductions)
    s0 = warp_sum_u64(s0);
    s1 = warp_sum_u64(s1);
    s2 = warp_sum_u64(s2);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_tot[0], s0);
        atomicAdd(&s_tot[1], s1);
        atomicAdd(&s_tot[2], s2);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        c.tile_sum[blockIdx.x] = s_tot[0];
        c.tile_s1[blockIdx.x] = s_tot[1];
        c.tile_s2[blockIdx.x] = s_tot[2];
    }
    // One GPU and few tiles (in-sweep, systematic / stratified): every block of k_resample sums the
    // tile totals itself while its TMA load is in flight and derives the plan, so this kernel ends
    // here -- no ticket, no serial last-block phase (3.4 us of 16 at N = 1e6).
    if (c.defer_plan) {
        if (multi && blockIdx.x == 0 && threadIdx.x == 0) {  // what block 0 of k_resample publishes with the totals
            acc->tot[3] = max_enc;
            acc->pad1 = (acc->bad | bad_in) ? 1u : 0u;
        }
        probe.end();
        return;
    }
    if (threadIdx.x == 0) {
        const unsigned ticket = ticket_release(&acc->done_ctr, 1u);
        s_last = (ticket == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    fence_acq_rel_gpu();

    // ---- last block: exclusive scan of the tile totals (each thread owns a contiguous chunk)
    const long long nt = c.num_tiles;
    const long long per = (nt + APS_K2_THREADS - 1) / APS_K2_THREADS;
    const long long lo = (long long)threadIdx.x * per;
    const long long hi = lo + per < nt ? lo + per : nt;
    u64 a0 = 0, a1 = 0, a2 = 0;
    for (long long k = lo; k < hi; ++k) {
        a0 += __ldcg(&c.tile_sum[k]);
        a1 += __ldcg(&c.tile_s1[k]);
        a2 += __ldcg(&c.tile_s2[k]);
    }
    u64 Q;
    u64 run = block_excl_scan_u64<APS_K2_WARPS>(a0, red, &Q);
    u64 Q1 = block_sum_u64<APS_K2_WARPS>(a1, red);
    u64 Q2 = block_sum_u64<APS_K2_WARPS>(a2, red);
    for (long long k = lo; k < hi; ++k) {
        c.tile_prefix[k] = run;  // sharded: local prefix; the rank offset is added by the consumer
        run += __ldcg(&c.tile_sum[k]);
    }
    if (multi) {
        // this rank's totals, published by block 0 of the resample kernel (Q2 < 2^63: its top bit
        // carries the NaN flag; tot[3] = the already global maximum)
        __shared__ u64 s_tot[4];
        if (threadIdx.x == 0) {
            s_tot[0] = acc->tot[0] = Q;
            s_tot[1] = acc->tot[1] = Q1;
            s_tot[2] = acc->tot[2] = Q2 | ((u64)((acc->bad | bad_in) ? 1 : 0) << 63);
            s_tot[3] = acc->tot[3] = max_enc;
        }
        __syncthreads();
        mail_post(c.peers, c.rank, c.world, 1, seq, s_tot, 4);  // all-gather of the totals, producer side
        return;
    }
    if (threadIdx.x == 0) {
        StepPlan p;
        int err = (acc->bad) ? APS_ERR_WEIGHTS : 0;
        if (INPUT != IN_Q && acc->max_enc == 0) err = APS_ERR_WEIGHTS;
        make_plan<INPUT>(c, s, M, Q, Q1, Q2, err, &p);
        record_plan(c, s, p);
    }
}

// ---------------------------------------------------------------- sharded: the plan from the shard totals
// One block posts `nv` values (shared memory) of this rank under (kind, seq) and collects every
// rank's values into out[r][k]. Returns false when a peer did not answer (timeout).
__device__ __forceinline__ bool block_exchange(const DevCtx &c, int kind, u64 seq, const u64 *v, int nv, u64 (*out)[4]) {
    __syncthreads();
    mail_post(c.peers, c.rank, c.world, kind, seq, v, nv);
    bool ok = true;
    if (!(c.dbg & 1)) ok = mail_wait(c.peers, c.rank, c.world, kind, seq, out, nv, c.st ? c.st->spin : nullptr, c.st ? &c.st->err : nullptr);
    return __syncthreads_and(ok ? 1 : 0) != 0;
}
__device__ __forceinline__ u64 step_seq(const DevCtx &c, long long s) { return c.sp->epoch * (u64)(c.T + 2) + (u64)s + 1; }

// thread 0 of a block: combine the shard totals of every rank (rank order; integers) into the
// plan of decision point s -- identical on every rank -- and this rank's exclusive weight offset
__device__ __forceinline__ void multi_plan(const DevCtx &c, long long s, const u64 (*s_t)[4], bool comm_ok, StepPlan *plan,
                                           u64 *rank_off) {
    u64 Q = 0, Q1 = 0, Q2 = 0, off = 0;
    int bad = 0;
    for (int r = 0; r < c.world; ++r) {
        if (r < c.rank) off += s_t[r][0];
        Q += s_t[r][0];
        Q1 += s_t[r][1];
        Q2 += s_t[r][2] & 0x7FFFFFFFFFFFFFFFULL;
        bad |= (int)(s_t[r][2] >> 63);
    }
    const u64 menc = s_t[0][3];
    int err = (bad || menc == 0) ? APS_ERR_WEIGHTS : 0;
    if (!comm_ok) err = APS_ERR_COMM;
    make_plan<IN_LOGW>(c, s, aps_decode_ordered(menc), Q, Q1, Q2, err, plan);
    *rank_off = off;
}

// sharded multinomial / residual: one block exchanges the shard totals and records the plan and
// this rank's weight offset in global memory for the kernels that follow
__global__ void __launch_bounds__(32) k_plan_multi(const __grid_constant__ DevCtx c, const long long s) {
    __shared__ u64 s_t[APS_MAX_RANKS][4];
    bool ok = true;  // the totals were published by the last block of every rank's normalise kernel
    if (!(c.dbg & 1)) ok = mail_wait(c.peers, c.rank, c.world, 1, step_seq(c, s), s_t, 4, c.st->spin, &c.st->err);
    ok = __syncthreads_and(ok ? 1 : 0) != 0;
    if (threadIdx.x == 0) {
        StepPlan p;
        u64 off;
        multi_plan(c, s, s_t, ok, &p, &off);
        c.acc[s].rank_off = off;
        record_plan(c, s, p);
        if (c.rank_woff) {   // weight prefix of every rank: where a routed multinomial draw belongs
            u64 run = 0;
            for (int r = 0; r < c.world; ++r) {
                c.rank_woff[s * (APS_MAX_RANKS + 1) + r] = run;
                run += s_t[r][0];
            }
            c.rank_woff[s * (APS_MAX_RANKS + 1) + c.world] = run;
        }
    }
}

// ---------------------------------------------------------------- K3: resample
// K(C) = #{ children i in [0,n) : i Q + R_i <= C n }: the number of children whose threshold lies
// at or below cumulative weight C. Parent j owns children [K(C_{j-1}), K(C_j)).
__device__ __forceinline__ bool thr_le(u64 i, u64 Q, u64 R, u128 Cn) {
    return le_128(add_128_64(mul_64_64(i, Q), R), Cn);
}

// exact: first i in [0,n] with !(i < n && i Q + R <= C n), starting from a guess k (rare path)
__device__ __noinline__ int first_above_exact(int k, u64 C, u64 Q, u64 R, int n) {
    const u128 Cn = mul_64_64(C, (u64)n);
    if (k < 0) k = 0;
    if (k > n) k = n;
    int lo = k - 2 < 0 ? 0 : k - 2, hi = k + 2 > n ? n : k + 2;
    if (lo > 0 && !thr_le((u64)(lo - 1), Q, R, Cn)) lo = 0;
    if (hi < n && thr_le((u64)hi, Q, R, Cn)) hi = n;
    while (lo < hi) {
        const int mid = (int)(((long long)lo + hi) >> 1);
        if (thr_le((u64)mid, Q, R, Cn)) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// est approximates 2^24 (C n - R) / Q. Error budget: (double)C, (double)Q, the division, the
// 2^24 scaling (exact) and the fma each contribute <= 2^-53 relative on a value <= 2^24 n, and the
// floor to 2^-24 units one more unit, so the true value is within (1 + n 2^-26.5) units of the
// integer v below. floor(est)+1 is therefore K unless the 24-bit fraction lies within
// guard = 2 + ceil(n / 2^25) units of an integer; those cases are flagged (*unsafe is OR-ed) and
// settled with exact 128-bit arithmetic (about 5e-7 of all evaluations at n = 2^20).
#define APS_KFRAC_BITS 24
__device__ __forceinline__ int first_above_est(double est20, int n, int guard, bool *unsafe) {
    const long long v = __double2ll_rd(est20);
    const int kf = (int)(v >> APS_KFRAC_BITS);
    const unsigned fr = (unsigned)v & ((1u << APS_KFRAC_BITS) - 1u);
    *unsafe = *unsafe || (fr - (unsigned)guard >= (1u << APS_KFRAC_BITS) - 2u * (unsigned)guard);
    return min(kf + 1, n);
}

// stratified: child i draws its own offset R_i; only stratum i* = floor(C n / Q) is undecided.
// F = first i with i Q > C n (clamped to n).
__device__ __forceinline__ int strat_children_below(int F, u64 C, u64 Q, int n, u64 key, u64 step) {
    const u128 Cn = mul_64_64(C, (u64)n);
    if (F == n && le_128(mul_64_64((u64)n, Q), Cn)) return n;
    const int istar = F - 1;
    uint64_t w0, w1;
    aps_philox2x64((u64)istar, aps_ctr1(step, APS_DOM_RESAMPLE, 0), key, &w0, &w1);
    const u64 Ri = ceil_uq53(aps_u53(w0), Q);
    return istar + (thr_le((u64)istar, Q, Ri, Cn) ? 1 : 0);
}

// Expand phase shared by all resamplers. Parents drop a marker (their tile-local id) at their
// first child slot in shared memory; a max-scan (20 consecutive slots per thread, 128-bit shared
// accesses) turns the markers into parent ids, which leave as 128-bit global stores. Cost follows
// the number of children, not the skew of the weights.
//
// destination of ancestor entries: child slot g (global index) -> address. On one GPU the local
// slab; when sharded, the slab of the rank that owns slot g (a peer-mapped pointer: the ancestor
// indices are scattered over NVLink with plain stores).
struct AncDst {
    int32_t *base;            // local slab
    const PeerTable *peers;   // null on one GPU
    long long slab_off;       // offset of the slab inside each rank's ancestor store
    int nl;                   // slots per rank
    int lo;                   // first global slot of this rank
    __device__ __forceinline__ int32_t *at(int g) const {
        if (!peers) return base + g;
        const unsigned gl = (unsigned)(g - lo);
        if (gl < (unsigned)nl) return base + gl;  // own shard (the common case)
        const int owner = (int)((unsigned)g / (unsigned)nl);
        return peers->anc[owner] + slab_off + (g - owner * nl);
    }
};

// ---- fat parents (see aps_device.cuh)
struct FatSink {
    int *cnt;                // this rank's counter of decision point s
    FatEntry *ent;           // this rank's entries of decision point s
    const PeerTable *peers;  // null on one GPU
    size_t off_cnt, off_ent; // byte offsets of (cnt, ent) from the mailbox base, for the peers' copies
    int fat_min, nl, rank, pad;
};
__device__ __forceinline__ FatSink make_fat_sink(const DevCtx &c, long long s, bool multi) {
    FatSink f;
    f.cnt = c.fat_cnt + s;
    f.ent = c.fat + s * APS_FAT_MAX;
    f.peers = multi ? c.peers : nullptr;
    f.off_cnt = aps_mail_bytes() + (size_t)s * sizeof(int);
    f.off_ent = aps_mail_bytes() + aps_fatcnt_bytes(c.fat_steps) + (size_t)s * APS_FAT_MAX * sizeof(FatEntry);
    f.fat_min = c.fat_min;
    f.nl = (int)c.N;
    f.rank = c.rank;
    f.pad = 0;
    return f;
}
// record parent `parent` as the owner of the global child slots [lo, hi) on every rank that owns some of them
__device__ __forceinline__ void fat_push(const FatSink &f, int lo, int hi, int parent) {
    FatEntry e;
    e.lo = lo;
    e.hi = hi;
    e.parent = parent;
    e.pad = 0;
    if (!f.peers) {
        const int idx = atomicAdd(f.cnt, 1);
        if (idx < APS_FAT_MAX) f.ent[idx] = e;
        return;
    }
    const int r0 = lo / f.nl, r1 = (hi - 1) / f.nl;
    for (int r = r0; r <= r1; ++r) {
        char *mb = reinterpret_cast<char *>(f.peers->mail[r]);
        int *cnt = r == f.rank ? f.cnt : reinterpret_cast<int *>(mb + f.off_cnt);
        FatEntry *ent = r == f.rank ? f.ent : reinterpret_cast<FatEntry *>(mb + f.off_ent);
        const int idx = atomicAdd_system(cnt, 1);
        if (idx < APS_FAT_MAX) *reinterpret_cast<int4 *>(ent + idx) = *reinterpret_cast<const int4 *>(&e);
    }
}
// Marker array layout. Parents are thread-blocked (16 consecutive parents per thread), so in one
// marker round the lanes of a warp write child positions about 16 apart: with a plain layout they
// fall into two banks (16-way conflicts; ncu: 1/3 of all shared-memory wavefronts of the kernel).
// One pad word per 32 positions spreads a stride-16 pattern over all 32 banks.
#ifndef APS_OWN_PAD
#define APS_OWN_PAD 0   // measured: 98.2 us padded (scalar read-back) vs 95.9 us plain at N = 2^25
#endif
#if APS_OWN_PAD
#define APS_OWN_WORDS(cap) ((((cap) + ((cap) >> 5) + 1) + 3) & ~3)
__device__ __forceinline__ int own_idx(int p) { return p + (p >> 5); }
#else
#define APS_OWN_WORDS(cap) (cap)
__device__ __forceinline__ int own_idx(int p) { return p; }
#endif

// scan + store for the child slots [cb, cb + cnt) whose markers are already in own[]
template <int TH, int CPT>
__device__ __forceinline__ void expand_scan_store(int cb, int cnt, int kA, int kB, int base, const AncDst &dst, int *own,
                                                  int *wmax) {
    constexpr int WARPS = TH / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool active = tid * CPT < cnt;
    int v[CPT];
    int run = 0;
    if (active) {
#if APS_OWN_PAD
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            run = max(run, own[own_idx(tid * CPT + j)]);
            v[j] = run;
        }
#else
        const int4 *own4 = reinterpret_cast<const int4 *>(own);
#pragma unroll
        for (int m = 0; m < CPT / 4; ++m) {
            const int4 t = own4[tid * (CPT / 4) + m];
            run = max(run, t.x); v[4 * m] = run;
            run = max(run, t.y); v[4 * m + 1] = run;
            run = max(run, t.z); v[4 * m + 2] = run;
            run = max(run, t.w); v[4 * m + 3] = run;
        }
#endif
    }
    int inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int tt = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, tt);
    }
    if (lane == 31) wmax[warp] = inc;
    int excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 0;
    __syncthreads();
    {
        int w = wmax[lane & (WARPS - 1)];
#pragma unroll
        for (int o = 1; o < WARPS; o <<= 1) {
            const int tt = __shfl_up_sync(0xffffffffu, w, o, WARPS);
            if ((lane & (WARPS - 1)) >= o) w = max(w, tt);
        }
        const int prev = __shfl_sync(0xffffffffu, w, (warp + WARPS - 1) & (WARPS - 1), WARPS);
        if (warp > 0) excl = max(excl, prev);
    }
    if (active) {
        const int add = base - 1;
#pragma unroll
        for (int m = 0; m < CPT / 4; ++m) {
            const int g = cb + tid * CPT + 4 * m;  // global child index of this vector, multiple of 4
            int4 o;
            o.x = max(v[4 * m], excl) + add;
            o.y = max(v[4 * m + 1], excl) + add;
            o.z = max(v[4 * m + 2], excl) + add;
            o.w = max(v[4 * m + 3], excl) + add;
            int32_t *p = dst.at(g);  // 4-aligned groups never straddle two ranks (slots per rank % 32 == 0)
            if (g >= kA && g + 4 <= kB) {
                *reinterpret_cast<int4 *>(p) = o;
            } else {
                if (g >= kA && g < kB) p[0] = o.x;
                if (g + 1 >= kA && g + 1 < kB) p[1] = o.y;
                if (g + 2 >= kA && g + 2 < kB) p[2] = o.z;
                if (g + 3 >= kA && g + 3 < kB) p[3] = o.w;
            }
        }
    }
}

template <int TH, int CPT>
__device__ __forceinline__ void zero_own(int *own) {
    int4 *own4 = reinterpret_cast<int4 *>(own);
    const int4 z = make_int4(0, 0, 0, 0);
#if APS_OWN_PAD
    constexpr int NV = APS_OWN_WORDS(TH * CPT) / 4;  // consecutive threads, consecutive vectors
#pragma unroll
    for (int m = 0; m < (NV + TH - 1) / TH; ++m)
        if (m * TH + (int)threadIdx.x < NV) own4[m * TH + threadIdx.x] = z;
#else
#pragma unroll
    for (int m = 0; m < CPT / 4; ++m) own4[threadIdx.x * (CPT / 4) + m] = z;
#endif
}

// general (rare) path: any number of children, clipped chunk by chunk. khi: inclusive child
// counts of this thread's parents, klo0: count below its first parent. Parents with at least
// fat_min children are pushed to the fat list and their child range is skipped (the consumer
// fills it in), so the cost follows the number of parents, not the number of children.
template <int TH, int IPT, int CPT>
__device__ __noinline__ void expand_tile_general(const int *khi, int klo0, int kA, int kB, int base, const AncDst &dst,
                                                 int *own, int *wmax, const FatSink &fat) {
    constexpr int CAP = TH * CPT;
    __shared__ int s_to;
    const int tid = threadIdx.x;
    {
        int klo = klo0;
        for (int j = 0; j < IPT; ++j) {
            const int kh = khi[j];
            if (kh - klo >= fat.fat_min) fat_push(fat, klo, kh, base + tid * IPT + j);
            klo = kh;
        }
    }
    for (int cb = kA & ~3; cb < kB;) {
        if (tid == 0) s_to = 0;
        __syncthreads();
        {   // does this chunk start inside a deferred range? then continue behind it
            int klo = klo0;
            for (int j = 0; j < IPT; ++j) {
                const int kh = khi[j];
                if (kh - klo >= fat.fat_min && klo <= cb && (kh & ~3) > cb) s_to = kh & ~3;
                klo = kh;
            }
        }
        __syncthreads();
        const int to = s_to;
        if (to > cb) {
            cb = to;
            continue;
        }
        const int cnt = (kB - cb) < CAP ? (kB - cb) : CAP;
        zero_own<TH, CPT>(own);
        __syncthreads();
        int klo = klo0;
        for (int j = 0; j < IPT; ++j) {
            const int kh = khi[j];
            const int lo_rel = klo - cb, hi_rel = kh - cb;
            if (kh > klo && hi_rel > 0 && lo_rel < cnt) own[own_idx(lo_rel > 0 ? lo_rel : 0)] = tid * IPT + j + 1;
            klo = kh;
        }
        __syncthreads();
        expand_scan_store<TH, CPT>(cb, cnt, kA, kB, base, dst, own, wmax);
        cb += CAP;
    }
}

// estimate with an immediate exact fix-up when it cannot be trusted (retry path of a tile)
template <int KIND>
__device__ __forceinline__ int children_below_checked(u64 C, u64 Q, u64 R, int n, double ratio, double roff, int guard,
                                                      u64 key, u64 step) {
    bool u = false;
    int k = first_above_est(__fma_rn((double)C, ratio, -roff), n, guard, &u);
    if (u) k = first_above_exact(k, C, Q, KIND == APS_RESAMPLE_SYSTEMATIC ? R : 0ull, n);
    if (KIND == APS_RESAMPLE_STRATIFIED) k = strat_children_below(k, C, Q, n, key, step);
    return k;
}

// children below cumulative weight C from the estimate; *unsafe is OR-ed when it cannot be trusted
template <int KIND>
__device__ __forceinline__ int children_below_fast(u64 C, u64 Q, int n, double ratio, double roff, int guard,
                                                   u64 key, u64 step, bool *unsafe) {
    int k = first_above_est(__fma_rn((double)C, ratio, -roff), n, guard, unsafe);
    if (KIND == APS_RESAMPLE_STRATIFIED) k = strat_children_below(k, C, Q, n, key, step);
    return k;
}

// One tile of APS_TILE parents per block: reads their integer weights (8 B each), writes the
// sorted ancestor indices of the children they own (4 B each).
//
// The tile's weights arrive through one TMA bulk-tensor copy: q is described to the TMA unit as a
// [rows][16] u64 tensor (128-byte rows), the box is 256 rows, and the 128-byte hardware swizzle
// (16-byte chunk index XOR row & 7) makes the thread-blocked read-back -- thread t owns row t,
// i.e. 16 consecutive weights -- free of bank conflicts. Rows past the end of the tensor are
// zero-filled by the hardware, so ragged tails need no special case.
#define APS_TILE_BYTES (APS_TILE * 8)
#define APS_K3_DYN_SMEM (APS_TILE_BYTES + APS_OWN_WORDS(APS_K3_CAP) * 4)
template <int KIND, bool MULTI, bool DEFER>
__global__ void __launch_bounds__(APS_K3_THREADS, MULTI ? APS_K3_MINBLOCKS - 2 : APS_K3_MINBLOCKS) k_resample(const __grid_constant__ DevCtx c, const long long s,
                                                             int32_t *__restrict__ anc_out,
                                                             const __grid_constant__ CUtensorMap tmap_q) {
    extern __shared__ __align__(1024) unsigned char dynsmem[];  // [tile: APS_TILE u64, swizzled][own: APS_K3_CAP int]
    __shared__ u64 red[APS_K3_WARPS];
    __shared__ int wmax[APS_K3_WARPS];
    __shared__ __align__(8) uint64_t mbar;
    unsigned char *tilebuf = dynsmem;
    int *own = reinterpret_cast<int *>(dynsmem + APS_TILE_BYTES);
    if (c.dbg & 512) APS_PDL_TRIGGER();
    const long long N = c.N;
    const StepPlan *pp = c.plan + s;
    const int tid = threadIdx.x;
    if (DEFER && !MULTI) zero_own<APS_K3_THREADS, APS_K3_CPT>(own);   // (shared memory only: before the wait)
    APS_PDL_WAIT();
    SpanProbe probe(&c.acc[s], 2, APS_TIMELINE && (c.dbg & 16));
    const long long base = (long long)blockIdx.x * APS_TILE;
    AncDst dst;
    dst.base = anc_out;
    dst.peers = (MULTI && !(c.dbg & 8)) ? c.peers : nullptr;
    dst.slab_off = anc_out - c.anc;
    dst.nl = (int)N;
    dst.lo = (int)c.slot0;
    const int gbase = (int)(c.slot0 + base);  // global index of the tile's first parent
    // update_keys! branch (src/container.jl:247): every particle continues, weights kept
    auto identity_ancestors = [&]() {
#pragma unroll
        for (int r = 0; r < APS_K3_IPT; ++r) {
            const long long i = base + r * APS_K3_THREADS + tid;
            if (i < N) anc_out[i] = (int32_t)(c.slot0 + i);
        }
    };
    u64 Q, R, key;
    int n, guard;
    double ratio, roff;
    auto load_plan = [&]() {
        Q = pp->Q;
        R = pp->R;
        n = (int)pp->n;
        guard = pp->guard;
        ratio = pp->ratio;
        roff = KIND == APS_RESAMPLE_SYSTEMATIC ? pp->roff : 0.0;
        key = KIND == APS_RESAMPLE_STRATIFIED ? c.sp->key : 0ull;
    };
    constexpr bool defer = DEFER;
    if (!MULTI && !defer) {
        if (!pp->resampled || pp->err) {
            identity_ancestors();
            return;
        }
        load_plan();
    }
    if (tid == 0) {  // only warp 0 touches the mbarrier (init, TMA issue, wait): no block barrier needed here
        if (smem_u32(tilebuf) & 1023u) __trap();  // the 128-byte swizzle pattern assumes a 1 KB aligned tile
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, APS_TILE_BYTES);
        tma_load_2d(tilebuf, &tmap_q, &mbar, 0, (int)(base / APS_ROW));
        // pull the tile that a block ~2 residency waves later will need from HBM into L2 now
#if APS_K3_PF_WAVES > 0
        const long long pf = (long long)blockIdx.x + (long long)APS_K3_PF_WAVES * APS_K3_MINBLOCKS * 148;
        if (pf < (long long)gridDim.x) tma_prefetch_2d(&tmap_q, 0, (int)(pf * (APS_TILE / APS_ROW)));
#endif
    }
    const u64 step = (u64)(s + c.ctr_offset);
    u64 tprefix = 0, shard_q = 0;   // (shard_q: total weight of this rank's tiles, deferred plan only)
    if (!defer) tprefix = c.tile_prefix[blockIdx.x];

    if (!(DEFER && !MULTI)) zero_own<APS_K3_THREADS, APS_K3_CPT>(own);
    if (defer) {
        // Deferred plan (see k_normalise): while the TMA load is in flight, sum the tile totals
        // (integers: any order), take the part below this tile as its prefix, and derive the plan of
        // decision point s -- the same arithmetic the normalise kernel's last block would have run.
        __shared__ u64 s_acc4[4];
        __shared__ StepPlan s_plan1;
        if (tid < 4) s_acc4[tid] = 0;
        __syncthreads();
        const long long nt = c.num_tiles;
        u64 t0 = 0, t1 = 0, t2 = 0, p0 = 0;
        for (long long k = tid; k < nt; k += APS_K3_THREADS) {
            const u64 v = __ldcg(&c.tile_sum[k]);
            t0 += v;
            if (k < (long long)blockIdx.x) p0 += v;
            t1 += __ldcg(&c.tile_s1[k]);
            t2 += __ldcg(&c.tile_s2[k]);
        }
        t0 = warp_sum_u64(t0);
        t1 = warp_sum_u64(t1);
        t2 = warp_sum_u64(t2);
        p0 = warp_sum_u64(p0);
        if ((tid & 31) == 0) {
            atomicAdd(&s_acc4[0], t0);
            atomicAdd(&s_acc4[1], t1);
            atomicAdd(&s_acc4[2], t2);
            atomicAdd(&s_acc4[3], p0);
        }
        __syncthreads();
        tprefix = s_acc4[3];
        shard_q = s_acc4[0];
        if (tid == 0) c.tile_prefix[blockIdx.x] = s_acc4[3];  // (local) prefix, kept for the final pick (k_pick)
        if (MULTI) {
            // sharded: these are the SHARD totals. Block 0 publishes them to every rank (all-gather,
            // consumer side: the normalise kernel has no last block any more); the wait and the plan
            // follow after the tile load and the local scan.
            __shared__ u64 s_pub[4];
            if (blockIdx.x == 0) {
                if (tid == 0) {
                    StepAcc *acc = &c.acc[s];
                    s_pub[0] = acc->tot[0] = s_acc4[0];
                    s_pub[1] = acc->tot[1] = s_acc4[1];
                    s_pub[2] = acc->tot[2] = s_acc4[2] | ((u64)(acc->pad1 ? 1 : 0) << 63);  // NaN flag (left by k_normalise)
                    s_pub[3] = acc->tot[3];                                                 // global maximum (left by k_normalise)
                }
                __syncthreads();
                mail_post(c.peers, c.rank, c.world, 1, step_seq(c, s), s_pub, 4);
            }
        } else {
            if (tid == 0) {  // the two halves of the plan on two warps
                const StepAcc *acc = &c.acc[s];
                int err = acc->bad ? APS_ERR_WEIGHTS : 0;
                if (acc->max_enc == 0) err = APS_ERR_WEIGHTS;
                make_plan_a<IN_LOGW>(c, s, aps_decode_ordered(acc->max_enc), s_acc4[0], s_acc4[1], s_acc4[2], err, &s_plan1);
            } else if (tid == 32) {
                make_plan_b(c, s, s_acc4[0], &s_plan1);
            }
            __syncthreads();
            if (blockIdx.x == 0 && tid == 0) record_plan(c, s, s_plan1);
            pp = &s_plan1;
            if (!pp->resampled || pp->err) {
                if (tid < 32) {
                    __syncwarp();
                    mbar_wait(&mbar, 0);  // the tile is still landing in shared memory: do not leave before it has
                }
                __syncthreads();
                identity_ancestors();
                return;
            }
            load_plan();
        }
    }
    if (tid < 32) {  // one warp polls the mbarrier, the others park on the block barrier
        __syncwarp();
        mbar_wait(&mbar, 0);
    }
    __syncthreads();

    // ---- this thread's APS_K3_IPT consecutive integer weights (its part of one 128-byte row of the
    //      swizzled tile: 16-byte chunk c of row r sits at chunk position c ^ (r & 7)), local inclusive sums
    u64 cum[APS_K3_IPT];
    {
        const int rowi = (tid * APS_K3_IPT) / APS_ROW;
        const int chunk0 = ((tid * APS_K3_IPT) % APS_ROW) / 2;
        const unsigned char *row = tilebuf + rowi * (APS_ROW * 8);
#pragma unroll
        for (int r = 0; r < APS_K3_IPT / 2; ++r) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(row + (((chunk0 + r) ^ (rowi & 7)) << 4));
            cum[2 * r] = v.x;
            cum[2 * r + 1] = v.y;
        }
    }
#pragma unroll
    for (int r = 1; r < APS_K3_IPT; ++r) cum[r] += cum[r - 1];
    u64 tile_total;
    u64 excl = block_excl_scan_u64<APS_K3_WARPS>(cum[APS_K3_IPT - 1], red, &tile_total);  // syncs: own[] is zeroed

    if (MULTI) {
        // sharded: combine the shard totals published by the normalise kernels of every rank
        // (rank order; integers) and derive the plan locally -- identical on every rank. The wait
        // comes after the tile load and the local scan, which do not depend on the peers.
        __shared__ StepPlan s_plan;
        __shared__ u64 s_t[APS_MAX_RANKS][4];
        __shared__ u64 s_off;
        __shared__ int s_okt;
        if (tid == 0) s_okt = 1;
        __syncthreads();
        if (!(c.dbg & 1) && !mail_wait(c.peers, c.rank, c.world, 1, step_seq(c, s), s_t, 4, c.st->spin, &c.st->err)) s_okt = 0;
        if (c.dbg & 1) { if (tid < c.world) { s_t[tid][0] = 1ull << 50; s_t[tid][1] = 1ull << 30; s_t[tid][2] = 1ull << 40; s_t[tid][3] = c.acc[s].max_enc; } }
        __syncthreads();
        if (tid == 0 || tid == 32) {  // the two halves of the plan on two warps (same arithmetic as multi_plan)
            u64 Q = 0, Q1 = 0, Q2 = 0, off = 0;
            int bad = 0;
            for (int r = 0; r < c.world; ++r) {
                if (r < c.rank) off += s_t[r][0];
                Q += s_t[r][0];
                Q1 += s_t[r][1];
                Q2 += s_t[r][2] & 0x7FFFFFFFFFFFFFFFULL;
                bad |= (int)(s_t[r][2] >> 63);
            }
            if (tid == 0) {
                const u64 menc = s_t[0][3];
                int err = (bad || menc == 0) ? APS_ERR_WEIGHTS : 0;
                if (!s_okt) err = APS_ERR_COMM;
                make_plan_a<IN_LOGW>(c, s, aps_decode_ordered(menc), Q, Q1, Q2, err, &s_plan);
                s_off = off;
            } else {
                make_plan_b(c, s, Q, &s_plan);
            }
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) {
            c.acc[s].rank_off = s_off;
            record_plan(c, s, s_plan);
        }
        pp = &s_plan;
        tprefix += s_off;
        if (!pp->resampled || pp->err) {
            if (blockIdx.x == 0 && tid == 0 && !pp->err) {   // identity ancestors: every slot of this rank is its own parent
                c.acc[s].safe_lo = (int)c.slot0;
                c.acc[s].safe_hi = (int)(c.slot0 + N);
            }
            identity_ancestors();
            return;
        }
        load_plan();
        if (blockIdx.x == 0 && tid == 64) {
            // the child slots that descend from this rank's parents, [K(offset), K(offset + shard total)): exact
            // (128-bit fix-up where the estimate is within its guard band). k_propagate runs them before the barrier.
            const u64 shard_total = defer ? shard_q : c.acc[s].tot[0];
            c.acc[s].safe_lo = (c.slot0 == 0) ? 0 : children_below_checked<KIND>(s_off, Q, R, n, ratio, roff, guard, key, step);
            c.acc[s].safe_hi = children_below_checked<KIND>(s_off + shard_total, Q, R, n, ratio, roff, guard, key, step);
        }
    }
    excl += tprefix;

    if (!(c.dbg & 512)) APS_PDL_TRIGGER();
    // ---- child range of the tile and of this thread (every thread evaluates the three bounds itself)
    const bool first_tile = blockIdx.x == 0 && c.slot0 == 0;  // K(C_{-1}) := 0 for the globally first parent
    bool unsafe = false;
    const int kA = first_tile ? 0 : children_below_fast<KIND>(tprefix, Q, n, ratio, roff, guard, key, step, &unsafe);
    const int kB = children_below_fast<KIND>(tprefix + tile_total, Q, n, ratio, roff, guard, key, step, &unsafe);
    int klo = tid == 0 ? kA : children_below_fast<KIND>(excl, Q, n, ratio, roff, guard, key, step, &unsafe);
    const int cb = kA & ~3;
    const bool single = kB - cb <= APS_K3_CAP;

    // ---- fast path: estimate the children below each parent and drop its marker at once
    if (single) {
#pragma unroll
        for (int r = 0; r < APS_K3_IPT; ++r) {
            const int k = children_below_fast<KIND>(excl + cum[r], Q, n, ratio, roff, guard, key, step, &unsafe);
            if (k > klo) own[own_idx(klo - cb)] = tid * APS_K3_IPT + r + 1;
            klo = k;
        }
    }
    const int slow = __syncthreads_or((unsafe || !single) ? 1 : 0);
    if (!slow) {
        expand_scan_store<APS_K3_THREADS, APS_K3_CPT>(cb, kB - cb, kA, kB, gbase, dst, own, wmax);
    } else {
        // ---- retry (some estimate fell within the guard band of an integer, or the tile owns
        //      more children than one pass holds): same walk with exact fix-ups where needed
        const int kAx = first_tile ? 0 : children_below_checked<KIND>(tprefix, Q, R, n, ratio, roff, guard, key, step);
        const int kBx = children_below_checked<KIND>(tprefix + tile_total, Q, R, n, ratio, roff, guard, key, step);
        int klx = tid == 0 ? kAx : children_below_checked<KIND>(excl, Q, R, n, ratio, roff, guard, key, step);
        const int cbx = kAx & ~3;
        if (kBx - cbx <= APS_K3_CAP) {
            zero_own<APS_K3_THREADS, APS_K3_CPT>(own);  // every thread is past the marker loop (the __syncthreads_or above)
            __syncthreads();
            for (int r = 0; r < APS_K3_IPT; ++r) {
                const int k = children_below_checked<KIND>(excl + cum[r], Q, R, n, ratio, roff, guard, key, step);
                if (k > klx) own[own_idx(klx - cbx)] = tid * APS_K3_IPT + r + 1;
                klx = k;
            }
            __syncthreads();
            expand_scan_store<APS_K3_THREADS, APS_K3_CPT>(cbx, kBx - cbx, kAx, kBx, gbase, dst, own, wmax);
        } else {
            int khi[APS_K3_IPT];
            for (int r = 0; r < APS_K3_IPT; ++r)
                khi[r] = children_below_checked<KIND>(excl + cum[r], Q, R, n, ratio, roff, guard, key, step);
            expand_tile_general<APS_K3_THREADS, APS_K3_IPT, APS_K3_CPT>(khi, klx, kAx, kBx, gbase, dst, own, wmax, make_fat_sink(c, s, MULTI));
        }
    }

    // reference particle keeps the last slot (src/container.jl:219-224); PGAS may overwrite it
    if (blockIdx.x == gridDim.x - 1 && tid == 0 && n < c.Ng && c.rank == c.world - 1) anc_out[N - 1] = (int32_t)(c.Ng - 1);
    probe.end();
}

// ---------------------------------------------------------------- multinomial / residual resampling
// resample_multinomial (src/resampling.jl:31-35): child i draws U_i and takes the first parent j
// with C_j > floor(U_i Q / 2^53) -- i.i.d. categorical draws by inverse CDF on the integer
// weights (upstream uses Distributions' alias table; same law). resample_residual (:53-81):
// floor(n q_j / Q) copies of j, the rest i.i.d. on the residuals n q_j - floor(.) Q (shifted so
// their sum fits 62 bits). Inside the sweep only the offspring counts matter (children are laid
// out grouped by parent, src/container.jl:185-217), so draws are histogrammed with
// warp-aggregated integer atomics and expanded like the systematic case.
struct MultiArgs {
    const u64 *qsrc;        // integer weights the draws are made on (q, or shifted residuals)
    u64 *cum;               // inclusive cumulative sums of qsrc (global)
    const u64 *tile_prefix; // exclusive tile prefix of qsrc
    int *counts;            // offspring count per parent
    int *tile_count;        // offspring count per tile (becomes its exclusive scan)
    int *tile_cprefix;
    const StepPlan *plan;   // decision point plan (resampled / err / n)
    const StepPlan *wplan;  // plan carrying Q of qsrc (plan itself, or the residual stage plan)
    const long long *n_draws; // device: number of i.i.d. draws (n, or the residual count Rc)
    int32_t *out32;         // operator level: 0-based parent of draw r goes to out32[out_offset + r]
    const long long *out_offset;
    long long N;
    long long num_tiles;
    long long step;         // Philox step counter of the draws
    // sharded sweep (null / zero on one GPU): every rank makes all draws and keeps those that fall
    // into its own weight range [*range_lo, *range_lo + *range_len)
    const u64 *range_lo, *range_len;
    long long *child_off;   // out: children owned by parents of lower ranks (k_scan_tile_counts)
    // cut-point (guide) table of the inverse CDF, built by k_cumsum: entry b of a tile is the first
    // tile-local element whose cumulative weight exceeds b << cut_sh[tile] (tile-relative units)
    unsigned short *cut;    // [num_tiles][APS_CUT]
    unsigned char *cut_sh;  // [num_tiles]
};
#define APS_CUT_BITS 10
#define APS_CUT ((1 << APS_CUT_BITS) + 1)   // entries per tile: <= 1024 buckets + one sentinel

// inclusive cumulative sums of one tile of qsrc (thread-blocked, 16 per thread) and the tile's
// cut-point table: the tile's weight range [0, S) is cut into <= 1024 equal buckets of 2^sh units;
// element j covers the tile-relative interval [C_{j-1}, C_j) and is the answer ("first element
// whose cumulative weight exceeds v") for every bucket boundary v = b << sh inside it. A draw then
// needs one table read and a binary search inside a bracket of a few elements instead of 11 steps
// over the whole tile (measured first: a forward scan from the cut point -- the longest scan of a
// warp's 32 lanes was ~70 elements with 256 buckets, no faster than the plain binary search).
__global__ void __launch_bounds__(APS_THREADS) k_cumsum(const __grid_constant__ MultiArgs a) {
    __shared__ u64 red[APS_WARPS];
    if (!a.plan->resampled || a.plan->err) return;
    const long long base = (long long)blockIdx.x * APS_TILE + (long long)threadIdx.x * APS_IPT;
    u64 cum[APS_IPT];
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r) cum[r] = (base + r < a.N) ? a.qsrc[base + r] : 0ull;
#pragma unroll
    for (int r = 1; r < APS_IPT; ++r) cum[r] += cum[r - 1];
    u64 tot;
    const u64 excl_rel = block_excl_scan_u64<APS_WARPS>(cum[APS_IPT - 1], red, &tot);
    const u64 excl = excl_rel + a.tile_prefix[blockIdx.x];
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r)
        if (base + r < a.N) a.cum[base + r] = excl + cum[r];
    // cut points
    const int bits = tot ? 64 - __clzll((long long)tot) : 0;
    const int sh = bits > APS_CUT_BITS ? bits - APS_CUT_BITS : 0;
    if (threadIdx.x == 0) a.cut_sh[blockIdx.x] = (unsigned char)sh;
    unsigned short *cut = a.cut + (long long)blockIdx.x * APS_CUT;
    if (threadIdx.x == 0) {  // sentinel behind the last used bucket: upper bracket of draws in that bucket
        const long long last = (base + APS_TILE < a.N ? (long long)APS_TILE : a.N - base) - 1;
        cut[tot ? ((tot - 1) >> sh) + 1 : 0] = (unsigned short)(last > 0 ? last : 0);
    }
    u64 prev = excl_rel;
#pragma unroll 1
    for (int r = 0; r < APS_IPT; ++r) {
        const u64 cur = excl_rel + cum[r];
        if (cur > prev) {
            const u64 b0 = (prev + ((1ull << sh) - 1)) >> sh, b1 = (cur - 1) >> sh;  // buckets with prev <= b << sh < cur
            for (u64 b = b0; b <= b1; ++b) cut[b] = (unsigned short)(threadIdx.x * APS_IPT + r);
        }
        prev = cur;
    }
}

// parent of a draw at position tau of this rank's weight range: two-level search (tile prefix, then
// the tile's cut-point table brackets a binary search of a few elements)
__device__ __forceinline__ long long multi_find(const MultiArgs &a, u64 tau) {
    long long lo = 0, hi = a.num_tiles - 1;
    while (lo < hi) {
        const long long mid = (lo + hi + 1) >> 1;
        if (a.tile_prefix[mid] <= tau) lo = mid;
        else hi = mid - 1;
    }
    const long long tile = lo;
    const u64 trel = tau - a.tile_prefix[tile];
    const unsigned short *cp = a.cut + tile * APS_CUT + (long long)(trel >> a.cut_sh[tile]);
    long long jl = tile * APS_TILE + cp[0], jh = tile * APS_TILE + cp[1];
    while (jl < jh) {
        const long long mid = (jl + jh) >> 1;
        if (a.cum[mid] > tau) jh = mid;
        else jl = mid + 1;
    }
    return jl;
}

// i.i.d. draws: draw i is word (i & 1) of Philox block i >> 1, so one thread makes two draws; each
// is a two-level binary search (tile prefix, then the tile's cumulative sums)
template <int TO_COUNTS>
__global__ void __launch_bounds__(APS_K1_THREADS) k_multi_search(const __grid_constant__ MultiArgs a, const u64 *keyp) {
    if (!a.plan->resampled || a.plan->err) return;
    const long long nd = *a.n_draws;
    const u64 Q = a.wplan->Q;
    const u64 key = *keyp;
    const long long off = a.out_offset ? *a.out_offset : 0;
    const u64 lo_w = a.range_lo ? *a.range_lo : 0ull;
    const u64 len_w = a.range_len ? *a.range_len : Q;
    const long long npairs = (nd + 1) >> 1;
    for (long long p = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; p < npairs;
         p += (long long)gridDim.x * APS_K1_THREADS) {
        uint64_t w[2];
        aps_philox2x64((u64)p, aps_ctr1((u64)a.step, APS_DOM_RESAMPLE, 0), key, &w[0], &w[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long i = 2 * p + h;
            const u64 tau = floor_uq53(aps_u53(w[h]), Q) - lo_w;  // relative to this rank's weight range
            if (i >= nd || tau >= len_w) continue;                // (unsigned: also rejects tau below the range)
            const long long jl = multi_find(a, tau);
            if (TO_COUNTS) {
                // integer histogram of the offspring counts. (Per-tile totals are NOT accumulated here:
                // the 489 tile counters share 16 cache lines, and 10^6 atomics on them serialised in
                // L2 -- 150 us of this kernel's 170 us; k_tile_counts sums them afterwards.)
                atomicAdd(&a.counts[jl], 1);
            } else {
                a.out32[off + i] = (int32_t)jl;
            }
        }
    }
}

// Sharded: rank r makes the draws of the Philox blocks [r pp, (r+1) pp) (pp = ceil(ceil(nd / 2) / G)) and
// appends each draw's position -- relative to the OWNER's weight range -- to the owner's receive
// buffer. Peer atomics are expensive (an NVLink round trip each), so a block reserves its slots with
// ONE peer atomic per owner: pass 1 counts the block's draws per owner in shared memory, `world`
// threads reserve the ranges, pass 2 recomputes the draws (Philox is cheaper than a second round of
// remote traffic) and fills the slots with peer stores. (First version: one peer atomic per warp,
// owner and iteration -- 18.5 ms per 2-GPU multinomial sweep against 11.2 ms with replicated draws.)
__global__ void __launch_bounds__(APS_K1_THREADS) k_multi_route(const __grid_constant__ MultiArgs a, const u64 *keyp,
                                                               const __grid_constant__ DevCtx c, const long long s) {
    __shared__ unsigned s_cnt[APS_MAX_RANKS];
    __shared__ unsigned long long s_base[APS_MAX_RANKS];
    if (!a.plan->resampled || a.plan->err) return;
    const long long nd = *a.n_draws;
    if (nd <= 0) return;
    const u64 Q = a.wplan->Q;
    const u64 key = *keyp;
    const long long npairs = (nd + 1) >> 1;
    const long long pp = (npairs + c.world - 1) / c.world;
    const long long p0 = (long long)c.rank * pp, p1 = p0 + pp < npairs ? p0 + pp : npairs;
    const u64 *woff = c.rank_woff + s * (APS_MAX_RANKS + 1);
    u64 off[APS_MAX_RANKS + 1];
#pragma unroll
    for (int r = 0; r <= APS_MAX_RANKS; ++r) off[r] = r <= c.world ? woff[r] : ~0ull;
    const size_t cnt_off = aps_recvcnt_off(c.fat_steps) + (size_t)s * 8, buf_off = aps_recv_off(c.fat_steps);
    if (threadIdx.x < APS_MAX_RANKS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * APS_K1_THREADS;
    const long long first = p0 + (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x;
    auto owner_of = [&](u64 tau) {
        int o = 0;
#pragma unroll
        for (int r = 1; r < APS_MAX_RANKS; ++r)
            if (r < c.world && tau >= off[r]) o = r;
        return o;
    };
    // pass 1: how many of this block's draws go to each owner
    for (long long p = first; p < p1; p += stride) {
        uint64_t w[2];
        aps_philox2x64((u64)p, aps_ctr1((u64)a.step, APS_DOM_RESAMPLE, 0), key, &w[0], &w[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (2 * p + h < nd) atomicAdd(&s_cnt[owner_of(floor_uq53(aps_u53(w[h]), Q))], 1u);
    }
    __syncthreads();
    // one peer atomic per owner reserves this block's slots in the owner's receive buffer
    if (threadIdx.x < c.world) {
        const unsigned n = s_cnt[threadIdx.x];
        char *mb = reinterpret_cast<char *>(c.peers->mail[threadIdx.x]);
        s_base[threadIdx.x] = n ? atomicAdd_system(reinterpret_cast<unsigned long long *>(mb + cnt_off), (unsigned long long)n) : 0ull;
        s_cnt[threadIdx.x] = 0;   // becomes the running offset inside the reserved range
    }
    __syncthreads();
    // pass 2: the same draws again, each stored at its reserved slot
    for (long long p = first; p < p1; p += stride) {
        uint64_t w[2];
        aps_philox2x64((u64)p, aps_ctr1((u64)a.step, APS_DOM_RESAMPLE, 0), key, &w[0], &w[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (2 * p + h < nd) {
                const u64 tau = floor_uq53(aps_u53(w[h]), Q);
                const int o = owner_of(tau);
                const unsigned long long pos = s_base[o] + atomicAdd(&s_cnt[o], 1u);
                char *mb = reinterpret_cast<char *>(c.peers->mail[o]);
                if ((long long)pos < c.recv_cap) reinterpret_cast<u64 *>(mb + buf_off)[pos] = tau - off[o];
                else c.st->err = APS_ERR_INVALID;
            }
        }
    }
}

// all ranks: the routed draws (plain peer stores + peer atomics of k_multi_route, which is complete on
// this rank by stream order) are published with a release post; returns when every rank has done so
__global__ void __launch_bounds__(32) k_route_barrier(const __grid_constant__ DevCtx c, const long long s, const StepPlan *plan) {
    __shared__ u64 s_v[1];
    __shared__ u64 s_t[APS_MAX_RANKS][4];
    if (!plan->resampled || plan->err) return;   // identical on every rank
    if (threadIdx.x == 0) s_v[0] = 0;
    const bool ok = block_exchange(c, 9, step_seq(c, s), s_v, 1, s_t);
    if (!ok && threadIdx.x == 0) c.st->err = APS_ERR_COMM;
}

// owner side: the draws this rank received, histogrammed into the offspring counts of their parents
__global__ void __launch_bounds__(APS_K1_THREADS) k_multi_search_recv(const __grid_constant__ MultiArgs a,
                                                                     const __grid_constant__ DevCtx c, const long long s) {
    if (!a.plan->resampled || a.plan->err) return;
    long long n = (long long)__ldcg(c.recv_cnt + s);
    if (n > c.recv_cap) n = c.recv_cap;
    for (long long k = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; k < n; k += (long long)gridDim.x * APS_K1_THREADS) {
        const u64 tau = __ldcg(c.recv + k);
        atomicAdd(&a.counts[multi_find(a, tau)], 1);
    }
}

// per-tile totals of the offspring counts (one tile per block, coalesced)
__global__ void __launch_bounds__(APS_THREADS) k_tile_counts(const __grid_constant__ MultiArgs a) {
    __shared__ u64 red[APS_WARPS];
    if (!a.plan->resampled || a.plan->err) return;
    const long long base = (long long)blockIdx.x * APS_TILE;
    u64 sum = 0;
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r) {
        const long long j = base + r * APS_THREADS + threadIdx.x;
        if (j < a.N) sum += (u64)a.counts[j];
    }
    sum = block_sum_u64<APS_WARPS>(sum, red);
    if (threadIdx.x == 0) a.tile_count[blockIdx.x] = (int)sum;
}

// exclusive scan of the per-tile offspring counts (one block). Sharded: the ranks exchange their
// offspring totals; the children of this rank's parents start after those of the lower ranks.
__global__ void __launch_bounds__(APS_THREADS) k_scan_tile_counts(const __grid_constant__ MultiArgs a, long long *total_out,
                                                                  const __grid_constant__ DevCtx c, const long long s) {
    __shared__ u64 red[APS_WARPS];
    __shared__ u64 s_v[1];
    __shared__ u64 s_t[APS_MAX_RANKS][4];
    if (!a.plan->resampled || a.plan->err) return;  // the plan is identical on every rank
    const long long nt = a.num_tiles;
    const long long per = (nt + APS_THREADS - 1) / APS_THREADS;
    const long long lo = (long long)threadIdx.x * per;
    const long long hi = lo + per < nt ? lo + per : nt;
    u64 acc = 0;
    for (long long k = lo; k < hi; ++k) acc += (u64)a.tile_count[k];
    u64 tot;
    u64 run = block_excl_scan_u64<APS_WARPS>(acc, red, &tot);
    for (long long k = lo; k < hi; ++k) {
        a.tile_cprefix[k] = (int)run;
        run += (u64)a.tile_count[k];
    }
    if (threadIdx.x == 0 && total_out) *total_out = (long long)tot;
    if (a.child_off) {
        if (threadIdx.x == 0) s_v[0] = tot;
        const bool ok = block_exchange(c, 3, step_seq(c, s), s_v, 1, s_t);
        if (threadIdx.x == 0) {
            u64 off = 0;
            for (int r = 0; r < c.rank; ++r) off += s_t[r][0];
            *a.child_off = (long long)off;
            if (!ok) c.st->err = APS_ERR_COMM;
        }
    }
}

// offspring counts -> sorted ancestor indices (same expand machinery as k_resample). Sharded:
// parent ids and child slots are global, children are scattered to the rank that owns their slot.
__global__ void __launch_bounds__(APS_THREADS) k_expand_counts(const __grid_constant__ MultiArgs a, int32_t *__restrict__ anc_out,
                                                               const int identity_if_not_resampled,
                                                               const __grid_constant__ DevCtx c, const long long fat_step) {
    __shared__ u64 red[APS_WARPS];
    __shared__ __align__(16) int own[APS_OWN_WORDS(APS_CAP)];
    __shared__ int wmax[APS_WARPS];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * APS_TILE;
    const bool multi = a.child_off != nullptr;
    const long long slot0 = multi ? c.slot0 : 0;
    if (!a.plan->resampled || a.plan->err) {
        if (identity_if_not_resampled) {
#pragma unroll
            for (int r = 0; r < APS_IPT; ++r) {
                const long long i = base + r * APS_THREADS + tid;
                if (i < a.N) anc_out[i] = (int32_t)(slot0 + i);
            }
        }
        return;
    }
    const long long i0 = base + (long long)tid * APS_IPT;
    int khi[APS_IPT];
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r) khi[r] = (i0 + r < a.N) ? a.counts[i0 + r] : 0;
#pragma unroll
    for (int r = 1; r < APS_IPT; ++r) khi[r] += khi[r - 1];
    u64 tot;
    const int kA = a.tile_cprefix[blockIdx.x] + (multi ? (int)*a.child_off : 0);
    const int excl = (int)block_excl_scan_u64<APS_WARPS>((u64)khi[APS_IPT - 1], red, &tot) + kA;
    const int kB = kA + (int)tot;
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r) khi[r] += excl;
    AncDst dst;
    dst.base = anc_out;
    dst.peers = (multi && !(c.dbg & 8)) ? c.peers : nullptr;
    dst.slab_off = multi ? anc_out - c.anc : 0;
    dst.nl = (int)a.N;
    dst.lo = (int)slot0;
    expand_tile_general<APS_THREADS, APS_IPT, APS_CPT>(khi, excl, kA, kB, (int)(slot0 + base), dst, own, wmax, make_fat_sink(c, fat_step, multi));
    const long long n = a.plan->n;
    if (identity_if_not_resampled && blockIdx.x == gridDim.x - 1 && tid == 0) {  // reference particle: globally last slot
        if (!multi && n < a.N) anc_out[a.N - 1] = (int32_t)(a.N - 1);
        if (multi && n < c.Ng && c.rank == c.world - 1) anc_out[a.N - 1] = (int32_t)(c.Ng - 1);
    }
}

// residual stage 1: deterministic copies d_j = floor(n q_j / Q) and raw residuals n q_j - d_j Q
struct ResidualState {
    u64 sum_d;         // deterministic copies of this rank's parents (atomic)
    long long n_rest;  // Rc = n - (deterministic copies of all ranks)
    long long n_det;   // deterministic copies as a signed count (output offset of the residual draws)
    int shift;         // ceil_log2(Rc + 1)
    unsigned done_ctr;
    u64 q_local;       // sharded: residual-weight total of this rank ...
    u64 q_off;         // ... and of the lower ranks
};

__global__ void __launch_bounds__(APS_THREADS) k_residual_split(const __grid_constant__ MultiArgs a, const u64 *__restrict__ q,
                                                                u64 *__restrict__ rq, ResidualState *rs) {
    __shared__ u64 red[APS_WARPS];
    if (!a.plan->resampled || a.plan->err) return;
    const u64 Q = a.plan->Q;
    const int n = (int)a.plan->n;
    const double ratio = a.plan->ratio;  // 2^24 n / Q
    const int guard = a.plan->guard;
    const long long base = (long long)blockIdx.x * APS_TILE;
    u64 sd = 0;
#pragma unroll 4
    for (int r = 0; r < APS_IPT; ++r) {
        const long long j = base + r * APS_THREADS + threadIdx.x;
        if (j < a.N) {
            const u64 qj = q[j];
            // F = first i in [0,n] with !(i < n && i Q <= qj n); d = F - 1, or n when qj == Q
            bool u = false;
            int F = first_above_est((double)qj * ratio, n, guard, &u);
            if (u) F = first_above_exact(F, qj, Q, 0ull, n);
            int d = F - 1;
            if (F == n && le_128(mul_64_64((u64)n, Q), mul_64_64(qj, (u64)n))) d = n;
            a.counts[j] = d;
            rq[j] = qj * (u64)n - (u64)d * Q;  // < Q: exact modulo 2^64
            sd += (u64)d;
        }
    }
    sd = block_sum_u64<APS_WARPS>(sd, red);
    if (threadIdx.x == 0) {
        a.tile_count[blockIdx.x] = (int)sd;
        atomicAdd(&rs->sum_d, sd);
        const unsigned ticket = ticket_release(&rs->done_ctr, 1u);
        if (ticket == gridDim.x - 1 && !a.child_off) {  // sharded: k_residual_exchange<0> combines the ranks
            fence_acq_rel_gpu();
            const u64 tot = atomicAdd(&rs->sum_d, 0ull);
            const long long rc = (long long)n - (long long)tot;
            rs->n_rest = rc;
            rs->n_det = (long long)tot;
            rs->shift = aps_ceil_log2((uint64_t)rc + 1);
        }
    }
}

// residual stage 2: shifted residual weights, their tile totals and exclusive tile prefix
__global__ void __launch_bounds__(APS_THREADS) k_residual_weights(const __grid_constant__ MultiArgs a, u64 *__restrict__ rq,
                                                                  const ResidualState *rs, u64 *tile_sum, u64 *tile_prefix,
                                                                  StepPlan *wplan, unsigned *done_ctr, int *err_out) {
    __shared__ u64 red[APS_WARPS];
    __shared__ unsigned s_last;
    if (!a.plan->resampled || a.plan->err) return;
    if (rs->n_rest <= 0) return;
    const int sh = rs->shift;
    const long long base = (long long)blockIdx.x * APS_TILE;
    u64 s0 = 0;
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r) {
        const long long j = base + r * APS_THREADS + threadIdx.x;
        if (j < a.N) {
            const u64 v = rq[j] >> sh;
            rq[j] = v;
            s0 += v;
        }
    }
    s0 = block_sum_u64<APS_WARPS>(s0, red);
    if (threadIdx.x == 0) {
        tile_sum[blockIdx.x] = s0;
        const unsigned ticket = ticket_release(done_ctr, 1u);
        s_last = (ticket == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    fence_acq_rel_gpu();
    const long long nt = a.num_tiles;
    const long long per = (nt + APS_THREADS - 1) / APS_THREADS;
    const long long lo = (long long)threadIdx.x * per;
    const long long hi = lo + per < nt ? lo + per : nt;
    u64 acc = 0;
    for (long long k = lo; k < hi; ++k) acc += __ldcg(&tile_sum[k]);
    u64 tot;
    u64 run = block_excl_scan_u64<APS_WARPS>(acc, red, &tot);
    for (long long k = lo; k < hi; ++k) {
        tile_prefix[k] = run;
        run += __ldcg(&tile_sum[k]);
    }
    if (threadIdx.x == 0) {
        if (a.child_off) {  // sharded: k_residual_exchange<1> combines the ranks
            const_cast<ResidualState *>(rs)->q_local = tot;
        } else {
            wplan->Q = tot;
            if (tot == 0 && err_out) *err_out = APS_ERR_WEIGHTS;
        }
    }
}

// sharded residual resampling, one block. STAGE 0: exchange the deterministic-copy counts -> the
// number of residual draws and the residual shift. STAGE 1: exchange the residual-weight totals
// -> their global total (wplan->Q) and this rank's offset.
template <int STAGE>
__global__ void __launch_bounds__(32) k_residual_exchange(const __grid_constant__ DevCtx c, const long long s,
                                                          const StepPlan *plan, ResidualState *rs, StepPlan *wplan) {
    __shared__ u64 s_v[1];
    __shared__ u64 s_t[APS_MAX_RANKS][4];
    if (!plan->resampled || plan->err) return;
    if (STAGE == 1 && rs->n_rest <= 0) return;  // identical on every rank
    if (threadIdx.x == 0) s_v[0] = STAGE == 0 ? rs->sum_d : rs->q_local;
    const bool ok = block_exchange(c, 4 + STAGE, step_seq(c, s), s_v, 1, s_t);
    if (threadIdx.x == 0) {
        u64 tot = 0, off = 0;
        for (int r = 0; r < c.world; ++r) {
            if (r < c.rank) off += s_t[r][0];
            tot += s_t[r][0];
        }
        if (STAGE == 0) {
            const long long rc = plan->n - (long long)tot;
            rs->n_rest = rc;
            rs->n_det = (long long)tot;
            rs->shift = aps_ceil_log2((uint64_t)rc + 1);
        } else {
            wplan->Q = tot;
            rs->q_off = off;
            if (tot == 0) c.st->err = APS_ERR_WEIGHTS;
            if (c.rank_woff) {   // the residual draws are routed by the residual-weight prefix of the ranks
                u64 run = 0;
                for (int r = 0; r < c.world; ++r) {
                    c.rank_woff[s * (APS_MAX_RANKS + 1) + r] = run;
                    run += s_t[r][0];
                }
                c.rank_woff[s * (APS_MAX_RANKS + 1) + c.world] = run;
            }
        }
        if (!ok) c.st->err = APS_ERR_COMM;
    }
}

// ---------------------------------------------------------------- categorical draw (PGAS ancestor, final pick)
// lw_i for the PGAS ancestor weights: log f(X_ref[c-1] | X_i[c-2]) + logW_i   (src/pgas.jl:26-46)
// xpp: states of time s-1 (= c-2); anc_cur: ancestors of set s (= c-1)
template <int D>
__device__ __forceinline__ double pgas_logweight(const DevCtx &c, long long s, long long i,
                                                 const double *__restrict__ xpp, const int32_t *__restrict__ anc_cur) {
    const long long N = c.NS;
    long long a = anc_cur[i];  // global index of the parent in set s-1
    const double *xsrc = xpp;
    if (c.world > 1) {
        const unsigned al = (unsigned)(a - c.slot0);
        if (al < (unsigned)c.N) {
            a = al;  // own shard
        } else {     // the parent lives on a peer: read its state over NVLink
            const int owner = (int)((unsigned)a / (unsigned)c.N);
            a -= (long long)owner * c.N;
            xsrc = c.peers->x[owner] + (xpp - c.x);
        }
    }
    double xp[D], xr[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        xp[k] = xsrc[(long long)k * N + a];
        xr[k] = c.ref[(s - 1) * D + k];                                   // X_ref[c-1], c = s+1
    }
    return aps_trans_logpdf<D>(&c.md, xp, xr) + c.logw[i];
}

__device__ __forceinline__ bool pgas_active(const DevCtx &c, long long s) {
    // update_ref! (src/pgas.jl:113-128) runs inside resample_propagate! only when resampling
    // happens, the reference counter c = s+1 is > 2 and the model is not done (c <= T).
    return c.sampler == APS_PGAS && c.sp->has_ref && s >= 2 && s <= c.T - 1 && c.plan[s].resampled &&
           !c.plan[s].err;
}

template <int D>
__global__ void __launch_bounds__(APS_K1_THREADS) k_pgas_max(const __grid_constant__ DevCtx c, const long long s,
                                                          const double *__restrict__ xpp,
                                                          const int32_t *__restrict__ anc_cur, int32_t *anc_out) {
    __shared__ u64 red[APS_K1_THREADS / 32];
    if (!pgas_active(c, s)) return;
    u64 bmax = 0;
    unsigned bad = 0;
    for (long long i = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; i < c.N;
         i += (long long)gridDim.x * APS_K1_THREADS) {
        const double lw = pgas_logweight<D>(c, s, i, xpp, anc_cur);
        if (lw != lw) bad = 1;
        else {
            const u64 e = aps_encode_ordered(lw);
            bmax = e > bmax ? e : bmax;
        }
    }
    bmax = block_max_u64<APS_K1_THREADS / 32>(bmax, red);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (bmax) atomicMax(&c.acc[s].sel_max_enc, bmax);
        if (bad) atomicOr(&c.acc[s].bad, 2u);
    }
}

// find the first element j of a tile with (prefix + inclusive_sum_j) > tau; all threads return it
// (or -1). Each thread supplies its 8 consecutive weights.
__device__ __forceinline__ int tile_find_first_above(const u64 *w8, u64 prefix, u64 tau, u64 *red, int *s_found) {
    u64 cum[APS_IPT];
    cum[0] = w8[0];
#pragma unroll
    for (int r = 1; r < APS_IPT; ++r) cum[r] = cum[r - 1] + w8[r];
    u64 tot;
    const u64 excl = block_excl_scan_u64<APS_WARPS>(cum[APS_IPT - 1], red, &tot) + prefix;
    if (threadIdx.x == 0) *s_found = 0x7fffffff;
    __syncthreads();
    int mine = 0x7fffffff;
#pragma unroll
    for (int r = APS_IPT - 1; r >= 0; --r)
        if (excl + cum[r] > tau) mine = threadIdx.x * APS_IPT + r;
    if (mine != 0x7fffffff) atomicMin(s_found, mine);
    __syncthreads();
    const int f = *s_found;
    return f == 0x7fffffff ? -1 : f;
}

// PGAS: tile totals of the quantised ancestor weights; the last block locates the drawn tile,
// rescans it and rewires the reference's ancestor pointer (the splice of src/pgas.jl:125-127).
// Sharded: the ranks exchange the maximum (every block) and their weight totals (last block);
// the rank whose weight range holds the draw writes the ancestor into the last rank's store.
template <int D>
__global__ void __launch_bounds__(APS_THREADS) k_pgas_select(const __grid_constant__ DevCtx c, const long long s,
                                                             const double *__restrict__ xpp,
                                                             const int32_t *__restrict__ anc_cur, int32_t *anc_out) {
    __shared__ u64 red[APS_THREADS / 32];
    __shared__ unsigned s_last;
    __shared__ long long s_tile;
    __shared__ u64 s_pref, s_tau;
    __shared__ int s_found;
    __shared__ u64 s_m[APS_MAX_RANKS][4];
    __shared__ u64 s_v[2];
    if (!pgas_active(c, s)) return;  // identical on every rank
    const long long N = c.N;
    const bool multi = c.world > 1;
    StepAcc *acc = &c.acc[s];
    u64 menc = acc->sel_max_enc;
    unsigned bad_in = acc->bad & 2u;
    if (multi) {  // all-reduce(max) of the ancestor log-weights: k_pgas_max of this rank is complete
        if (threadIdx.x == 0) {
            s_v[0] = menc;
            s_v[1] = bad_in;
        }
        __syncthreads();
        if (blockIdx.x == 0) mail_post(c.peers, c.rank, c.world, 6, step_seq(c, s), s_v, 2);
        bool ok = true;
        if (!(c.dbg & 1)) ok = mail_wait(c.peers, c.rank, c.world, 6, step_seq(c, s), s_m, 2, nullptr, &c.st->err);
        if (!__syncthreads_and(ok ? 1 : 0) && threadIdx.x == 0) c.st->err = APS_ERR_COMM;
        menc = 0;
        for (int r = 0; r < c.world; ++r) {
            menc = s_m[r][0] > menc ? s_m[r][0] : menc;
            bad_in |= (unsigned)s_m[r][1];
        }
        __syncthreads();
    }
    const double M = aps_decode_ordered(menc);
    const double scale = aps_pow2i(c.S);
    const long long base = (long long)blockIdx.x * APS_TILE;
    u64 s0 = 0;
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r) {
        const long long i = base + r * APS_THREADS + threadIdx.x;
        if (i < N) {
            const double e = aps_exp(pgas_logweight<D>(c, s, i, xpp, anc_cur) - M);
            s0 += (e > 0.0) ? (u64)__double2ull_rz(e * scale) : 0ull;
        }
    }
    s0 = block_sum_u64<APS_WARPS>(s0, red);
    // tile_s1 is free at this point of the step (the plan of s is already written)
    if (threadIdx.x == 0) {
        c.tile_s1[blockIdx.x] = s0;
        const unsigned ticket = ticket_release(&acc->sel_done_ctr, 1u);
        s_last = (ticket == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    fence_acq_rel_gpu();
    const long long nt = c.num_tiles;
    const long long per = (nt + APS_THREADS - 1) / APS_THREADS;
    const long long lo = (long long)threadIdx.x * per;
    const long long hi = lo + per < nt ? lo + per : nt;
    u64 a0 = 0;
    for (long long k = lo; k < hi; ++k) a0 += __ldcg(&c.tile_s1[k]);
    u64 Qloc;
    u64 run = block_excl_scan_u64<APS_WARPS>(a0, red, &Qloc);
    u64 Qs = Qloc, off = 0;
    if (multi) {  // all-gather of the ancestor-weight totals, combined in rank order
        if (threadIdx.x == 0) s_v[0] = Qloc;
        const bool ok = block_exchange(c, 7, step_seq(c, s), s_v, 1, s_m);
        Qs = 0;
        for (int r = 0; r < c.world; ++r) {
            if (r < c.rank) off += s_m[r][0];
            Qs += s_m[r][0];
        }
        if (!ok && threadIdx.x == 0) c.st->err = APS_ERR_COMM;
    }
    if (threadIdx.x == 0) {
        s_tile = -1;
        uint64_t w0, w1;
        aps_philox2x64(0, aps_ctr1((u64)s, APS_DOM_PGAS, 0), c.sp->key, &w0, &w1);
        s_tau = floor_uq53(aps_u53(w0), Qs) - off;  // relative to this rank's weight range
        const double Mv = M;
        if (bad_in || Qs == 0 || !(Mv == Mv) || Mv == aps_bits2d(0xFFF0000000000000ULL)) c.st->err = APS_ERR_WEIGHTS;
    }
    __syncthreads();
    const u64 tau = s_tau;
    if (Qs == 0 || tau >= Qloc) return;  // (unsigned: also rejects draws below this rank's range)
    for (long long k = lo; k < hi; ++k) {
        const u64 t = __ldcg(&c.tile_s1[k]);
        if (run <= tau && tau < run + t) {
            s_tile = k;
            s_pref = run;
        }
        run += t;
    }
    __syncthreads();
    const long long tile = s_tile;
    if (tile < 0) return;
    u64 w8[APS_IPT];
#pragma unroll
    for (int r = 0; r < APS_IPT; ++r) {
        const long long i = tile * APS_TILE + (long long)threadIdx.x * APS_IPT + r;
        w8[r] = 0;
        if (i < N) {
            const double e = aps_exp(pgas_logweight<D>(c, s, i, xpp, anc_cur) - M);
            w8[r] = (e > 0.0) ? (u64)__double2ull_rz(e * scale) : 0ull;
        }
    }
    const int f = tile_find_first_above(w8, s_pref, tau, red, &s_found);
    if (threadIdx.x == 0 && f >= 0) {
        // the reference particle is the globally last slot: last slot of the last rank's store
        int32_t *dst = multi ? c.peers->anc[c.world - 1] + (anc_out - c.anc) + (N - 1) : anc_out + (N - 1);
        *dst = (int32_t)(c.slot0 + tile * APS_TILE + f);
    }
}

// final pick over the weights of the final set: rand(pc.rng, pc) (src/container.jl:33-36).
// One block. If the last decision point resampled, the weights are uniform. picked_slot is a
// GLOBAL slot index. Sharded: every rank looks for the draw in its own weight range and the ranks
// exchange their candidates (pick_seq: number of picks made on this handle so far).
__global__ void __launch_bounds__(APS_THREADS) k_pick(const __grid_constant__ DevCtx c, const long long plan_idx,
                                                      const long long step_ctr, const unsigned dom, const u64 pick_seq) {
    __shared__ u64 red[APS_THREADS / 32];
    __shared__ long long s_tile;
    __shared__ u64 s_pref;
    __shared__ int s_found;
    __shared__ u64 s_v[1];
    __shared__ u64 s_t[APS_MAX_RANKS][4];
    const long long N = c.N;
    const bool multi = c.world > 1;
    const StepPlan &p = c.plan[plan_idx];
    uint64_t w0, w1;
    aps_philox2x64(0, aps_ctr1((u64)step_ctr, dom, 0), c.sp->key, &w0, &w1);
    const u64 U = aps_u53(w0);
    if (p.resampled) {
        if (threadIdx.x == 0) {
            const u64 Q = (u64)c.Ng << c.S;
            long long slot = (long long)(floor_uq53(U, Q) >> c.S);
            if (slot >= c.Ng) slot = c.Ng - 1;
            c.st->picked_slot = slot;
        }
        return;
    }
    const u64 off = multi ? c.acc[plan_idx].rank_off : 0ull;
    const u64 Qloc = multi ? c.acc[plan_idx].tot[0] : p.Q;
    const u64 tau = floor_uq53(U, p.Q) - off;  // relative to this rank's weight range
    const long long nt = c.num_tiles;
    if (threadIdx.x == 0) s_tile = -1;
    __syncthreads();
    if (tau < Qloc) {
        for (long long k = threadIdx.x; k < nt; k += APS_THREADS) {
            const u64 pre = c.tile_prefix[k];
            const u64 t = c.tile_sum[k];
            if (pre <= tau && tau < pre + t) {
                s_tile = k;
                s_pref = pre;
            }
        }
    }
    __syncthreads();
    const long long tile = s_tile;
    long long cand = -1;
    if (tile >= 0) {
        u64 w8[APS_IPT];
#pragma unroll
        for (int r = 0; r < APS_IPT; ++r) {
            const long long i = tile * APS_TILE + (long long)threadIdx.x * APS_IPT + r;
            w8[r] = i < N ? c.q[i] : 0ull;
        }
        const int f = tile_find_first_above(w8, s_pref, tau, red, &s_found);
        if (f >= 0) cand = c.slot0 + tile * APS_TILE + f;
    }
    if (multi) {
        if (threadIdx.x == 0) s_v[0] = (u64)(cand + 1);
        const bool ok = block_exchange(c, 8, pick_seq, s_v, 1, s_t);
        cand = -1;
        for (int r = 0; r < c.world; ++r)
            if (s_t[r][0]) cand = (long long)s_t[r][0] - 1;
        if (!ok) cand = -1;
    }
    if (threadIdx.x == 0 && cand >= 0) c.st->picked_slot = cand;
}

// state / ancestor of a GLOBAL slot index (sharded: through the owner's peer-mapped store)
__device__ __forceinline__ double load_state(const DevCtx &c, long long slab, int k, long long g) {
    const long long off = slab * (long long)c.d * c.NS + (long long)k * c.NS;
    if (c.world == 1) return c.x[off + g];
    const int owner = (int)(g / c.N);
    return c.peers->x[owner][off + (g - (long long)owner * c.N)];
}
__device__ __forceinline__ long long load_anc(const DevCtx &c, long long slab, long long g) {
    if (c.world == 1) return c.anc[slab * c.NS + g];
    const int owner = (int)(g / c.N);
    return c.peers->anc[owner][slab * c.NS + (g - (long long)owner * c.N)];
}

// trajectory of one final-set slot (global index): T pointer hops through the ancestor store
__global__ void k_backtrace(const __grid_constant__ DevCtx c, const long long slot_in, double *__restrict__ traj) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long T = c.T;
    const int D = c.d;
    const long long slot = slot_in >= 0 ? slot_in : c.st->picked_slot;
    if (slot < 0 || slot >= c.Ng) return;
    long long j = load_anc(c, T % c.anc_slabs, slot);
    for (long long t = T; t >= 1; --t) {
        for (int k = 0; k < D; ++k) traj[(t - 1) * D + k] = load_state(c, (t - 1) % c.x_slabs, k, j);
        if (t > 1) j = load_anc(c, (t - 1) % c.anc_slabs, j);
    }
}

// final particle set (this rank's slots) as N x d row-major: x_T[anc_{T+1}[i]]   (collect(pc), src/smc.jl:56)
__global__ void __launch_bounds__(APS_THREADS) k_gather_final(const __grid_constant__ DevCtx c, double *__restrict__ out) {
    const long long N = c.N, T = c.T;
    const int D = c.d;
    const int32_t *anc = c.anc + (T % c.anc_slabs) * c.NS;
    for (long long i = (long long)blockIdx.x * APS_THREADS + threadIdx.x; i < N;
         i += (long long)gridDim.x * APS_THREADS) {
        const long long a = anc[i];
        for (int k = 0; k < D; ++k) out[i * D + k] = load_state(c, (T - 1) % c.x_slabs, k, a);
    }
}

// ---------------------------------------------------------------- smoothing summaries over the genealogy
// One backward step of the weighted trajectory mean E[x_t | y_1:T] ~ sum_i W_i x_t[b_t(i)], where
// b_t(i) is the time-t ancestor of final particle i (b_T = anc_{T+1}, b_{t-1} = anc_t[b_t]).
// idx holds b_t(i) for this rank's slots on entry and b_{t-1}(i) on exit. Block partial sums are
// combined by the last block in block order, so the result does not depend on scheduling.
#define APS_SMOOTH_BLOCKS 296
__global__ void __launch_bounds__(APS_K1_THREADS) k_smooth_step(const __grid_constant__ DevCtx c, const long long t,
                                                             int32_t *__restrict__ idx, const u64 *__restrict__ q,
                                                             const int uniform, double *__restrict__ partial,
                                                             unsigned *done_ctr, double *__restrict__ mean_out) {
    __shared__ double red[APS_K1_THREADS / 32][APS_MAX_D];
    __shared__ unsigned s_last;
    const long long N = c.N, T = c.T;
    const int D = c.d;
    const StepPlan &p = c.plan[T];
    const double Qd = uniform ? (double)c.Ng : (double)p.Q;
    double acc[APS_MAX_D] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; i < N;
         i += (long long)gridDim.x * APS_K1_THREADS) {
        const long long b = t == T ? (long long)c.anc[(T % c.anc_slabs) * c.NS + i] : (long long)idx[i];
        const double w = (uniform ? 1.0 : (double)q[i]) / Qd;
        for (int k = 0; k < D; ++k) acc[k] += w * load_state(c, (t - 1) % c.x_slabs, k, b);
        idx[i] = t > 1 ? (int32_t)load_anc(c, (t - 1) % c.anc_slabs, b) : 0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < APS_MAX_D; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < APS_MAX_D) {
        double v = 0.0;
        for (int w = 0; w < APS_K1_THREADS / 32; ++w) v += red[w][threadIdx.x];
        partial[(long long)blockIdx.x * APS_MAX_D + threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = ticket_release(done_ctr + (t - 1), 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    fence_acq_rel_gpu();
    if (threadIdx.x < D) {
        double v = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(&partial[(long long)b * APS_MAX_D + threadIdx.x]);
        mean_out[(t - 1) * D + threadIdx.x] = v;
    }
}

// One backward step of "all N trajectories" (SMCSample(collect(pc), ...), src/smc.jl:56): out[i][k] =
// x_t[b_t(i)][k] for this rank's final-set slots i; idx carries b_t(i) -> b_{t-1}(i) like k_smooth_step.
__global__ void __launch_bounds__(APS_K1_THREADS) k_traj_step(const __grid_constant__ DevCtx c, const long long t,
                                                           int32_t *__restrict__ idx, double *__restrict__ out) {
    const long long N = c.N, T = c.T;
    const int D = c.d;
    for (long long i = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; i < N;
         i += (long long)gridDim.x * APS_K1_THREADS) {
        const long long b = t == T ? (long long)c.anc[(T % c.anc_slabs) * c.NS + i] : (long long)idx[i];
        for (int k = 0; k < D; ++k) out[i * D + k] = load_state(c, (t - 1) % c.x_slabs, k, b);
        idx[i] = t > 1 ? (int32_t)load_anc(c, (t - 1) % c.anc_slabs, b) : 0;
    }
}

// Stepwise container (aps_pc_*): decision point s before resample_propagate! has run on it -- every
// particle continues as itself and the weights are kept (what update_keys! leaves, container.jl:247)
__global__ void __launch_bounds__(APS_K1_THREADS) k_pc_provisional(const __grid_constant__ DevCtx c, const long long s,
                                                                int32_t *__restrict__ anc_out) {
    for (long long i = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; i < c.N;
         i += (long long)gridDim.x * APS_K1_THREADS)
        anc_out[i] = (int32_t)(c.slot0 + i);
    if (blockIdx.x == 0 && threadIdx.x == 0) c.plan[s].resampled = 0;
}

// normalised weights W_i = q_i / Q (getweights, src/container.jl:95) or 1/N after a resample
__global__ void __launch_bounds__(APS_THREADS) k_weights_out(const u64 *__restrict__ q, const StepPlan *p, long long N,
                                                             long long Ng, int S, int force_uniform,
                                                             double *__restrict__ out) {
    const bool uni = force_uniform && p->resampled;
    const double Qd = uni ? (double)((u64)Ng << S) : (double)p->Q;
    const double qu = (double)(1ull << S);
    for (long long i = (long long)blockIdx.x * APS_THREADS + threadIdx.x; i < N;
         i += (long long)gridDim.x * APS_THREADS)
        out[i] = (uni ? qu : (double)q[i]) / Qd;
}

// int32 0-based -> int64 1-based (operator boundary, Julia Vector{Int})
__global__ void __launch_bounds__(APS_THREADS) k_to_one_based(const int32_t *__restrict__ in, long long n,
                                                              long long *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * APS_THREADS + threadIdx.x; i < n;
         i += (long long)gridDim.x * APS_THREADS)
        out[i] = (long long)in[i] + 1;
}

// fills the deferred child ranges of decision point s into this rank's ancestor slab (after the
// final decision point, where no propagate kernel follows, and at the operator level)
__global__ void __launch_bounds__(APS_K1_THREADS) k_fill_fat(const __grid_constant__ DevCtx c, const long long s,
                                                          int32_t *__restrict__ anc_out, const int barrier) {
    if (barrier && c.world > 1) {
        // sharded, after the final decision point: the peers' resample kernels (which push into this
        // rank's list and scatter into its ancestor store) must be complete -- same exchange as the
        // one at the start of k_propagate
        __shared__ u64 s_w[APS_MAX_RANKS][4];
        const u64 v0 = 0;
        const u64 seq = c.sp->epoch * (u64)(c.T + 2) + (u64)s + 1;
        if (blockIdx.x == 0) mail_post(c.peers, c.rank, c.world, 2, seq, &v0, 1);
        if (!(c.dbg & 1) && !mail_wait(c.peers, c.rank, c.world, 2, seq, s_w, 1, nullptr, &c.st->err)) c.st->err = APS_ERR_COMM;
        __syncthreads();
    }
    int nfat = __ldcg(&c.fat_cnt[s]);
    if (nfat > APS_FAT_MAX) nfat = APS_FAT_MAX;
    // this rank's child slots (operator level: n_override children drawn from N weights)
    const long long lo_r = c.slot0, hi_r = c.slot0 + (c.n_override > 0 ? c.n_override : c.N);
    for (int e = 0; e < nfat; ++e) {
        const int4 fv = __ldcg(reinterpret_cast<const int4 *>(c.fat + s * APS_FAT_MAX + e));
        FatEntry f;
        f.lo = fv.x;
        f.hi = fv.y;
        f.parent = fv.z;
        const long long lo = f.lo > lo_r ? f.lo : lo_r, hi = f.hi < hi_r ? f.hi : hi_r;
        for (long long g = lo + (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; g < hi;
             g += (long long)gridDim.x * APS_K1_THREADS)
            anc_out[g - lo_r] = f.parent;
    }
}

// second half of the L2 flush of aps_bench_resample: streaming reads replace the dirty lines the
// memset left in L2 by clean ones, so the timed kernel does not pay for their write-back
__global__ void __launch_bounds__(APS_K1_THREADS) k_read_flush(const uint4 *__restrict__ buf, long long n16, unsigned *sink) {
    unsigned acc = 0;
    for (long long i = (long long)blockIdx.x * APS_K1_THREADS + threadIdx.x; i < n16; i += (long long)gridDim.x * APS_K1_THREADS) {
        const uint4 v = __ldcs(buf + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x9E3779B9u) *sink = acc;
}

// synthetic integer weights for aps_bench_resample (hash of the index, roughly log-normal spread)
__global__ void __launch_bounds__(APS_THREADS) k_bench_weights(u64 *__restrict__ q, long long n, int S, u64 seed) {
    for (long long i = (long long)blockIdx.x * APS_THREADS + threadIdx.x; i < n;
         i += (long long)gridDim.x * APS_THREADS) {
        uint64_t w0, w1;
        aps_philox2x64((u64)i, 0x42, seed, &w0, &w1);
        double z0, z1;
        aps_normal_pair(w0, w1, &z0, &z1);
        const double e = aps_exp(-0.5 * z0 * z0);
        q[i] = (u64)__double2ull_rz(e * aps_pow2i(S));
    }
}
