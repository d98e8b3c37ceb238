// aps_fused.cuh -- the whole particle sweep as ONE persistent cooperative kernel.
//
//   sweep! (src/container.jl:316-363): T+1 rounds of resample_propagate! (:171-251) -> logZ ->
//   reweight! (:259-302) -> logZ, with advance! (src/pgas.jl:53-89), the systematic / stratified
//   resamplers (src/resampling.jl:98-183), the ESS rule (:193-204) and the PGAS ancestor draw
//   (src/pgas.jl:113-128) -- everything aps_sweep runs -- inside a single launch.
//
// Every CTA owns one contiguous chunk of particle slots for the whole sweep (grid = co-resident
// CTAs, one per SM). A time step is three phases separated by TWO grid-wide exchanges:
//
//   A  propagate + reweight my slots (gather parent state, transition draw, observation density),
//      block maximum of the new log-weights                        -> exchange 1: all-reduce(max)
//   B  canonical integer weights q = floor(exp(logw - M) 2^S) of my slots, chunk / tile totals
//                                                                   -> exchange 2: all-gather(totals)
//   C  every CTA derives the same plan (logZ, ESS, decision, offset) from the totals and then
//      resolves the ancestors OF ITS OWN SLOTS ("pull"): it locates the parents whose cumulative
//      weight range covers its children -- its own chunk and, usually, a neighbour's -- scans their
//      integer weights and expands them into its slots.
//
// Pull instead of push is what removes the third exchange of the three-kernel path: the ancestors a
// CTA needs in phase A of the next step are the ones it has just written itself, so nothing has to
// be globally visible between C and A. It also makes the cost independent of weight degeneracy:
// a parent with a million children costs each CTA one marker, no "fat parent" lists are needed.
//
// An exchange is an array with one entry per CTA plus an arrival counter: a CTA stores its entry,
// fences, and bumps the counter (one L2 atomic); ONE thread per CTA polls the counter, fences, and
// after a block barrier every thread reads the entries it needs straight from L2. (Measured first:
// every CTA polling every entry -- "the data is the barrier" -- put 148 x 148 polling loads on the
// lines being written and cost 4.4 us (exchange 1) / 10 us (exchange 2) per step at N = 1e6.)
//
// Results are bit-identical to the three-kernel path and to the oracle: all sums that feed a
// comparison are integer sums (associative), so the chunking does not matter.
#pragma once
#include "aps_kernels.cuh"

#ifndef APS_FUSED_MAX_THREADS
#define APS_FUSED_MAX_THREADS 768   // 80 registers per thread: the propagate loop does not spill (1024 -> 64 registers does)
#endif
#define APS_FUSED_MAX_CTAS 304      // >= 2 x 148 SMs
#define APS_FUSED_CPT 16            // child slots per thread in one expand pass
#define APS_FUSED_IPT 4             // parents per lane in one scan sub-tile
#define APS_FUSED_GRP 4             // chunks whose sub-tile prefixes are staged in shared memory at once
#define APS_FUSED_SUB (32 * APS_FUSED_IPT)   // parents per sub-tile: one warp scans one sub-tile without any block barrier

struct FusedArgs {
    ulonglong2 *ex_max;     // [G]     exchange 1: (encoded max log-weight or ~0 for NaN, seq)
    ulonglong2 *ex_pmax;    // [G]     PGAS: (encoded max ancestor log-weight, unused); published before ex_max
    ulonglong2 *ex_tot;     // [G][3]  exchange 2: (Q, seq), (Q1, Q2), (PGAS ancestor-weight total, unused)
    u64 *sub_prefix;        // [G][nsub + 1] exclusive integer-weight prefix of every 128-parent sub-tile inside its
                            //         chunk (entry nsub = chunk total), written by the owner in phase B
    u64 *qp;                // [2][NS] PGAS: integer ancestor weights, double-buffered by step parity (the owner of
                            //         the reference slot scans another CTA's chunk in phase C while that CTA may
                            //         already be writing the next step's values in phase A)
    u64 *ctr;               // [2]     arrival counters of the two exchanges, zeroed before every launch
    u64 *dbg;               // [G][8]  diagnostics (APS_DEBUG_MULTI & 32): ns per phase and CTA, summed over the sweep
    int chunk;              // slots per CTA (multiple of APS_FUSED_SUB)
    int nsub;               // sub-tiles per chunk
};

// ---------------------------------------------------------------- block primitives for any warp count <= 32
// smem: 33 u64. All threads get the result.
__device__ __forceinline__ u64 fblk_sum(u64 v, u64 *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_u64(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        u64 t = lane < nw ? sm[lane] : 0ull;
        t = warp_sum_u64(t);
        if (lane == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}
__device__ __forceinline__ u64 fblk_max(u64 v, u64 *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max_u64(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        u64 t = lane < nw ? sm[lane] : 0ull;
        t = warp_max_u64(t);
        if (lane == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}
// exclusive scan of one u64 per thread; *total = block sum
__device__ __forceinline__ u64 fblk_excl_scan(u64 v, u64 *sm, u64 *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const u64 inc = warp_incl_scan_u64(v, lane);
    __syncthreads();
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const u64 w = lane < nw ? sm[lane] : 0ull;
        const u64 wi = warp_incl_scan_u64(w, lane);
        sm[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    *total = sm[32];
    return sm[warp] + inc - v;
}

// ---------------------------------------------------------------- exchange entries (device scope)
__device__ __forceinline__ void st_pair_rel_gpu(ulonglong2 *p, u64 v, u64 seq) {
    asm volatile("{ .reg .b128 t; mov.b128 t, {%1, %2}; st.release.gpu.global.b128 [%0], t; }" ::"l"(p), "l"(v), "l"(seq)
                 : "memory");
}
__device__ __forceinline__ ulonglong2 ld_pair_acq_gpu(const ulonglong2 *p) {
    ulonglong2 r;
    asm volatile("{ .reg .b128 t; ld.acquire.gpu.global.b128 t, [%2]; mov.b128 {%0, %1}, t; }"
                 : "=l"(r.x), "=l"(r.y)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ ulonglong2 ld_pair_rlx_gpu(const ulonglong2 *p) {
    ulonglong2 r;
    asm volatile("{ .reg .b128 t; ld.relaxed.gpu.global.b128 t, [%2]; mov.b128 {%0, %1}, t; }"
                 : "=l"(r.x), "=l"(r.y)
                 : "l"(p)
                 : "memory");
    return r;
}
// spin until the entry carries sequence number >= seq; ~3 s budget (a CTA that never arrives means
// the launch was not co-resident -- a bug, not a runtime condition; fail loudly instead of hanging)
// The poll itself is relaxed (an acquire load costs a whole-L1 invalidate per iteration); the caller
// issues ONE fence after its last successful poll, which together with the block barrier that
// follows orders every later read of the CTA after the producers' release stores.
__device__ __forceinline__ ulonglong2 wait_pair(const ulonglong2 *p, u64 seq, int *err) {
    ulonglong2 r = ld_pair_rlx_gpu(p);
    if (r.y >= seq) return r;
    const long long t0 = clock64();
    unsigned it = 0;
    while (r.y < seq) {
        if ((++it & 255u) == 0 && clock64() - t0 > 6000000000LL) {
            *err = APS_ERR_COMM;
            break;
        }
        r = ld_pair_rlx_gpu(p);
    }
    return r;
}


// arrival: (block barrier first, by the caller) entry stores by this thread, fence, one L2 atomic
__device__ __forceinline__ void ex_arrive(u64 *ctr) {
    // release: MEMBAR.ALL.GPU + RED (no sequentially consistent fence, no L1 invalidate on the producer side)
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ull) : "memory");
}
// one thread polls the counter until `target` CTAs have arrived; ~3 s budget (a CTA that never
// arrives means the launch was not co-resident -- a bug, not a runtime condition: fail loudly)
__device__ __forceinline__ void ex_wait(const u64 *ctr, u64 target, int *err) {
    const volatile u64 *p = ctr;
    if (*p < target) {
        const long long t0 = clock64();
        unsigned it = 0;
        while (*p < target) {
            if ((++it & 1023u) == 0 && clock64() - t0 > 6000000000LL) {
                *err = APS_ERR_COMM;
                break;
            }
        }
    }
    fence_acq_rel_gpu();   // acquire (with the block barrier that follows in the caller)
}

// ---------------------------------------------------------------- the sweep
// Shared accumulators (64-bit shared-memory atomics; one block barrier publishes them)
enum { FA_BMAX = 0, FA_PMAX, FA_Q1, FA_Q2, FA_QP, FA_MENC, FA_PENC, FA_GQ1, FA_GQ2, FA_N };

template <int D, int DY, int OBS, int KIND>
__global__ void __launch_bounds__(APS_FUSED_MAX_THREADS, 1) k_sweep_fused(const __grid_constant__ DevCtx c,
                                                                         const __grid_constant__ FusedArgs f) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ u64 red[33];
    __shared__ u64 s_acc[FA_N];
    __shared__ u64 s_T[APS_FUSED_MAX_CTAS];     // chunk totals as gathered
    __shared__ u64 s_E[APS_FUSED_MAX_CTAS];     // inclusive chunk ends E_k
    __shared__ u64 s_P[APS_FUSED_MAX_CTAS];     // PGAS: ancestor-weight totals of the chunks, then inclusive ends
    __shared__ int s_kend[APS_FUSED_MAX_CTAS];  // K(E_k): children below the end of chunk k
    __shared__ StepPlan s_plan;
    __shared__ u64 s_pre;
    __shared__ unsigned s_flags;                // bit 0: NaN log-weight, bit 1: NaN ancestor weight
    __shared__ int s_err, s_k0, s_found;
    int *own = reinterpret_cast<int *>(fsm);                                        // [blockDim.x * APS_FUSED_CPT]
    u64 *s_sub = reinterpret_cast<u64 *>(fsm + (size_t)blockDim.x * APS_FUSED_CPT * 4);  // [nsub + 1] my sub-tile totals
    u64 *s_pf = s_sub + (f.nsub + 1);                                                     // [GRP][nsub + 1] staged prefixes
    int *s_ks = reinterpret_cast<int *>(s_pf + APS_FUSED_GRP * (f.nsub + 1));             // [GRP][nsub] K at sub-tile ends

    const int tid = threadIdx.x, NT = blockDim.x, cta = blockIdx.x, G = gridDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = (NT + 31) >> 5;
    const long long NS = c.NS, T = c.T;
    const int N = (int)c.N;
    const int Nc = f.chunk, nsubc = f.nsub;
    const int i0 = cta * Nc < N ? cta * Nc : N;                       // my slots [i0, i1)
    const int i1 = i0 + Nc < N ? i0 + Nc : N;
    const int nsub_mine = (i1 - i0 + APS_FUSED_SUB - 1) / APS_FUSED_SUB;
    const int has_ref = c.sp->has_ref;
    const u64 key = c.sp->key;
    const u64 seq0 = c.sp->epoch * (u64)(T + 2);
    const bool pgas = c.sampler == APS_PGAS && has_ref;
    const int CAP = NT * APS_FUSED_CPT;
    const int ref_cta = (N - 1) / Nc;                                 // owner of the reference slot
    const double scale = aps_pow2i(c.S);
    u64 *my_prefix = f.sub_prefix + (long long)cta * (nsubc + 1);

    if (tid == 0) {
        s_err = 0;
        s_flags = 0;
    }
    if (tid < FA_N) s_acc[tid] = 0;

    // ---- decision point 0 (what k_init_sweep records): all log-weights are zero
    int resampled_prev = c.bare ? 1 : ((double)c.Ng <= c.ess_threshold * (double)c.Ng ? 1 : 0);
    double logz_prev = c.logN, logev = 0.0;
    int sweep_err = 0;
    if (cta == 0 && tid == 0) {
        StepPlan p;
        p.M = 0.0;
        p.logZ = c.logN;
        p.ess = (double)c.Ng;
        p.Q = (u64)c.Ng << c.S;
        p.R = 0;
        p.ratio = 0.0;
        p.roff = 0.0;
        p.n = c.Ng - (has_ref ? 1 : 0);
        p.resampled = resampled_prev;
        p.err = 0;
        p.guard = 8;
        p.pad = 0;
        c.plan[0] = p;
        c.st->err = 0;
        c.st->picked_slot = -1;
        c.st->spin[0] = c.st->spin[1] = c.st->spin[2] = c.st->spin[3] = 0;
    }
    __syncthreads();

    // diagnostics (APS_DEBUG_MULTI & 32): ns this CTA spent in A | exchange 1 | B | exchange 2 + plan | C, summed over the sweep
    const bool prof = (c.dbg & 32) && tid == 0;
    u64 tp0 = 0, tacc[5] = {0, 0, 0, 0, 0};
#define APS_FPROF(k_)                          \
    if (prof) {                                \
        const u64 n_ = global_timer_ns();      \
        tacc[k_] += n_ - tp0;                  \
        tp0 = n_;                              \
    }

    for (long long t = 1; t <= T; ++t) {
        if (prof) tp0 = global_timer_ns();
        double *__restrict__ xt = c.x + ((t - 1) % c.x_slabs) * (long long)D * NS;
        const double *xp = c.x + ((t - 2 + c.x_slabs) % c.x_slabs) * (long long)D * NS;
        const int32_t *anc_prev = c.anc + ((t - 1) % c.anc_slabs) * NS;   // ancestors of set t (written in phase C of t-1)
        int32_t *anc_out = c.anc + (t % c.anc_slabs) * NS;                // ancestors of set t+1
        const double *__restrict__ y = c.Y + (t - 1) * c.dy;
        const u64 seq = seq0 + (u64)t + 1;
        const bool reset = t == 1 || resampled_prev != 0;
        const bool pgas_step = pgas && t >= 2 && t <= T - 1;              // update_ref! can run at this decision point
        u64 *qp_t = f.qp + (t & 1) * NS;

        // =============================================================== phase A: propagate + reweight
        {
            u64 bmax = 0, pmax = 0;
            unsigned bad = 0;
            double xref[D];
            if (pgas_step) {
#pragma unroll
                for (int k = 0; k < D; ++k) xref[k] = c.ref[(t - 1) * D + k];   // X_ref[c-1], c = t+1
            }
            const int p_end = (i1 + 1) >> 1;
            double mx = aps_bits2d(0xFFF0000000000000ULL), pmx = mx;   // running maxima (-inf), encoded once after the loop
            bool any = false, pany = false;
            // order inside one iteration: ancestors, Philox rounds (integer only) while they arrive, the
            // parent-state gather, the floating-point half of the draw while THAT is in flight
            for (int p = (i0 >> 1) + tid; p < p_end; p += NT) {
                const int j0 = 2 * p;
                int2 a2 = make_int2(0, 0);
                if (t > 1) a2 = *reinterpret_cast<const int2 *>(anc_prev + j0);
                double2 lw2 = make_double2(0.0, 0.0);
                if (!reset) lw2 = *reinterpret_cast<const double2 *>(c.logw + j0);
                uint64_t w[2 * D];
                aps_pair_words<D>(key, (u64)((c.slot0 >> 1) + p), (u64)t, w);
                double xg[2][D];
                if (t > 1) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int a = j0 + h < N ? (h ? a2.y : a2.x) : 0;   // (padding slot of an odd N: entry not written)
#pragma unroll
                        for (int k = 0; k < D; ++k) xg[h][k] = __ldcg(xp + (long long)k * NS + a);
                    }
                }
                double z[2 * D];
                aps_words_to_normals<D>(w, z);
                double xo[2][D];
                double lwo[2], lpo[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = j0 + h;
                    double x[D];
                    lwo[h] = 0.0;
                    lpo[h] = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = 0.0;
                    if (i < N) {
                        const bool is_ref = has_ref && c.slot0 + i == c.Ng - 1;   // the reference keeps the globally last slot
                        if (is_ref) {
#pragma unroll
                            for (int k = 0; k < D; ++k) x[k] = c.ref[(t - 1) * D + k];
                        } else if (t == 1) {
                            aps_prior_draw<D>(&c.md, z + h * D, x);
                        } else {
                            aps_trans_draw<D>(&c.md, xg[h], z + h * D, x);
                        }
                        const double ll = aps_obs_logpdf<D, DY, OBS>(&c.md, x, y);
                        const double lw = (reset ? 0.0 : (h ? lw2.y : lw2.x)) + ll;
                        lwo[h] = lw;
                        if (lw != lw) bad = 1;
                        else {
                            mx = lw > mx ? lw : mx;
                            any = true;
                        }
                        if (pgas_step) {   // log f(X_ref[c-1] | X_i[c-2]) + logW_i   (src/pgas.jl:26-46)
                            const double lp = aps_trans_logpdf<D>(&c.md, xg[h], xref) + lw;
                            lpo[h] = lp;
                            if (lp != lp) bad |= 2u;
                            else {
                                pmx = lp > pmx ? lp : pmx;
                                pany = true;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) xo[h][k] = x[k];
                }
#pragma unroll
                for (int k = 0; k < D; ++k)
                    *reinterpret_cast<double2 *>(xt + (long long)k * NS + j0) = make_double2(xo[0][k], xo[1][k]);
                *reinterpret_cast<double2 *>(c.logw + j0) = make_double2(lwo[0], lwo[1]);
                if (pgas_step) *reinterpret_cast<double2 *>(reinterpret_cast<double *>(qp_t) + j0) = make_double2(lpo[0], lpo[1]);
            }
            if (any) bmax = aps_encode_ordered(mx);
            if (pany) pmax = aps_encode_ordered(pmx);
            // block maxima: warp shuffles, then one 64-bit shared-memory atomic per warp
            bmax = warp_max_u64(bmax);
            if (lane == 0 && bmax) atomicMax(&s_acc[FA_BMAX], bmax);
            if (pgas_step) {
                pmax = warp_max_u64(pmax);
                if (lane == 0 && pmax) atomicMax(&s_acc[FA_PMAX], pmax);
            }
            if (bad) atomicOr(&s_flags, bad);
        }
        APS_FPROF(0)
        // ---- exchange 1: all-reduce(max)
        __syncthreads();   // (also: every x / logw store of this CTA precedes the arrival)
        if (tid == 0) {
            const unsigned fl = s_flags;
            if (pgas_step) f.ex_pmax[cta] = make_ulonglong2((fl & 2u) ? ~0ull : s_acc[FA_PMAX], 0ull);
            f.ex_max[cta] = make_ulonglong2((fl & 1u) ? ~0ull : s_acc[FA_BMAX], seq);
            ex_arrive(f.ctr);
            int err = 0;
            ex_wait(f.ctr, (u64)G * (u64)t, &err);
            if (err) s_err = err;
        }
        __syncthreads();
        {
            u64 m = 0, pm = 0;
            for (int k = tid; k < G; k += NT) {
                const ulonglong2 v = __ldcg(&f.ex_max[k]);
                m = v.x > m ? v.x : m;
                if (pgas_step) {
                    const ulonglong2 pv = __ldcg(&f.ex_pmax[k]);
                    pm = pv.x > pm ? pv.x : pm;
                }
            }
            if (warp * 32 < G) {
                m = warp_max_u64(m);
                if (lane == 0) atomicMax(&s_acc[FA_MENC], m);
                if (pgas_step) {
                    pm = warp_max_u64(pm);
                    if (lane == 0) atomicMax(&s_acc[FA_PENC], pm);
                }
            }
        }
        __syncthreads();
        const u64 menc = s_acc[FA_MENC], penc = s_acc[FA_PENC];
        const bool bad_w = menc == ~0ull;         // some log-weight was NaN
        const double M = aps_decode_ordered(menc);
        APS_FPROF(1)

        // =============================================================== phase B: integer weights, sub-tile totals
        // one warp per 128-slot sub-tile, no block barrier inside
        {
            const double Mp = aps_decode_ordered(penc);
            u64 a1 = 0, a2 = 0, ap = 0;
            for (int g = warp; g < nsub_mine; g += nw) {
                const int base = i0 + g * APS_FUSED_SUB;
                u64 s0 = 0;
#pragma unroll
                for (int r = 0; r < APS_FUSED_IPT; ++r) {
                    const int i = base + r * 32 + lane;
                    if (i < i1) {
                        const double e = aps_exp(c.logw[i] - M);
                        const u64 qi = (e > 0.0) ? (u64)__double2ull_rz(e * scale) : 0ull;
                        c.q[i] = qi;
                        const u64 qs = qi >> c.Hs;
                        s0 += qi;
                        a1 += qs;
                        a2 += qs * qs;
                        if (pgas_step) {
                            const double ep = aps_exp(reinterpret_cast<const double *>(qp_t)[i] - Mp);
                            const u64 qpi = (ep > 0.0) ? (u64)__double2ull_rz(ep * scale) : 0ull;
                            qp_t[i] = qpi;
                            ap += qpi;
                        }
                    }
                }
                s0 = warp_sum_u64(s0);
                if (lane == 0) s_sub[g] = s0;
            }
            a1 = warp_sum_u64(a1);
            a2 = warp_sum_u64(a2);
            if (lane == 0) {
                atomicAdd(&s_acc[FA_Q1], a1);
                atomicAdd(&s_acc[FA_Q2], a2);
            }
            if (pgas_step) {
                ap = warp_sum_u64(ap);
                if (lane == 0) atomicAdd(&s_acc[FA_QP], ap);
            }
        }
        __syncthreads();   // all q stores and sub-tile totals of this CTA are complete
        APS_FPROF(2)
        // ---- exchange 2: all-gather(totals). Warp 0 turns the sub-tile totals into exclusive prefixes
        //      (what the pulling CTAs read) and publishes the chunk totals.
        if (warp == 0) {
            const int per = (nsubc + 1 + 31) >> 5;       // consecutive prefix entries per lane (nsubc + 1 entries in all)
            const int lo = lane * per;
            u64 sum = 0;
            for (int g = lo; g < lo + per && g < nsub_mine; ++g) sum += s_sub[g];
            const u64 inc = warp_incl_scan_u64(sum, lane);
            const u64 cq = __shfl_sync(0xffffffffu, inc, 31);
            u64 run = inc - sum;
            for (int g = lo; g < lo + per && g <= nsubc; ++g) {
                my_prefix[g] = run;                      // (entries past my last sub-tile repeat the chunk total)
                if (g < nsub_mine) run += s_sub[g];
            }
            __syncwarp();
            if (lane == 0) {
                f.ex_tot[3 * cta + 1] = make_ulonglong2(s_acc[FA_Q1], s_acc[FA_Q2]);
                if (pgas_step) f.ex_tot[3 * cta + 2] = make_ulonglong2(s_acc[FA_QP], 0ull);
                f.ex_tot[3 * cta] = make_ulonglong2(cq, seq);
                ex_arrive(f.ctr + 1);
                int err = 0;
                ex_wait(f.ctr + 1, (u64)G * (u64)t, &err);
                if (err) s_err = err;
            }
        }
        __syncthreads();
        {
            u64 q1 = 0, q2 = 0;
            for (int k = tid; k < G; k += NT) {
                const ulonglong2 v = __ldcg(&f.ex_tot[3 * k]);
                const ulonglong2 w = __ldcg(&f.ex_tot[3 * k + 1]);
                s_T[k] = v.x;
                q1 += w.x;
                q2 += w.y;
                if (pgas_step) s_P[k] = __ldcg(&f.ex_tot[3 * k + 2]).x;
            }
            if (warp * 32 < G) {
                q1 = warp_sum_u64(q1);
                q2 = warp_sum_u64(q2);
                if (lane == 0) {
                    atomicAdd(&s_acc[FA_GQ1], q1);
                    atomicAdd(&s_acc[FA_GQ2], q2);
                }
            }
        }
        __syncthreads();
        // ---- inclusive chunk ends (warp 0) and the plan of decision point t, derived identically by
        //      every CTA: weights summary on warp 0, resampling offsets on warp 1
        if (warp < 2) {
            const int per = (G + 31) >> 5;
            const int lo = lane * per;
            u64 sum = 0;
            for (int k = lo; k < lo + per && k < G; ++k) sum += s_T[k];
            const u64 inc = warp_incl_scan_u64(sum, lane);
            const u64 Q = __shfl_sync(0xffffffffu, inc, 31);
            if (warp == 0) {
                u64 run = inc - sum;
                for (int k = lo; k < lo + per && k < G; ++k) {
                    run += s_T[k];
                    s_E[k] = run;
                }
                if (pgas_step) {   // same for the ancestor-weight totals
                    u64 psum = 0;
                    for (int k = lo; k < lo + per && k < G; ++k) psum += s_P[k];
                    const u64 pinc = warp_incl_scan_u64(psum, lane);
                    u64 prun = pinc - psum;
                    for (int k = lo; k < lo + per && k < G; ++k) {
                        prun += s_P[k];
                        s_P[k] = prun;
                    }
                }
                if (lane == 0) {
                    int err = bad_w ? APS_ERR_WEIGHTS : 0;
                    if (menc == 0) err = APS_ERR_WEIGHTS;
                    if (s_err) err = s_err;
                    make_plan_a<IN_LOGW>(c, t, M, Q, s_acc[FA_GQ1], s_acc[FA_GQ2], err, &s_plan);
                }
            } else if (lane == 0) {
                make_plan_b(c, t, Q, &s_plan);
            }
        }
        __syncthreads();
        const StepPlan pl = s_plan;
        if (cta == 0 && tid == 0) {
            if (pl.err) sweep_err = pl.err;
            else logev += pl.logZ - (resampled_prev ? c.logN : logz_prev);   // src/container.jl:341,359
            c.plan[t] = pl;
        }
        logz_prev = pl.logZ;
        resampled_prev = pl.resampled;
        APS_FPROF(3)

        // =============================================================== phase C: ancestors of my slots (pull)
        const int ni = (int)pl.n;                                            // children drawn: Ng, or Ng - 1 with a reference
        if (!pl.resampled || pl.err) {
            // update_keys! branch (src/container.jl:247): every particle continues, weights kept
            for (int i = i0 + tid; i < i1; i += NT) anc_out[i] = (int32_t)(c.slot0 + i);
        } else {
            const u64 Q = pl.Q, R = pl.R;
            const double ratio = pl.ratio, roff = KIND == APS_RESAMPLE_SYSTEMATIC ? pl.roff : 0.0;
            const int guard = pl.guard;
            const u64 skey = KIND == APS_RESAMPLE_STRATIFIED ? key : 0ull;
            const u64 step = (u64)(t + c.ctr_offset);
            // children below the end of every chunk (exact); the last chunk ends at n by definition
            if (tid < G) s_kend[tid] = tid == G - 1 ? ni : children_below_checked<KIND>(s_E[tid], Q, R, ni, ratio, roff, guard, skey, step);
            const int g0 = (int)c.slot0 + i0;                                    // my children: global slots [g0, g1)
            const int g1 = (int)c.slot0 + i1 < ni ? (int)c.slot0 + i1 : ni;
            for (int cb = g0; cb < g1; cb += CAP) {
                const int ce = cb + CAP < g1 ? cb + CAP : g1;
                {   // clear the marker array
                    int4 *own4 = reinterpret_cast<int4 *>(own);
                    const int4 z4 = make_int4(0, 0, 0, 0);
#pragma unroll
                    for (int m = 0; m < APS_FUSED_CPT / 4; ++m) own4[m * NT + tid] = z4;
                }
                __syncthreads();   // markers cleared, s_kend complete
                // chunks whose children intersect [cb, ce): k0 = first chunk with K(E_k) > cb, k1 = first with
                // K(E_k) >= ce (K is non-decreasing in k); every warp finds them itself
                int k0 = G - 1, k1 = G - 1;
                for (int b = 0; b < G; b += 32) {
                    const int k = b + lane;
                    const unsigned m0 = __ballot_sync(0xffffffffu, k < G && s_kend[k] > cb);
                    if (m0) {
                        k0 = b + __ffs(m0) - 1;
                        break;
                    }
                }
                for (int b = k0 & ~31; b < G; b += 32) {
                    const int k = b + lane;
                    const unsigned m1 = __ballot_sync(0xffffffffu, k < G && s_kend[k] >= ce);
                    if (m1) {
                        k1 = b + __ffs(m1) - 1;
                        break;
                    }
                }
                // chunks are handled in groups of APS_FUSED_GRP (one group unless the weights are very sparse)
                for (int kg = k0; kg <= k1; kg += APS_FUSED_GRP) {
                    const int nch = k1 - kg + 1 < APS_FUSED_GRP ? k1 - kg + 1 : APS_FUSED_GRP;
                    const int nF = nch * nsubc;                                // sub-tiles of the group, flattened
                    // (1) the group's sub-tile prefixes -> shared memory, one coalesced batch
                    for (int idx = tid; idx < nch * (nsubc + 1); idx += NT)
                        s_pf[idx] = __ldcg(f.sub_prefix + (long long)kg * (nsubc + 1) + idx);
                    __syncthreads();
                    // (2) children below the END of every sub-tile, exactly, one evaluation per thread
                    for (int F = tid; F < nF; F += NT) {
                        const int kc = F / nsubc, g = F - kc * nsubc;
                        const u64 Eb = kg + kc == 0 ? 0ull : s_E[kg + kc - 1];
                        s_ks[F] = g == nsubc - 1 ? s_kend[kg + kc]
                                                 : children_below_checked<KIND>(Eb + s_pf[kc * (nsubc + 1) + g + 1], Q, R, ni, ratio, roff, guard, skey, step);
                    }
                    __syncthreads();
                    // (3) sub-tiles whose children intersect [cb, ce): FA = first with K(end) > cb, FB = first with
                    //     K(end) >= ce (K is non-decreasing); every warp finds them itself
                    int FA = nF, FB = nF - 1;
                    for (int b = 0; b < nF; b += 32) {
                        const int F = b + lane;
                        const unsigned m0 = __ballot_sync(0xffffffffu, F < nF && s_ks[F] > cb);
                        if (m0) {
                            FA = b + __ffs(m0) - 1;
                            break;
                        }
                    }
                    for (int b = FA & ~31; b < nF; b += 32) {
                        const int F = b + lane;
                        const unsigned m1 = __ballot_sync(0xffffffffu, F < nF && s_ks[F] >= ce);
                        if (m1) {
                            FB = b + __ffs(m1) - 1;
                            break;
                        }
                    }
                    // (4) one warp per relevant sub-tile: scan its 128 integer weights, drop the markers
                    for (int F = FA + warp; F <= FB; F += nw) {
                        const int kc = F / nsubc, g = F - kc * nsubc, k = kg + kc;
                        const u64 pre0 = s_pf[kc * (nsubc + 1) + g], pre1 = s_pf[kc * (nsubc + 1) + g + 1];
                        if (pre1 == pre0) continue;                            // no weight, no children (warp-uniform)
                        const int pbase = k * Nc;                              // first parent of chunk k (local index)
                        const int pend = pbase + Nc < N ? pbase + Nc : N;
                        const u64 Cs = (k == 0 ? 0ull : s_E[k - 1]) + pre0;
                        // K at the start of the sub-tile; K(C_{-1}) := 0 for the globally first parent
                        const int ka = F > 0 ? s_ks[F - 1] : (k == 0 ? 0 : s_kend[k - 1]);
                        const int j0 = pbase + g * APS_FUSED_SUB + lane * APS_FUSED_IPT;
                        u64 cum[APS_FUSED_IPT];
                        if (j0 + APS_FUSED_IPT <= pend) {
                            const ulonglong2 v0 = __ldcg(reinterpret_cast<const ulonglong2 *>(c.q + j0));
                            const ulonglong2 v1 = __ldcg(reinterpret_cast<const ulonglong2 *>(c.q + j0 + 2));
                            cum[0] = v0.x; cum[1] = v0.y; cum[2] = v1.x; cum[3] = v1.y;
                        } else {
#pragma unroll
                            for (int r = 0; r < APS_FUSED_IPT; ++r) cum[r] = j0 + r < pend ? __ldcg(c.q + j0 + r) : 0ull;
                        }
#pragma unroll
                        for (int r = 1; r < APS_FUSED_IPT; ++r) cum[r] += cum[r - 1];
                        const u64 inc = warp_incl_scan_u64(cum[APS_FUSED_IPT - 1], lane);
                        const u64 excl = Cs + (inc - cum[APS_FUSED_IPT - 1]);
                        bool unsafe = false;
                        int kk[APS_FUSED_IPT];
#pragma unroll
                        for (int r = 0; r < APS_FUSED_IPT; ++r)
                            kk[r] = children_below_fast<KIND>(excl + cum[r], Q, ni, ratio, roff, guard, skey, step, &unsafe);
                        if (__any_sync(0xffffffffu, unsafe)) {   // an estimate fell into the guard band: exact values
#pragma unroll
                            for (int r = 0; r < APS_FUSED_IPT; ++r)
                                kk[r] = children_below_checked<KIND>(excl + cum[r], Q, R, ni, ratio, roff, guard, skey, step);
                        }
                        // parent j owns children [K(C_{j-1}), K(C_j)): marker at its first child inside [cb, ce)
                        int klo = __shfl_up_sync(0xffffffffu, kk[APS_FUSED_IPT - 1], 1);
                        if (lane == 0) klo = ka;
                        const int gj0 = (int)c.slot0 + j0;
#pragma unroll
                        for (int r = 0; r < APS_FUSED_IPT; ++r) {
                            const int khi = kk[r];
                            if (khi > klo) {
                                const int lo = klo > cb ? klo : cb;
                                const int hi = khi < ce ? khi : ce;
                                if (lo < hi) own[lo - cb] = gj0 + r + 1;
                            }
                            klo = khi;
                        }
                    }
                    if (kg + APS_FUSED_GRP <= k1) __syncthreads();   // s_pf / s_ks are reloaded for the next group
                }
                __syncthreads();
                // ---- markers -> ancestor ids: running maximum over the child slots, 16 consecutive per thread
                {
                    const int cnt = ce - cb;
                    const bool active = tid * APS_FUSED_CPT < cnt;
                    int v[APS_FUSED_CPT];
                    int run = 0;
                    if (active) {
                        const int4 *own4 = reinterpret_cast<const int4 *>(own);
#pragma unroll
                        for (int m = 0; m < APS_FUSED_CPT / 4; ++m) {
                            const int4 q4 = own4[tid * (APS_FUSED_CPT / 4) + m];
                            run = max(run, q4.x); v[4 * m] = run;
                            run = max(run, q4.y); v[4 * m + 1] = run;
                            run = max(run, q4.z); v[4 * m + 2] = run;
                            run = max(run, q4.w); v[4 * m + 3] = run;
                        }
                    }
                    int inc = run;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int tt = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc = max(inc, tt);
                    }
                    int excl = __shfl_up_sync(0xffffffffu, inc, 1);
                    if (lane == 0) excl = 0;
                    int *wm = reinterpret_cast<int *>(red);     // 33 u64 = 66 ints
                    if (lane == 31) wm[warp] = inc;
                    __syncthreads();
                    if (warp == 0) {
                        int w = lane < nw ? wm[lane] : 0;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int tt = __shfl_up_sync(0xffffffffu, w, o);
                            if (lane >= o) w = max(w, tt);
                        }
                        wm[32 + lane] = w;                        // inclusive maxima of the warps
                    }
                    __syncthreads();
                    if (warp > 0) excl = max(excl, wm[32 + warp - 1]);
                    if (active) {
                        int32_t *dst = anc_out + (cb - (int)c.slot0) + tid * APS_FUSED_CPT;
#pragma unroll
                        for (int m = 0; m < APS_FUSED_CPT / 4; ++m) {
                            int4 o;
                            o.x = max(v[4 * m], excl) - 1;
                            o.y = max(v[4 * m + 1], excl) - 1;
                            o.z = max(v[4 * m + 2], excl) - 1;
                            o.w = max(v[4 * m + 3], excl) - 1;
                            const int pos = tid * APS_FUSED_CPT + 4 * m;
                            if (pos + 4 <= cnt) {
                                *reinterpret_cast<int4 *>(dst + 4 * m) = o;
                            } else {
                                if (pos < cnt) dst[4 * m] = o.x;
                                if (pos + 1 < cnt) dst[4 * m + 1] = o.y;
                                if (pos + 2 < cnt) dst[4 * m + 2] = o.z;
                            }
                        }
                    }
                    // (own[] and wm[] are reused by the next pass: the barrier after its marker clear... own is
                    //  cleared by every thread for its OWN int4 slots only after all threads read theirs)
                    __syncthreads();
                }
            }
            // ---- the reference particle keeps the globally last slot (src/container.jl:219-224)
            if (has_ref && cta == ref_cta && c.rank == c.world - 1) {
                if (tid == 0) anc_out[N - 1] = (int32_t)(c.Ng - 1);
                // update_ref! (src/pgas.jl:113-128): one categorical draw over the ancestor weights; the
                // owner of the reference slot locates the drawn chunk from the totals and scans it
                if (pgas_step) {
                    const u64 Qp = s_P[G - 1];
                    uint64_t w0, w1;
                    aps_philox2x64(0, aps_ctr1((u64)t, APS_DOM_PGAS, 0), key, &w0, &w1);
                    const double Mp = aps_decode_ordered(penc);
                    const bool okp = Qp != 0 && penc != ~0ull && penc != 0 && Mp == Mp && Mp != aps_bits2d(0xFFF0000000000000ULL);
                    if (!okp) {
                        if (tid == 0) c.st->err = APS_ERR_WEIGHTS;
                    } else {
                        const u64 tau = floor_uq53(aps_u53(w0), Qp);
                        if (tid == 0) s_k0 = -1;
                        __syncthreads();
                        if (tid < G) {
                            const u64 pex = tid == 0 ? 0ull : s_P[tid - 1];
                            if (pex <= tau && tau < s_P[tid]) {
                                s_k0 = tid;
                                s_pre = pex;
                            }
                        }
                        __syncthreads();
                        const int ks = s_k0;
                        if (ks >= 0) {
                            const int pbase = ks * Nc;
                            const int pend = pbase + Nc < N ? pbase + Nc : N;
                            u64 run = s_pre;
                            if (tid == 0) s_found = 0x7fffffff;
                            __syncthreads();
                            for (int b = pbase; b < pend; b += NT * APS_FUSED_IPT) {   // uniform loop: first element with cum > tau
                                const int j0 = b + tid * APS_FUSED_IPT;
                                u64 cum[APS_FUSED_IPT];
#pragma unroll
                                for (int r = 0; r < APS_FUSED_IPT; ++r) cum[r] = j0 + r < pend ? __ldcg(qp_t + j0 + r) : 0ull;
#pragma unroll
                                for (int r = 1; r < APS_FUSED_IPT; ++r) cum[r] += cum[r - 1];
                                u64 tsum;
                                const u64 excl = fblk_excl_scan(cum[APS_FUSED_IPT - 1], red, &tsum) + run;
                                int mine = 0x7fffffff;
#pragma unroll
                                for (int r = APS_FUSED_IPT - 1; r >= 0; --r)
                                    if (j0 + r < pend && excl + cum[r] > tau) mine = j0 + r;
                                if (mine != 0x7fffffff) atomicMin(&s_found, mine);
                                __syncthreads();
                                if (s_found != 0x7fffffff) break;
                                run += tsum;
                            }
                            if (tid == 0 && s_found != 0x7fffffff) anc_out[N - 1] = (int32_t)((int)c.slot0 + s_found);
                        }
                    }
                }
            }
        }
        if (tid < FA_N) s_acc[tid] = 0;
        if (tid == 32) s_flags = 0;
        __syncthreads();   // phase A of the next step reads this CTA's ancestors; accumulators are clear
        APS_FPROF(4)
    }
#undef APS_FPROF
    if (prof && f.dbg) {
        for (int k = 0; k < 5; ++k) f.dbg[cta * 8 + k] = tacc[k];
    }
    if (cta == 0 && tid == 0) {
        c.st->logev = logev;
        if (sweep_err) c.st->err = sweep_err;
    }
}
