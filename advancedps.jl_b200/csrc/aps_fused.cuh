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
// An exchange is an array with one 16-byte (value, sequence number) entry per CTA: a CTA publishes
// its entry with a release store and then reads everybody's (acquire) until all carry the current
// sequence number -- the data IS the barrier, there is no separate counter or second round trip.
// Sequence numbers grow across steps and sweeps, so the arrays are never reset.
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
#define APS_FUSED_IPT 4             // parents per thread in one scan tile

struct FusedArgs {
    ulonglong2 *ex_max;     // [G]     exchange 1: (encoded max log-weight or ~0 for NaN, seq)
    ulonglong2 *ex_pmax;    // [G]     PGAS: (encoded max ancestor log-weight, unused); published before ex_max
    ulonglong2 *ex_tot;     // [G][3]  exchange 2: (Q, seq), (Q1, Q2), (PGAS ancestor-weight total, unused)
    u64 *tile_tot;          // [G][tpc] integer weight totals of the scan tiles of every chunk
    u64 *qp;                // [NS]    PGAS: integer ancestor weights
    int chunk;              // slots per CTA (multiple of 64)
    int tpc;                // scan tiles per chunk
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
// spin until the entry carries sequence number >= seq; ~3 s budget (a CTA that never arrives means
// the launch was not co-resident -- a bug, not a runtime condition; fail loudly instead of hanging)
__device__ __forceinline__ ulonglong2 wait_pair(const ulonglong2 *p, u64 seq, int *err) {
    ulonglong2 r = ld_pair_acq_gpu(p);
    if (r.y >= seq) return r;
    const long long t0 = clock64();
    unsigned it = 0;
    while (r.y < seq) {
        if ((++it & 255u) == 0 && clock64() - t0 > 6000000000LL) {
            *err = APS_ERR_COMM;
            break;
        }
        r = ld_pair_acq_gpu(p);
    }
    return r;
}

// ---------------------------------------------------------------- the sweep
template <int D, int DY, int OBS, int KIND>
__global__ void __launch_bounds__(APS_FUSED_MAX_THREADS, 1) k_sweep_fused(const __grid_constant__ DevCtx c,
                                                                         const __grid_constant__ FusedArgs f) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ u64 red[33];
    __shared__ u64 s_tot[APS_FUSED_MAX_CTAS];   // chunk totals, then inclusive chunk ends E_k
    __shared__ int s_kend[APS_FUSED_MAX_CTAS];  // K(E_k): children below the end of chunk k
    __shared__ StepPlan s_plan;
    __shared__ u64 s_q12[2];
    __shared__ int s_err, s_k0, s_k1, s_found;
    int *own = reinterpret_cast<int *>(fsm);    // [blockDim.x * APS_FUSED_CPT]

    const int tid = threadIdx.x, NT = blockDim.x, cta = blockIdx.x, G = gridDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = (NT + 31) >> 5;
    const long long N = c.N, NS = c.NS, T = c.T;
    const int Nc = f.chunk;
    const long long i0 = (long long)cta * Nc;                         // my slots [i0, i1)
    const long long i1 = i0 + Nc < N ? i0 + Nc : (i0 < N ? N : i0);
    const int has_ref = c.sp->has_ref;
    const u64 key = c.sp->key;
    const u64 seq0 = c.sp->epoch * (u64)(T + 2);
    const bool pgas = c.sampler == APS_PGAS && has_ref;
    const int CAP = NT * APS_FUSED_CPT;
    const int TP = NT * APS_FUSED_IPT;                                // parents per scan tile
    const int ref_cta = (int)((N - 1) / Nc);                          // owner of the reference slot
    const double scale = aps_pow2i(c.S);

    if (tid == 0) s_err = 0;

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

    for (long long t = 1; t <= T; ++t) {
        double *__restrict__ xt = c.x + ((t - 1) % c.x_slabs) * (long long)D * NS;
        const double *xp = c.x + ((t - 2 + c.x_slabs) % c.x_slabs) * (long long)D * NS;
        const int32_t *anc_prev = c.anc + ((t - 1) % c.anc_slabs) * NS;   // ancestors of set t (written in phase C of t-1)
        int32_t *anc_out = c.anc + (t % c.anc_slabs) * NS;                // ancestors of set t+1
        const double *__restrict__ y = c.Y + (t - 1) * c.dy;
        const u64 seq = seq0 + (u64)t + 1;
        const bool reset = t == 1 || resampled_prev != 0;
        const bool pgas_step = pgas && t >= 2 && t <= T - 1;              // update_ref! can run at this decision point

        // =============================================================== phase A: propagate + reweight
        u64 bmax = 0, pmax = 0;
        unsigned bad = 0;
        {
            double xref[D];
            if (pgas_step) {
#pragma unroll
                for (int k = 0; k < D; ++k) xref[k] = c.ref[(t - 1) * D + k];   // X_ref[c-1], c = t+1
            }
            const long long p_end = (i1 + 1) >> 1;
            for (long long p = (i0 >> 1) + tid; p < p_end; p += NT) {
                double z[2 * D];
                aps_pair_normals<D>(key, (u64)((c.slot0 >> 1) + p), (u64)t, z);
                const long long j0 = 2 * p;
                int2 a2 = make_int2(0, 0);
                if (t > 1) a2 = *reinterpret_cast<const int2 *>(anc_prev + j0);
                double2 lw2 = make_double2(0.0, 0.0);
                if (!reset) lw2 = *reinterpret_cast<const double2 *>(c.logw + j0);
                double xo[2][D];
                double lwo[2], lpo[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long i = j0 + h;
                    double x[D];
                    lwo[h] = 0.0;
                    lpo[h] = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = 0.0;
                    if (i < N) {
                        const bool is_ref = has_ref && c.slot0 + i == c.Ng - 1;   // the reference keeps the globally last slot
                        double xpv[D];
                        if (t > 1 && (!is_ref || pgas_step)) {
                            const long long a = h ? a2.y : a2.x;
#pragma unroll
                            for (int k = 0; k < D; ++k) xpv[k] = __ldcg(xp + (long long)k * NS + a);
                        }
                        if (is_ref) {
#pragma unroll
                            for (int k = 0; k < D; ++k) x[k] = c.ref[(t - 1) * D + k];
                        } else if (t == 1) {
                            aps_prior_draw<D>(&c.md, z + h * D, x);
                        } else {
                            aps_trans_draw<D>(&c.md, xpv, z + h * D, x);
                        }
                        const double ll = aps_obs_logpdf<D, DY, OBS>(&c.md, x, y);
                        const double lw = (reset ? 0.0 : (h ? lw2.y : lw2.x)) + ll;
                        lwo[h] = lw;
                        if (lw != lw) bad = 1;
                        else {
                            const u64 e = aps_encode_ordered(lw);
                            bmax = e > bmax ? e : bmax;
                        }
                        if (pgas_step) {   // log f(X_ref[c-1] | X_i[c-2]) + logW_i   (src/pgas.jl:26-46)
                            const double lp = aps_trans_logpdf<D>(&c.md, xpv, xref) + lw;
                            lpo[h] = lp;
                            if (lp != lp) bad |= 2u;
                            else {
                                const u64 e = aps_encode_ordered(lp);
                                pmax = e > pmax ? e : pmax;
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
                if (pgas_step) *reinterpret_cast<double2 *>(reinterpret_cast<double *>(f.qp) + j0) = make_double2(lpo[0], lpo[1]);
            }
        }
        // ---- exchange 1: all-reduce(max)
        bmax = fblk_max(bmax, red);
        if (pgas_step) pmax = fblk_max(pmax, red);
        const int bad1 = __syncthreads_or((int)(bad & 1u));   // (also: every x / logw store of this CTA precedes the post)
        const int bad2 = pgas_step ? __syncthreads_or((int)(bad & 2u)) : 0;
        if (tid == 0) {
            if (pgas_step) f.ex_pmax[cta] = make_ulonglong2(bad2 ? ~0ull : pmax, 0ull);
            st_pair_rel_gpu(&f.ex_max[cta], bad1 ? ~0ull : bmax, seq);
        }
        u64 menc = 0, penc = 0;
        {
            int err = 0;
            for (int k = tid; k < G; k += NT) {
                const ulonglong2 v = wait_pair(&f.ex_max[k], seq, &err);
                menc = v.x > menc ? v.x : menc;
                if (pgas_step) {
                    const ulonglong2 pv = __ldcg(&f.ex_pmax[k]);
                    penc = pv.x > penc ? pv.x : penc;
                }
            }
            if (err) s_err = err;
            menc = fblk_max(menc, red);
            if (pgas_step) penc = fblk_max(penc, red);
        }
        const bool bad_w = menc == ~0ull;         // some log-weight was NaN
        const double M = aps_decode_ordered(menc);

        // =============================================================== phase B: integer weights, totals
        u64 cq = 0, cq1 = 0, cq2 = 0, cqp = 0;
        {
            const double Mp = aps_decode_ordered(penc);
            for (int tl = 0; tl < f.tpc; ++tl) {
                u64 s0 = 0;
#pragma unroll
                for (int r = 0; r < APS_FUSED_IPT; ++r) {
                    const long long i = i0 + (long long)tl * TP + r * NT + tid;
                    if (i < i1) {
                        const double e = aps_exp(c.logw[i] - M);
                        const u64 qi = (e > 0.0) ? (u64)__double2ull_rz(e * scale) : 0ull;
                        c.q[i] = qi;
                        const u64 qs = qi >> c.Hs;
                        s0 += qi;
                        cq1 += qs;
                        cq2 += qs * qs;
                        if (pgas_step) {
                            const double ep = aps_exp(reinterpret_cast<const double *>(f.qp)[i] - Mp);
                            const u64 qpi = (ep > 0.0) ? (u64)__double2ull_rz(ep * scale) : 0ull;
                            f.qp[i] = qpi;
                            cqp += qpi;
                        }
                    }
                }
                s0 = fblk_sum(s0, red);
                if (tid == 0) f.tile_tot[(long long)cta * f.tpc + tl] = s0;
                cq += s0;
            }
            cq1 = fblk_sum(cq1, red);
            cq2 = fblk_sum(cq2, red);
            if (pgas_step) cqp = fblk_sum(cqp, red);
        }
        // ---- exchange 2: all-gather(totals). (fblk_sum ends with a barrier: all q / tile_tot stores precede the post)
        if (tid == 0) {
            f.ex_tot[3 * cta + 1] = make_ulonglong2(cq1, cq2);
            if (pgas_step) f.ex_tot[3 * cta + 2] = make_ulonglong2(cqp, 0ull);
            st_pair_rel_gpu(&f.ex_tot[3 * cta], cq, seq);
        }
        u64 Q1 = 0, Q2 = 0, my_tot = 0, my_ptot = 0;
        {
            int err = 0;
            for (int k = tid; k < G; k += NT) {
                const ulonglong2 v = wait_pair(&f.ex_tot[3 * k], seq, &err);
                const ulonglong2 w = __ldcg(&f.ex_tot[3 * k + 1]);
                s_tot[k] = v.x;
                Q1 += w.x;
                Q2 += w.y;
                if (k == tid) my_tot = v.x;
                if (pgas_step && k == tid) my_ptot = __ldcg(&f.ex_tot[3 * k + 2]).x;
            }
            if (err) s_err = err;
        }
        Q1 = fblk_sum(Q1, red);
        Q2 = fblk_sum(Q2, red);
        // inclusive chunk ends E_k (G <= blockDim.x is guaranteed by the launcher)
        u64 Q;
        {
            const u64 ex = fblk_excl_scan(tid < G ? my_tot : 0ull, red, &Q);
            if (tid < G) s_tot[tid] = ex + my_tot;
        }
        // ---- the plan of decision point t, derived identically by every CTA (two warps in parallel)
        if (tid == 0) {
            int err = bad_w ? APS_ERR_WEIGHTS : 0;
            if (menc == 0) err = APS_ERR_WEIGHTS;
            if (s_err) err = s_err;
            make_plan_a<IN_LOGW>(c, t, M, Q, Q1, Q2, err, &s_plan);
        } else if (tid == 32) {   // (the launcher guarantees at least 64 threads)
            make_plan_b(c, t, Q, &s_plan);
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

        // =============================================================== phase C: ancestors of my slots (pull)
        const long long n = pl.n;                                            // children drawn: Ng, or Ng - 1 with a reference
        if (!pl.resampled || pl.err) {
            // update_keys! branch (src/container.jl:247): every particle continues, weights kept
            for (long long i = i0 + tid; i < i1; i += NT) anc_out[i] = (int32_t)(c.slot0 + i);
        } else {
            const u64 R = pl.R;
            const double ratio = pl.ratio, roff = KIND == APS_RESAMPLE_SYSTEMATIC ? pl.roff : 0.0;
            const int guard = pl.guard, ni = (int)n;
            const u64 skey = KIND == APS_RESAMPLE_STRATIFIED ? key : 0ull;
            const u64 step = (u64)(t + c.ctr_offset);
            // children below the end of every chunk (exact); the last chunk ends at n by definition
            if (tid < G) s_kend[tid] = tid == G - 1 ? ni : children_below_checked<KIND>(s_tot[tid], Q, R, ni, ratio, roff, guard, skey, step);
            __syncthreads();
            const long long g1 = c.slot0 + i1 < n ? c.slot0 + i1 : n;        // my children: global slots [slot0 + i0, g1)
            for (long long cb = c.slot0 + i0; cb < g1; cb += CAP) {
                const int cbi = (int)cb;
                const int cei = (int)(cb + CAP < g1 ? cb + CAP : g1);
                {   // clear the marker array
                    int4 *own4 = reinterpret_cast<int4 *>(own);
                    const int4 z4 = make_int4(0, 0, 0, 0);
#pragma unroll
                    for (int m = 0; m < APS_FUSED_CPT / 4; ++m) own4[m * NT + tid] = z4;
                }
                if (tid == 0) {
                    s_k0 = G;
                    s_k1 = -1;
                }
                __syncthreads();
                // chunks whose children intersect [cb, ce): K(E_{k-1}) < ce and K(E_k) > cb
                if (tid < G) {
                    const int ka = tid == 0 ? 0 : s_kend[tid - 1], kb = s_kend[tid];
                    if (ka < cei && kb > cbi) {
                        atomicMin(&s_k0, tid);
                        atomicMax(&s_k1, tid);
                    }
                }
                __syncthreads();
                const int k0 = s_k0, k1 = s_k1;
                for (int k = k0; k <= k1; ++k) {
                    const long long pbase = (long long)k * Nc;                 // first parent of chunk k (local index)
                    const long long pend = pbase + Nc < N ? pbase + Nc : N;
                    u64 tprefix = k == 0 ? 0ull : s_tot[k - 1];
                    int ka = k == 0 ? 0 : s_kend[k - 1];                       // K at the start of the tile
                    for (int tl = 0; tl < f.tpc && pbase + (long long)tl * TP < pend; ++tl) {
                        const u64 ttot = __ldcg(&f.tile_tot[(long long)k * f.tpc + tl]);
                        // K at the end of the tile: every thread evaluates it (uniform), exactly
                        const bool last_tile = pbase + (long long)(tl + 1) * TP >= pend;
                        const int kb = last_tile ? s_kend[k]
                                                 : children_below_checked<KIND>(tprefix + ttot, Q, R, ni, ratio, roff, guard, skey, step);
                        if (ka < cei && kb > cbi) {
                            // ---- scan the tile: APS_FUSED_IPT consecutive parents per thread
                            const long long j0 = pbase + (long long)tl * TP + (long long)tid * APS_FUSED_IPT;
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
                            u64 tsum;
                            const u64 excl = fblk_excl_scan(cum[APS_FUSED_IPT - 1], red, &tsum) + tprefix;
                            bool unsafe = false;
                            int kk[APS_FUSED_IPT];
                            int klo = tid == 0 ? ka : children_below_fast<KIND>(excl, Q, ni, ratio, roff, guard, skey, step, &unsafe);
#pragma unroll
                            for (int r = 0; r < APS_FUSED_IPT; ++r)
                                kk[r] = children_below_fast<KIND>(excl + cum[r], Q, ni, ratio, roff, guard, skey, step, &unsafe);
                            if (__syncthreads_or(unsafe ? 1 : 0)) {   // an estimate fell into the guard band: exact values
                                if (tid != 0) klo = children_below_checked<KIND>(excl, Q, R, ni, ratio, roff, guard, skey, step);
#pragma unroll
                                for (int r = 0; r < APS_FUSED_IPT; ++r)
                                    kk[r] = children_below_checked<KIND>(excl + cum[r], Q, R, ni, ratio, roff, guard, skey, step);
                            }
                            // parent j owns children [K(C_{j-1}), K(C_j)): marker at its first child inside [cb, ce)
                            const int gj0 = (int)(c.slot0 + j0);
#pragma unroll
                            for (int r = 0; r < APS_FUSED_IPT; ++r) {
                                const int khi = kk[r];
                                if (khi > klo) {
                                    const int lo = klo > cbi ? klo : cbi;
                                    const int hi = khi < cei ? khi : cei;
                                    if (lo < hi) own[lo - cbi] = gj0 + r + 1;
                                }
                                klo = khi;
                            }
                        }
                        tprefix += ttot;
                        ka = kb;
                    }
                }
                __syncthreads();
                // ---- markers -> ancestor ids: running maximum over the child slots, 16 consecutive per thread
                {
                    const int cnt = cei - cbi;
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
                    __syncthreads();
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
                        int32_t *dst = anc_out + (cb - c.slot0) + tid * APS_FUSED_CPT;
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
                    __syncthreads();   // own[] and red[] are reused by the next pass
                }
            }
            // ---- the reference particle keeps the globally last slot (src/container.jl:219-224)
            if (has_ref && cta == ref_cta && c.rank == c.world - 1) {
                if (tid == 0) anc_out[N - 1] = (int32_t)(c.Ng - 1);
                // update_ref! (src/pgas.jl:113-128): one categorical draw over the ancestor weights; the
                // owner of the reference slot locates the drawn chunk from the totals and scans it
                if (pgas_step) {
                    u64 Qp;
                    const u64 pex = fblk_excl_scan(tid < G ? my_ptot : 0ull, red, &Qp);
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
                        if (tid < G && pex <= tau && tau < pex + my_ptot) {
                            s_k0 = tid;
                            s_q12[0] = pex;
                        }
                        __syncthreads();
                        const int ks = s_k0;
                        if (ks >= 0) {
                            const long long pbase = (long long)ks * Nc;
                            const long long pend = pbase + Nc < N ? pbase + Nc : N;
                            u64 run = s_q12[0];
                            if (tid == 0) s_found = 0x7fffffff;
                            __syncthreads();
                            for (long long b = pbase; b < pend; b += TP) {   // uniform loop: first element with cum > tau
                                const long long j0 = b + (long long)tid * APS_FUSED_IPT;
                                u64 cum[APS_FUSED_IPT];
#pragma unroll
                                for (int r = 0; r < APS_FUSED_IPT; ++r) cum[r] = j0 + r < pend ? __ldcg(f.qp + j0 + r) : 0ull;
#pragma unroll
                                for (int r = 1; r < APS_FUSED_IPT; ++r) cum[r] += cum[r - 1];
                                u64 tsum;
                                const u64 excl = fblk_excl_scan(cum[APS_FUSED_IPT - 1], red, &tsum) + run;
                                int mine = 0x7fffffff;
#pragma unroll
                                for (int r = APS_FUSED_IPT - 1; r >= 0; --r)
                                    if (j0 + r < pend && excl + cum[r] > tau) mine = (int)(j0 + r);
                                if (mine != 0x7fffffff) atomicMin(&s_found, mine);
                                __syncthreads();
                                if (s_found != 0x7fffffff) break;
                                run += tsum;
                            }
                            if (tid == 0 && s_found != 0x7fffffff) anc_out[N - 1] = (int32_t)(c.slot0 + s_found);
                        }
                    }
                }
            }
        }
        __syncthreads();   // phase A of the next step reads this CTA's ancestors
    }
    if (cta == 0 && tid == 0) {
        c.st->logev = logev;
        if (sweep_err) c.st->err = sweep_err;
    }
}
