// aps_device.cuh -- device-side structures and warp/block primitives shared by the kernels.
//
// Data layout in HBM (one handle, single GPU; N particles, T steps, d state dims):
//   x      [slab t-1][k][i]  f64   SoA per time step, coalesced along the particle index i
//   anc    [slab s][i]       i32   ancestors (0-based, in set s) of slot i of set s+1; s = 0..T
//   logw   [i]               f64   unnormalised log-weights (pc.logWs, src/container.jl:9)
//   q      [i]               u64   canonical integer weights floor(exp(logw - M) 2^S)
//   tile_sum / tile_prefix [tile]  u64   per-2048-particle tile totals and their exclusive scan
//   plan   [s]                     per-decision-point scalars (max, total, logZ, ESS, decision, thresholds)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "aps_b200.h"

typedef unsigned long long u64;

#define APS_THREADS 128         // threads per block of the tile kernels (normalise, resample, select)
#ifndef APS_IPT
#define APS_IPT 16              // items per thread in a tile: 16 u64 = one 128-byte row of the TMA box
#endif
#define APS_TILE (APS_THREADS * APS_IPT)   // 2048 particles per tile
#define APS_CPT 20              // child slots per thread in one expand pass
#define APS_CAP (APS_THREADS * APS_CPT)    // 2560 children staged per pass
#define APS_WARPS (APS_THREADS / 32)
#define APS_ROW 16              // u64 per 128-byte row of the TMA box; a tile is APS_TILE / APS_ROW = 128 rows
// geometry of the systematic / stratified resample kernel. Measured at N = 2^25 (L2 flushed):
// 128 threads x 16 parents, 8 blocks/SM: 95.9 us; 256 x 8 at 4 / 5 / 6 blocks/SM: 105 / 110 / 119 us
// (more warps in flight, but twice the block-scan and expand-scan overhead per parent).
#ifndef APS_K3_THREADS
#define APS_K3_THREADS 128
#endif
#define APS_K3_IPT (APS_TILE / APS_K3_THREADS)      // parents per thread
#ifndef APS_K3_CPT
#define APS_K3_CPT 20                               // child slots per thread in one expand pass
#endif
#define APS_K3_CAP (APS_K3_THREADS * APS_K3_CPT)    // children staged per pass
#define APS_K3_WARPS (APS_K3_THREADS / 32)
#ifndef APS_K3_PF_WAVES
#define APS_K3_PF_WAVES 0       // L2 prefetch distance of the tile loads, in residency waves (measured at N = 2^25: 0: 84.2 us, 1: 86.1, 2: 93.2, 4: 99.9)
#endif
#ifndef APS_K3_MINBLOCKS
#define APS_K3_MINBLOCKS 8
#endif
// 128 threads x 8 blocks per SM (64 registers): measured best at N = 1e6 (2.64 ms per sweep; 256 x 4: 2.74,
// 128 x 10 at 48 registers: 2.72, 256 x 5 at 48 registers: 2.84 -- scripts/time_sweep_kernels.py, round 2)
#ifndef APS_K1_THREADS
#define APS_K1_THREADS 128      // threads per block of the grid-stride kernels (propagate, maxima)
#endif
#ifndef APS_K1_MINBLOCKS
#define APS_K1_MINBLOCKS 8
#endif
// Programmatic dependent launch (single-GPU graph of the systematic / stratified path): a kernel
// launched with the programmatic-serialization attribute may start while its predecessor drains;
// it must not touch anything the predecessor reads or writes before APS_PDL_WAIT() returns (the
// predecessor is then complete and flushed). APS_PDL_TRIGGER() at the top of a kernel lets ITS
// successor launch as soon as every block of this kernel is resident. Both are no-ops in a kernel
// launched without the attribute.
// Measured (round 2, C2): trigger at kernel start 3.08 ms per sweep, trigger after the main loop 2.68 ms,
// no PDL 2.69 ms -- no gain, and the extra state costs the resample kernel its spill-free register
// allocation (isolated N = 2^25: 85.8 -> 89.3 us). Hence a build option (-DAPS_PDL=1 and APS_PDL=1 in
// the environment), off by default: the instructions are not even emitted.
#ifndef APS_PDL
#define APS_PDL 0
#endif
#if APS_PDL
#define APS_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define APS_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#else
#define APS_PDL_WAIT() ((void)0)
#define APS_PDL_TRIGGER() ((void)0)
#endif
// In-graph timeline probes of K2 / K3 (APS_DEBUG_MULTI & 16 prints first-block-start -> last-block-end per
// kernel): -DAPS_TIMELINE=1 only, for the same reason. (K1 always carries its probe.)
#ifndef APS_TIMELINE
#define APS_TIMELINE 0
#endif

// ---------------------------------------------------------------- multi-GPU sharding (one process per GPU)
// Rank r owns the contiguous global slots [r Nl, (r+1) Nl). Peers' state / ancestor stores and a
// small mailbox are mapped through CUDA IPC, so kernels address them with plain loads and stores
// over NVLink. A few tiny exchanges per step run inside the kernels (no NCCL on the data path):
//   kind 0  all-reduce(max) of the log-weight maximum       (posted by k_normalise)
//   kind 1  all-gather of the integer weight totals           (posted by k_resample / k_plan_multi)
//   kind 2  barrier after the ancestor scatter                (posted by k_propagate)
//   kind 3  multinomial / residual: offspring totals per rank (k_scan_tile_counts)
//   kind 4  residual: deterministic copies per rank           (k_residual_exchange<0>)
//   kind 5  residual: residual-weight totals per rank         (k_residual_exchange<1>)
//   kind 6  PGAS: maximum of the ancestor log-weights         (k_pgas_select)
//   kind 7  PGAS: ancestor-weight totals per rank             (last block of k_pgas_select)
//   kind 8  final pick: candidate slot per rank               (k_pick)
//   kind 9  multinomial / residual: barrier after the draws were routed to their owners (k_route_barrier)
// All combined quantities are integers, so every rank derives the identical plan.
#define APS_MAX_RANKS 8
#define APS_MAIL_KINDS 10
struct MailSlot {
    ulonglong2 pair[4];   // .x = value, .y = sequence number (epoch * stride + step + 1)
};
struct PeerTable {
    double *x[APS_MAX_RANKS];
    int32_t *anc[APS_MAX_RANKS];
    MailSlot *mail[APS_MAX_RANKS];  // mail[r][kind * APS_MAX_RANKS + src]
};

// Fat parents. Under weight degeneracy a single parent can own a large share of all children; the
// block that owns its tile would have to write them all (O(children) serial chunks). Instead,
// parents with at least fat_min children are recorded as (child range, parent) entries in a small
// per-decision-point list that lives next to the mailbox (so peers can push into it), their child
// range is skipped by the expand, and the consumer fills it in: the propagate kernel of the next
// step resolves its own slots against the list and patches the ancestor store; k_fill_fat does the
// same after the final decision point and at the operator level. fat_min >= Ng / 128, so a list
// never holds more than APS_FAT_MAX entries.
#define APS_FAT_MAX 128
struct FatEntry {
    int lo, hi;   // global child slots [lo, hi)
    int parent;   // global parent index
    int pad;
};
// byte layout of the mailbox allocation: MailSlot[KINDS * RANKS] | int fat_cnt[steps] | FatEntry[steps][APS_FAT_MAX]
__host__ __device__ __forceinline__ size_t aps_mail_bytes() { return sizeof(MailSlot) * APS_MAIL_KINDS * APS_MAX_RANKS; }
__host__ __device__ __forceinline__ size_t aps_fatcnt_bytes(long long steps) { return ((size_t)steps * sizeof(int) + 15) & ~(size_t)15; }
__host__ __device__ __forceinline__ size_t aps_recvcnt_off(long long steps) {
    return aps_mail_bytes() + aps_fatcnt_bytes(steps) + (size_t)steps * APS_FAT_MAX * sizeof(FatEntry);
}
// Sharded multinomial / residual: the i.i.d. draws are made once (rank r makes draws [r n/G, (r+1) n/G))
// and ROUTED to the rank whose weight range they fall into -- appended to that rank's receive
// buffer with a peer atomic (warp-aggregated) and a peer store -- instead of every rank making all
// n G draws and keeping its own. Layout behind the fat lists: recv_cnt[steps] u64 | recv[recv_cap] u64.
__host__ __device__ __forceinline__ size_t aps_recv_off(long long steps) { return aps_recvcnt_off(steps) + (((size_t)steps * 8 + 15) & ~(size_t)15); }
__host__ __device__ __forceinline__ size_t aps_mailbox_alloc_bytes(long long steps, long long recv_cap) {
    return aps_recv_off(steps) + (size_t)recv_cap * 8;
}

// parent of global child slot g if it lies in a deferred range, else -1. `ent` is the block's
// shared-memory copy of the list (fat_stage): the list is read once per block at L2 -- sharded, peers
// push into it, so neither ld.global.nc nor L1 may serve it -- and looked up from shared memory.
__device__ __forceinline__ int fat_lookup(const int4 *ent, int nfat, int g) {
    int a = -1;
#pragma unroll 1
    for (int e = 0; e < nfat; ++e) {
        const int4 v = ent[e];
        if (g >= v.x && g < v.y) a = v.z;
    }
    return a;
}
// all threads of the block; ends with a block barrier
__device__ __forceinline__ void fat_stage(int4 *s_ent, const FatEntry *ent, int nfat) {
    for (int e = threadIdx.x; e < nfat; e += blockDim.x) s_ent[e] = __ldcg(reinterpret_cast<const int4 *>(ent + e));
    __syncthreads();
}

__device__ __forceinline__ u64 ld_sys_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(u64 *p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exchanges are push + poll. Scalars that a kernel reduces (the log-weight maximum, the integer
// totals) are pushed to every rank's mailbox by the LAST block of the producing kernel
// (device-scope ticket), so they travel over NVLink while the kernel drains and the next one
// launches. The barrier after the ancestor scatter is different: the peer stores of all blocks
// must have landed, which holds when the NEXT kernel of the stream starts, so block 0 of the
// consumer kernel posts it. In both cases all blocks of the consumer kernel, on every rank, spin on
// their LOCAL mailbox until every rank's values for the expected sequence number are there.
//
// Memory model. A value travels with its sequence number in ONE 128-bit access (st / ld .b128:
// single-copy atomic), so an exchange of VALUES (maxima, integer totals, candidates: every kind
// but 2) needs no ordering against other memory at all: the consumer validates each pair by its
// sequence number and uses nothing else the producer wrote. Those posts and polls are relaxed.
//
// Kind 2 is different: it is a barrier that publishes DATA written with plain peer stores -- the
// ancestor scatter and the fat-list pushes of the previous resample kernel. It is posted by block
// 0 of the NEXT kernel of the stream (the previous kernel, including its peer stores, is complete
// by then) with a RELEASE store at system scope, which is cheap there: a block that has just
// started has nothing outstanding to drain. On the consumer side the poll is relaxed and the
// peer-written data (ancestor slab, fat entries and counts) is read with L1-bypassing loads
// (ld.global.cg): L2 is the point of coherence for peer writes, and the writes were performed
// there before the flag was even sent.
//
// The textbook alternative -- an acquire per poll, or one fence.acq_rel.sys after the wait in every
// block -- was built and measured (round 2, 2 x B200, bench workload): acquire loads 4.70 ms per
// sweep, relaxed poll + one fence per block 7.55 ms, this scheme 3.6 ms (profiles/README.md). A
// system-scope fence in each of the ~1200 blocks of a step drains the whole SM every time.
// -DAPS_STRICT_FENCES=1 builds the fenced variant; tests/test_gpu_stress.py passes under both.
#ifndef APS_STRICT_FENCES
#define APS_STRICT_FENCES 0
#endif
__device__ __forceinline__ void st_pair_sys(ulonglong2 *p, u64 v, u64 seq) {
    asm volatile("{ .reg .b128 t; mov.b128 t, {%1, %2}; st.relaxed.sys.global.b128 [%0], t; }" ::"l"(p), "l"(v), "l"(seq)
                 : "memory");
}
__device__ __forceinline__ void st_pair_sys_release(ulonglong2 *p, u64 v, u64 seq) {
    asm volatile("{ .reg .b128 t; mov.b128 t, {%1, %2}; st.release.sys.global.b128 [%0], t; }" ::"l"(p), "l"(v), "l"(seq)
                 : "memory");
}
__device__ __forceinline__ ulonglong2 ld_pair_sys(const ulonglong2 *p) {
    ulonglong2 r;
    asm volatile("{ .reg .b128 t; ld.relaxed.sys.global.b128 t, [%2]; mov.b128 {%0, %1}, t; }"
                 : "=l"(r.x), "=l"(r.y)
                 : "l"(p)
                 : "memory");
    return r;
}
// Last-block detection (device scope). A block publishes its results (plain stores / relaxed atomics), then takes a
// ticket with RELEASE semantics; the block that draws the last ticket fences (acquire) and reads what the others
// published. `__threadfence(); atomicAdd()` does the same but compiles to MEMBAR.SC.GPU + CCTL.IVALL -- a sequentially
// consistent fence plus an invalidate of the SM's whole L1, paid by EVERY block and felt by its neighbours on the SM --
// where the release atomic is MEMBAR.ALL.GPU + ATOMG and only the last block invalidates (cuobjdump, round 2).
__device__ __forceinline__ unsigned ticket_release(unsigned *ctr, unsigned v) {
    unsigned r;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(r) : "l"(ctr), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// publish nv values of this rank (threads 0..world-1 of ONE block; thread r writes to rank r)
__device__ __forceinline__ void mail_post(const PeerTable *pt, int rank, int world, int kind, u64 seq, const u64 *v,
                                          int nv) {
    const int r = threadIdx.x;
    if (r < world) {
        MailSlot *dst = pt->mail[r] + kind * APS_MAX_RANKS + rank;
        if (kind == 2 || kind == 9 || APS_STRICT_FENCES)   // 2 and 9 publish data written with plain peer stores
            for (int k = 0; k < nv; ++k) st_pair_sys_release(&dst->pair[k], v[k], seq);
        else
            for (int k = 0; k < nv; ++k) st_pair_sys(&dst->pair[k], v[k], seq);
    }
}
// Spin budget of one wait, in SM cycles (~2 GHz): APS_COMM_TIMEOUT_MS, default 3000 ms. A dead peer
// costs ONE such timeout per sweep: the first wait that expires raises the sweep's error flag and
// every later wait gives up as soon as it sees the flag.
__device__ long long g_spin_limit = 6000000000LL;
// wait for every rank's nv values of sequence number `seq` (threads 0..world-1 of a block; thread
// r reads slot r of the local mailbox into out[r]). Returns false on timeout (a peer is gone) or
// when *errflag is already set.
__device__ __forceinline__ bool mail_wait(const PeerTable *pt, int rank, int world, int kind, u64 seq, u64 (*out)[4],
                                          int nv, unsigned long long *spin = nullptr, const int *errflag = nullptr,
                                          int limit_mul = 1) {
    const int r = threadIdx.x;
    bool ok = true;
    if (r < world) {
        const MailSlot *src = pt->mail[rank] + kind * APS_MAX_RANKS + r;
        const long long t0 = clock64();
        const long long limit = g_spin_limit * limit_mul;
        for (int k = 0; k < nv; ++k) {
            ulonglong2 pr = ld_pair_sys(&src->pair[k]);
            unsigned it = 0;
            while (pr.y < seq) {
                if ((++it & 63u) == 0) {
                    if (clock64() - t0 > limit || (errflag && *reinterpret_cast<const volatile int *>(errflag) == APS_ERR_COMM)) {
                        ok = false;
                        break;
                    }
                }
                pr = ld_pair_sys(&src->pair[k]);
            }
            out[r][k] = pr.x;
        }
#if APS_STRICT_FENCES
        fence_acq_rel_sys();   // acquire: orders everything after the wait behind the producers' release stores
#endif
        if (spin && kind < 4 && blockIdx.x == 0 && r == (rank + 1) % world)
            atomicAdd(&spin[kind], (unsigned long long)(clock64() - t0));
    }
    return ok;
}

// per-decision-point accumulators, zeroed at sweep start (order-free integer atomics only)
struct StepAcc {
    u64 max_enc;       // atomicMax of aps_encode_ordered(logw)
    u64 sel_max_enc;   // same for the PGAS ancestor log-weights
    unsigned int bad;  // NaN seen
    unsigned int done_ctr;      // last-block detection, normalise kernel
    unsigned int sel_done_ctr;  // last-block detection, categorical kernel
    unsigned int k1_done;       // sharded: last-block detection of the propagate kernel (it posts the maximum)
    unsigned int pad1, pad2;
    u64 tot[4];                 // sharded: this rank's integer totals (Q, Q1, Q2 | nan-flag) and the global max
    u64 rank_off;               // sharded: weight total of the lower ranks at this decision point
    long long child_off;        // sharded multinomial / residual: children owned by parents of lower ranks
    int safe_lo, safe_hi;       // sharded: global child slots [safe_lo, safe_hi) descend from THIS rank's parents (block 0 of
                                // k_resample; empty when unknown) -- the next propagate kernel handles them before the scatter barrier
    u64 t_first_neg[3];         // diagnostics (APS_DEBUG_SPAN): ~globaltimer of the first block start per kernel
    u64 t_last[3];              // globaltimer of the last block end per kernel
};

// per-decision-point plan, written by the last block of the normalise kernel
struct StepPlan {
    double M;        // max log-weight
    double logZ;     // logZ(pc) after reweight (src/container.jl:109)
    double ess;      // effective sample size (src/container.jl:116-119)
    u64 Q;           // integer weight total
    u64 R;           // ceil(U Q / 2^53): systematic offset in integer weight units
    double ratio;    // 2^20 n / Q   (estimate only; results are verified exactly)
    double roff;     // 2^20 R / Q
    long long n;     // children to draw (N, or N-1 with a reference particle)
    int resampled;   // decision of resample_propagate! (src/container.jl:233-251)
    int err;         // aps_status if the weights could not be normalised
    int guard;       // half-width (units of 2^-20) of the band around integers where est is not trusted
    int pad;
};

struct SweepState {
    double logev;    // running log-evidence (src/container.jl:341,359)
    int err;
    int pad;
    long long picked_slot;
    unsigned long long spin[4];  // diagnostics: cycles block 0 spent waiting in each mailbox wait kind
};

// ---------------------------------------------------------------- warp / block primitives
__device__ __forceinline__ u64 warp_incl_scan_u64(u64 v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// block-wide reductions / scan for NW warps (NW a power of two <= 8); results valid in every
// thread. smem: >= NW u64. The NW warp partials are combined with NW-lane shuffle steps.
template <int NW>
__device__ __forceinline__ u64 block_sum_u64(u64 v, u64 *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum_u64(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    u64 t = smem[lane & (NW - 1)];
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

template <int NW>
__device__ __forceinline__ u64 block_max_u64(u64 v, u64 *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max_u64(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    u64 t = smem[lane & (NW - 1)];
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) {
        const u64 x = __shfl_xor_sync(0xffffffffu, t, o);
        t = x > t ? x : t;
    }
    return t;
}

// block-wide exclusive scan of one u64 per thread; *total gets the block sum
template <int NW>
__device__ __forceinline__ u64 block_excl_scan_u64(u64 v, u64 *smem, u64 *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 inc = warp_incl_scan_u64(v, lane);
    __syncthreads();
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    u64 w = smem[lane & (NW - 1)];
#pragma unroll
    for (int o = 1; o < NW; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, w, o, NW);
        if ((lane & (NW - 1)) >= o) w += t;
    }
    *total = __shfl_sync(0xffffffffu, w, NW - 1, NW);
    u64 off = __shfl_sync(0xffffffffu, w, (warp + NW - 1) & (NW - 1), NW);
    if (warp == 0) off = 0;
    return off + inc - v;
}

// ---------------------------------------------------------------- exact 128-bit helpers
struct u128 {
    u64 hi, lo;
};
__device__ __forceinline__ u128 mul_64_64(u64 a, u64 b) {
    u128 r;
    r.hi = __umul64hi(a, b);
    r.lo = a * b;
    return r;
}
__device__ __forceinline__ u128 add_128_64(u128 a, u64 b) {
    u128 r;
    r.lo = a.lo + b;
    r.hi = a.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
__device__ __forceinline__ bool le_128(u128 a, u128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }

// ceil(U Q / 2^53), U < 2^53
__device__ __forceinline__ u64 ceil_uq53(u64 U, u64 Q) {
    u128 p = add_128_64(mul_64_64(U, Q), (1ull << 53) - 1);
    return (p.hi << 11) | (p.lo >> 53);
}
// floor(U Q / 2^53)
__device__ __forceinline__ u64 floor_uq53(u64 U, u64 Q) {
    u128 p = mul_64_64(U, Q);
    return (p.hi << 11) | (p.lo >> 53);
}

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "APS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra APS_DONE;\n"
        "bra APS_WAIT;\n"
        "APS_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 2-D tiled bulk tensor load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// L2 prefetch of a tile that a later block will load (keeps HBM busy across block boundaries)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ u64 global_timer_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct SpanProbe {  // first-block-start / last-block-end of a kernel, for diagnostics only
    StepAcc *a;
    int k;
    __device__ __forceinline__ SpanProbe(StepAcc *acc, int kind, int enabled) : a(enabled ? acc : nullptr), k(kind) {
        if (a && threadIdx.x == 0) atomicMax(&a->t_first_neg[k], ~global_timer_ns());
    }
    __device__ __forceinline__ void end() {
        if (a && threadIdx.x == 0) atomicMax(&a->t_last[k], global_timer_ns());
    }
};
