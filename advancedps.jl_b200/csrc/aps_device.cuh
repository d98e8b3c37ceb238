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
#include <cuda_runtime.h>
#include <stdint.h>
#include "aps_b200.h"

typedef unsigned long long u64;

#define APS_TILE 2048          // particles per tile (normalise + resample kernels)
#define APS_THREADS 256
#define APS_IPT 8              // items per thread in a tile
#define APS_CAP 3072           // children staged per expand pass (12 per thread)
#define APS_CPT 12

// per-decision-point accumulators, zeroed at sweep start (order-free integer atomics only)
struct StepAcc {
    u64 max_enc;       // atomicMax of aps_encode_ordered(logw)
    u64 sel_max_enc;   // same for the PGAS ancestor log-weights
    unsigned int bad;  // NaN seen
    unsigned int done_ctr;      // last-block detection, normalise kernel
    unsigned int sel_done_ctr;  // last-block detection, categorical kernel
    unsigned int pad;
};

// per-decision-point plan, written by the last block of the normalise kernel
struct StepPlan {
    double M;        // max log-weight
    double logZ;     // logZ(pc) after reweight (src/container.jl:109)
    double ess;      // effective sample size (src/container.jl:116-119)
    u64 Q;           // integer weight total
    u64 R;           // ceil(U Q / 2^53): systematic offset in integer weight units
    double ratio;    // n / Q   (estimate only; results are verified exactly)
    double roff;     // R / Q
    long long n;     // children to draw (N, or N-1 with a reference particle)
    int resampled;   // decision of resample_propagate! (src/container.jl:233-251)
    int err;         // aps_status if the weights could not be normalised
};

struct SweepState {
    double logev;    // running log-evidence (src/container.jl:341,359)
    int err;
    int pad;
    long long picked_slot;
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

// block-wide sum for APS_THREADS threads; result valid in every thread. smem: >= 8 u64
__device__ __forceinline__ u64 block_sum_u64(u64 v, u64 *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum_u64(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    u64 t = 0;
#pragma unroll
    for (int w = 0; w < APS_THREADS / 32; ++w) t += smem[w];
    return t;
}

__device__ __forceinline__ u64 block_max_u64(u64 v, u64 *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max_u64(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    u64 t = 0;
#pragma unroll
    for (int w = 0; w < APS_THREADS / 32; ++w) t = smem[w] > t ? smem[w] : t;
    return t;
}

// block-wide exclusive scan of one u64 per thread; *total gets the block sum. smem: >= 8 u64
__device__ __forceinline__ u64 block_excl_scan_u64(u64 v, u64 *smem, u64 *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = warp_incl_scan_u64(v, lane);
    __syncthreads();
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    u64 off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < APS_THREADS / 32; ++w) {
        u64 s = smem[w];
        if (w < warp) off += s;
        tot += s;
    }
    *total = tot;
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
