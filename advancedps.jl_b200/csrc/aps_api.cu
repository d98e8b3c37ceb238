// aps_api.cu -- C ABI of libaps_b200.so (include/aps_b200.h): handle lifetime, sweep
// orchestration (CUDA graph of the per-step kernels), result accessors, operator-level entry
// points and the resample-kernel micro-benchmark. No torch types; plain pointers and sizes.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <math.h>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

#include "aps_kernels.cuh"
#include "aps_fused.cuh"

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(e_ == cudaErrorMemoryAllocation ? APS_ERR_NOMEM : APS_ERR_CUDA,                \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                           \
    } while (0)

extern "C" const char *aps_last_error(void) { return g_err.c_str(); }
extern "C" const char *aps_version(void) { return "aps_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------ kernel dispatch tables
typedef void (*prop_fn)(const DevCtx, const long long, double *, const double *, const int32_t *);
typedef void (*res_fn)(const DevCtx, const long long, int32_t *, const CUtensorMap);
typedef void (*pgas_fn)(const DevCtx, const long long, const double *, const int32_t *, int32_t *);

// PRE: the variant that loads pre-drawn normals (pair kernel, d <= 3; see k_draw_normals)
template <int D, int OBS, bool MULTI, bool PRE>
static prop_fn prop_for_dy(int dy) {
    switch (dy) {
        case 1: return k_propagate<D, 1, OBS, MULTI, PRE>;
        case 2: return k_propagate<D, 2, OBS, MULTI, PRE>;
        case 3: return k_propagate<D, 3, OBS, MULTI, PRE>;
        default: return k_propagate<D, 4, OBS, MULTI, PRE>;
    }
}
// d = 4: one thread per slot (k_propagate1); APS_K1_PAIRS=1 selects the pair kernel for comparison
template <int OBS, bool MULTI, bool PRE>
static prop_fn prop1_for_dy4(int dy) {
    switch (dy) {
        case 1: return k_propagate1<4, 1, OBS, MULTI, PRE>;
        case 2: return k_propagate1<4, 2, OBS, MULTI, PRE>;
        case 3: return k_propagate1<4, 3, OBS, MULTI, PRE>;
        default: return k_propagate1<4, 4, OBS, MULTI, PRE>;
    }
}
template <int OBS, bool MULTI, bool PRE>
static prop_fn prop_for_dim(int d, int dy) {
    switch (d) {
        case 1: return prop_for_dy<1, OBS, MULTI, PRE>(dy);
        case 2: return prop_for_dy<2, OBS, MULTI, PRE>(dy);
        case 3: return prop_for_dy<3, OBS, MULTI, PRE>(dy);
        default: return getenv("APS_K1_PAIRS") ? prop_for_dy<4, OBS, MULTI, false>(dy) : prop1_for_dy4<OBS, MULTI, PRE>(dy);
    }
}
template <bool MULTI, bool PRE>
static prop_fn pick_propagate_m(int obs, int d, int dy) {
    switch (obs) {
        case APS_OBS_LINEAR_GAUSS: return prop_for_dim<APS_OBS_LINEAR_GAUSS, MULTI, PRE>(d, dy);
        case APS_OBS_STOCH_VOL: return k_propagate<1, 1, APS_OBS_STOCH_VOL, MULTI, PRE>;  // d = dy = 1 (aps_model_prepare)
        default:  // constant log-likelihood: dy is not used
            switch (d) {
                case 1: return k_propagate<1, 1, APS_OBS_CONST, MULTI, PRE>;
                case 2: return k_propagate<2, 1, APS_OBS_CONST, MULTI, PRE>;
                case 3: return k_propagate<3, 1, APS_OBS_CONST, MULTI, PRE>;
                default: return k_propagate1<4, 1, APS_OBS_CONST, MULTI, PRE>;
            }
    }
}
static prop_fn pick_propagate(int obs, int d, int dy, bool multi, bool pre = false) {
    if (pre) return multi ? pick_propagate_m<true, true>(obs, d, dy) : pick_propagate_m<false, true>(obs, d, dy);
    return multi ? pick_propagate_m<true, false>(obs, d, dy) : pick_propagate_m<false, false>(obs, d, dy);
}
static pgas_fn pick_pgas_max(int d) {
    switch (d) {
        case 1: return k_pgas_max<1>;
        case 2: return k_pgas_max<2>;
        case 3: return k_pgas_max<3>;
        default: return k_pgas_max<4>;
    }
}
static pgas_fn pick_pgas_select(int d) {
    switch (d) {
        case 1: return k_pgas_select<1>;
        case 2: return k_pgas_select<2>;
        case 3: return k_pgas_select<3>;
        default: return k_pgas_select<4>;
    }
}
static res_fn pick_resample(int kind, bool multi = false, bool defer = false) {
    const bool strat = kind == APS_RESAMPLE_STRATIFIED;
    if (multi && defer) return strat ? k_resample<APS_RESAMPLE_STRATIFIED, true, true> : k_resample<APS_RESAMPLE_SYSTEMATIC, true, true>;
    if (multi) return strat ? k_resample<APS_RESAMPLE_STRATIFIED, true, false> : k_resample<APS_RESAMPLE_SYSTEMATIC, true, false>;
    if (defer) return strat ? k_resample<APS_RESAMPLE_STRATIFIED, false, true> : k_resample<APS_RESAMPLE_SYSTEMATIC, false, true>;
    return strat ? k_resample<APS_RESAMPLE_STRATIFIED, false, false> : k_resample<APS_RESAMPLE_SYSTEMATIC, false, false>;
}

// the fused persistent sweep (aps_fused.cuh): one instantiation per (model family, resampler)
typedef void (*fused_fn)(const DevCtx, const FusedArgs);
template <int KIND>
static fused_fn pick_fused_kind(int obs, int d, int dy) {
    if (d != 1) return nullptr;   // state dimension 1 (LG1 / SV / constant likelihood); d > 1 runs the three-kernel path
    switch (obs) {
        case APS_OBS_LINEAR_GAUSS: return dy == 1 ? k_sweep_fused<1, 1, APS_OBS_LINEAR_GAUSS, KIND> : nullptr;
        case APS_OBS_STOCH_VOL: return k_sweep_fused<1, 1, APS_OBS_STOCH_VOL, KIND>;
        default: return k_sweep_fused<1, 1, APS_OBS_CONST, KIND>;
    }
}
static fused_fn pick_fused(int kind, int obs, int d, int dy) {
    if (kind == APS_RESAMPLE_SYSTEMATIC) return pick_fused_kind<APS_RESAMPLE_SYSTEMATIC>(obs, d, dy);
    if (kind == APS_RESAMPLE_STRATIFIED) return pick_fused_kind<APS_RESAMPLE_STRATIFIED>(obs, d, dy);
    return nullptr;
}

// ------------------------------------------------------------------ TMA descriptor of the integer-weight array
// q viewed as [rows][APS_IPT] u64 (128-byte rows); box = APS_THREADS rows; 128-byte swizzle.
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_q_tensormap(CUtensorMap *out, u64 *q, long long n_padded) {
    static encode_tiled_fn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(APS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        enc = (encode_tiled_fn)fn;
    }
    const cuuint64_t rows = (cuuint64_t)((n_padded + APS_ROW - 1) / APS_ROW);
    const cuuint64_t gdim[2] = {APS_ROW, rows};
    const cuuint64_t gstride[1] = {APS_ROW * 8};
    const cuuint32_t box[2] = {APS_ROW, APS_TILE / APS_ROW};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, q, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(APS_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return APS_OK;
}
template <typename F>
static void prefer_max_smem(F f) {
    // every kernel of a step asks for the same shared-memory carve-out, so the SMs never have to
    // drain to re-partition L1 / shared memory between consecutive launches
    cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
static int enable_k3_smem() {
    static bool done = false;
    if (done) return APS_OK;
    for (int kind : {APS_RESAMPLE_SYSTEMATIC, APS_RESAMPLE_STRATIFIED})
        for (int variant = 0; variant < 4; ++variant) {
            res_fn f = pick_resample(kind, (variant & 1) != 0, (variant & 2) != 0);
            prefer_max_smem(f);
            CU(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, APS_K3_DYN_SMEM));
        }
    prefer_max_smem(k_normalise<IN_LOGW>);
    prefer_max_smem(k_normalise<IN_W>);
    prefer_max_smem(k_normalise<IN_Q>);
    prefer_max_smem(k_bench_weights);
    prefer_max_smem(k_to_one_based);
    prefer_max_smem(k_vector_max<IN_W>);
    prefer_max_smem(k_vector_max<IN_LOGW>);
    done = true;
    return APS_OK;
}

// CUDA loads kernels lazily, and a first-time load may synchronise the context. Kernels that spin
// on a peer (the sharded exchanges) must therefore never be launched for the first time while the
// peer's matching kernel is still to be loaded by another thread of the same process: load the
// kernels that run outside the captured sweep graph up front.
static void preload_kernels() {
    static bool done = false;
    if (done) return;
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_pick);
    cudaFuncGetAttributes(&a, k_backtrace);
    cudaFuncGetAttributes(&a, k_gather_final);
    cudaFuncGetAttributes(&a, k_weights_out);
    cudaFuncGetAttributes(&a, k_init_sweep);
    cudaFuncGetAttributes(&a, k_plan_multi);
    cudaFuncGetAttributes(&a, k_fill_fat);
    cudaFuncGetAttributes(&a, k_smooth_step);
    cudaFuncGetAttributes(&a, k_traj_step);
    cudaFuncGetAttributes(&a, k_pc_provisional);
    cudaFuncGetAttributes(&a, k_route_barrier);
    cudaGetLastError();
    done = true;
}

// children from which a parent is deferred to the fat list: well above what one expand pass holds,
// and at least Ng / 128 so that a list never exceeds APS_FAT_MAX entries (APS_FAT_MIN: test override)
static int fat_min_for(long long n_global) {
    long long m = (n_global + APS_FAT_MAX - 1) / APS_FAT_MAX;
    long long want = 16LL * APS_K3_CAP;
    if (const char *e = getenv("APS_FAT_MIN")) want = atoll(e);
    if (m < want) m = want;
    if (m < 8) m = 8;
    return (int)(m > 2147483647LL ? 2147483647LL : m);
}

static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}
// grid for the grid-stride kernels: a multiple of the SM count, 8 resident CTAs of 256 per SM
static int stride_grid(long long n) {
    long long need = (n + APS_K1_THREADS - 1) / APS_K1_THREADS;
    long long cap = (long long)sm_count() * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#ifndef APS_PREDRAW_MIN_TILES
#define APS_PREDRAW_MIN_TILES 190
#endif
// ------------------------------------------------------------------ handle
struct aps_handle {
    aps_config cfg;
    DevCtx ctx;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    // pre-drawn normals (k_draw_normals) on a parallel, low-priority branch of the sweep's graph
    cudaStream_t stream_draw[2];
    cudaEvent_t ev_fork, ev_join[2];
    void (*f_draw)(DevCtx, long long);
    int grid_draw, threads_draw, draw_ahead;   // draw_ahead: 1 or 2 steps (2: two buffers, the draws of step t + 2 are forked behind propagate(t))
    long long zbuf_stride;
    SweepParams *d_sp;
    double *d_ref, *d_traj, *d_Y, *d_scratch;  // d_scratch: N x d doubles for accessors
    // multinomial / residual resampling scratch
    u64 *d_cum, *d_rq;
    unsigned short *d_cut;
    unsigned char *d_cut_sh;
    int *d_counts, *d_tile_count, *d_tile_cprefix;
    ResidualState *d_rs;   // one per decision point
    unsigned *d_done2;     // one per decision point
    double *h_weights;                          // pinned staging of aps_get_weights_view (lazily allocated)
    SweepParams *h_sp;                          // pinned
    SweepState *h_st;                           // pinned
    cudaGraphExec_t graph;
    bool graph_ready, has_obs, swept, ref_valid;
    // multi-GPU
    MailSlot *d_mail;
    PeerTable *d_peers;
    void *ipc_opened[3 * APS_MAX_RANKS];
    int n_ipc_opened;
    bool comm_ready;
    unsigned long long epoch, pick_seq;
    float last_ms;
    long long last_launches, graph_nodes;
    // fused persistent sweep (single GPU, systematic / stratified): kernel, geometry, exchange arrays
    fused_fn f_fused;
    FusedArgs fa;
    int fused_grid, fused_threads, fused_smem;
    bool last_fused, fused_forced;
    int grid_prop;   // grid of the propagate kernel (propagate_grid), computed on first use
    bool prop_per_slot;   // k_propagate1 (d = 4): one thread per slot
    bool pdl;        // programmatic dependent launch between the three kernels of a step (single GPU, systematic / stratified, SMC / PG)
    // stepwise container (aps_pc_*): reweights done so far, decision points settled so far
    bool pc_active;
    long long pc_t, pc_decided;
    CUtensorMap tmap_q;
    prop_fn f_prop;
    prop_fn f_prop_pre;   // the variant that loads pre-drawn normals (null: no pre-draw on this handle)
    int grid_prop_pre;
    res_fn f_res;
    pgas_fn f_pmax, f_psel;
};

static void free_handle(aps_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->graph_ready) cudaGraphExecDestroy(h->graph);
    cudaFree(h->ctx.zbuf);
    for (int k = 0; k < 2; ++k) {
        if (h->stream_draw[k]) cudaStreamDestroy(h->stream_draw[k]);
        if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    cudaFree(h->ctx.x);
    cudaFree(h->ctx.anc);
    cudaFree(h->ctx.logw);
    cudaFree(h->ctx.q);
    cudaFree(h->ctx.tile_sum);
    cudaFree(h->ctx.tile_s1);
    cudaFree(h->ctx.tile_s2);
    cudaFree(h->ctx.tile_prefix);
    cudaFree(h->ctx.acc);
    cudaFree(h->ctx.plan);
    cudaFree(h->ctx.st);
    cudaFree(h->d_sp);
    cudaFree(h->d_ref);
    cudaFree(h->d_traj);
    cudaFree(h->d_Y);
    cudaFree(h->d_scratch);
    cudaFree(h->d_cum);
    cudaFree(h->d_cut);
    cudaFree(h->d_cut_sh);
    cudaFree(h->d_rq);
    cudaFree(h->d_counts);
    cudaFree(h->d_tile_count);
    cudaFree(h->d_tile_cprefix);
    cudaFree(h->d_rs);
    cudaFree(h->d_done2);
    for (int i = 0; i < h->n_ipc_opened; ++i) cudaIpcCloseMemHandle(h->ipc_opened[i]);
    cudaFree(h->d_mail);
    cudaFree(h->d_peers);
    cudaFree(h->ctx.rank_woff);
    cudaFree(h->fa.ex_max);
    cudaFree(h->fa.ex_pmax);
    cudaFree(h->fa.ex_tot);
    cudaFree(h->fa.sub_prefix);
    cudaFree(h->fa.dbg);
    cudaFree(h->fa.ctr);
    cudaFree(h->fa.qp);
    if (h->h_weights) cudaFreeHost(h->h_weights);
    if (h->h_sp) cudaFreeHost(h->h_sp);
    if (h->h_st) cudaFreeHost(h->h_st);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int aps_create(const aps_config *cfg, aps_handle **out) {
    if (!cfg || !out) return fail(APS_ERR_INVALID, "aps_create: null argument");
    *out = nullptr;
    const long long N = cfg->n_particles, T = cfg->n_steps;
    if (N < 1 || N > 2147483647LL) return fail(APS_ERR_INVALID, "aps_create: n_particles must be in 1..2^31-1");
    if (T < 1) return fail(APS_ERR_INVALID, "aps_create: n_steps must be >= 1");
    if (cfg->sampler < APS_SMC || cfg->sampler > APS_PGAS) return fail(APS_ERR_INVALID, "aps_create: unknown sampler");
    if (cfg->resampler < APS_RESAMPLE_MULTINOMIAL || cfg->resampler > APS_RESAMPLE_SYSTEMATIC)
        return fail(APS_ERR_INVALID, "aps_create: unknown resampler");
    if (cfg->sampler != APS_SMC && !cfg->keep_history)
        return fail(APS_ERR_INVALID, "aps_create: PG / PGAS need keep_history = 1 (trajectory extraction)");
    const int world = cfg->world_size;
    if (world < 1 || world > APS_MAX_RANKS || cfg->rank < 0 || cfg->rank >= world)
        return fail(APS_ERR_INVALID, "aps_create: rank / world_size out of range (1..8 ranks)");
    if (world > 1) {
        if (N % ((long long)world * 32) != 0)
            return fail(APS_ERR_INVALID, "aps_create: n_particles must be a multiple of 32 * world_size");
    }
    aps_handle *h = new aps_handle();
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    if (aps_model_prepare(&cfg->model, &h->ctx.md)) {
        delete h;
        return fail(APS_ERR_INVALID, "aps_create: invalid model (dimensions 1..4, positive noise scales)");
    }
#define CUH(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) {                                           \
            free_handle(h);                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? APS_ERR_NOMEM : APS_ERR_CUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e_)); \
        }                                                                  \
    } while (0)
    CUH(cudaSetDevice(cfg->device));
    CUH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUH(cudaEventCreate(&h->ev0));
    CUH(cudaEventCreate(&h->ev1));
    DevCtx &c = h->ctx;
    const int d = cfg->model.d;
    const long long Nl = N / world;  // this rank's shard of the N global particles
    c.N = Nl;
    c.Ng = N;
    c.slot0 = (long long)cfg->rank * Nl;
    c.rank = cfg->rank;
    c.world = world;
    c.peers = nullptr;
    c.dbg = getenv("APS_DEBUG_MULTI") ? atoi(getenv("APS_DEBUG_MULTI")) : 0;
    c.NS = (Nl + 31) & ~31LL;
    c.T = T;
    c.d = d;
    c.dy = cfg->model.dy;
    c.S = aps_weight_shift((uint64_t)N);
    c.Hs = aps_ess_shift((uint64_t)N);
    c.sampler = cfg->sampler;
    c.resampler = cfg->resampler;
    c.bare = (cfg->ess_threshold != cfg->ess_threshold) ? 1 : 0;
    c.ess_threshold = cfg->ess_threshold;
    c.logN = aps_log((double)N);
    c.n_override = 0;
    c.x_slabs = cfg->keep_history ? T : 2;
    c.anc_slabs = cfg->keep_history ? T + 1 : 2;
    c.num_tiles = (Nl + APS_TILE - 1) / APS_TILE;
    CUH(cudaMalloc(&c.x, sizeof(double) * (size_t)c.x_slabs * d * c.NS));
    CUH(cudaMalloc(&c.anc, sizeof(int32_t) * (size_t)c.anc_slabs * c.NS));
    CUH(cudaMalloc(&c.logw, sizeof(double) * (size_t)c.NS));
    CUH(cudaMalloc(&c.q, sizeof(u64) * (size_t)c.NS));
    CUH(cudaMalloc(&c.tile_sum, sizeof(u64) * (size_t)c.num_tiles));
    CUH(cudaMalloc(&c.tile_s1, sizeof(u64) * (size_t)c.num_tiles));
    CUH(cudaMalloc(&c.tile_s2, sizeof(u64) * (size_t)c.num_tiles));
    CUH(cudaMalloc(&c.tile_prefix, sizeof(u64) * (size_t)c.num_tiles));
    CUH(cudaMalloc(&c.acc, sizeof(StepAcc) * (size_t)(T + 2)));
    CUH(cudaMalloc(&c.plan, sizeof(StepPlan) * (size_t)(T + 2)));
    CUH(cudaMalloc(&c.st, sizeof(SweepState)));
    CUH(cudaMalloc(&h->d_sp, sizeof(SweepParams)));
    CUH(cudaMalloc(&h->d_ref, sizeof(double) * (size_t)T * d));
    CUH(cudaMalloc(&h->d_traj, sizeof(double) * (size_t)T * d));
    CUH(cudaMalloc(&h->d_Y, sizeof(double) * (size_t)T * c.dy));
    CUH(cudaMalloc(&h->d_scratch, sizeof(double) * (size_t)Nl * d));
    // mailbox + fat-parent lists in one allocation (one IPC handle covers both)
    c.fat_steps = T + 2;
    // sharded multinomial / residual: receive buffer of the routed draws (worst case: every draw lands here)
    const bool iid = cfg->resampler == APS_RESAMPLE_MULTINOMIAL || cfg->resampler == APS_RESAMPLE_RESIDUAL;
    c.recv_cap = (world > 1 && iid) ? N : 0;
    CUH(cudaMalloc(&h->d_mail, aps_mailbox_alloc_bytes(c.fat_steps, c.recv_cap)));
    CUH(cudaMemset(h->d_mail, 0, aps_recv_off(c.fat_steps)));
    c.recv_cnt = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(h->d_mail) + aps_recvcnt_off(c.fat_steps));
    c.recv = reinterpret_cast<u64 *>(reinterpret_cast<char *>(h->d_mail) + aps_recv_off(c.fat_steps));
    c.rank_woff = nullptr;
    if (c.recv_cap) CUH(cudaMalloc(&c.rank_woff, sizeof(u64) * (size_t)c.fat_steps * (APS_MAX_RANKS + 1)));
    c.fat_cnt = reinterpret_cast<int *>(reinterpret_cast<char *>(h->d_mail) + aps_mail_bytes());
    c.fat = reinterpret_cast<FatEntry *>(reinterpret_cast<char *>(h->d_mail) + aps_mail_bytes() + aps_fatcnt_bytes(c.fat_steps));
    c.fat_min = fat_min_for(N);
    // deferred plan: one GPU, systematic / stratified, few enough tiles that every resample block can
    // afford to sum them (<= 8 loads per thread and array)
    c.defer_plan = (c.num_tiles <= 1024 && getenv("APS_NO_DEFER_PLAN") == nullptr &&
                    (cfg->resampler == APS_RESAMPLE_SYSTEMATIC || cfg->resampler == APS_RESAMPLE_STRATIFIED))
                       ? 1 : 0;
    if (world > 1) {
        CUH(cudaMalloc(&h->d_peers, sizeof(PeerTable)));
        if (const char *e = getenv("APS_COMM_TIMEOUT_MS")) {  // spin budget of one exchange wait (default 3000 ms)
            const long long cyc = atoll(e) * 2000000LL;
            if (cyc > 0) CUH(cudaMemcpyToSymbol(g_spin_limit, &cyc, sizeof(cyc)));
        }
    }
    if (cfg->resampler == APS_RESAMPLE_MULTINOMIAL || cfg->resampler == APS_RESAMPLE_RESIDUAL) {
        CUH(cudaMalloc(&h->d_cum, sizeof(u64) * (size_t)c.NS));
        CUH(cudaMalloc(&h->d_cut, sizeof(unsigned short) * (size_t)c.num_tiles * APS_CUT));
        CUH(cudaMalloc(&h->d_cut_sh, (size_t)c.num_tiles));
        CUH(cudaMalloc(&h->d_counts, sizeof(int) * (size_t)c.NS));
        CUH(cudaMalloc(&h->d_tile_count, sizeof(int) * (size_t)c.num_tiles));
        CUH(cudaMalloc(&h->d_tile_cprefix, sizeof(int) * (size_t)c.num_tiles));
        CUH(cudaMalloc(&h->d_rs, sizeof(ResidualState) * (size_t)(T + 2)));
        CUH(cudaMalloc(&h->d_done2, sizeof(unsigned) * (size_t)(T + 2)));
        if (cfg->resampler == APS_RESAMPLE_RESIDUAL) CUH(cudaMalloc(&h->d_rq, sizeof(u64) * (size_t)c.NS));
    }
    CUH(cudaMallocHost(&h->h_sp, sizeof(SweepParams)));
    CUH(cudaMallocHost(&h->h_st, sizeof(SweepState)));
    CUH(cudaMemset(c.plan, 0, sizeof(StepPlan) * (size_t)(T + 2)));
    CUH(cudaMemset(c.q, 0, sizeof(u64) * (size_t)c.NS));  // the padding past N must read as zero weight
    if (make_q_tensormap(&h->tmap_q, c.q, c.NS) || enable_k3_smem()) {
        free_handle(h);
        return APS_ERR_CUDA;
    }
    preload_kernels();
    c.Y = h->d_Y;
    c.ref = h->d_ref;
    c.sp = h->d_sp;
    h->f_prop = pick_propagate(cfg->model.obs_kind, d, cfg->model.dy, world > 1);
    h->prop_per_slot = d == 4 && getenv("APS_K1_PAIRS") == nullptr;
    // Pre-drawn normals (pair kernel, d <= 3): k_draw_normals(t + 1) runs beside k_normalise(t) / k_resample(t)
    // Measured at the headline shape (N = 1e6 per GPU, profiles/README.md): one 128-thread block per SM, forked
    // behind the propagate kernel, one step ahead; more threads per SM and the neighbouring kernel's span
    // jumps to the draw kernel's duration. Enabled where the normalise / resample kernels are a single
    // wave (<= 740 tiles, N <= 1.5e6 per GPU) -- beyond that the window beside them is shorter than the draw
    // kernel at this size and the propagate kernel would wait for it -- and large enough that a step is not
    // launch-bound (>= 190 tiles, N >= ~4e5: at 1e5 / 3e5 the extra graph node and its two edges cost 2-3 %; measured gain 3 % at 4e5, 8.5 % at 1e6, 5 % at 1.5e6).
    // APS_PREDRAW=1 forces, APS_NO_PREDRAW=1 disables.
    // (systematic / stratified steps only: beside the longer multinomial / residual decision kernels a first
    //  measurement showed no gain -- 8.32 / 8.88 ms per sweep against 8.26 / 8.63)
    const bool predraw_ok = (c.num_tiles >= APS_PREDRAW_MIN_TILES && c.num_tiles <= 740 &&
                             (cfg->resampler == APS_RESAMPLE_SYSTEMATIC || cfg->resampler == APS_RESAMPLE_STRATIFIED)) ||
                            (getenv("APS_PREDRAW") != nullptr && atoi(getenv("APS_PREDRAW")) != 0);
    // d = 4 (k_propagate1, one thread per slot): measured only when forced (APS_PREDRAW=1), see DESIGN section 4
    const bool predraw_d4 = d == 4 && h->prop_per_slot && getenv("APS_PREDRAW") != nullptr && atoi(getenv("APS_PREDRAW")) != 0;
    if (((!h->prop_per_slot && d <= 3 && predraw_ok) || predraw_d4) && getenv("APS_NO_PREDRAW") == nullptr) {
        h->f_draw = d == 1 ? k_draw_normals<1> : d == 2 ? k_draw_normals<2> : d == 3 ? k_draw_normals<3> : k_draw_normals<4>;
        h->f_prop_pre = pick_propagate(cfg->model.obs_kind, d, cfg->model.dy, world > 1, true);
        prefer_max_smem(h->f_prop_pre);
        prefer_max_smem(h->f_draw);   // same carve-out as its neighbours: an SM that had to re-partition would serialise them
        const long long npairs = (Nl + 1) / 2;
        h->draw_ahead = 1;
        if (const char *e = getenv("APS_DRAW_AHEAD")) h->draw_ahead = atoi(e) == 2 ? 2 : 1;
        h->zbuf_stride = npairs * 2 * d;
        CUH(cudaMalloc(&c.zbuf, sizeof(double) * (size_t)h->zbuf_stride * 2));
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);   // (numerically: lo = least urgent)
        for (int k = 0; k < 2; ++k) {
            CUH(cudaStreamCreateWithPriority(&h->stream_draw[k], cudaStreamNonBlocking, lo_prio));
            CUH(cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming));
        }
        CUH(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        int bps = 1;   // blocks per SM of the draw kernel: a background trickle that fills idle issue slots, not SMs
        if (const char *e = getenv("APS_DRAW_BPS")) bps = atoi(e) > 0 ? atoi(e) : bps;
        h->threads_draw = 128;
        if (const char *e = getenv("APS_DRAW_THREADS")) h->threads_draw = atoi(e) >= 32 && atoi(e) <= 512 ? (atoi(e) & ~31) : 128;
        long long g = (npairs + h->threads_draw - 1) / h->threads_draw;
        if (g > (long long)bps * sm_count()) g = (long long)bps * sm_count();
        h->grid_draw = (int)(g < 1 ? 1 : g);
    }
    prefer_max_smem(h->f_prop);
    h->f_res = pick_resample(cfg->resampler, world > 1, c.defer_plan != 0);
    h->f_pmax = pick_pgas_max(d);
    h->f_psel = pick_pgas_select(d);
    prefer_max_smem(h->f_pmax);
    prefer_max_smem(h->f_psel);
    // ---- fused persistent sweep: one CTA per SM, every CTA owns a contiguous chunk of slots
    // PDL: only where a step is exactly K1 -> K2 -> K3 (every kernel of the chain carries the wait)
    h->pdl = world == 1 && (cfg->resampler == APS_RESAMPLE_SYSTEMATIC || cfg->resampler == APS_RESAMPLE_STRATIFIED) &&
             cfg->sampler != APS_PGAS && APS_PDL && getenv("APS_PDL") != nullptr && atoi(getenv("APS_PDL")) != 0;
    h->f_fused = nullptr;
    h->fused_forced = getenv("APS_FUSED") != nullptr && atoi(getenv("APS_FUSED")) != 0;
    if (world == 1 && getenv("APS_NO_FUSED") == nullptr) {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device);
        fused_fn fn = coop ? pick_fused(cfg->resampler, cfg->model.obs_kind, d, cfg->model.dy) : nullptr;
        if (fn) {
            const long long SUB = APS_FUSED_SUB;
            long long G = (Nl + SUB - 1) / SUB;
            if (G > sm_count()) G = sm_count();
            long long Nc = ((Nl + G - 1) / G + SUB - 1) / SUB * SUB;
            G = (Nl + Nc - 1) / Nc;
            const long long P = Nc / 2;                          // slot pairs per chunk
            const long long rounds = (P + APS_FUSED_MAX_THREADS - 1) / APS_FUSED_MAX_THREADS;
            long long NT = (((P + rounds - 1) / rounds) + 31) & ~31LL;   // fewest idle lanes in the last round
            if (NT < ((G + 31) & ~31LL)) NT = (G + 31) & ~31LL;
            if (NT < 64) NT = 64;
            if (NT > APS_FUSED_MAX_THREADS) NT = APS_FUSED_MAX_THREADS;
            if (const char *e = getenv("APS_FUSED_THREADS")) NT = atoll(e);   // tuning experiments
            const long long nsub = Nc / SUB;
            const long long smem_ll = NT * APS_FUSED_CPT * (long long)sizeof(int) + (nsub + 1) * (long long)sizeof(u64) * (1 + APS_FUSED_GRP) +
                                      APS_FUSED_GRP * nsub * (long long)sizeof(int) + 16;
            const int smem = (int)smem_ll;
            int occ = 0;
            if (Nl <= (1LL << 30) && smem_ll <= 200 * 1024 && G <= APS_FUSED_MAX_CTAS && NT >= G && NT <= APS_FUSED_MAX_THREADS &&
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, (int)NT, smem) == cudaSuccess &&
                (long long)occ * sm_count() >= G) {
                h->f_fused = fn;
                h->fused_grid = (int)G;
                h->fused_threads = (int)NT;
                h->fused_smem = smem;
                h->fa.chunk = (int)Nc;
                h->fa.nsub = (int)nsub;
                CUH(cudaMalloc(&h->fa.ex_max, sizeof(ulonglong2) * (size_t)G));
                CUH(cudaMalloc(&h->fa.ex_pmax, sizeof(ulonglong2) * (size_t)G));
                CUH(cudaMalloc(&h->fa.ex_tot, sizeof(ulonglong2) * 3 * (size_t)G));
                CUH(cudaMalloc(&h->fa.sub_prefix, sizeof(u64) * (size_t)(G * (nsub + 1))));
                CUH(cudaMalloc(&h->fa.dbg, sizeof(u64) * 8 * (size_t)G));
                CUH(cudaMalloc(&h->fa.ctr, sizeof(u64) * 2));
                CUH(cudaMemset(h->fa.ex_max, 0, sizeof(ulonglong2) * (size_t)G));
                CUH(cudaMemset(h->fa.ex_pmax, 0, sizeof(ulonglong2) * (size_t)G));
                CUH(cudaMemset(h->fa.ex_tot, 0, sizeof(ulonglong2) * 3 * (size_t)G));
                CUH(cudaMemset(h->fa.dbg, 0, sizeof(u64) * 8 * (size_t)G));
                if (cfg->sampler == APS_PGAS) {
                    CUH(cudaMalloc(&h->fa.qp, sizeof(u64) * 2 * (size_t)c.NS));
                    CUH(cudaMemset(h->fa.qp, 0, sizeof(u64) * 2 * (size_t)c.NS));
                }
            }
            cudaGetLastError();
        }
    }
#undef CUH
    *out = h;
    return APS_OK;
}

extern "C" int aps_destroy(aps_handle *h) {
    free_handle(h);
    return APS_OK;
}

extern "C" int aps_set_observations(aps_handle *h, const double *Y, int64_t T, int64_t dy) {
    if (!h || !Y) return fail(APS_ERR_INVALID, "aps_set_observations: null argument");
    if (T != h->cfg.n_steps || dy != h->cfg.model.dy)
        return fail(APS_ERR_INVALID, "aps_set_observations: shape does not match the handle (T x dy)");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h->d_Y, Y, sizeof(double) * (size_t)T * dy, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->has_obs = true;
    return APS_OK;
}

// optional per-launch CUDA events (aps_sweep_profiled): class 0 propagate, 1 normalise, 2 resample, 3 PGAS
struct LaunchProf {
    std::vector<cudaEvent_t> ev;
    std::vector<int> cls;
    cudaStream_t st;
    void begin(int c) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        cls.push_back(c);
    }
    void end() {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
    }
};

// the kernels of decision point t (resample_propagate! after the t-th reweight!): the resampler of
// the handle and, for PGAS, the ancestor draw. `c` may differ from h->ctx (stepwise container).
struct Launcher {
    LaunchProf *prof;
    long long n;
    void begin(int cls) { if (prof) prof->begin(cls); }
    void end() { if (prof) prof->end(); ++n; }
};
#define APS_LAUNCH(cls_, ...) \
    do {                      \
        L.begin(cls_);        \
        __VA_ARGS__;          \
        L.end();              \
    } while (0)

// launch with the programmatic-stream-serialization attribute (PDL): the kernel may begin while the
// previous kernel of the stream drains; see APS_PDL_WAIT in csrc/aps_device.cuh
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*fn)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, fn, KArgs(args)...);
}

static double *x_slab_of(const DevCtx &c, long long t) { return c.x + ((t - 1 + c.x_slabs) % c.x_slabs) * (long long)c.d * c.NS; }
static int32_t *anc_slab_of(const DevCtx &c, long long sidx) { return c.anc + ((sidx + c.anc_slabs) % c.anc_slabs) * c.NS; }

// Grid of the propagate kernel: one thread per slot pair, grid-stride. The grid is (a) no larger than
// what is resident at once (occupancy x SMs: a second wave of a few blocks would double the kernel's
// tail) and (b) sized so that every thread runs the SAME number of iterations: with npairs / resident
// threads = 2.64 (N = 1e6) a maximal grid leaves a third of the blocks idle for the last third of
// the kernel; ceil(npairs / (iterations x threads)) blocks spread the same work evenly.
static int propagate_grid(aps_handle *h, long long n_local, prop_fn fn) {
    // work items: slot pairs, or slots for the one-thread-per-slot kernel of d = 4
    const long long npairs = h->prop_per_slot ? n_local : (n_local + 1) / 2;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, APS_K1_THREADS, 0) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
#ifdef APS_K1_GRID_PER_SM   // tuning experiments
    if (occ > APS_K1_GRID_PER_SM) occ = APS_K1_GRID_PER_SM;
#endif
    const long long resident = (long long)occ * sm_count() * APS_K1_THREADS;
    long long iters = (npairs + resident - 1) / resident;
    if (const char *e = getenv("APS_K1_ITERS")) {   // tuning experiments (single GPU only: sharded grids must be resident)
        if (h->ctx.world == 1 && atoll(e) > 0) iters = atoll(e);
    }
    long long g = (npairs + iters * APS_K1_THREADS - 1) / (iters * APS_K1_THREADS);
    if (g < 1) g = 1;
    return (int)g;
}

// with_draw: the normals of step t are drawn right here, on the same stream, and count as part of the propagate
// launch (per-launch profiling and APS_DRAW_SERIAL: no parallel branch)
static void launch_propagate(aps_handle *h, const DevCtx &c, long long t, Launcher &L, bool with_draw = false) {
    cudaStream_t st = h->stream;
    if (h->grid_prop == 0) h->grid_prop = propagate_grid(h, c.N, h->f_prop);
    if (h->f_prop_pre && h->grid_prop_pre == 0) h->grid_prop_pre = propagate_grid(h, c.N, h->f_prop_pre);
    const bool pre = c.zbuf != nullptr && h->f_prop_pre != nullptr;   // the normals of this step were drawn ahead
    // slab addressing is resolved here, once per launch (no 64-bit modulo in the kernels)
    APS_LAUNCH(0, {
        if (with_draw) {
            h->f_draw<<<h->grid_draw, h->threads_draw, 0, st>>>(c, t);
            ++L.n;
        }
        launch_pdl(pre ? h->f_prop_pre : h->f_prop, pre ? h->grid_prop_pre : h->grid_prop, APS_K1_THREADS, 0, st, h->pdl && t > 1, c,
                   (long long)t, x_slab_of(c, t), (const double *)x_slab_of(c, t - 1), (const int32_t *)anc_slab_of(c, t - 1));
    });
}

static void launch_decision(aps_handle *h, const DevCtx &c, long long t, res_fn f_res, Launcher &L) {
    cudaStream_t st = h->stream;
    const int gp = stride_grid(c.N);
    const int gt = (int)c.num_tiles;
    if (c.resampler == APS_RESAMPLE_SYSTEMATIC || c.resampler == APS_RESAMPLE_STRATIFIED) {
        APS_LAUNCH(2, launch_pdl(f_res, gt, APS_K3_THREADS, APS_K3_DYN_SMEM, st, h->pdl, c, (long long)t, anc_slab_of(c, t), h->tmap_q));
    } else {
        const bool multi = c.world > 1;
        MultiArgs a;
        memset(&a, 0, sizeof(a));
        a.cum = h->d_cum;
        a.cut = h->d_cut;
        a.cut_sh = h->d_cut_sh;
        a.counts = h->d_counts;
        a.tile_count = h->d_tile_count;
        a.tile_cprefix = h->d_tile_cprefix;
        a.tile_prefix = c.tile_prefix;
        a.plan = c.plan + t;
        a.N = c.N;
        a.num_tiles = c.num_tiles;
        a.step = t;
        // sharded: the plan of this decision point comes from the exchanged shard totals
        if (multi) {
            APS_LAUNCH(2, k_plan_multi<<<1, 32, 0, st>>>(c, t));
            a.child_off = &c.acc[t].child_off;
        }
        if (c.resampler == APS_RESAMPLE_MULTINOMIAL) {
            a.qsrc = c.q;
            a.wplan = c.plan + t;
            a.n_draws = &c.plan[t].n;
            if (multi) {
                a.range_lo = &c.acc[t].rank_off;
                a.range_len = &c.acc[t].tot[0];
            }
            cudaMemsetAsync(h->d_counts, 0, sizeof(int) * (size_t)c.N, st);
        } else {
            a.qsrc = h->d_rq;
            a.wplan = c.plan + c.T + 1;
            a.n_draws = &h->d_rs[t].n_rest;
            if (multi) {
                a.range_lo = &h->d_rs[t].q_off;
                a.range_len = &h->d_rs[t].q_local;
            }
            APS_LAUNCH(2, k_residual_split<<<gt, APS_THREADS, 0, st>>>(a, c.q, h->d_rq, h->d_rs + t));
            if (multi) APS_LAUNCH(2, k_residual_exchange<0><<<1, 32, 0, st>>>(c, t, c.plan + t, h->d_rs + t, c.plan + c.T + 1));
            APS_LAUNCH(2, k_residual_weights<<<gt, APS_THREADS, 0, st>>>(a, h->d_rq, h->d_rs + t, c.tile_sum, c.tile_prefix,
                                                                       c.plan + c.T + 1, h->d_done2 + t, &c.st->err));
            if (multi) APS_LAUNCH(2, k_residual_exchange<1><<<1, 32, 0, st>>>(c, t, c.plan + t, h->d_rs + t, c.plan + c.T + 1));
        }
        APS_LAUNCH(2, k_cumsum<<<gt, APS_THREADS, 0, st>>>(a));
        if (multi && c.recv_cap && getenv("APS_NO_ROUTE") == nullptr) {
            // sharded: every rank makes 1/G of the draws and routes each to the rank that owns its weight range
            const int gr = stride_grid(((c.Ng + 1) / 2 + c.world - 1) / c.world);
            APS_LAUNCH(2, k_multi_route<<<gr, APS_K1_THREADS, 0, st>>>(a, &h->d_sp->key, c, t));
            APS_LAUNCH(2, k_route_barrier<<<1, 32, 0, st>>>(c, t, c.plan + t));
            APS_LAUNCH(2, k_multi_search_recv<<<stride_grid(c.N), APS_K1_THREADS, 0, st>>>(a, c, t));
        } else {
            // (single GPU; or APS_NO_ROUTE: every rank makes all Ng draws and keeps those in its own weight range)
            const int gs = stride_grid((c.Ng + 1) / 2);
            APS_LAUNCH(2, k_multi_search<1><<<gs, APS_K1_THREADS, 0, st>>>(a, &h->d_sp->key));
        }
        APS_LAUNCH(2, k_tile_counts<<<gt, APS_THREADS, 0, st>>>(a));
        APS_LAUNCH(2, k_scan_tile_counts<<<1, APS_THREADS, 0, st>>>(a, nullptr, c, t));
        APS_LAUNCH(2, k_expand_counts<<<gt, APS_THREADS, 0, st>>>(a, anc_slab_of(c, t), 1, c, t));
    }
    if (c.sampler == APS_PGAS && t >= 2 && t <= c.T - 1) {
        APS_LAUNCH(3, h->f_pmax<<<gp, APS_K1_THREADS, 0, st>>>(c, t, x_slab_of(c, t - 1), anc_slab_of(c, t - 1), anc_slab_of(c, t)));
        APS_LAUNCH(3, h->f_psel<<<gt, APS_THREADS, 0, st>>>(c, t, x_slab_of(c, t - 1), anc_slab_of(c, t - 1), anc_slab_of(c, t)));
    }
}

static void launch_fill_fat(aps_handle *h, const DevCtx &c, long long s, Launcher &L) {
    // children of fat parents at the final decision point (no propagate kernel follows to resolve them)
    // (sharded: every block spins on the peers first, so the grid stays small -- ranks emulated on one
    // GPU must all fit at once; the fill itself is a few MB at most, once per sweep)
    APS_LAUNCH(2, k_fill_fat<<<c.world > 1 ? 16 : sm_count() * 2, APS_K1_THREADS, 0, h->stream>>>(c, s, anc_slab_of(c, s), 1));
}

static void reset_sweep_scratch(aps_handle *h) {
    const DevCtx &c = h->ctx;
    cudaStream_t st = h->stream;
    cudaMemsetAsync(c.acc, 0, sizeof(StepAcc) * (size_t)(c.T + 2), st);
    cudaMemsetAsync(c.fat_cnt, 0, sizeof(int) * (size_t)c.fat_steps, st);
    if (c.recv_cap) cudaMemsetAsync(c.recv_cnt, 0, sizeof(unsigned long long) * (size_t)c.fat_steps, st);
    if (h->d_rs) {
        cudaMemsetAsync(h->d_rs, 0, sizeof(ResidualState) * (size_t)(c.T + 2), st);
        cudaMemsetAsync(h->d_done2, 0, sizeof(unsigned) * (size_t)(c.T + 2), st);
    }
}

// enqueue every kernel of one sweep on the handle's stream; returns the number of launches
static long long enqueue_sweep(aps_handle *h, LaunchProf *prof = nullptr) {
    const DevCtx &c = h->ctx;
    cudaStream_t st = h->stream;
    Launcher L{prof, 0};
    reset_sweep_scratch(h);
    k_init_sweep<<<1, 32, 0, st>>>(c);
    ++L.n;
    const int gt = (int)c.num_tiles;
    // Pre-drawn normals (k_draw_normals). The draws of step u are forked behind propagate(u - A) -- which has
    // consumed the buffer they go to -- onto a low-priority stream, run beside the kernels in between, and are
    // joined before propagate(u); A = draw_ahead = 2 gives them more than a whole step (two buffers, two
    // streams: consecutive draw kernels may overlap). Under stream capture these are parallel branches of the
    // graph. Per-launch profiling (aps_sweep_profiled) times the classical step instead -- the propagate kernel
    // draws its own normals -- so that its per-kernel figures describe whole kernels, not a background kernel
    // run in the foreground; APS_DRAW_SERIAL=1 keeps the draw kernel but launches it on the sweep's stream.
    const bool draw = c.zbuf != nullptr && prof == nullptr;
    const bool fork = draw && getenv("APS_DRAW_SERIAL") == nullptr;
    const int A = fork ? h->draw_ahead : 1;
    const bool fork_late = fork && getenv("APS_DRAW_FORK") != nullptr && atoi(getenv("APS_DRAW_FORK")) == 2;
    auto ctx_for = [&](long long u) {   // the context of step u: its buffer of normals
        DevCtx cc = c;
        cc.zbuf = draw ? c.zbuf + (A == 2 ? (u & 1) * h->zbuf_stride : 0) : nullptr;
        return cc;
    };
    if (fork)
        for (long long u = 1; u <= A && u <= c.T; ++u) {
            h->f_draw<<<h->grid_draw, h->threads_draw, 0, st>>>(ctx_for(u), u);
            ++L.n;
        }
    for (long long t = 1; t <= c.T; ++t) {
        launch_propagate(h, ctx_for(t), t, L, draw && !fork);
        const long long u = t + A;   // the draws that may now overwrite the buffer propagate(t) has read
        auto fork_draws = [&]() {
            cudaStream_t sd = h->stream_draw[u & 1];
            cudaEventRecord(h->ev_fork, st);
            cudaStreamWaitEvent(sd, h->ev_fork, 0);
            h->f_draw<<<h->grid_draw, h->threads_draw, 0, sd>>>(ctx_for(u), u);
            cudaEventRecord(h->ev_join[u & 1], sd);
            ++L.n;
        };
        if (fork && u <= c.T && !fork_late) fork_draws();
        APS_LAUNCH(1, launch_pdl(k_normalise<IN_LOGW>, gt, APS_K2_THREADS, 0, st, h->pdl, c, (const double *)c.logw, (long long)t));
        if (fork && u <= c.T && fork_late) fork_draws();   // beside the resample kernel only (its issue slots are 70 % idle)
        launch_decision(h, c, t, h->f_res, L);
        if (fork && t + 1 <= c.T && t + 1 > A) cudaStreamWaitEvent(st, h->ev_join[(t + 1) & 1], 0);   // the draws of step t + 1
    }
    launch_fill_fat(h, c, c.T, L);
    return L.n;
}

static int sweep_impl(aps_handle *h, uint64_t master_seed, const double *ref_traj, double *logevidence,
                      float *class_ms, int64_t *class_launches) {
    if (!h || !logevidence) return fail(APS_ERR_INVALID, "aps_sweep: null argument");
    if (!h->has_obs) return fail(APS_ERR_INVALID, "aps_sweep: observations not set");
    CU(cudaSetDevice(h->cfg.device));
    const DevCtx &c = h->ctx;
    int has_ref = 0;
    if (ref_traj != nullptr) {
        if (h->cfg.sampler == APS_SMC) return fail(APS_ERR_INVALID, "aps_sweep: SMC takes no reference trajectory");
        if (c.N < 2) return fail(APS_ERR_INVALID, "aps_sweep: a conditional sweep needs at least 2 particles");
        if (ref_traj == APS_REF_ON_DEVICE) {
            if (!h->ref_valid) return fail(APS_ERR_INVALID, "aps_sweep: no trajectory has been picked yet");
        } else {
            CU(cudaMemcpyAsync(h->d_ref, ref_traj, sizeof(double) * (size_t)c.T * c.d, cudaMemcpyHostToDevice, h->stream));
            h->ref_valid = true;
        }
        has_ref = 1;
    }
    if (c.world > 1 && !h->comm_ready) return fail(APS_ERR_COMM, "aps_sweep: peers not attached (aps_ipc_export / aps_ipc_import)");
    h->h_sp->key = master_seed;
    h->h_sp->has_ref = has_ref;
    h->h_sp->pad = 0;
    h->h_sp->epoch = h->epoch++;
    CU(cudaMemcpyAsync(h->d_sp, h->h_sp, sizeof(SweepParams), cudaMemcpyHostToDevice, h->stream));
    LaunchProf prof;
    prof.st = h->stream;
    const bool profiled = class_ms != nullptr;
    // Which path runs this sweep. The fused persistent kernel wins where a step is many small launches
    // (PGAS ancestor sampling: 5 kernels per step) or the per-parent work is heavy (stratified);
    // for the plain systematic sweep the three-kernel graph is ~20 % faster at every N measured
    // (profiles/fused_vs_three_kernel_r02.txt), so that stays the default there.
    // APS_FUSED=1 forces the fused kernel wherever it is eligible, APS_NO_FUSED=1 disables it.
    bool fused = !profiled && h->f_fused != nullptr;
    if (fused && !h->fused_forced)
        fused = c.resampler == APS_RESAMPLE_STRATIFIED || (c.sampler == APS_PGAS && has_ref);
    h->last_fused = fused;
    if (fused) {
        CU(cudaMemsetAsync(c.fat_cnt, 0, sizeof(int) * (size_t)c.fat_steps, h->stream));   // (no fat lists on this path)
        CU(cudaMemsetAsync(h->fa.ctr, 0, sizeof(u64) * 2, h->stream));
        CU(cudaEventRecord(h->ev0, h->stream));
        void *args[2] = {(void *)&h->ctx, (void *)&h->fa};
        CU(cudaLaunchCooperativeKernel((const void *)h->f_fused, dim3(h->fused_grid), dim3(h->fused_threads), args,
                                       (size_t)h->fused_smem, h->stream));
        h->last_launches = 1;
        CU(cudaEventRecord(h->ev1, h->stream));
    } else {
    // APS_NO_GRAPH=1: launch the per-step kernels directly instead of replaying a CUDA graph. Needed
    // only where several ranks are EMULATED on one GPU with long sweeps (tests): graphs of > ~1000
    // nodes from different streams were observed not to run concurrently, and ranks that spin on
    // each other then never meet. One process per GPU -- the product path -- replays the graph.
    const bool no_graph = getenv("APS_NO_GRAPH") != nullptr;
    if (!profiled && !no_graph && !h->graph_ready) {
        cudaGraph_t g;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->graph_nodes = enqueue_sweep(h);
        CU(cudaStreamEndCapture(h->stream, &g));
        CU(cudaGraphInstantiate(&h->graph, g, 0));
        CU(cudaGraphDestroy(g));
        h->graph_ready = true;
    }
    CU(cudaEventRecord(h->ev0, h->stream));
    if (profiled) h->last_launches = enqueue_sweep(h, &prof);
    else if (no_graph) h->last_launches = enqueue_sweep(h);
    else {
        CU(cudaGraphLaunch(h->graph, h->stream));
        h->last_launches = h->graph_nodes;
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    }
    CU(cudaMemcpyAsync(h->h_st, c.st, sizeof(SweepState), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    if (profiled) {
        for (int k = 0; k < 4; ++k) {
            class_ms[k] = 0.f;
            if (class_launches) class_launches[k] = 0;
        }
        for (size_t i = 0; i < prof.cls.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, prof.ev[2 * i], prof.ev[2 * i + 1]);
            class_ms[prof.cls[i]] += ms;
            if (class_launches) class_launches[prof.cls[i]] += 1;
        }
        for (cudaEvent_t e : prof.ev) cudaEventDestroy(e);
    }
    h->swept = true;
    h->pc_active = false;
    if (h->h_st->err)
        return fail(h->h_st->err, h->h_st->err == APS_ERR_COMM
                                      ? (fused ? "aps_sweep: a CTA of the persistent sweep kernel never arrived at an exchange (launch not co-resident?)"
                                               : "aps_sweep: a peer rank did not answer within the exchange timeout")
                                      : "aps_sweep: particle weights could not be normalised (all -Inf or NaN log-weights)");
    *logevidence = h->h_st->logev;
    if (c.dbg & 16) {   // in-graph timeline of the three kernels (globaltimer: first block start -> last block end), us per step
        std::vector<StepAcc> acc((size_t)c.T + 2);
        cudaMemcpy(acc.data(), c.acc, sizeof(StepAcc) * acc.size(), cudaMemcpyDeviceToHost);
        double span[3] = {0, 0, 0}, gap[3] = {0, 0, 0};
        long long cnt = 0;
        for (long long t = 2; t <= c.T; ++t) {
            const double b0 = (double)~acc[t].t_first_neg[0], e0 = (double)acc[t].t_last[0];
            const double b1 = (double)~acc[t].t_first_neg[1], e1 = (double)acc[t].t_last[1];
            const double b2 = (double)~acc[t].t_first_neg[2], e2 = (double)acc[t].t_last[2];
            const double e2p = (double)acc[t - 1].t_last[2];
            span[0] += e0 - b0; span[1] += e1 - b1; span[2] += e2 - b2;
            gap[0] += b0 - e2p; gap[1] += b1 - e0; gap[2] += b2 - e1;
            ++cnt;
        }
        fprintf(stderr, "[aps rank %d] per step (us): gap %.2f | K1 %.2f | gap %.2f | K2 %.2f | gap %.2f | K3 %.2f  (first block start -> last block end)\n",
                c.rank, gap[0] / cnt * 1e-3, span[0] / cnt * 1e-3, gap[1] / cnt * 1e-3, span[1] / cnt * 1e-3, gap[2] / cnt * 1e-3, span[2] / cnt * 1e-3);
    }
    if (fused && (c.dbg & 32)) {   // per-phase time of every CTA (ns, summed over the sweep): min / median / max over the CTAs
        std::vector<u64> d((size_t)h->fused_grid * 8);
        cudaMemcpy(d.data(), h->fa.dbg, sizeof(u64) * d.size(), cudaMemcpyDeviceToHost);
        static const char *nm[5] = {"A propagate", "exchange 1", "B quantise", "exchange 2 + plan", "C pull-resample"};
        fprintf(stderr, "[aps fused] %d CTAs x %d threads, chunk %d, %d sub-tiles; us per step (min / median / max over CTAs)\n",
                h->fused_grid, h->fused_threads, h->fa.chunk, h->fa.nsub);
        for (int k = 0; k < 5; ++k) {
            std::vector<double> v;
            for (int g = 0; g < h->fused_grid; ++g) v.push_back(d[(size_t)g * 8 + k] * 1e-3 / (double)c.T);
            std::sort(v.begin(), v.end());
            fprintf(stderr, "[aps fused]   %-18s %7.2f %7.2f %7.2f\n", nm[k], v.front(), v[v.size() / 2], v.back());
        }
    }
    if (getenv("APS_DEBUG_SPIN") && c.world > 1)
        fprintf(stderr, "[aps rank %d] block-0 wait cycles per sweep: max-exchange %llu, totals %llu, scatter-done %llu (T=%lld)\n",
                c.rank, h->h_st->spin[0], h->h_st->spin[1], h->h_st->spin[2], c.T);
    return APS_OK;
}

extern "C" int aps_sweep(aps_handle *h, uint64_t master_seed, const double *ref_traj, double *logevidence) {
    return sweep_impl(h, master_seed, ref_traj, logevidence, nullptr, nullptr);
}

extern "C" int aps_sweep_profiled(aps_handle *h, uint64_t master_seed, const double *ref_traj, double *logevidence,
                                  float *class_ms, int64_t *class_launches) {
    if (!class_ms) return fail(APS_ERR_INVALID, "aps_sweep_profiled: null class_ms");
    return sweep_impl(h, master_seed, ref_traj, logevidence, class_ms, class_launches);
}

// ================================================================== container level (stepwise)
// The device-resident ParticleContainer driven call by call (include/aps_b200.h). Uses the
// per-step kernels with the plan written by the normalise kernel (no deferred plan), so that every
// call leaves the container in a state the accessors can read.
static DevCtx pc_ctx(aps_handle *h) {
    DevCtx c = h->ctx;
    c.defer_plan = 0;
    c.zbuf = nullptr;   // (the stepwise container draws inside the propagate kernel)
    return c;
}
#define NEED_PC(name)                                                                               \
    if (!h) return fail(APS_ERR_INVALID, name ": null handle");                                     \
    if (!h->pc_active) return fail(APS_ERR_INVALID, name ": aps_pc_begin has not been called");     \
    CU(cudaSetDevice(h->cfg.device));

// max + normalise of the current log-weights into (acc[s], plan[s]); s = T + 1 is scratch
static int pc_normalise(aps_handle *h, const DevCtx &c, long long s) {
    CU(cudaMemsetAsync(c.acc + s, 0, sizeof(StepAcc), h->stream));
    k_vector_max<IN_LOGW><<<stride_grid(c.N), APS_K1_THREADS, 0, h->stream>>>(c.logw, c.N, c.acc + s);
    k_normalise<IN_LOGW><<<(int)c.num_tiles, APS_K2_THREADS, 0, h->stream>>>(c, c.logw, s);
    return APS_OK;
}

static int pc_sync_state(aps_handle *h, const char *what) {
    CU(cudaMemcpyAsync(h->h_st, h->ctx.st, sizeof(SweepState), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (h->h_st->err)
        return fail(h->h_st->err, std::string(what) + ": particle weights could not be normalised (all -Inf or NaN log-weights)");
    return APS_OK;
}

extern "C" int aps_pc_begin(aps_handle *h, uint64_t master_seed, const double *ref_traj) {
    if (!h) return fail(APS_ERR_INVALID, "aps_pc_begin: null handle");
    if (!h->has_obs) return fail(APS_ERR_INVALID, "aps_pc_begin: observations not set");
    if (h->ctx.world > 1) return fail(APS_ERR_INVALID, "aps_pc_begin: the stepwise container is single-GPU only");
    CU(cudaSetDevice(h->cfg.device));
    const DevCtx &c = h->ctx;
    int has_ref = 0;
    if (ref_traj != nullptr) {
        if (h->cfg.sampler == APS_SMC) return fail(APS_ERR_INVALID, "aps_pc_begin: SMC takes no reference trajectory");
        if (c.N < 2) return fail(APS_ERR_INVALID, "aps_pc_begin: a conditional sweep needs at least 2 particles");
        if (ref_traj == APS_REF_ON_DEVICE) {
            if (!h->ref_valid) return fail(APS_ERR_INVALID, "aps_pc_begin: no trajectory has been picked yet");
        } else {
            CU(cudaMemcpyAsync(h->d_ref, ref_traj, sizeof(double) * (size_t)c.T * c.d, cudaMemcpyHostToDevice, h->stream));
            h->ref_valid = true;
        }
        has_ref = 1;
    }
    h->h_sp->key = master_seed;
    h->h_sp->has_ref = has_ref;
    h->h_sp->pad = 0;
    h->h_sp->epoch = h->epoch++;
    CU(cudaMemcpyAsync(h->d_sp, h->h_sp, sizeof(SweepParams), cudaMemcpyHostToDevice, h->stream));
    reset_sweep_scratch(h);
    CU(cudaMemsetAsync(c.logw, 0, sizeof(double) * (size_t)c.NS, h->stream));
    k_init_sweep<<<1, 32, 0, h->stream>>>(c);
    // decision point 0 is provisional until resample_propagate! runs on it
    k_pc_provisional<<<1, APS_K1_THREADS, 0, h->stream>>>(c, 0, anc_slab_of(c, 0));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    h->pc_active = true;
    h->pc_t = 0;
    h->pc_decided = 0;
    h->swept = false;
    return APS_OK;
}

extern "C" int aps_pc_reweight(aps_handle *h, int32_t *isdone_out) {
    NEED_PC("aps_pc_reweight");
    const DevCtx c = pc_ctx(h);
    if (h->pc_t >= c.T) {  // every particle is done (src/container.jl:288): nothing changes
        if (isdone_out) *isdone_out = 1;
        return APS_OK;
    }
    const long long t = ++h->pc_t;
    Launcher L{nullptr, 0};
    launch_propagate(h, c, t, L);
    // provisional decision point t: plan (logZ, ESS, ...) of the new weights, identity ancestors,
    // weights kept -- until aps_pc_resample_propagate settles it
    int rc = pc_normalise(h, c, t);
    if (rc) return rc;
    k_pc_provisional<<<stride_grid(c.N), APS_K1_THREADS, 0, h->stream>>>(c, t, anc_slab_of(c, t));
    if (isdone_out) *isdone_out = 0;
    h->swept = h->pc_t == c.T;
    // a weight vector that cannot be normalised is reported by the call that needs it normalised
    // (resample_propagate! / logZ), like upstream
    CU(cudaMemsetAsync(&h->ctx.st->err, 0, sizeof(int), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return APS_OK;
}

extern "C" int aps_pc_resample_propagate(aps_handle *h, int32_t *resampled_out) {
    NEED_PC("aps_pc_resample_propagate");
    const DevCtx c = pc_ctx(h);
    const long long s = h->pc_t;
    int resampled = 0;
    if (s == 0) {
        // particles carry no state yet: resampling them only re-keys (src/container.jl:325 at c = 1);
        // with position-derived counters that is the identity. The decision is the one k_init_sweep made.
        StepPlan p;
        k_init_sweep<<<1, 32, 0, h->stream>>>(c);
        CU(cudaMemcpyAsync(&p, c.plan, sizeof(StepPlan), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        resampled = p.resampled;
    } else {
        CU(cudaMemsetAsync(c.fat_cnt + s, 0, sizeof(int), h->stream));
        if (h->d_rs) {
            CU(cudaMemsetAsync(h->d_rs + s, 0, sizeof(ResidualState), h->stream));
            CU(cudaMemsetAsync(h->d_done2 + s, 0, sizeof(unsigned), h->stream));
        }
        int rc = pc_normalise(h, c, s);
        if (rc) return rc;
        Launcher L{nullptr, 0};
        launch_decision(h, c, s, pick_resample(c.resampler, false, false), L);
        k_fill_fat<<<sm_count() * 2, APS_K1_THREADS, 0, h->stream>>>(c, s, anc_slab_of(c, s), 0);
        StepPlan p;
        CU(cudaMemcpyAsync(&p, c.plan + s, sizeof(StepPlan), cudaMemcpyDeviceToHost, h->stream));
        rc = pc_sync_state(h, "aps_pc_resample_propagate");
        if (rc) return rc;
        resampled = p.resampled;
        if (resampled) CU(cudaMemsetAsync(c.logw, 0, sizeof(double) * (size_t)c.NS, h->stream));  // reset_logweights!, :228
        CU(cudaStreamSynchronize(h->stream));
    }
    h->pc_decided = s + 1;
    if (resampled_out) *resampled_out = resampled;
    return APS_OK;
}

extern "C" int aps_pc_logz(aps_handle *h, double *logz_out) {
    NEED_PC("aps_pc_logz");
    if (!logz_out) return fail(APS_ERR_INVALID, "aps_pc_logz: null output");
    const DevCtx c = pc_ctx(h);
    int rc = pc_normalise(h, c, c.T + 1);
    if (rc) return rc;
    StepPlan p;
    CU(cudaMemcpyAsync(&p, c.plan + c.T + 1, sizeof(StepPlan), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemsetAsync(&h->ctx.st->err, 0, sizeof(int), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (p.err) {
        if (p.M == -INFINITY) {  // logsumexp of all -Inf is -Inf, not an error
            *logz_out = -INFINITY;
            return APS_OK;
        }
        return fail(APS_ERR_WEIGHTS, "aps_pc_logz: NaN or +Inf log-weight");
    }
    *logz_out = p.logZ;
    return APS_OK;
}

extern "C" int aps_set_logweights(aps_handle *h, const double *logw) {
    NEED_PC("aps_set_logweights");
    if (!logw) return fail(APS_ERR_INVALID, "aps_set_logweights: null argument");
    if (h->pc_t == 0) return fail(APS_ERR_INVALID, "aps_set_logweights: call it after the first aps_pc_reweight");
    const DevCtx c = pc_ctx(h);
    CU(cudaMemcpyAsync(c.logw, logw, sizeof(double) * (size_t)c.N, cudaMemcpyHostToDevice, h->stream));
    // keep the provisional plan of the current decision point in step with the new weights
    int rc = pc_normalise(h, c, h->pc_t);
    if (rc) return rc;
    k_pc_provisional<<<stride_grid(c.N), APS_K1_THREADS, 0, h->stream>>>(c, h->pc_t, anc_slab_of(c, h->pc_t));
    CU(cudaMemsetAsync(&h->ctx.st->err, 0, sizeof(int), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    h->pc_decided = h->pc_t;  // the decision point is open again
    return APS_OK;
}

#define NEED_SWEEP(name)                                                        \
    if (!h) return fail(APS_ERR_INVALID, name ": null handle");                 \
    if (!h->swept) return fail(APS_ERR_INVALID, name ": no sweep has run yet"); \
    CU(cudaSetDevice(h->cfg.device));

extern "C" int aps_pick_trajectory(aps_handle *h, double *traj_out, int64_t *index_out) {
    NEED_SWEEP("aps_pick_trajectory");
    if (!h->cfg.keep_history) return fail(APS_ERR_INVALID, "aps_pick_trajectory: handle was created with keep_history = 0");
    const DevCtx &c = h->ctx;
    // sharded: a collective call -- every rank picks (the ranks exchange their candidates) and
    // walks the genealogy through the peer-mapped stores, so each ends up with the trajectory
    if (h->last_fused && !h->pc_active) {
        // the fused sweep keeps per-chunk totals only; the pick over non-uniform final weights walks the
        // 2048-particle tile totals / prefixes: rebuild them from the integer weights (scratch plan / acc slot)
        DevCtx cc = c;
        cc.st = nullptr;
        cc.defer_plan = 0;
        cc.acc = c.acc + (c.T + 1);
        cc.plan = c.plan + (c.T + 1);
        CU(cudaMemsetAsync(cc.acc, 0, sizeof(StepAcc), h->stream));
        k_normalise<IN_Q><<<(int)c.num_tiles, APS_K2_THREADS, 0, h->stream>>>(cc, nullptr, 0);
    }
    k_pick<<<1, APS_THREADS, 0, h->stream>>>(c, c.T, c.T + 1, APS_DOM_PICK, ++h->pick_seq);
    k_backtrace<<<1, 32, 0, h->stream>>>(c, -1, h->d_traj);
    CU(cudaMemcpyAsync(h->d_ref, h->d_traj, sizeof(double) * (size_t)c.T * c.d, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->h_st, c.st, sizeof(SweepState), cudaMemcpyDeviceToHost, h->stream));
    if (traj_out)
        CU(cudaMemcpyAsync(traj_out, h->d_traj, sizeof(double) * (size_t)c.T * c.d, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (h->h_st->picked_slot < 0) return fail(APS_ERR_WEIGHTS, "aps_pick_trajectory: no particle could be selected");
    h->ref_valid = true;
    if (index_out) *index_out = h->h_st->picked_slot;
    return APS_OK;
}

static int final_resampled(aps_handle *h, int *out) {
    StepPlan p;
    CU(cudaMemcpyAsync(&p, h->ctx.plan + h->ctx.T, sizeof(StepPlan), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out = p.resampled;
    return APS_OK;
}

extern "C" int aps_host_alloc(int64_t bytes, void **out) {
    if (!out || bytes <= 0) return fail(APS_ERR_INVALID, "aps_host_alloc: bad argument");
    *out = nullptr;
    CU(cudaMallocHost(out, (size_t)bytes));
    return APS_OK;
}

extern "C" int aps_host_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return APS_OK;
}

extern "C" int aps_get_weights(aps_handle *h, double *w_out) {
    NEED_SWEEP("aps_get_weights");
    if (!w_out) return fail(APS_ERR_INVALID, "aps_get_weights: null output");
    const DevCtx &c = h->ctx;
    k_weights_out<<<stride_grid(c.N), APS_THREADS, 0, h->stream>>>(c.q, c.plan + c.T, c.N, c.Ng, c.S, 1, h->d_scratch);
    CU(cudaMemcpyAsync(w_out, h->d_scratch, sizeof(double) * (size_t)c.N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return APS_OK;
}

extern "C" int aps_get_weights_view(aps_handle *h, const double **w_out) {
    NEED_SWEEP("aps_get_weights_view");
    if (!w_out) return fail(APS_ERR_INVALID, "aps_get_weights_view: null output");
    const DevCtx &c = h->ctx;
    if (!h->h_weights) CU(cudaMallocHost(&h->h_weights, sizeof(double) * (size_t)c.N));  // pinned, owned by the handle
    k_weights_out<<<stride_grid(c.N), APS_THREADS, 0, h->stream>>>(c.q, c.plan + c.T, c.N, c.Ng, c.S, 1, h->d_scratch);
    CU(cudaMemcpyAsync(h->h_weights, h->d_scratch, sizeof(double) * (size_t)c.N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    *w_out = h->h_weights;
    return APS_OK;
}

extern "C" int aps_get_logweights(aps_handle *h, double *logw_out) {
    if (h && h->pc_active) {  // stepwise container: pc.logWs as it stands (reset_logweights! zeroes the buffer there)
        if (!logw_out) return fail(APS_ERR_INVALID, "aps_get_logweights: null output");
        CU(cudaSetDevice(h->cfg.device));
        CU(cudaMemcpyAsync(logw_out, h->ctx.logw, sizeof(double) * (size_t)h->ctx.N, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return APS_OK;
    }
    NEED_SWEEP("aps_get_logweights");
    if (!logw_out) return fail(APS_ERR_INVALID, "aps_get_logweights: null output");
    int res = 0;
    int rc = final_resampled(h, &res);
    if (rc) return rc;
    if (res) {  // reset_logweights! after the final resampling (src/container.jl:228)
        memset(logw_out, 0, sizeof(double) * (size_t)h->ctx.N);
        return APS_OK;
    }
    CU(cudaMemcpyAsync(logw_out, h->ctx.logw, sizeof(double) * (size_t)h->ctx.N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return APS_OK;
}

extern "C" int aps_get_final_states(aps_handle *h, double *x_out) {
    NEED_SWEEP("aps_get_final_states");
    if (!x_out) return fail(APS_ERR_INVALID, "aps_get_final_states: null output");
    const DevCtx &c = h->ctx;
    k_gather_final<<<stride_grid(c.N), APS_THREADS, 0, h->stream>>>(c, h->d_scratch);
    CU(cudaMemcpyAsync(x_out, h->d_scratch, sizeof(double) * (size_t)c.N * c.d, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return APS_OK;
}

extern "C" int aps_get_trajectory(aps_handle *h, int64_t slot, double *traj_out) {
    NEED_SWEEP("aps_get_trajectory");
    if (!traj_out) return fail(APS_ERR_INVALID, "aps_get_trajectory: null output");
    if (!h->cfg.keep_history) return fail(APS_ERR_INVALID, "aps_get_trajectory: handle was created with keep_history = 0");
    const DevCtx &c = h->ctx;
    if (slot < 0 || slot >= c.N) return fail(APS_ERR_INVALID, "aps_get_trajectory: slot out of range");
    k_backtrace<<<1, 32, 0, h->stream>>>(c, c.slot0 + slot, h->d_traj);  // slot: index in this rank's shard
    CU(cudaMemcpyAsync(traj_out, h->d_traj, sizeof(double) * (size_t)c.T * c.d, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return APS_OK;
}

extern "C" int aps_get_trajectories(aps_handle *h, double *traj_out) {
    NEED_SWEEP("aps_get_trajectories");
    if (!traj_out) return fail(APS_ERR_INVALID, "aps_get_trajectories: null output");
    if (!h->cfg.keep_history) return fail(APS_ERR_INVALID, "aps_get_trajectories: handle was created with keep_history = 0");
    const DevCtx &c = h->ctx;
    int32_t *d_idx = nullptr;
    double *d_out[2] = {nullptr, nullptr};  // double-buffered: the copy of step t overlaps the gather of step t-1
    cudaEvent_t ev[2] = {nullptr, nullptr};
    CU(cudaMalloc(&d_idx, sizeof(int32_t) * (size_t)c.N));
    const size_t step_bytes = sizeof(double) * (size_t)c.N * c.d;
    int rc = APS_OK;
    for (int b = 0; b < 2 && rc == APS_OK; ++b) {
        if (cudaMalloc(&d_out[b], step_bytes) != cudaSuccess || cudaEventCreate(&ev[b]) != cudaSuccess)
            rc = fail(APS_ERR_NOMEM, "aps_get_trajectories: scratch allocation failed");
    }
    cudaStream_t cp = nullptr;
    if (rc == APS_OK && cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking) != cudaSuccess)
        rc = fail(APS_ERR_CUDA, "aps_get_trajectories: stream creation failed");
    if (rc == APS_OK) {
        cudaEvent_t done_k;
        cudaEventCreate(&done_k);
        for (long long t = c.T; t >= 1; --t) {
            const int b = (int)(t & 1);
            cudaStreamWaitEvent(h->stream, ev[b], 0);  // the copy that last used this buffer is complete
            k_traj_step<<<stride_grid(c.N), APS_K1_THREADS, 0, h->stream>>>(c, t, d_idx, d_out[b]);
            cudaEventRecord(done_k, h->stream);
            cudaStreamWaitEvent(cp, done_k, 0);
            cudaMemcpyAsync(traj_out + (size_t)(t - 1) * c.N * c.d, d_out[b], step_bytes, cudaMemcpyDeviceToHost, cp);
            cudaEventRecord(ev[b], cp);
        }
        cudaStreamSynchronize(h->stream);
        cudaStreamSynchronize(cp);
        cudaEventDestroy(done_k);
        if (cudaGetLastError() != cudaSuccess) rc = fail(APS_ERR_CUDA, "aps_get_trajectories: kernel or copy failed");
    }
    if (cp) cudaStreamDestroy(cp);
    for (int b = 0; b < 2; ++b) {
        if (ev[b]) cudaEventDestroy(ev[b]);
        cudaFree(d_out[b]);
    }
    cudaFree(d_idx);
    return rc;
}

extern "C" int aps_get_step_stats(aps_handle *h, double *logz_out, double *ess_out, uint8_t *resampled_out) {
    NEED_SWEEP("aps_get_step_stats");
    const long long T = h->ctx.T;
    std::vector<StepPlan> p((size_t)T + 1);
    CU(cudaMemcpyAsync(p.data(), h->ctx.plan, sizeof(StepPlan) * (size_t)(T + 1), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (long long s = 0; s <= T; ++s) {
        if (logz_out && s >= 1) logz_out[s - 1] = p[(size_t)s].logZ;
        if (ess_out) ess_out[s] = p[(size_t)s].ess;
        if (resampled_out) resampled_out[s] = (uint8_t)p[(size_t)s].resampled;
    }
    return APS_OK;
}

extern "C" int aps_get_states(aps_handle *h, int64_t t, double *x_out) {
    NEED_SWEEP("aps_get_states");
    const DevCtx &c = h->ctx;
    if (!x_out || t < 1 || t > c.T) return fail(APS_ERR_INVALID, "aps_get_states: t out of range");
    if (!h->cfg.keep_history && t < c.T - 1) return fail(APS_ERR_INVALID, "aps_get_states: history not kept");
    std::vector<double> soa((size_t)c.NS * c.d);
    CU(cudaMemcpyAsync(soa.data(), c.x + ((t - 1) % c.x_slabs) * (long long)c.d * c.NS, sizeof(double) * soa.size(),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (long long i = 0; i < c.N; ++i)
        for (int k = 0; k < c.d; ++k) x_out[i * c.d + k] = soa[(size_t)k * c.NS + i];
    return APS_OK;
}

extern "C" int aps_get_ancestors(aps_handle *h, int64_t t, int32_t *anc_out) {
    NEED_SWEEP("aps_get_ancestors");
    const DevCtx &c = h->ctx;
    if (!anc_out || t < 2 || t > c.T + 1) return fail(APS_ERR_INVALID, "aps_get_ancestors: t must be in 2..T+1");
    if (!h->cfg.keep_history && t < c.T) return fail(APS_ERR_INVALID, "aps_get_ancestors: history not kept");
    CU(cudaMemcpyAsync(anc_out, c.anc + ((t - 1) % c.anc_slabs) * c.NS, sizeof(int32_t) * (size_t)c.N,
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return APS_OK;
}

extern "C" int aps_smoothing_mean(aps_handle *h, double *mean_out) {
    NEED_SWEEP("aps_smoothing_mean");
    if (!mean_out) return fail(APS_ERR_INVALID, "aps_smoothing_mean: null output");
    if (!h->cfg.keep_history) return fail(APS_ERR_INVALID, "aps_smoothing_mean: handle was created with keep_history = 0");
    const DevCtx &c = h->ctx;
    int res = 0;
    int rc = final_resampled(h, &res);
    if (rc) return rc;
    // scratch: ancestor cursor per slot, block partials, per-step tickets, the T x d result
    int32_t *d_idx = nullptr;
    double *d_partial = nullptr, *d_mean = nullptr;
    unsigned *d_ctr = nullptr;
    CU(cudaMalloc(&d_idx, sizeof(int32_t) * (size_t)c.N));
    CU(cudaMalloc(&d_partial, sizeof(double) * APS_SMOOTH_BLOCKS * APS_MAX_D));
    CU(cudaMalloc(&d_mean, sizeof(double) * (size_t)c.T * c.d));
    CU(cudaMalloc(&d_ctr, sizeof(unsigned) * (size_t)c.T));
    CU(cudaMemsetAsync(d_ctr, 0, sizeof(unsigned) * (size_t)c.T, h->stream));
    long long g = (c.N + APS_K1_THREADS - 1) / APS_K1_THREADS;
    if (g > APS_SMOOTH_BLOCKS) g = APS_SMOOTH_BLOCKS;
    for (long long t = c.T; t >= 1; --t)
        k_smooth_step<<<(int)g, APS_K1_THREADS, 0, h->stream>>>(c, t, d_idx, c.q, res, d_partial, d_ctr, d_mean);
    CU(cudaMemcpyAsync(mean_out, d_mean, sizeof(double) * (size_t)c.T * c.d, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d_idx);
    cudaFree(d_partial);
    cudaFree(d_mean);
    cudaFree(d_ctr);
    CU(cudaGetLastError());
    return APS_OK;
}

extern "C" int aps_get_fat_counts(aps_handle *h, int32_t *counts_out) {
    NEED_SWEEP("aps_get_fat_counts");
    if (!counts_out) return fail(APS_ERR_INVALID, "aps_get_fat_counts: null output");
    CU(cudaMemcpyAsync(counts_out, h->ctx.fat_cnt, sizeof(int) * (size_t)(h->ctx.T + 1), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return APS_OK;
}

extern "C" int aps_last_sweep_ms(aps_handle *h, float *ms_out) {
    NEED_SWEEP("aps_last_sweep_ms");
    if (ms_out) *ms_out = h->last_ms;
    return APS_OK;
}

extern "C" int aps_last_sweep_launches(aps_handle *h, int64_t *n_out) {
    NEED_SWEEP("aps_last_sweep_launches");
    if (n_out) *n_out = h->last_launches;
    return APS_OK;
}

// ================================================================== operator level
// A grow-only device workspace shared by the operator-level entry points (serialised by a mutex).
struct OpWorkspace {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    long long cap_m = 0, cap_n = 0;
    CUtensorMap tmap_q;
    double *d_in = nullptr, *d_wout = nullptr;
    u64 *d_q = nullptr, *tile_sum = nullptr, *tile_s1 = nullptr, *tile_s2 = nullptr, *tile_prefix = nullptr;
    int32_t *d_idx32 = nullptr;
    long long *d_idx64 = nullptr;
    u64 *d_cum = nullptr, *d_rq = nullptr;
    unsigned short *d_cut = nullptr;
    unsigned char *d_cut_sh = nullptr;
    int *d_counts = nullptr, *d_tile_count = nullptr, *d_tile_cprefix = nullptr;
    ResidualState *d_rs = nullptr;
    unsigned *d_done2 = nullptr;
    StepAcc *acc = nullptr;
    StepPlan *plan = nullptr, *plan2 = nullptr;
    int *d_err = nullptr;
    SweepState *st = nullptr;
    SweepParams *sp = nullptr;
    int *d_fat_cnt = nullptr;     // fat-parent list of the one decision point an operator call has
    FatEntry *d_fat = nullptr;
};
static OpWorkspace g_ws;

static int ws_reserve(OpWorkspace &w, long long m, long long n) {
    if (!w.stream) {
        CU(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&w.acc, sizeof(StepAcc)));
        CU(cudaMalloc(&w.plan, sizeof(StepPlan)));
        CU(cudaMalloc(&w.plan2, sizeof(StepPlan)));
        CU(cudaMalloc(&w.d_err, sizeof(int)));
        CU(cudaMalloc(&w.st, sizeof(SweepState)));
        CU(cudaMalloc(&w.sp, sizeof(SweepParams)));
        CU(cudaMalloc(&w.d_fat_cnt, 16));
        CU(cudaMalloc(&w.d_fat, sizeof(FatEntry) * APS_FAT_MAX));
        CU(cudaMalloc(&w.d_rs, sizeof(ResidualState)));
        CU(cudaMalloc(&w.d_done2, sizeof(unsigned)));
    }
    if (m > w.cap_m) {
        cudaFree(w.d_in); cudaFree(w.d_wout); cudaFree(w.d_q);
        cudaFree(w.tile_sum); cudaFree(w.tile_s1); cudaFree(w.tile_s2); cudaFree(w.tile_prefix);
        cudaFree(w.d_cum); cudaFree(w.d_rq); cudaFree(w.d_cut); cudaFree(w.d_cut_sh); cudaFree(w.d_counts); cudaFree(w.d_tile_count); cudaFree(w.d_tile_cprefix);
        w.cap_m = 0;
        const long long nt = (m + APS_TILE - 1) / APS_TILE;
        CU(cudaMalloc(&w.d_in, sizeof(double) * (size_t)m));
        CU(cudaMalloc(&w.d_wout, sizeof(double) * (size_t)m));
        CU(cudaMalloc(&w.d_q, sizeof(u64) * (size_t)(m + 32)));
        CU(cudaMalloc(&w.tile_sum, sizeof(u64) * (size_t)nt));
        CU(cudaMalloc(&w.tile_s1, sizeof(u64) * (size_t)nt));
        CU(cudaMalloc(&w.tile_s2, sizeof(u64) * (size_t)nt));
        CU(cudaMalloc(&w.tile_prefix, sizeof(u64) * (size_t)nt));
        CU(cudaMalloc(&w.d_cum, sizeof(u64) * (size_t)(m + 32)));
        CU(cudaMalloc(&w.d_cut, sizeof(unsigned short) * (size_t)nt * APS_CUT));
        CU(cudaMalloc(&w.d_cut_sh, (size_t)nt));
        CU(cudaMalloc(&w.d_rq, sizeof(u64) * (size_t)(m + 32)));
        CU(cudaMalloc(&w.d_counts, sizeof(int) * (size_t)(m + 32)));
        CU(cudaMalloc(&w.d_tile_count, sizeof(int) * (size_t)nt));
        CU(cudaMalloc(&w.d_tile_cprefix, sizeof(int) * (size_t)nt));
        w.cap_m = m;
    }
    if (n > w.cap_n) {
        cudaFree(w.d_idx32); cudaFree(w.d_idx64);
        w.cap_n = 0;
        CU(cudaMalloc(&w.d_idx32, sizeof(int32_t) * (size_t)n));
        CU(cudaMalloc(&w.d_idx64, sizeof(long long) * (size_t)n));
        w.cap_n = n;
    }
    return APS_OK;
}

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static void op_ctx(OpWorkspace &w, DevCtx &c, long long m, long long n_draw) {
    memset(&c, 0, sizeof(c));
    c.q = w.d_q;
    c.tile_sum = w.tile_sum;
    c.tile_s1 = w.tile_s1;
    c.tile_s2 = w.tile_s2;
    c.tile_prefix = w.tile_prefix;
    c.acc = w.acc;
    c.plan = w.plan;
    c.st = nullptr;
    c.sp = w.sp;
    c.anc = w.d_idx32;
    c.N = m;
    c.Ng = m;
    c.world = 1;
    c.NS = (m + 31) & ~31LL;
    c.T = 0;
    c.x_slabs = 1;
    c.anc_slabs = 1;
    c.num_tiles = (m + APS_TILE - 1) / APS_TILE;
    c.d = 1;
    c.dy = 1;
    c.S = aps_weight_shift((uint64_t)m);
    c.Hs = aps_ess_shift((uint64_t)m);
    c.bare = 1;
    c.ess_threshold = NAN;
    c.logN = 0.0;
    c.n_override = n_draw;
    c.fat_cnt = w.d_fat_cnt;
    c.fat = w.d_fat;
    c.fat_steps = 1;
    c.fat_min = fat_min_for(n_draw > m ? n_draw : m);
}

// max + normalise of a weight / log-weight vector into the workspace; fills *plan_host
template <int INPUT>
static int op_normalise(OpWorkspace &w, DevCtx &c, const double *in, long long m, uint64_t key, uint64_t ctr,
                        StepPlan *plan_host) {
    const double *d_in = in;
    if (!is_device_ptr(in)) {
        CU(cudaMemcpyAsync(w.d_in, in, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, w.stream));
        d_in = w.d_in;
    }
    SweepParams sp;
    sp.key = key;
    sp.has_ref = 0;
    sp.pad = 0;
    CU(cudaMemcpyAsync(w.sp, &sp, sizeof(sp), cudaMemcpyHostToDevice, w.stream));
    CU(cudaMemsetAsync(w.acc, 0, sizeof(StepAcc), w.stream));
    CU(cudaMemsetAsync(w.d_fat_cnt, 0, 16, w.stream));
    if (c.NS > m) CU(cudaMemsetAsync(w.d_q + m, 0, sizeof(u64) * (size_t)(c.NS - m), w.stream));  // zero-weight padding
    k_vector_max<INPUT><<<stride_grid(m), APS_K1_THREADS, 0, w.stream>>>(d_in, m, w.acc);
    c.ctr_offset = (long long)ctr;
    k_normalise<INPUT><<<(int)c.num_tiles, APS_K2_THREADS, 0, w.stream>>>(c, d_in, 0);
    CU(cudaMemcpyAsync(plan_host, w.plan, sizeof(StepPlan), cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    CU(cudaGetLastError());
    return APS_OK;
}

extern "C" int aps_resample(int kind, const double *wts, int64_t m, int64_t n, uint64_t key, uint64_t ctr,
                            int64_t *idx_out) {
    if (!wts || !idx_out) return fail(APS_ERR_INVALID, "aps_resample: null argument");
    if (m <= 0) return fail(APS_ERR_INVALID, "weight vector is empty");  // src/resampling.jl:103,154
    if (m > 2147483647LL || n < 0 || n > 2147483647LL) return fail(APS_ERR_INVALID, "aps_resample: size out of range");
    if (kind < APS_RESAMPLE_MULTINOMIAL || kind > APS_RESAMPLE_SYSTEMATIC)
        return fail(APS_ERR_INVALID, "aps_resample: unknown resampler kind");
    if (ctr >= (1ull << 40)) return fail(APS_ERR_INVALID, "aps_resample: ctr must be < 2^40");
    if (n == 0) return APS_OK;
    OpWorkspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    int rc = ws_reserve(w, m, n);
    if (rc) return rc;
    DevCtx c;
    op_ctx(w, c, m, n);
    StepPlan p;
    rc = op_normalise<IN_W>(w, c, wts, m, key, ctr, &p);
    if (rc) return rc;
    if (p.err) return fail(APS_ERR_WEIGHTS, "sample could not be selected (are the weights normalized?)");
    if (kind == APS_RESAMPLE_SYSTEMATIC || kind == APS_RESAMPLE_STRATIFIED) {
        rc = make_q_tensormap(&w.tmap_q, w.d_q, c.NS);
        if (rc) return rc;
        rc = enable_k3_smem();
        if (rc) return rc;
        pick_resample(kind)<<<(int)c.num_tiles, APS_K3_THREADS, APS_K3_DYN_SMEM, w.stream>>>(c, 0, w.d_idx32, w.tmap_q);
        k_fill_fat<<<sm_count() * 2, APS_K1_THREADS, 0, w.stream>>>(c, 0, w.d_idx32, 0);
    } else {
        const int gt = (int)c.num_tiles;
        MultiArgs a;
        memset(&a, 0, sizeof(a));
        a.cum = w.d_cum;
        a.cut = w.d_cut;
        a.cut_sh = w.d_cut_sh;
        a.counts = w.d_counts;
        a.tile_count = w.d_tile_count;
        a.tile_cprefix = w.d_tile_cprefix;
        a.tile_prefix = w.tile_prefix;
        a.plan = w.plan;
        a.N = m;
        a.num_tiles = c.num_tiles;
        a.step = (long long)ctr;
        a.out32 = w.d_idx32;
        if (kind == APS_RESAMPLE_MULTINOMIAL) {
            a.qsrc = w.d_q;
            a.wplan = w.plan;
            a.n_draws = &w.plan->n;
        } else {
            // deterministic copies first (sorted), then the residual draws in draw order (src/resampling.jl:62-78)
            CU(cudaMemsetAsync(w.d_err, 0, sizeof(int), w.stream));
            CU(cudaMemsetAsync(w.d_rs, 0, sizeof(ResidualState), w.stream));
            CU(cudaMemsetAsync(w.d_done2, 0, sizeof(unsigned), w.stream));
            a.qsrc = w.d_rq;
            a.n_draws = &w.d_rs->n_rest;
            a.out_offset = &w.d_rs->n_det;
            k_residual_split<<<gt, APS_THREADS, 0, w.stream>>>(a, w.d_q, w.d_rq, w.d_rs);
            k_tile_counts<<<gt, APS_THREADS, 0, w.stream>>>(a);
            k_scan_tile_counts<<<1, APS_THREADS, 0, w.stream>>>(a, nullptr, c, 0);
            k_expand_counts<<<gt, APS_THREADS, 0, w.stream>>>(a, w.d_idx32, 0, c, 0);
            k_fill_fat<<<sm_count() * 2, APS_K1_THREADS, 0, w.stream>>>(c, 0, w.d_idx32, 0);
            a.wplan = w.plan2;
            k_residual_weights<<<gt, APS_THREADS, 0, w.stream>>>(a, w.d_rq, w.d_rs, w.tile_sum, w.tile_prefix, w.plan2,
                                                                w.d_done2, w.d_err);
        }
        k_cumsum<<<gt, APS_THREADS, 0, w.stream>>>(a);
        k_multi_search<0><<<stride_grid((n + 1) / 2), APS_K1_THREADS, 0, w.stream>>>(a, &w.sp->key);
    }
    if (kind == APS_RESAMPLE_RESIDUAL) {
        int herr = 0;
        CU(cudaMemcpyAsync(&herr, w.d_err, sizeof(int), cudaMemcpyDeviceToHost, w.stream));
        CU(cudaStreamSynchronize(w.stream));
        if (herr) return fail(APS_ERR_WEIGHTS, "sample could not be selected (residual weights vanish)");
    }
    const bool dev_out = is_device_ptr(idx_out);
    long long *d_out = dev_out ? (long long *)idx_out : w.d_idx64;
    k_to_one_based<<<stride_grid(n), APS_THREADS, 0, w.stream>>>(w.d_idx32, n, d_out);
    if (!dev_out) CU(cudaMemcpyAsync(idx_out, w.d_idx64, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    CU(cudaGetLastError());
    return APS_OK;
}

static int op_logw(const double *logw, int64_t n, StepPlan *p, OpWorkspace &w, DevCtx &c) {
    if (!logw) return fail(APS_ERR_INVALID, "null log-weight vector");
    if (n <= 0 || n > 2147483647LL) return fail(APS_ERR_INVALID, "log-weight vector is empty or too long");
    int rc = ws_reserve(w, n, 1);
    if (rc) return rc;
    op_ctx(w, c, n, 0);
    return op_normalise<IN_LOGW>(w, c, logw, n, 0, 0, p);
}

extern "C" int aps_logsumexp(const double *logw, int64_t n, double *out) {
    OpWorkspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    DevCtx c;
    StepPlan p;
    int rc = op_logw(logw, n, &p, w, c);
    if (rc) return rc;
    if (p.err) {
        if (p.M == -INFINITY) {  // logsumexp of all -Inf is -Inf, not an error
            *out = -INFINITY;
            return APS_OK;
        }
        return fail(APS_ERR_WEIGHTS, "aps_logsumexp: NaN or +Inf log-weight");
    }
    *out = p.logZ;
    return APS_OK;
}

extern "C" int aps_ess(const double *logw, int64_t n, double *out) {
    OpWorkspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    DevCtx c;
    StepPlan p;
    int rc = op_logw(logw, n, &p, w, c);
    if (rc) return rc;
    if (p.err) return fail(APS_ERR_WEIGHTS, "aps_ess: weights not normalisable");
    *out = p.ess;
    return APS_OK;
}

extern "C" int aps_softmax(const double *logw, int64_t n, double *w_out) {
    if (!w_out) return fail(APS_ERR_INVALID, "aps_softmax: null output");
    OpWorkspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    DevCtx c;
    StepPlan p;
    int rc = op_logw(logw, n, &p, w, c);
    if (rc) return rc;
    if (p.err) return fail(APS_ERR_WEIGHTS, "aps_softmax: weights not normalisable");
    const bool dev_out = is_device_ptr(w_out);
    double *d_out = dev_out ? w_out : w.d_wout;
    k_weights_out<<<stride_grid(n), APS_THREADS, 0, w.stream>>>(w.d_q, w.plan, n, n, c.S, 0, d_out);
    if (!dev_out) CU(cudaMemcpyAsync(w_out, w.d_wout, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    CU(cudaGetLastError());
    return APS_OK;
}

extern "C" int aps_randcat(const double *wts, int64_t n, uint64_t key, uint64_t ctr, int64_t *idx_out) {
    if (!wts || !idx_out) return fail(APS_ERR_INVALID, "aps_randcat: null argument");
    if (n <= 0 || n > 2147483647LL) return fail(APS_ERR_INVALID, "weight vector is empty");
    if (ctr >= (1ull << 40)) return fail(APS_ERR_INVALID, "aps_randcat: ctr must be < 2^40");
    OpWorkspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    int rc = ws_reserve(w, n, 1);
    if (rc) return rc;
    DevCtx c;
    op_ctx(w, c, n, 1);
    StepPlan p;
    rc = op_normalise<IN_W>(w, c, wts, n, key, ctr, &p);
    if (rc) return rc;
    if (p.err) return fail(APS_ERR_WEIGHTS, "aps_randcat: weights not normalisable");
    // force the non-uniform branch of k_pick: it reads plan[0].resampled
    StepPlan p2 = p;
    p2.resampled = 0;
    CU(cudaMemcpyAsync(w.plan, &p2, sizeof(p2), cudaMemcpyHostToDevice, w.stream));
    DevCtx cc = c;
    cc.st = w.st;
    SweepState st0;
    memset(&st0, 0, sizeof(st0));
    st0.picked_slot = -1;
    CU(cudaMemcpyAsync(w.st, &st0, sizeof(st0), cudaMemcpyHostToDevice, w.stream));
    k_pick<<<1, APS_THREADS, 0, w.stream>>>(cc, 0, (long long)ctr, APS_DOM_RESAMPLE, 0);
    CU(cudaMemcpyAsync(&st0, w.st, sizeof(st0), cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    CU(cudaGetLastError());
    if (st0.picked_slot < 0) return fail(APS_ERR_WEIGHTS, "aps_randcat: no index could be selected");
    *idx_out = st0.picked_slot + 1;
    return APS_OK;
}

// ================================================================== micro-benchmark of the resample kernel
extern "C" int aps_bench_resample(int kind, int64_t n, int iters, int flush_l2, uint64_t seed, float *avg_ms_out,
                                  float *min_ms_out) {
    if (n <= 0 || n > 2147483647LL || iters < 1) return fail(APS_ERR_INVALID, "aps_bench_resample: bad size");
    if (kind != APS_RESAMPLE_SYSTEMATIC && kind != APS_RESAMPLE_STRATIFIED)
        return fail(APS_ERR_INVALID, "aps_bench_resample: kind not built yet");
    OpWorkspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    int rc = ws_reserve(w, n, n);
    if (rc) return rc;
    DevCtx c;
    op_ctx(w, c, n, n);
    SweepParams sp;
    sp.key = seed;
    sp.has_ref = 0;
    sp.pad = 0;
    CU(cudaMemcpyAsync(w.sp, &sp, sizeof(sp), cudaMemcpyHostToDevice, w.stream));
    CU(cudaMemsetAsync(w.acc, 0, sizeof(StepAcc), w.stream));
    CU(cudaMemsetAsync(w.d_fat_cnt, 0, 16, w.stream));
    if (c.NS > n) CU(cudaMemsetAsync(w.d_q + n, 0, sizeof(u64) * (size_t)(c.NS - n), w.stream));
    k_bench_weights<<<stride_grid(n), APS_THREADS, 0, w.stream>>>(w.d_q, n, c.S, seed);
    k_normalise<IN_Q><<<(int)c.num_tiles, APS_K2_THREADS, 0, w.stream>>>(c, nullptr, 0);
    StepPlan p;
    CU(cudaMemcpyAsync(&p, w.plan, sizeof(p), cudaMemcpyDeviceToHost, w.stream));
    CU(cudaStreamSynchronize(w.stream));
    CU(cudaGetLastError());
    if (p.err) return fail(APS_ERR_WEIGHTS, "aps_bench_resample: synthetic weights not normalisable");
    // L2 flush between launches: write 512 MB (evicts everything), then stream-read another 256 MB
    // so that L2 holds clean lines only -- otherwise the timed kernel pays the write-back of up to
    // 126 MB of the memset's dirty lines (measured: 95.9 us vs 86 us under ncu's own cache control)
    void *flush = nullptr, *flush2 = nullptr;
    const size_t flush_bytes = 512ull << 20, flush2_bytes = 256ull << 20;
    if (flush_l2) {
        CU(cudaMalloc(&flush, flush_bytes));
        CU(cudaMalloc(&flush2, flush2_bytes));
        CU(cudaMemsetAsync(flush2, 1, flush2_bytes, w.stream));
    }
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    res_fn f = pick_resample(kind);
    rc = make_q_tensormap(&w.tmap_q, w.d_q, c.NS);
    if (rc) return rc;
    rc = enable_k3_smem();
    if (rc) return rc;
    float tot = 0.f, mn = 1e30f;
    for (int it = -3; it < iters; ++it) {  // 3 warm-up launches
        if (flush_l2) {
            CU(cudaMemsetAsync(flush, it & 0xff, flush_bytes, w.stream));
            if (flush_l2 > 1)
                k_read_flush<<<sm_count() * 8, APS_K1_THREADS, 0, w.stream>>>((const uint4 *)flush2, (long long)(flush2_bytes / 16),
                                                                             (unsigned *)flush);
        }
        CU(cudaEventRecord(e0, w.stream));
        f<<<(int)c.num_tiles, APS_K3_THREADS, APS_K3_DYN_SMEM, w.stream>>>(c, 0, w.d_idx32, w.tmap_q);
        CU(cudaEventRecord(e1, w.stream));
        CU(cudaStreamSynchronize(w.stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 0) {
            tot += ms;
            mn = ms < mn ? ms : mn;
        }
    }
    CU(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (flush) cudaFree(flush);
    if (flush2) cudaFree(flush2);
    if (avg_ms_out) *avg_ms_out = tot / iters;
    if (min_ms_out) *min_ms_out = mn;
    return APS_OK;
}

// ================================================================== multi-GPU plumbing
// Blob exchanged between ranks: CUDA IPC handles of the state store, the ancestor store and the
// mailbox (plus raw pointers, used when both "ranks" live in one process, e.g. in tests).
struct IpcBlob {
    unsigned long long magic;
    long long pid;
    int device, rank;
    void *raw[3];
    cudaIpcMemHandle_t mem[3];
};
static_assert(sizeof(IpcBlob) <= APS_IPC_BLOB_BYTES, "IpcBlob must fit the ABI blob");

extern "C" int aps_ipc_export(aps_handle *h, uint8_t *blob_out) {
    if (!h || !blob_out) return fail(APS_ERR_INVALID, "aps_ipc_export: null argument");
    if (h->ctx.world < 2) return fail(APS_ERR_INVALID, "aps_ipc_export: handle is not sharded (world_size == 1)");
    CU(cudaSetDevice(h->cfg.device));
    IpcBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = 0x4150534950433031ULL;
    b.pid = (long long)getpid();
    b.device = h->cfg.device;
    b.rank = h->ctx.rank;
    b.raw[0] = h->ctx.x;
    b.raw[1] = h->ctx.anc;
    b.raw[2] = h->d_mail;
    for (int k = 0; k < 3; ++k) CU(cudaIpcGetMemHandle(&b.mem[k], b.raw[k]));
    memset(blob_out, 0, APS_IPC_BLOB_BYTES);
    memcpy(blob_out, &b, sizeof(b));
    return APS_OK;
}

extern "C" int aps_ipc_import(aps_handle *h, const uint8_t *blobs) {
    if (!h || !blobs) return fail(APS_ERR_INVALID, "aps_ipc_import: null argument");
    DevCtx &c = h->ctx;
    if (c.world < 2) return fail(APS_ERR_INVALID, "aps_ipc_import: handle is not sharded (world_size == 1)");
    CU(cudaSetDevice(h->cfg.device));
    PeerTable pt;
    memset(&pt, 0, sizeof(pt));
    for (int r = 0; r < c.world; ++r) {
        IpcBlob b;
        memcpy(&b, blobs + (size_t)r * APS_IPC_BLOB_BYTES, sizeof(b));
        if (b.magic != 0x4150534950433031ULL || b.rank != r)
            return fail(APS_ERR_COMM, "aps_ipc_import: blob " + std::to_string(r) + " is not an aps_ipc_export blob of that rank");
        void *p[3];
        if (r == c.rank) {
            p[0] = c.x; p[1] = c.anc; p[2] = h->d_mail;
        } else if (b.pid == (long long)getpid()) {
            // both ranks in one process: use the pointers directly (enable peer access across devices)
            if (b.device != h->cfg.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(APS_ERR_COMM, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            for (int k = 0; k < 3; ++k) p[k] = b.raw[k];
        } else {
            for (int k = 0; k < 3; ++k) {
                cudaError_t e = cudaIpcOpenMemHandle(&p[k], b.mem[k], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) return fail(APS_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
                h->ipc_opened[h->n_ipc_opened++] = p[k];
            }
        }
        pt.x[r] = (double *)p[0];
        pt.anc[r] = (int32_t *)p[1];
        pt.mail[r] = (MailSlot *)p[2];
    }
    CU(cudaMemcpy(h->d_peers, &pt, sizeof(pt), cudaMemcpyHostToDevice));
    c.peers = h->d_peers;
    if (h->graph_ready) {  // the captured graph holds the old context by value
        cudaGraphExecDestroy(h->graph);
        h->graph_ready = false;
    }
    h->comm_ready = true;
    return APS_OK;
}
