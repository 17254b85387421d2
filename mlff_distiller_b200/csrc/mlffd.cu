// libmlffd.so -- C-ABI front end (include/mlffd.h) and step orchestration.
//
// One context = one device: weights re-laid out for the kernels, and a workspace sized once by
// mlffd_workspace_reserve so the hot path never allocates.  A step is a fixed sequence of kernel
// launches on the caller's stream with no host synchronisation; sizes discovered on the device
// (number of edges / pairs) stay on the device (DeviceStatus) and are read by later kernels.
#include "../../include/mlffd.h"

#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cell_list.cuh"
#include "common.cuh"
#include "edge_features.cuh"
#include "filter.cuh"
#include "filter_umma.cuh"
#include "umma_rows.cuh"
#include "md.cuh"
#include "message.cuh"
#include "message_pipe.cuh"
#include "message_team.cuh"
#include "message_spline.cuh"
#include "neighbor.cuh"
#include "readout.cuh"
#include "skin_list.cuh"
#include "update.cuh"

using namespace mlffd;

namespace {

thread_local std::string g_create_error;

struct LayerWeights {
    FilterWeights filter;
    UpdateWeights update;
};

struct Workspace {
    int64_t cap_atoms = 0, cap_edges = 0, cap_structs = 0, cap_pairs = 0;
    int64_t req_edges = 0;   // edge capacity the caller asked for (cap_edges is larger when a skin list needs room)
    void* arena = nullptr;
    size_t arena_bytes = 0;
    // graph
    int *atom_struct = nullptr, *deg = nullptr, *deg_low = nullptr, *rowptr = nullptr, *lowptr = nullptr;
    int *col = nullptr, *edge_dst = nullptr, *rev = nullptr, *pair = nullptr;
    float4 *geo = nullptr, *edge_adj = nullptr;
    float4* erec = nullptr;   // [E][4] per-edge records of the spline message kernels (message_spline.cuh)
    // Verlet-skin candidate list (skin_list.cuh), only when mlffd_set_skin > 0
    SkinState* skin = nullptr;
    int *cand_deg = nullptr, *cand_rowptr = nullptr, *cand_col = nullptr;
    float* pos_ref = nullptr;
    double* virial64 = nullptr;      // [cap_structs][9] FP64 accumulators of mlffd_virial
    float* pair_dist = nullptr;
    // cell list (large structures)
    GridInfo* grids = nullptr;
    int *ncells = nullptr, *total_cells = nullptr, *atom_cell = nullptr, *cell_count = nullptr,
        *cell_start = nullptr, *cell_cursor = nullptr, *cell_atoms = nullptr;
    int64_t cap_cells = 0;
    // per layer
    float* filt[kMaxLayers] = {};
    float* dfilt[kMaxLayers] = {};
    float* s_in[kMaxLayers + 1] = {};   // s_in[L] = final scalar features
    float* v_in[kMaxLayers] = {};       // v_in[0] unused (zero)
    float* s_msg[kMaxLayers] = {};
    float* v_msg[kMaxLayers] = {};
    float* y1[kMaxLayers] = {};
    float* gates[kMaxLayers] = {};
    float* eps = nullptr;
    float* sbar[kMaxLayers] = {};       // adjoint sets (ping-pong unless debug)
    float* vbar[kMaxLayers] = {};
    void* cub_temp = nullptr;
    size_t cub_bytes = 0;
};

}  // namespace

struct mlffd_ctx {
    int device = 0;
    mlffd_config cfg{};
    int H = 0, K = 0, L = 0;
    bool debug_keep = false;
    int neighbor_mode = 0;   // 0 auto, 1 sweep, 2 cells (env MLFFD_NEIGHBOR)
    float skin = 0.f;                // mlffd_set_skin: Verlet-skin width in Angstrom (0 = exact rebuild every step)
    int64_t skin_atoms = -1;         // system the candidate list belongs to (atoms, structures); a change invalidates it
    int skin_structs = -1;
    bool dense_fallback = false;     // mlffd_set_dense_fallback: dense layers on the FP32 FFMA kernels whatever cfg.precision
    int readout_mode = 0;            // env MLFFD_READOUT: 0 = by size, 1 = tile, 2 = warp
    bool readout_configured = false; // smem attribute of readout_tile_kernel set on this device
    uint32_t pipe_configured = 0;    // bit per pipelined-kernel instantiation whose smem attribute is set on this device
    int last_adj_slabs = 0;          // edge-adjoint slabs written by the last force evaluation (0 = none)
    int last_adj_pairs = 0, last_adj_direct = 0;   // slab kinds (readout.cuh): [0, pairs) pair sums, [pairs, direct) adjoint of e, the rest that of rev(e)
    int msg_bwd_mode = 2;            // env MLFFD_MSG_BWD = edges (0) | pairs (1) | pipe (2)
    int msg_fwd_mode = 1;            // env MLFFD_MSG_FWD = rows (0) | pipe (1)
    int pipe_depth_fwd = 2;          // env MLFFD_PIPE_DEPTH_FWD: ring slots per warp
    int pipe_depth_bwd = 2;          // env MLFFD_PIPE_DEPTH_BWD
    std::string err;
    float* weights_d = nullptr;
    uint8_t* w2_images_d = nullptr;   // swizzled fp16 hi/lo 64 KB weight images (tensor-core path)
    bool spline = true;               // cfg.filter_mode == MLFFD_FILTER_SPLINE: no per-step filter tables
    float* spline_d = nullptr;        // [L][H/32][kSplineRows][3][32] filter spline coefficients (always built)
    float spline_inv_h = 0.f;         // intervals per Angstrom
    int adj_slabs_per_layer = 1;      // edge-adjoint slabs one layer's reverse kernel writes (H/32 in spline mode)
    int spline_fwd_shape = 0, spline_bwd_shape = 0;   // env MLFFD_SPLINE_FWD / _BWD: launch shape of the spline message kernels
    int skinny8_rows = 1024;        // env MLFFD_SKINNY8_ROWS: atoms at or below which ffma_rows_kernel uses 8-row tiles
    int msg_team = 8;               // env MLFFD_MSG_TEAM (4 | 8; 0 = row-per-warp kernels): warps sharing a CSR row in the small-system message kernels (0 / 1 = off)
    int filter_batch = 2;           // env MLFFD_FILTER_BATCH: 1 = all layers' filter tables in one launch, 0 = one launch per layer, 2 = one launch only for small systems
    int small_rows = 2048;          // env MLFFD_SMALL_ROWS: at or below this many atoms the update block runs on ffma_rows_kernel
    int tc_mode = 0;                // kTcSplit | kTcF16 | kTcBF16 (filter_umma.cuh), from cfg.precision
    bool use_umma = false;          // tensor-core update block (H = 128)
    bool use_umma_filter = false;   // tensor-core filter table (H = 128, 64, 32)
    struct ImageOffsets { size_t filter1, filter, upd_f1, upd_f2, upd_b1, upd_b2; } img[kMaxLayers] = {};
    const float *emb = nullptr, *centers = nullptr, *gammas = nullptr;
    LayerWeights layer[kMaxLayers];
    HeadWeights head{};
    Workspace ws;
    DeviceStatus* status_d = nullptr;
    int64_t last_atoms = 0;
    int last_structs = 0;
    bool last_had_forces = false;
    cudaStream_t last_stream = nullptr;
    // optional per-stage timing: an event after every launch group (mlffd_profile_enable)
    bool profiling = false;
    int64_t launches = 0;
    int64_t stage_launches[MLFFD_NUM_STAGES] = {};
    double stage_ms[MLFFD_NUM_STAGES] = {};
    std::vector<cudaEvent_t> event_pool;
    std::vector<std::pair<cudaEvent_t, int>> marks;  // (event, stage it closes; -1 = step start)
    size_t events_used = 0;
};

namespace {

// Entry points run on the context's device and leave the calling thread's current device as they found
// it (one process may drive several GPUs, and the caller's framework relies on its own current device).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int fail(mlffd_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

// Count launches; in profiling mode also drop an event that closes `stage`.
void mark(mlffd_ctx* ctx, int stage, cudaStream_t st, int launches = 1) {
    ctx->launches += launches;
    if (stage >= 0) ctx->stage_launches[stage] += launches;
    if (!ctx->profiling) return;
    if (ctx->events_used == ctx->event_pool.size()) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        ctx->event_pool.push_back(ev);
    }
    cudaEvent_t ev = ctx->event_pool[ctx->events_used++];
    cudaEventRecord(ev, st);
    ctx->marks.emplace_back(ev, stage);
}

int drain_marks(mlffd_ctx* ctx) {
    for (size_t i = 1; i < ctx->marks.size(); ++i) {
        const int stage = ctx->marks[i].second;
        if (stage < 0) continue;  // a step-start mark: gap between steps is not attributed
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->marks[i - 1].first, ctx->marks[i].first) == cudaSuccess)
            ctx->stage_ms[stage] += ms;
    }
    ctx->marks.clear();
    ctx->events_used = 0;
    return 0;
}

#define CUDA_TRY(ctx, expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(ctx, MLFFD_ECUDA,                                                     \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
    } while (0)

#define LAUNCH_CHECK(ctx, what)                                                               \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return fail(ctx, MLFFD_ECUDA, std::string(what) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// launch check + launch accounting for `stage` on stream `st`
#define LAUNCHED(ctx, what, stage, st)                                                        \
    do {                                                                                      \
        LAUNCH_CHECK(ctx, what);                                                              \
        mark(ctx, stage, st);                                                                 \
    } while (0)

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int clamp_grid(int64_t wanted, int max_blocks) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(wanted, max_blocks));
}

// ---- host-side weight re-layout -------------------------------------------------------------
struct Stager {
    std::vector<float> buf;
    size_t push(const float* src, size_t n) {
        const size_t off = (buf.size() + 63) / 64 * 64;  // 256-byte alignment
        buf.resize(off + n);
        std::memcpy(buf.data() + off, src, n * sizeof(float));
        return off;
    }
    // src is [rows][cols] row-major; stores its transpose [cols][rows]
    size_t push_transposed(const float* src, int rows, int cols) {
        std::vector<float> t((size_t)rows * cols);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = src[(size_t)r * cols + c];
        return push(t.data(), t.size());
    }
};

template <int H>
int set_kernel_attributes(mlffd_ctx* ctx) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(filter_table_kernel<H>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)filter_smem_bytes<H>()));
    CUDA_TRY(ctx, cudaFuncSetAttribute(update_forward_kernel<H, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)update_fwd_smem_bytes<H>()));
    CUDA_TRY(ctx, cudaFuncSetAttribute(update_forward_kernel<H, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)update_fwd_smem_bytes<H>()));
    CUDA_TRY(ctx, cudaFuncSetAttribute(update_backward_kernel<H, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)update_bwd_smem_bytes<H>()));
    CUDA_TRY(ctx, cudaFuncSetAttribute(update_backward_kernel<H, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)update_bwd_smem_bytes<H>()));
    return MLFFD_OK;
}

// Run a statement with the tensor-core arithmetic mode of the context as the constant MODE.
#define TC_DISPATCH(mode, ...)                                                             \
    do {                                                                                   \
        if ((mode) == kTcF16) { constexpr int MODE = kTcF16; __VA_ARGS__; }                \
        else if ((mode) == kTcBF16) { constexpr int MODE = kTcBF16; __VA_ARGS__; }         \
        else { constexpr int MODE = kTcSplit; __VA_ARGS__; }                               \
    } while (0)

// ---- per-H launch sequences -----------------------------------------------------------------
// Tensor-core filter tables of layers [l0, l1) in one launch (filter_umma.cuh).  filt / dfilt
// override the workspace tables for the single-layer stage entry point.
template <int H>
int launch_filter_umma(mlffd_ctx* ctx, int l0, int l1, const float* dist, const int* num_pairs_ptr,
                       int num_pairs_arg, const DeviceStatus* status, float* filt, float* dfilt,
                       int64_t pair_bound, cudaStream_t st) {
    const int nl = l1 - l0;
    const int tiles = (int)ceil_div(std::max<int64_t>(pair_bound, 1), kUmmaPairs);
    FilterBatchArgs batch{};
    batch.num_layers = nl;
    // persistent CTAs per layer in proportion to the layer's 128-channel chunks (layer 0 skips the b gate)
    int weight[kFilterMaxLayers], total_w = 0;
    for (int i = 0; i < nl; ++i) {
        const bool skip = (l0 + i == 0) && status != nullptr;
        // measured on C2: a layer-0 tile (two of three chunks) costs ~0.75 of a full tile -- the RBF,
        // first-layer and publish phases do not shrink with the chunk count
        weight[i] = (skip && H == 128) ? 3 : 4;
        total_w += weight[i];
    }
    const int budget = std::max(kNumSMs, nl);
    int begin = 0;
    for (int i = 0; i < nl; ++i) {
        const int l = l0 + i;
        int share = (i == nl - 1) ? budget - begin : std::max(1, (budget * weight[i] + total_w / 2) / total_w);
        share = std::max(1, std::min(share, tiles));
        FilterLayerArgs& a = batch.layer[i];
        a.w = ctx->layer[l].filter;
        a.w1_image = ctx->w2_images_d + ctx->img[l].filter1;
        a.w2_images = ctx->w2_images_d + ctx->img[l].filter;
        a.filt = filt ? filt : ctx->ws.filt[l];
        a.dfilt = dfilt ? dfilt : ctx->ws.dfilt[l];
        a.skip_vector_gate = (l == 0 && status != nullptr) ? 1 : 0;
        a.cta_begin = begin;
        begin += share;
    }
    TC_DISPATCH(ctx->tc_mode, filter_table_umma_kernel<H, MODE><<<begin, kFilterUmmaThreads, UmmaGeom<H>::total(), st>>>(
        dist, num_pairs_ptr, num_pairs_arg, status, ctx->centers, ctx->gammas, ctx->K, ctx->cfg.cutoff, batch));
    LAUNCHED(ctx, "filter_table_umma_kernel", MLFFD_STAGE_FILTER, st);
    return MLFFD_OK;
}

template <int H>
int launch_filter(mlffd_ctx* ctx, int l, const float* dist, const int* num_pairs_ptr,
                  int num_pairs_arg, const DeviceStatus* status, float* filt, float* dfilt,
                  int64_t pair_bound, cudaStream_t st) {
    if (ctx->use_umma_filter && !ctx->dense_fallback)
        return launch_filter_umma<H>(ctx, l, l + 1, dist, num_pairs_ptr, num_pairs_arg, status, filt, dfilt, pair_bound, st);
    const int blocks_per_sm = (filter_smem_bytes<H>() <= 110 * 1024) ? 2 : 1;
    const int grid = clamp_grid(ceil_div(std::max<int64_t>(pair_bound, 1), kFilterPairs),
                                kNumSMs * blocks_per_sm);
    filter_table_kernel<H><<<grid, kGemmThreads, filter_smem_bytes<H>(), st>>>(
        dist, num_pairs_ptr, num_pairs_arg, status, ctx->centers, ctx->gammas, ctx->K,
        ctx->cfg.cutoff, ctx->layer[l].filter, (l == 0 && status != nullptr) ? 1 : 0, filt, dfilt);
    LAUNCHED(ctx, "filter_table_kernel", MLFFD_STAGE_FILTER, st);
    return MLFFD_OK;
}

template <bool LAYER0, int D>
void launch_forward_pipe_d(mlffd_ctx* ctx, int l, int grid, int N, cudaStream_t st) {
    Workspace& ws = ctx->ws;
    auto kernel = message_forward_pipe_kernel<LAYER0, D>;
    constexpr size_t smem = message_forward_pipe_smem<D>();
    // per context (device), not per process: a second context on another GPU needs the attribute too
    constexpr uint32_t bit = 1u << (2 * D + (LAYER0 ? 1 : 0) + 0);
    if (!(ctx->pipe_configured & bit)) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ctx->pipe_configured |= bit;
    }
    kernel<<<grid, 32 * kPipeWarps, smem, st>>>(ws.rowptr, ws.col, ws.pair, ws.geo, ws.filt[l], ws.s_in[l],
                                              LAYER0 ? nullptr : ws.v_in[l], ws.s_msg[l], ws.v_msg[l], N,
                                              ctx->status_d);
}
void launch_forward_pipe(mlffd_ctx* ctx, int l, int grid, int N, cudaStream_t st) {
#define FWD_PIPE(D) (l == 0 ? launch_forward_pipe_d<true, D>(ctx, l, grid, N, st) : launch_forward_pipe_d<false, D>(ctx, l, grid, N, st))
    switch (ctx->pipe_depth_fwd) {
        case 3: FWD_PIPE(3); break;
        case 4: FWD_PIPE(4); break;
        case 8: FWD_PIPE(8); break;
        default: FWD_PIPE(2); break;
    }
#undef FWD_PIPE
}

template <bool LAYER0, int D>
void launch_backward_pipe_d(mlffd_ctx* ctx, int l, int grid, const float* sb, const float* vb, float* sb_in,
                            float* vb_in, int N, cudaStream_t st) {
    Workspace& ws = ctx->ws;
    auto kernel = message_backward_pipe_kernel<LAYER0, D>;
    constexpr size_t smem = message_backward_pipe_smem<D>();
    // per context (device), not per process: a second context on another GPU needs the attribute too
    constexpr uint32_t bit = 1u << (2 * D + (LAYER0 ? 1 : 0) + 14);
    if (!(ctx->pipe_configured & bit)) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ctx->pipe_configured |= bit;
    }
    kernel<<<grid, 32 * kPipeWarps, smem, st>>>(ws.rowptr, ws.col, ws.pair, ws.rev, ws.geo, ws.filt[l], ws.dfilt[l],
                                              ws.s_in[l], LAYER0 ? nullptr : ws.v_in[l], sb, vb, sb_in, vb_in,
                                              ws.edge_adj + (size_t)l * ws.cap_edges, N, ctx->status_d);
}
void launch_backward_pipe(mlffd_ctx* ctx, int l, int grid, const float* sb, const float* vb, float* sb_in,
                          float* vb_in, int N, cudaStream_t st) {
#define BWD_PIPE(D) (l == 0 ? launch_backward_pipe_d<true, D>(ctx, l, grid, sb, vb, sb_in, vb_in, N, st) \
                            : launch_backward_pipe_d<false, D>(ctx, l, grid, sb, vb, sb_in, vb_in, N, st))
    switch (ctx->pipe_depth_bwd) {
        case 4: BWD_PIPE(4); break;
        default: BWD_PIPE(2); break;
    }
#undef BWD_PIPE
}

// small systems: update-block op on the many-block FFMA kernel (umma_rows.cuh:ffma_rows_kernel)
template <class Op, int TR>
void launch_ffma_rows_t(mlffd_ctx* ctx, const Op& op, int N, const float* wt, int ld, cudaStream_t st) {
    static bool configured[16] = {};   // per device: the attribute is per (function, device)
    auto kernel = ffma_rows_kernel<Op, TR>;
    constexpr size_t smem = ffma_rows_smem_bytes<Op, TR>();
    if (!configured[ctx->device & 15]) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[ctx->device & 15] = true;
    }
    kernel<<<dim3(ceil_div(N, TR), 4), kSkinnyThreads, smem, st>>>(op, N, wt, ld, ctx->status_d);
}
template <class Op>
void launch_ffma_rows(mlffd_ctx* ctx, const Op& op, int N, const float* wt, int ld, cudaStream_t st) {
    // 8-row tiles (one row per epilogue warp) while they still fit the chip in about one wave
    if (N <= ctx->skinny8_rows) launch_ffma_rows_t<Op, 8>(ctx, op, N, wt, ld, st);
    else launch_ffma_rows_t<Op, 16>(ctx, op, N, wt, ld, st);
}

// ---- spline-filter message kernels (message_spline.cuh): one (slice, partition) per CTA -------
inline int spline_grid(int N, int slices, int groups, int ctas_per_sm) {
    const int blocks = ceil_div(std::max(N, 1), groups);
    const int parts = std::max(1, std::min(blocks, (kNumSMs * ctas_per_sm) / slices));
    return parts * slices;
}
const float* spline_layer_table(const mlffd_ctx* ctx, int l) {
    return ctx->spline_d + (size_t)l * 3 * ctx->H * kSplineRows;
}
template <int THREADS, int CTAS>
cudaError_t launch_spline_forward_t(mlffd_ctx* ctx, int l, int N, cudaStream_t st) {
    Workspace& ws = ctx->ws;
    const int grid = spline_grid(N, ctx->H / kSliceChannels, THREADS / 8, CTAS);
    static bool configured[16] = {};   // the attribute is per (function, device)
    if (!configured[ctx->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(spline_message_forward_kernel<true, THREADS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spline_message_forward_kernel<false, THREADS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e != cudaSuccess) return e;
        configured[ctx->device & 15] = true;
    }
    if (l == 0)
        spline_message_forward_kernel<true, THREADS, CTAS><<<grid, THREADS, kSplineSmemBytes, st>>>(
            spline_layer_table(ctx, l), ctx->H, ws.rowptr, ws.erec, ws.s_in[l], nullptr,
            ws.s_msg[l], ws.v_msg[l], N, ctx->status_d);
    else
        spline_message_forward_kernel<false, THREADS, CTAS><<<grid, THREADS, kSplineSmemBytes, st>>>(
            spline_layer_table(ctx, l), ctx->H, ws.rowptr, ws.erec, ws.s_in[l], ws.v_in[l],
            ws.s_msg[l], ws.v_msg[l], N, ctx->status_d);
    return cudaSuccess;
}
template <int THREADS, int CTAS>
cudaError_t launch_spline_backward_t(mlffd_ctx* ctx, int l, const float* sb, const float* vb, float* sb_in, float* vb_in,
                                     int N, cudaStream_t st) {
    Workspace& ws = ctx->ws;
    const int slices = ctx->H / kSliceChannels;
    const int grid = spline_grid(N, slices, THREADS / 8, CTAS);
    static bool configured[16] = {};
    if (!configured[ctx->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(spline_message_backward_kernel<true, THREADS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spline_message_backward_kernel<false, THREADS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e != cudaSuccess) return e;
        configured[ctx->device & 15] = true;
    }
    float4* slab = ws.edge_adj + (size_t)l * slices * ws.cap_edges;
    if (l == 0)
        spline_message_backward_kernel<true, THREADS, CTAS><<<grid, THREADS, kSplineSmemBytes, st>>>(
            spline_layer_table(ctx, l), ctx->H, ws.rowptr, ws.erec, ws.s_in[l], nullptr,
            sb, vb, sb_in, vb_in, slab, (size_t)ws.cap_edges, N, ws.lowptr, ctx->status_d);
    else
        spline_message_backward_kernel<false, THREADS, CTAS><<<grid, THREADS, kSplineSmemBytes, st>>>(
            spline_layer_table(ctx, l), ctx->H, ws.rowptr, ws.erec, ws.s_in[l], ws.v_in[l],
            sb, vb, sb_in, vb_in, slab, (size_t)ws.cap_edges, N, ws.lowptr, ctx->status_d);
    return cudaSuccess;
}
// small systems: a warp (four row groups) per CSR row, 8 rows per CTA
cudaError_t launch_spline_team(mlffd_ctx* ctx, int l, bool backward, const float* sb, const float* vb, float* sb_in,
                               float* vb_in, int N, cudaStream_t st) {
    Workspace& ws = ctx->ws;
    const int slices = ctx->H / kSliceChannels;
    const int grid = spline_grid(N, slices, kSplineTeamThreads / 32, 2);
    static bool configured[16] = {};
    if (!configured[ctx->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(spline_message_forward_team_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spline_message_forward_team_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spline_message_backward_team_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(spline_message_backward_team_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplineSmemBytes);
        if (e != cudaSuccess) return e;
        configured[ctx->device & 15] = true;
    }
    const float* table = spline_layer_table(ctx, l);
    if (!backward) {
        if (l == 0)
            spline_message_forward_team_kernel<true><<<grid, kSplineTeamThreads, kSplineSmemBytes, st>>>(
                table, ctx->H, ws.rowptr, ws.erec, ws.s_in[l], nullptr, ws.s_msg[l], ws.v_msg[l], N, ctx->status_d);
        else
            spline_message_forward_team_kernel<false><<<grid, kSplineTeamThreads, kSplineSmemBytes, st>>>(
                table, ctx->H, ws.rowptr, ws.erec, ws.s_in[l], ws.v_in[l], ws.s_msg[l], ws.v_msg[l], N, ctx->status_d);
    } else {
        float4* slab = ws.edge_adj + (size_t)l * slices * ws.cap_edges;
        if (l == 0)
            spline_message_backward_team_kernel<true><<<grid, kSplineTeamThreads, kSplineSmemBytes, st>>>(
                table, ctx->H, ws.rowptr, ws.erec, ws.s_in[l], nullptr, sb, vb, sb_in, vb_in, slab, (size_t)ws.cap_edges, N, ctx->status_d);
        else
            spline_message_backward_team_kernel<false><<<grid, kSplineTeamThreads, kSplineSmemBytes, st>>>(
                table, ctx->H, ws.rowptr, ws.erec, ws.s_in[l], ws.v_in[l], sb, vb, sb_in, vb_in, slab, (size_t)ws.cap_edges, N, ctx->status_d);
    }
    return cudaSuccess;
}

// launch shapes: (threads per CTA, resident CTAs per SM); env MLFFD_SPLINE_FWD / MLFFD_SPLINE_BWD pick one
bool spline_team_path(const mlffd_ctx* ctx, int N) { return N <= ctx->small_rows && ctx->msg_team > 1; }
cudaError_t launch_spline_forward(mlffd_ctx* ctx, int l, int N, cudaStream_t st) {
    if (spline_team_path(ctx, N)) return launch_spline_team(ctx, l, false, nullptr, nullptr, nullptr, nullptr, N, st);
    switch (ctx->spline_fwd_shape) {
        case 1: return launch_spline_forward_t<768, 1>(ctx, l, N, st);
        case 2: return launch_spline_forward_t<384, 2>(ctx, l, N, st);
        case 3: return launch_spline_forward_t<1024, 1>(ctx, l, N, st);
        default: return launch_spline_forward_t<512, 2>(ctx, l, N, st);
    }
}
cudaError_t launch_spline_backward(mlffd_ctx* ctx, int l, const float* sb, const float* vb, float* sb_in, float* vb_in,
                                   int N, cudaStream_t st) {
    if (spline_team_path(ctx, N)) return launch_spline_team(ctx, l, true, sb, vb, sb_in, vb_in, N, st);
    switch (ctx->spline_bwd_shape) {
        case 1: return launch_spline_backward_t<768, 1>(ctx, l, sb, vb, sb_in, vb_in, N, st);
        case 2: return launch_spline_backward_t<384, 2>(ctx, l, sb, vb, sb_in, vb_in, N, st);
        case 3: return launch_spline_backward_t<256, 2>(ctx, l, sb, vb, sb_in, vb_in, N, st);
        case 4: return launch_spline_backward_t<640, 1>(ctx, l, sb, vb, sb_in, vb_in, N, st);
        case 5: return launch_spline_backward_t<576, 1>(ctx, l, sb, vb, sb_in, vb_in, N, st);
        default: return launch_spline_backward_t<512, 1>(ctx, l, sb, vb, sb_in, vb_in, N, st);
    }
}

template <int H>
int run_model(mlffd_ctx* ctx, const int* z, int64_t n_atoms, const int* offsets, int n_structs,
              float* energy, float* forces, cudaStream_t st) {
    Workspace& ws = ctx->ws;
    const int N = (int)n_atoms, L = ctx->L;
    const DeviceStatus* status = ctx->status_d;
    using M = MsgTraits<H>;
    const int msg_grid = clamp_grid(ceil_div(N, 8 * M::APW), kNumSMs * 8);
    const int tiles = ceil_div(N, kTileRows);
    const int upd_fwd_grid = clamp_grid(tiles, kNumSMs * (update_fwd_smem_bytes<H>() <= 110 * 1024 ? 2 : 1));
    const int upd_bwd_grid = clamp_grid(tiles, kNumSMs * (update_bwd_smem_bytes<H>() <= 110 * 1024 ? 2 : 1));
    const int warp_grid = clamp_grid(ceil_div(N, 8), kNumSMs * 8);
    // reverse message kernels: 0 = one directed edge at a time (accumulates edge adjoints in place),
    // 1 = every pair once, 2 = pair once + cp.async ring (H = 128); 1 and 2 write per-layer slabs
    const bool bwd_pipe = ctx->msg_bwd_mode == 2 && H == 128;
    const bool fwd_pipe = ctx->msg_fwd_mode == 1 && H == 128;
    // small systems: a team of kTeam warps per CSR row (needs the per-layer adjoint slabs of the pair-once modes)
    const bool team = (ctx->msg_team == 4 || ctx->msg_team == 8) && N <= ctx->small_rows && ctx->msg_bwd_mode >= 1;
    const int team_grid = clamp_grid(ceil_div(N, M::APW), kNumSMs * 16);
    const bool adj_slabs = ctx->msg_bwd_mode >= 1;

    embedding_kernel<H><<<clamp_grid(ceil_div((int64_t)N * (H / 4), 256), kNumSMs * 8), 256, 0, st>>>(
        z, ctx->emb, ctx->cfg.max_z, ws.s_in[0], N);
    LAUNCHED(ctx, "embedding_kernel", MLFFD_STAGE_EMBEDDING, st);
    if (ctx->spline) {   // per-edge spline basis weights: once per step, read by every layer in both directions
        spline_basis_kernel<<<clamp_grid(ceil_div(std::max<int64_t>(ws.cap_edges, 1), 256), kNumSMs * 8), 256, 0, st>>>(
            ws.col, ws.geo, ctx->spline_inv_h, ws.erec, status);
        LAUNCHED(ctx, "spline_basis_kernel", MLFFD_STAGE_FILTER, st);
    }

    // the filters depend on the distances only: with the tensor-core kernel all layers' tables come
    // from one launch ahead of the layer loop
    const bool spline = ctx->spline;
    const bool use_umma = ctx->use_umma && !ctx->dense_fallback;   // tensor-core update block
    const bool filters_up_front = !spline && ctx->use_umma_filter && !ctx->dense_fallback &&
                                  (ctx->filter_batch == 1 || (ctx->filter_batch == 2 && N <= ctx->small_rows));
    if (filters_up_front) {
        int rc = launch_filter_umma<H>(ctx, 0, L, ws.pair_dist, &ctx->status_d->num_pairs, 0, status,
                                       nullptr, nullptr, ws.cap_pairs, st);
        if (rc) return rc;
    }
    for (int l = 0; l < L; ++l) {
        if (!spline && !filters_up_front) {
            int rc = launch_filter<H>(ctx, l, ws.pair_dist, &ctx->status_d->num_pairs, 0, status,
                                      ws.filt[l], ws.dfilt[l], ws.cap_pairs, st);
            if (rc) return rc;
        }
        if (spline) {
            CUDA_TRY(ctx, launch_spline_forward(ctx, l, N, st));
        } else if (team) {   // small system: a team of warps per CSR row (message_team.cuh)
#define MSG_FWD_TEAM(LAYER0, TT)                                                                       \
    message_forward_team_kernel<H, LAYER0, TT><<<team_grid, 32 * TT, 0, st>>>(                         \
        ws.rowptr, ws.col, ws.pair, ws.geo, ws.filt[l], ws.s_in[l], LAYER0 ? nullptr : ws.v_in[l],     \
        ws.s_msg[l], ws.v_msg[l], N, status)
            if (ctx->msg_team == 8) { if (l == 0) MSG_FWD_TEAM(true, 8); else MSG_FWD_TEAM(false, 8); }
            else                    { if (l == 0) MSG_FWD_TEAM(true, 4); else MSG_FWD_TEAM(false, 4); }
#undef MSG_FWD_TEAM
        } else if (fwd_pipe) {
            if constexpr (H == 128) launch_forward_pipe(ctx, l, msg_grid, N, st);
        } else if (l == 0)
            message_forward_kernel<H, true><<<msg_grid, 256, 0, st>>>(
                ws.rowptr, ws.col, ws.pair, ws.geo, ws.filt[l], ws.s_in[l], nullptr, ws.s_msg[l],
                ws.v_msg[l], N, status);
        else
            message_forward_kernel<H, false><<<msg_grid, 256, 0, st>>>(
                ws.rowptr, ws.col, ws.pair, ws.geo, ws.filt[l], ws.s_in[l], ws.v_in[l], ws.s_msg[l],
                ws.v_msg[l], N, status);
        LAUNCHED(ctx, "message_forward_kernel", MLFFD_STAGE_MESSAGE_FWD, st);
        bool upd_done = false;
        if constexpr (H == 128) {
            if (use_umma && N <= ctx->small_rows) {
                const UpdateWeights& uw = ctx->layer[l].update;
                launch_ffma_rows(ctx, UpdateFwd1Op{ws.s_msg[l], ws.v_msg[l], uw.m1, ws.y1[l]}, N, uw.M1t, H, st);
                LAUNCHED(ctx, "ffma_rows_kernel<UpdateFwd1Op>", MLFFD_STAGE_UPDATE_FWD, st);
                if (l == L - 1)
                    launch_ffma_rows(ctx, UpdateFwd2Op<true>{ws.y1[l], ws.s_msg[l], ws.v_msg[l], uw.m2, uw.U, ws.s_in[l + 1], nullptr, nullptr},
                                     N, uw.M2t, 3 * H, st);
                else
                    launch_ffma_rows(ctx, UpdateFwd2Op<false>{ws.y1[l], ws.s_msg[l], ws.v_msg[l], uw.m2, uw.U, ws.s_in[l + 1], ws.v_in[l + 1], ws.gates[l]},
                                     N, uw.M2t, 3 * H, st);
                upd_done = true;
            } else if (use_umma) {
                const UpdateWeights& uw = ctx->layer[l].update;
                const int rows_grid = clamp_grid(ceil_div(N, 128), kNumSMs);
                TC_DISPATCH(ctx->tc_mode, umma_rows_kernel<UpdateFwd1Op, MODE><<<rows_grid, kFilterUmmaThreads, UmmaRowsSmem::TOTAL, st>>>(
                    UpdateFwd1Op{ws.s_msg[l], ws.v_msg[l], uw.m1, ws.y1[l]}, N,
                    ctx->w2_images_d + ctx->img[l].upd_f1, status));
                LAUNCHED(ctx, "umma_rows_kernel<UpdateFwd1Op>", MLFFD_STAGE_UPDATE_FWD, st);
                if (l == L - 1)
                    TC_DISPATCH(ctx->tc_mode, umma_rows_kernel<UpdateFwd2Op<true>, MODE><<<rows_grid, kFilterUmmaThreads, UmmaRowsSmem::TOTAL, st>>>(
                        UpdateFwd2Op<true>{ws.y1[l], ws.s_msg[l], ws.v_msg[l], uw.m2, uw.U, ws.s_in[l + 1], nullptr, nullptr},
                        N, ctx->w2_images_d + ctx->img[l].upd_f2, status));
                else
                    TC_DISPATCH(ctx->tc_mode, umma_rows_kernel<UpdateFwd2Op<false>, MODE><<<rows_grid, kFilterUmmaThreads, UmmaRowsSmem::TOTAL, st>>>(
                        UpdateFwd2Op<false>{ws.y1[l], ws.s_msg[l], ws.v_msg[l], uw.m2, uw.U, ws.s_in[l + 1], ws.v_in[l + 1], ws.gates[l]},
                        N, ctx->w2_images_d + ctx->img[l].upd_f2, status));
                upd_done = true;
            }
        }
        if (upd_done) {
        } else if (l == L - 1)
            update_forward_kernel<H, true><<<upd_fwd_grid, kGemmThreads, update_fwd_smem_bytes<H>(), st>>>(
                ws.s_msg[l], ws.v_msg[l], ctx->layer[l].update, ws.s_in[l + 1], nullptr, ws.y1[l],
                nullptr, N, status);
        else
            update_forward_kernel<H, false><<<upd_fwd_grid, kGemmThreads, update_fwd_smem_bytes<H>(), st>>>(
                ws.s_msg[l], ws.v_msg[l], ctx->layer[l].update, ws.s_in[l + 1], ws.v_in[l + 1],
                ws.y1[l], ws.gates[l], N, status);
        LAUNCHED(ctx, "update_forward_kernel", MLFFD_STAGE_UPDATE_FWD, st);
    }

    const bool want_forces = forces != nullptr;
    auto adj = [&](int l) { return ctx->debug_keep ? l : (l & 1); };
    // enough 64-atom tiles to fill the GPU: register-tiled GEMMs (env MLFFD_READOUT = tile | warp forces one)
    if (ctx->readout_mode == 0 && N <= ctx->small_rows) {
        readout_block_kernel<H><<<ceil_div(N, 4), 256, 0, st>>>(ws.s_in[L], ctx->head, ws.eps,
                                                               want_forces ? ws.sbar[adj(L - 1)] : nullptr, N, status);
    } else if (H == 128 && (ctx->readout_mode == 1 || (ctx->readout_mode == 0 && N >= 64 * kNumSMs))) {
        if (!ctx->readout_configured) {
            cudaFuncSetAttribute(readout_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)readout_tile_smem_bytes());
            ctx->readout_configured = true;
        }
        readout_tile_kernel<<<clamp_grid(ceil_div(N, kTileRows), kNumSMs * 3), kGemmThreads, readout_tile_smem_bytes(), st>>>(
            ws.s_in[L], ctx->head, ws.eps, want_forces ? ws.sbar[adj(L - 1)] : nullptr, N, status);
    } else
        readout_kernel<H><<<clamp_grid(ceil_div(N, 8 * 4), kNumSMs * 8), 256, 0, st>>>(ws.s_in[L], ctx->head, ws.eps,
                                                     want_forces ? ws.sbar[adj(L - 1)] : nullptr, N, status);
    LAUNCHED(ctx, "readout_kernel", MLFFD_STAGE_READOUT, st);
    structure_energy_kernel<<<clamp_grid(ceil_div(n_structs, 8), kNumSMs * 8), 256, 0, st>>>(
        ws.eps, offsets, n_structs, energy, status);
    LAUNCHED(ctx, "structure_energy_kernel", MLFFD_STAGE_ENERGY_SUM, st);
    if (!want_forces) { ctx->last_adj_slabs = 0; return MLFFD_OK; }

    for (int l = L - 1; l >= 0; --l) {
        float* sb = ws.sbar[adj(l)];
        float* vb = ws.vbar[adj(l)];
        bool bwd_done = false;
        if constexpr (H == 128) {
            if (use_umma && N <= ctx->small_rows) {
                const UpdateWeights& uw = ctx->layer[l].update;
                if (l == L - 1) {
                    launch_ffma_rows(ctx, UpdateBwd1Op<true>{sb, vb, ws.v_msg[l], uw.U, ws.y1[l]}, N, uw.M2, H, st);
                    LAUNCHED(ctx, "ffma_rows_kernel<UpdateBwd1Op>", MLFFD_STAGE_UPDATE_BWD, st);
                    launch_ffma_rows(ctx, UpdateBwd2Op<true>{ws.y1[l], ws.v_msg[l], nullptr, uw.U, sb, vb}, N, uw.M1, 2 * H, st);
                } else {
                    launch_ffma_rows(ctx, UpdateBwd1Op<false>{sb, vb, ws.v_msg[l], uw.U, ws.y1[l]}, N, uw.M2, H, st);
                    LAUNCHED(ctx, "ffma_rows_kernel<UpdateBwd1Op>", MLFFD_STAGE_UPDATE_BWD, st);
                    launch_ffma_rows(ctx, UpdateBwd2Op<false>{ws.y1[l], ws.v_msg[l], ws.gates[l], uw.U, sb, vb}, N, uw.M1, 2 * H, st);
                }
                bwd_done = true;
            } else if (use_umma) {
                const UpdateWeights& uw = ctx->layer[l].update;
                const int rows_grid = clamp_grid(ceil_div(N, 128), kNumSMs);
                if (l == L - 1) {
                    TC_DISPATCH(ctx->tc_mode, umma_rows_kernel<UpdateBwd1Op<true>, MODE><<<rows_grid, kFilterUmmaThreads, UmmaRowsSmem::TOTAL, st>>>(
                        UpdateBwd1Op<true>{sb, vb, ws.v_msg[l], uw.U, ws.y1[l]}, N, ctx->w2_images_d + ctx->img[l].upd_b1, status));
                    LAUNCHED(ctx, "umma_rows_kernel<UpdateBwd1Op>", MLFFD_STAGE_UPDATE_BWD, st);
                    TC_DISPATCH(ctx->tc_mode, umma_rows_kernel<UpdateBwd2Op<true>, MODE><<<rows_grid, kFilterUmmaThreads, UmmaRowsSmem::TOTAL, st>>>(
                        UpdateBwd2Op<true>{ws.y1[l], ws.v_msg[l], nullptr, uw.U, sb, vb}, N, ctx->w2_images_d + ctx->img[l].upd_b2, status));
                } else {
                    TC_DISPATCH(ctx->tc_mode, umma_rows_kernel<UpdateBwd1Op<false>, MODE><<<rows_grid, kFilterUmmaThreads, UmmaRowsSmem::TOTAL, st>>>(
                        UpdateBwd1Op<false>{sb, vb, ws.v_msg[l], uw.U, ws.y1[l]}, N, ctx->w2_images_d + ctx->img[l].upd_b1, status));
                    LAUNCHED(ctx, "umma_rows_kernel<UpdateBwd1Op>", MLFFD_STAGE_UPDATE_BWD, st);
                    TC_DISPATCH(ctx->tc_mode, umma_rows_kernel<UpdateBwd2Op<false>, MODE><<<rows_grid, kFilterUmmaThreads, UmmaRowsSmem::TOTAL, st>>>(
                        UpdateBwd2Op<false>{ws.y1[l], ws.v_msg[l], ws.gates[l], uw.U, sb, vb}, N, ctx->w2_images_d + ctx->img[l].upd_b2, status));
                }
                bwd_done = true;
            }
        }
        if (bwd_done) {
        } else if (l == L - 1)
            update_backward_kernel<H, true><<<upd_bwd_grid, kGemmThreads, update_bwd_smem_bytes<H>(), st>>>(
                ws.v_msg[l], ws.y1[l], nullptr, ctx->layer[l].update, sb, vb, N, status);
        else
            update_backward_kernel<H, false><<<upd_bwd_grid, kGemmThreads, update_bwd_smem_bytes<H>(), st>>>(
                ws.v_msg[l], ws.y1[l], ws.gates[l], ctx->layer[l].update, sb, vb, N, status);
        LAUNCHED(ctx, "update_backward_kernel", MLFFD_STAGE_UPDATE_BWD, st);
        float* sb_in = (l > 0) ? ws.sbar[adj(l - 1)] : nullptr;
        float* vb_in = (l > 0) ? ws.vbar[adj(l - 1)] : nullptr;
        const bool first = (l == L - 1);
#define MSG_BWD(LAYER0, ACC)                                                                     \
    message_backward_kernel<H, LAYER0, ACC><<<msg_grid, 256, 0, st>>>(                           \
        ws.rowptr, ws.col, ws.pair, ws.geo, ws.filt[l], ws.dfilt[l], ws.s_in[l], ws.v_in[l], sb, \
        vb, sb_in, vb_in, ws.edge_adj, N, status)
#define MSG_BWD_PAIRS(LAYER0)                                                                    \
    message_backward_pairs_kernel<H, LAYER0, false><<<msg_grid, 256, 0, st>>>(                   \
        ws.rowptr, ws.col, ws.pair, ws.rev, ws.geo, ws.filt[l], ws.dfilt[l], ws.s_in[l],         \
        ws.v_in[l], sb, vb, sb_in, vb_in, ws.edge_adj + (size_t)l * ws.cap_edges, N, status)
        if (spline) {
            CUDA_TRY(ctx, launch_spline_backward(ctx, l, sb, vb, sb_in, vb_in, N, st));
        } else if (team) {
#define MSG_BWD_TEAM(LAYER0, TT)                                                                          \
    message_backward_pairs_team_kernel<H, LAYER0, TT><<<team_grid, 32 * TT, 0, st>>>(                     \
        ws.rowptr, ws.col, ws.pair, ws.rev, ws.geo, ws.filt[l], ws.dfilt[l], ws.s_in[l], ws.v_in[l], sb,  \
        vb, sb_in, vb_in, ws.edge_adj + (size_t)l * ws.cap_edges, N, status)
            if (ctx->msg_team == 8) { if (l == 0) MSG_BWD_TEAM(true, 8); else MSG_BWD_TEAM(false, 8); }
            else                    { if (l == 0) MSG_BWD_TEAM(true, 4); else MSG_BWD_TEAM(false, 4); }
#undef MSG_BWD_TEAM
        } else if (bwd_pipe) {
            if constexpr (H == 128) launch_backward_pipe(ctx, l, msg_grid, sb, vb, sb_in, vb_in, N, st);
        } else if (ctx->msg_bwd_mode >= 1) {
            if (l == 0) MSG_BWD_PAIRS(true); else MSG_BWD_PAIRS(false);
        } else if (l == 0) { if (first) MSG_BWD(true, false); else MSG_BWD(true, true); }
        else        { if (first) MSG_BWD(false, false); else MSG_BWD(false, true); }
#undef MSG_BWD_PAIRS
#undef MSG_BWD
        LAUNCHED(ctx, "message_backward_kernel", MLFFD_STAGE_MESSAGE_BWD, st);
    }
    ctx->last_adj_slabs = spline ? L * ctx->adj_slabs_per_layer : (adj_slabs ? L : 1);
    // spline reverse kernels: layer 0 writes the adjoint of e, the other layers that of rev(e) (message_spline.cuh)
    ctx->last_adj_direct = spline ? ctx->adj_slabs_per_layer : ctx->last_adj_slabs;
    // ... and layer 0 of the row kernels (not of the small-system team kernels) writes pair sums
    ctx->last_adj_pairs = (spline && !spline_team_path(ctx, N)) ? ctx->adj_slabs_per_layer : 0;
    force_kernel<<<warp_grid, 256, 0, st>>>(ws.rowptr, ws.rev, ws.geo, ws.edge_adj, ctx->last_adj_slabs,
                                            ctx->last_adj_pairs, ctx->last_adj_direct, (size_t)ws.cap_edges,
                                            ctx->debug_keep ? ws.edge_adj + (size_t)L * ctx->adj_slabs_per_layer * ws.cap_edges : nullptr,
                                            forces, N, status);
    LAUNCHED(ctx, "force_kernel", MLFFD_STAGE_FORCE, st);
    return MLFFD_OK;
}

int build_neighbors(mlffd_ctx* ctx, const float* pos, const int* offsets, int n_structs,
                    int64_t n_atoms, const float* cells, const uint8_t* pbc, cudaStream_t st) {
    Workspace& ws = ctx->ws;
    if (n_atoms > ws.cap_atoms || n_structs > ws.cap_structs)
        return fail(ctx, MLFFD_ECAPACITY, "workspace too small: call mlffd_workspace_reserve");
    if ((cells == nullptr) != (pbc == nullptr))
        return fail(ctx, MLFFD_EINVAL, "cells_d and pbc_d must both be given or both be NULL");
    const int N = (int)n_atoms;
    ctx->last_atoms = n_atoms;
    ctx->last_structs = n_structs;
    ctx->last_stream = st;
    mark(ctx, -1, st, 0);  // step start
    if (N <= kSmallNeighborAtoms && ctx->neighbor_mode == 0 && ctx->small_rows > 0) {   // latency path: one launch
        neighbor_small_kernel<<<1, 1024, 0, st>>>(pos, offsets, n_structs, cells, pbc, N, ctx->cfg.cutoff,
                                                  (int)ws.cap_edges, ws.atom_struct, ws.rowptr, ws.lowptr, ws.col,
                                                  ws.edge_dst, ws.geo, ws.rev, ws.pair, ws.pair_dist, ctx->status_d);
        LAUNCHED(ctx, "neighbor_small_kernel", MLFFD_STAGE_NEIGHBOR, st);
        return MLFFD_OK;
    }
    atom_structure_kernel<<<clamp_grid(ceil_div(N, 256), kNumSMs * 8), 256, 0, st>>>(
        offsets, n_structs, N, ws.atom_struct);
    LAUNCHED(ctx, "atom_structure_kernel", MLFFD_STAGE_NEIGHBOR, st);
    CUDA_TRY(ctx, cudaMemsetAsync(ws.deg + N, 0, sizeof(int), st));
    CUDA_TRY(ctx, cudaMemsetAsync(ws.deg_low + N, 0, sizeof(int), st));
    const int sweep_grid = clamp_grid(ceil_div(N, 8), kNumSMs * 16);
    const bool use_cells = ctx->neighbor_mode == 2 ||
                           (ctx->neighbor_mode == 0 && n_atoms >= (int64_t)2048 * n_structs);
    const bool skin = ctx->skin > 0.f && ws.skin != nullptr;
    // Candidate generator of the heavy count / fill passes: the whole structure (sweep), the 27 surrounding
    // cells, or -- with a skin -- the candidate list, itself rebuilt by the same sweep / cell kernels with
    // cutoff + skin whenever the device-side displacement check asks for it (skin_list.cuh).
    size_t bytes;
    auto candidate_pass = [&](bool fill, float cutoff, int* deg, int* deg_low, const int* rowptr, int* col,
                              const int* gate) -> int {
        if (use_cells) {
            if (!fill) {
                const int C = (int)ws.cap_cells;
                const int cap_per_struct = (int)std::max<int64_t>(8, 2 * (n_atoms / n_structs));
                grid_setup_kernel<<<n_structs, 256, 0, st>>>(pos, offsets, cells, pbc, cutoff, cap_per_struct, ws.grids,
                                                             ws.ncells, gate);
                LAUNCHED(ctx, "grid_setup_kernel", MLFFD_STAGE_NEIGHBOR, st);
                grid_offsets_kernel<<<1, 32, 0, st>>>(ws.grids, ws.ncells, n_structs, ws.total_cells, gate);
                LAUNCHED(ctx, "grid_offsets_kernel", MLFFD_STAGE_NEIGHBOR, st);
                CUDA_TRY(ctx, cudaMemsetAsync(ws.cell_count, 0, sizeof(int) * (C + 1), st));
                CUDA_TRY(ctx, cudaMemsetAsync(ws.cell_cursor, 0, sizeof(int) * (C + 1), st));
                const int atom_grid = clamp_grid(ceil_div(N, 256), kNumSMs * 8);
                cell_count_kernel<<<atom_grid, 256, 0, st>>>(pos, ws.atom_struct, ws.grids, N, ws.atom_cell, ws.cell_count, gate);
                LAUNCHED(ctx, "cell_count_kernel", MLFFD_STAGE_NEIGHBOR, st);
                size_t b2 = ws.cub_bytes;
                CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ws.cub_temp, b2, ws.cell_count, ws.cell_start, C + 1, st));
                mark(ctx, MLFFD_STAGE_NEIGHBOR, st, 2);
                cell_fill_kernel<<<atom_grid, 256, 0, st>>>(ws.atom_cell, ws.cell_start, ws.cell_cursor, N, ws.cell_atoms, gate);
                LAUNCHED(ctx, "cell_fill_kernel", MLFFD_STAGE_NEIGHBOR, st);
                neighbor_cells_kernel<false><<<sweep_grid, 256, 0, st>>>(
                    pos, ws.atom_struct, cells, ws.grids, ws.cell_start, ws.cell_atoms, N, cutoff,
                    deg, deg_low, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->status_d, gate);
                LAUNCHED(ctx, "neighbor_cells_kernel<count>", MLFFD_STAGE_NEIGHBOR, st);
            } else {
                // scratch for the unsorted rows: `rev` and `edge_adj` are only written by later kernels
                neighbor_cells_kernel<true><<<sweep_grid, 256, 0, st>>>(
                    pos, ws.atom_struct, cells, ws.grids, ws.cell_start, ws.cell_atoms, N, cutoff,
                    nullptr, nullptr, rowptr, ws.rev, ws.edge_adj, col, ws.edge_dst, ws.geo, ctx->status_d, gate);
                LAUNCHED(ctx, "neighbor_cells_kernel<fill>", MLFFD_STAGE_NEIGHBOR, st);
            }
        } else if (!fill) {
            neighbor_sweep_kernel<false><<<sweep_grid, 256, 0, st>>>(
                pos, offsets, ws.atom_struct, cells, pbc, N, cutoff, deg, deg_low, nullptr,
                nullptr, nullptr, nullptr, ctx->status_d, gate);
            LAUNCHED(ctx, "neighbor_sweep_kernel<count>", MLFFD_STAGE_NEIGHBOR, st);
        } else {
            neighbor_sweep_kernel<true><<<sweep_grid, 256, 0, st>>>(
                pos, offsets, ws.atom_struct, cells, pbc, N, cutoff, nullptr, nullptr, rowptr,
                col, ws.edge_dst, ws.geo, ctx->status_d, gate);
            LAUNCHED(ctx, "neighbor_sweep_kernel<fill>", MLFFD_STAGE_NEIGHBOR, st);
        }
        return MLFFD_OK;
    };
    if (skin) {
        if (ctx->skin_atoms != n_atoms || ctx->skin_structs != n_structs) {   // another system: no valid candidates
            CUDA_TRY(ctx, cudaMemsetAsync(ws.skin, 0, sizeof(SkinState), st));
            ctx->skin_atoms = n_atoms; ctx->skin_structs = n_structs;
        }
        const int* gate = &ws.skin->rebuild;
        skin_check_kernel<<<clamp_grid(ceil_div(N, 256), kNumSMs * 4), 256, 0, st>>>(pos, ws.pos_ref, N, ws.skin);
        LAUNCHED(ctx, "skin_check_kernel", MLFFD_STAGE_NEIGHBOR, st);
        skin_decide_kernel<<<1, 32, 0, st>>>(ws.skin, 0.25f * ctx->skin * ctx->skin, ctx->status_d);
        LAUNCHED(ctx, "skin_decide_kernel", MLFFD_STAGE_NEIGHBOR, st);
        const float wide = ctx->cfg.cutoff + ctx->skin;
        int rc = candidate_pass(false, wide, ws.cand_deg, ws.deg_low, nullptr, nullptr, gate);   // deg_low: scratch here
        if (rc) return rc;
        if (N <= 32768) {
            neighbor_scan_small_kernel<<<1, 1024, 0, st>>>(ws.cand_deg, ws.deg_low, ws.cand_rowptr, ws.lowptr, N,
                                                           (int)ws.cap_edges, nullptr);
            LAUNCHED(ctx, "neighbor_scan_small_kernel", MLFFD_STAGE_NEIGHBOR, st);
        } else {
            bytes = ws.cub_bytes;
            CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ws.cub_temp, bytes, ws.cand_deg, ws.cand_rowptr, N + 1, st));
            mark(ctx, MLFFD_STAGE_NEIGHBOR, st, 2);
        }
        skin_cand_finalize_kernel<<<1, 32, 0, st>>>(ws.cand_rowptr, N, (int)ws.cap_edges, ws.skin, ctx->status_d);
        LAUNCHED(ctx, "skin_cand_finalize_kernel", MLFFD_STAGE_NEIGHBOR, st);
        rc = candidate_pass(true, wide, nullptr, nullptr, ws.cand_rowptr, ws.cand_col, gate);
        if (rc) return rc;
        skin_copy_ref_kernel<<<clamp_grid(ceil_div(3 * N, 256), kNumSMs * 4), 256, 0, st>>>(pos, ws.pos_ref, 3 * N, ws.skin);
        LAUNCHED(ctx, "skin_copy_ref_kernel", MLFFD_STAGE_NEIGHBOR, st);
        neighbor_cand_kernel<false><<<sweep_grid, 256, 0, st>>>(
            pos, ws.atom_struct, cells, pbc, N, ctx->cfg.cutoff, ws.cand_rowptr, ws.cand_col, ws.deg, ws.deg_low,
            nullptr, nullptr, nullptr, nullptr, ws.skin, ctx->status_d);
        LAUNCHED(ctx, "neighbor_cand_kernel<count>", MLFFD_STAGE_NEIGHBOR, st);
    } else {
        int rc = candidate_pass(false, ctx->cfg.cutoff, ws.deg, ws.deg_low, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    if (N <= 32768) {   // latency path: one launch instead of five
        neighbor_scan_small_kernel<<<1, 1024, 0, st>>>(ws.deg, ws.deg_low, ws.rowptr, ws.lowptr, N,
                                                       (int)ws.cap_edges, ctx->status_d);
        LAUNCHED(ctx, "neighbor_scan_small_kernel", MLFFD_STAGE_NEIGHBOR, st);
    } else {
        bytes = ws.cub_bytes;
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ws.cub_temp, bytes, ws.deg, ws.rowptr, N + 1, st));
        bytes = ws.cub_bytes;
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ws.cub_temp, bytes, ws.deg_low, ws.lowptr, N + 1, st));
        mark(ctx, MLFFD_STAGE_NEIGHBOR, st, 4);  // two CUB scans = 2 x (init + scan) kernels
        neighbor_finalize_kernel<<<1, 32, 0, st>>>(ws.rowptr, ws.lowptr, N, (int)ws.cap_edges,
                                                   ctx->status_d);
        LAUNCHED(ctx, "neighbor_finalize_kernel", MLFFD_STAGE_NEIGHBOR, st);
    }
    if (skin) {
        skin_merge_status_kernel<<<1, 32, 0, st>>>(ws.skin, ctx->status_d);
        LAUNCHED(ctx, "skin_merge_status_kernel", MLFFD_STAGE_NEIGHBOR, st);
        neighbor_cand_kernel<true><<<sweep_grid, 256, 0, st>>>(
            pos, ws.atom_struct, cells, pbc, N, ctx->cfg.cutoff, ws.cand_rowptr, ws.cand_col, nullptr, nullptr,
            ws.rowptr, ws.col, ws.edge_dst, ws.geo, ws.skin, ctx->status_d);
        LAUNCHED(ctx, "neighbor_cand_kernel<fill>", MLFFD_STAGE_NEIGHBOR, st);
    } else {
        int rc = candidate_pass(true, ctx->cfg.cutoff, nullptr, nullptr, ws.rowptr, ws.col, nullptr);
        if (rc) return rc;
    }
    reverse_pair_kernel<<<clamp_grid(ceil_div(std::max<int64_t>(ws.cap_edges, 1), 256), kNumSMs * 8),
                          256, 0, st>>>(ws.rowptr, ws.lowptr, ws.col, ws.edge_dst, ws.geo, ws.rev,
                                        ws.pair, ws.pair_dist, ctx->status_d);
    LAUNCHED(ctx, "reverse_pair_kernel", MLFFD_STAGE_NEIGHBOR, st);
    return MLFFD_OK;
}

struct ArenaPlan {
    size_t total = 0;
    size_t take(size_t bytes) {
        const size_t off = total;
        total += (bytes + 255) / 256 * 256;
        return off;
    }
};

}  // namespace

// ================================ C ABI =====================================================

extern "C" int mlffd_version(void) { return MLFFD_ABI_VERSION; }

extern "C" const char* mlffd_last_error(const mlffd_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int mlffd_model_create(mlffd_ctx** out, int device, const mlffd_config* config,
                                  const float* weights_host, size_t num_floats) {
    if (!out || !config || !weights_host) return fail(nullptr, MLFFD_EINVAL, "null argument");
    *out = nullptr;
    const int H = config->hidden_dim, K = config->num_rbf, L = config->num_interactions;
    if (H != 32 && H != 64 && H != 128)
        return fail(nullptr, MLFFD_EINVAL, "hidden_dim must be 32, 64 or 128");
    if (K < 1 || K > kMaxRbf) return fail(nullptr, MLFFD_EINVAL, "num_rbf must be in 1..32");
    if (L < 1 || L > kMaxLayers) return fail(nullptr, MLFFD_EINVAL, "num_interactions must be in 1..8");
    if (config->max_z < 1 || !(config->cutoff > 0.f))
        return fail(nullptr, MLFFD_EINVAL, "max_z and cutoff must be positive");
    if (config->precision != MLFFD_PREC_FP32 && config->precision != MLFFD_PREC_TC_FP16X2 &&
        config->precision != MLFFD_PREC_TC_FP16 && config->precision != MLFFD_PREC_TC_BF16)
        return fail(nullptr, MLFFD_EINVAL, "precision must be MLFFD_PREC_FP32, _TC_FP16X2, _TC_FP16 or _TC_BF16");
    if (config->filter_mode != MLFFD_FILTER_SPLINE && config->filter_mode != MLFFD_FILTER_TABLE)
        return fail(nullptr, MLFFD_EINVAL, "filter_mode must be MLFFD_FILTER_SPLINE or MLFFD_FILTER_TABLE");
    const size_t per_layer = (size_t)H * K + H + 3 * H * H + 3 * H + 2 * H * H + H + 3 * H * H + 3 * H + 9;
    const size_t expect = (size_t)(config->max_z + 1) * H + 2 * K + L * per_layer +
                          (size_t)(H / 2) * H + H / 2 + (size_t)(H / 4) * (H / 2) + H / 4 + H / 4 + 1;
    if (num_floats != expect)
        return fail(nullptr, MLFFD_EINVAL, "weight blob has " + std::to_string(num_floats) +
                                               " floats, expected " + std::to_string(expect));
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MLFFD_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, MLFFD_EINVAL, "bad device index");
    DeviceGuard device_guard(device);
    {
        int current = -1;
        if (cudaGetDevice(&current) != cudaSuccess || current != device)
            return fail(nullptr, MLFFD_ECUDA, "cannot select the requested CUDA device");
    }

    mlffd_ctx* ctx = new (std::nothrow) mlffd_ctx();
    if (!ctx) return fail(nullptr, MLFFD_ENOMEM, "out of host memory");
    ctx->device = device;
    ctx->cfg = *config;
    ctx->H = H; ctx->K = K; ctx->L = L;
    const char* dbg = std::getenv("MLFFD_DEBUG_KEEP");
    ctx->debug_keep = dbg && dbg[0] == '1';
    // measured (C2 shapes, one B200): the pair-once reverse pass wins at H = 128 (1.49 vs 1.97 ms),
    // but with 2 - 4 atoms per warp its two edge classes diverge and the per-directed-edge kernel
    // is faster (Tiny 0.69 vs 0.88 ms, Ultra-tiny 0.44 vs 0.57 ms)
    ctx->msg_bwd_mode = (H == 128) ? 2 : 0;
    if (const char* ns = std::getenv("MLFFD_MSG_BWD"))
        ctx->msg_bwd_mode = !std::strcmp(ns, "edges") ? 0 : !std::strcmp(ns, "pairs") ? 1 : 2;
    if (const char* ns = std::getenv("MLFFD_READOUT")) ctx->readout_mode = !std::strcmp(ns, "tile") ? 1 : 2;
    if (const char* ns = std::getenv("MLFFD_MSG_FWD")) ctx->msg_fwd_mode = !std::strcmp(ns, "rows") ? 0 : 1;
    if (const char* ns = std::getenv("MLFFD_SMALL_ROWS")) ctx->small_rows = std::atoi(ns);
    if (const char* ns = std::getenv("MLFFD_FILTER_BATCH")) ctx->filter_batch = std::atoi(ns);
    if (const char* ns = std::getenv("MLFFD_MSG_TEAM")) ctx->msg_team = std::atoi(ns);
    if (const char* ns = std::getenv("MLFFD_SKINNY8_ROWS")) ctx->skinny8_rows = std::atoi(ns);
    if (const char* ns = std::getenv("MLFFD_PIPE_DEPTH_FWD")) ctx->pipe_depth_fwd = std::atoi(ns);
    if (const char* ns = std::getenv("MLFFD_PIPE_DEPTH_BWD")) ctx->pipe_depth_bwd = std::atoi(ns);
    if (const char* nm = std::getenv("MLFFD_NEIGHBOR"))
        ctx->neighbor_mode = (std::strcmp(nm, "sweep") == 0) ? 1 : (std::strcmp(nm, "cells") == 0) ? 2 : 0;

    // walk the blob in state_dict order and stage the device layout
    Stager st;
    const float* p = weights_host;
    auto take = [&](size_t n) { const float* q = p; p += n; return q; };
    const size_t o_emb = st.push(take((size_t)(config->max_z + 1) * H), (size_t)(config->max_z + 1) * H);
    const size_t o_centers = st.push(take(K), K);
    const float* widths = take(K);
    std::vector<float> gam(K);
    for (int k = 0; k < K; ++k) gam[k] = 1.0f / (widths[k] * widths[k]);  // student_model.py:252
    const size_t o_gammas = st.push(gam.data(), K);
    struct LOff { size_t W1t, b1, W2t, b2, M1t, m1, M2t, m2, M1, M2, U; } lo[kMaxLayers];
    for (int l = 0; l < L; ++l) {
        const float* W1 = take((size_t)H * K);
        lo[l].W1t = st.push_transposed(W1, H, K);
        lo[l].b1 = st.push(take(H), H);
        const float* W2 = take((size_t)3 * H * H);
        lo[l].W2t = st.push_transposed(W2, 3 * H, H);
        lo[l].b2 = st.push(take(3 * H), 3 * H);
        const float* M1 = take((size_t)H * 2 * H);
        lo[l].M1t = st.push_transposed(M1, H, 2 * H);
        lo[l].M1 = st.push(M1, (size_t)H * 2 * H);
        lo[l].m1 = st.push(take(H), H);
        const float* M2 = take((size_t)3 * H * H);
        lo[l].M2t = st.push_transposed(M2, 3 * H, H);
        lo[l].M2 = st.push(M2, (size_t)3 * H * H);
        lo[l].m2 = st.push(take(3 * H), 3 * H);
        lo[l].U = st.push(take(9), 9);
    }
    const float* A1 = take((size_t)(H / 2) * H);
    const size_t o_A1t = st.push_transposed(A1, H / 2, H);
    const size_t o_A1 = st.push(A1, (size_t)(H / 2) * H);
    const size_t o_a1 = st.push(take(H / 2), H / 2);
    const float* A2 = take((size_t)(H / 4) * (H / 2));
    const size_t o_A2t = st.push_transposed(A2, H / 4, H / 2);
    const size_t o_A2 = st.push(A2, (size_t)(H / 4) * (H / 2));
    const size_t o_a2 = st.push(take(H / 4), H / 4);
    const size_t o_A3 = st.push(take(H / 4), H / 4);
    const size_t o_a3 = st.push(take(1), 1);

    auto bail = [&](int code, const std::string& msg) {
        g_create_error = msg;
        if (ctx->weights_d) cudaFree(ctx->weights_d);
        if (ctx->w2_images_d) cudaFree(ctx->w2_images_d);
        if (ctx->spline_d) cudaFree(ctx->spline_d);
        if (ctx->status_d) cudaFree(ctx->status_d);
        delete ctx;
        return code;
    };
    e = cudaMalloc(&ctx->weights_d, st.buf.size() * sizeof(float));
    if (e != cudaSuccess) return bail(MLFFD_ENOMEM, cudaGetErrorString(e));
    e = cudaMemcpy(ctx->weights_d, st.buf.data(), st.buf.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(MLFFD_ECUDA, cudaGetErrorString(e));
    e = cudaMalloc(&ctx->status_d, sizeof(DeviceStatus));
    if (e != cudaSuccess) return bail(MLFFD_ENOMEM, cudaGetErrorString(e));
    cudaMemset(ctx->status_d, 0, sizeof(DeviceStatus));
    {   // per-model filter splines (spline_table.h), FP64 on the host, once
        ctx->spline = config->filter_mode == MLFFD_FILTER_SPLINE;
        ctx->adj_slabs_per_layer = ctx->spline ? H / kSliceChannels : 1;
        ctx->spline_inv_h = (float)kSplineIntervals / config->cutoff;
        const float* centers_h = weights_host + (size_t)(config->max_z + 1) * H;
        const float* q = centers_h + 2 * K;
        std::vector<float> all;
        for (int l = 0; l < L; ++l) {
            const float *W1 = q, *b1 = W1 + (size_t)H * K, *W2 = b1 + H, *b2 = W2 + (size_t)3 * H * H;
            const std::vector<float> t = build_filter_spline(H, K, config->cutoff, centers_h, gam.data(), W1, b1, W2, b2);
            all.insert(all.end(), t.begin(), t.end());
            q += per_layer;
        }
        e = cudaMalloc(&ctx->spline_d, all.size() * sizeof(float));
        if (e != cudaSuccess) return bail(MLFFD_ENOMEM, cudaGetErrorString(e));
        e = cudaMemcpy(ctx->spline_d, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(MLFFD_ECUDA, cudaGetErrorString(e));
        if (const char* ns = std::getenv("MLFFD_SPLINE_FWD")) ctx->spline_fwd_shape = std::atoi(ns);
        if (const char* ns = std::getenv("MLFFD_SPLINE_BWD")) ctx->spline_bwd_shape = std::atoi(ns);
    }
    if (config->precision != MLFFD_PREC_FP32) {
        ctx->tc_mode = config->precision == MLFFD_PREC_TC_FP16 ? kTcF16
                     : config->precision == MLFFD_PREC_TC_BF16 ? kTcBF16 : kTcSplit;
        const int tc_mode = ctx->tc_mode;
        // K-major SWIZZLE_128B images of 128-row blocks of an [out][in] matrix:
        // [hi kb0 .. | lo kb0 ..], each K block 128 rows x 64 halves (see filter_umma.cuh); rows or
        // columns beyond the matrix are zero.
        // (the single-pass modes read only the hi halves; BF16 mode stores them as BF16 patterns)
        std::vector<uint16_t> img;
        auto add_image = [&](const float* Wm, int ld, int rows, int cols, int r0, int c0, int kblk) {
            const size_t base = img.size();
            const size_t term = (size_t)kblk * (kKBlockBytes / 2);
            img.resize(base + 2 * term, (uint16_t)0);
            for (int r = 0; r < 128; ++r)
                for (int k = 0; k < kblk * 64; ++k) {
                    if (r0 + r >= rows || c0 + k >= cols) continue;
                    const float x = Wm[(size_t)(r0 + r) * ld + c0 + k] * kWeightScale;
                    uint16_t hi, lo = 0;
                    if (tc_mode == kTcBF16) {
                        hi = static_cast<__nv_bfloat16_raw>(__float2bfloat16_rn(x)).x;
                    } else {
                        const __half h = __float2half_rn(x);
                        hi = static_cast<__half_raw>(h).x;
                        lo = static_cast<__half_raw>(__float2half_rn(x - __half2float(h))).x;
                    }
                    const int kb = k / 64, kc = k % 64;
                    const size_t off = (size_t)(r / 8) * 512 + (size_t)(r % 8) * 64 +
                                       (size_t)(((kc / 8) ^ (r % 8)) * 8) + (size_t)(kc % 8);
                    img[base + (size_t)kb * (kKBlockBytes / 2) + off] = hi;
                    img[base + term + (size_t)kb * (kKBlockBytes / 2) + off] = lo;
                }
            return base * sizeof(uint16_t);
        };
        const float* q = weights_host + (size_t)(config->max_z + 1) * H + 2 * K;
        const int fkblk = (H > 64) ? H / 64 : 1;
        const int fchunks = (3 * H + 127) / 128;
        std::vector<float> M1T((size_t)2 * H * H), M2T((size_t)H * 3 * H);
        for (int l = 0; l < L; ++l) {
            ctx->img[l].filter1 = add_image(q, K, H, K, 0, 0, 1);          // filter layer 1 [H][K], K padded
            const float* W2 = q + (size_t)H * K + H;                       // filter layer 2 [3H][H]
            const float* M1 = W2 + (size_t)3 * H * H + 3 * H;              // update_mlp.0 [H][2H]
            const float* M2 = M1 + (size_t)H * 2 * H + H;                  // update_mlp.2 [3H][H]
            for (int c = 0; c < fchunks; ++c) {
                const size_t off = add_image(W2, H, 3 * H, H, c * 128, 0, fkblk);
                if (c == 0) ctx->img[l].filter = off;
            }
            if (H == 128) {
                for (int r = 0; r < H; ++r)
                    for (int c = 0; c < 2 * H; ++c) M1T[(size_t)c * H + r] = M1[(size_t)r * 2 * H + c];
                for (int r = 0; r < 3 * H; ++r)
                    for (int c = 0; c < H; ++c) M2T[(size_t)c * 3 * H + r] = M2[(size_t)r * H + c];
                ctx->img[l].upd_f1 = add_image(M1, 2 * H, H, 2 * H, 0, 0, 2);          // [ks][c]: W = M1 [H][2H]
                add_image(M1, 2 * H, H, 2 * H, 0, 128, 2);
                ctx->img[l].upd_f2 = add_image(M2, H, 3 * H, H, 0, 0, 2);              // W = M2 [3H][H]
                add_image(M2, H, 3 * H, H, 128, 0, 2); add_image(M2, H, 3 * H, H, 256, 0, 2);
                ctx->img[l].upd_b1 = add_image(M2T.data(), 3 * H, H, 3 * H, 0, 0, 2);  // W = M2^T [H][3H]
                add_image(M2T.data(), 3 * H, H, 3 * H, 0, 128, 2); add_image(M2T.data(), 3 * H, H, 3 * H, 0, 256, 2);
                ctx->img[l].upd_b2 = add_image(M1T.data(), H, 2 * H, H, 0, 0, 2);      // W = M1^T [2H][H]
                add_image(M1T.data(), H, 2 * H, H, 128, 0, 2);
            }
            q += per_layer;
        }
        e = cudaMalloc(&ctx->w2_images_d, img.size() * sizeof(uint16_t));
        if (e != cudaSuccess) return bail(MLFFD_ENOMEM, cudaGetErrorString(e));
        e = cudaMemcpy(ctx->w2_images_d, img.data(), img.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(MLFFD_ECUDA, cudaGetErrorString(e));
        TC_DISPATCH(tc_mode,
            if (H == 128) e = cudaFuncSetAttribute(filter_table_umma_kernel<128, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UmmaGeom<128>::total());
            else if (H == 64) e = cudaFuncSetAttribute(filter_table_umma_kernel<64, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UmmaGeom<64>::total());
            else e = cudaFuncSetAttribute(filter_table_umma_kernel<32, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UmmaGeom<32>::total()));
        if (e != cudaSuccess) return bail(MLFFD_ECUDA, cudaGetErrorString(e));
        // FP16 range guard for the pre-scaled weight images: a matrix with max|w| * 2^8 beyond the FP16
        // range cannot be split; such a model runs its dense layers on the FP32 FFMA kernels instead
        float wmax = 0.f;
        {
            const float* q2 = weights_host + (size_t)(config->max_z + 1) * H + 2 * K;
            for (size_t i = 0; i < (size_t)L * per_layer; ++i) wmax = std::max(wmax, std::fabs(q2[i]));
        }
        const bool weights_fit = tc_mode == kTcBF16 || (wmax * kWeightScale < 65000.0f && std::isfinite(wmax));
        ctx->use_umma_filter = weights_fit && K <= kUmmaMaxRbf;   // first layer runs as one 32-wide K block
        if (H == 128) {
#define SET_ROWS_ATTR(OP)                                                                            \
        TC_DISPATCH(tc_mode, e = cudaFuncSetAttribute(umma_rows_kernel<OP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)UmmaRowsSmem::TOTAL));                                         \
        if (e != cudaSuccess) return bail(MLFFD_ECUDA, cudaGetErrorString(e));
        SET_ROWS_ATTR(UpdateFwd1Op) SET_ROWS_ATTR(UpdateFwd2Op<false>) SET_ROWS_ATTR(UpdateFwd2Op<true>)
        SET_ROWS_ATTR(UpdateBwd1Op<false>) SET_ROWS_ATTR(UpdateBwd1Op<true>)
        SET_ROWS_ATTR(UpdateBwd2Op<false>) SET_ROWS_ATTR(UpdateBwd2Op<true>)
#undef SET_ROWS_ATTR
            ctx->use_umma = weights_fit;
        }
    }
    const float* W = ctx->weights_d;
    ctx->emb = W + o_emb; ctx->centers = W + o_centers; ctx->gammas = W + o_gammas;
    for (int l = 0; l < L; ++l) {
        ctx->layer[l].filter = FilterWeights{W + lo[l].W1t, W + lo[l].b1, W + lo[l].W2t, W + lo[l].b2};
        ctx->layer[l].update = UpdateWeights{W + lo[l].M1t, W + lo[l].m1, W + lo[l].M2t, W + lo[l].m2,
                                             W + lo[l].M1, W + lo[l].M2, W + lo[l].U};
    }
    ctx->head = HeadWeights{W + o_A1t, W + o_a1, W + o_A2t, W + o_a2, W + o_A3, W + o_a3, W + o_A1, W + o_A2};
    int rc = (H == 128) ? set_kernel_attributes<128>(ctx)
           : (H == 64)  ? set_kernel_attributes<64>(ctx)
                        : set_kernel_attributes<32>(ctx);
    if (rc != MLFFD_OK) return bail(rc, ctx->err);
    *out = ctx;
    return MLFFD_OK;
}

extern "C" void mlffd_model_destroy(mlffd_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard device_guard(ctx->device);
    for (cudaEvent_t ev : ctx->event_pool) cudaEventDestroy(ev);
    if (ctx->ws.arena) cudaFree(ctx->ws.arena);
    if (ctx->weights_d) cudaFree(ctx->weights_d);
    if (ctx->w2_images_d) cudaFree(ctx->w2_images_d);
    if (ctx->spline_d) cudaFree(ctx->spline_d);
    if (ctx->status_d) cudaFree(ctx->status_d);
    delete ctx;
}

extern "C" int mlffd_workspace_reserve(mlffd_ctx* ctx, int64_t max_atoms, int64_t max_edges,
                                       int64_t max_structures) {
    if (!ctx) return MLFFD_EINVAL;
    Workspace& ws = ctx->ws;
    if (max_atoms <= ws.cap_atoms && max_edges <= ws.req_edges && max_structures <= ws.cap_structs)
        return MLFFD_OK;
    max_atoms = std::max<int64_t>(std::max(max_atoms, ws.cap_atoms), 1);
    max_edges = std::max<int64_t>(std::max(max_edges, ws.req_edges), 2);
    const int64_t requested_edges = max_edges;
    if (ctx->skin > 0.f) {   // room for the candidate list: all pairs within cutoff + skin
        const double r = (ctx->cfg.cutoff + ctx->skin) / ctx->cfg.cutoff;
        max_edges = (int64_t)(max_edges * r * r * r * 1.15) + 64;
    }
    max_structures = std::max<int64_t>(std::max(max_structures, ws.cap_structs), 1);
    if (max_atoms >= (1ll << 30) || max_edges >= (1ll << 31) - 64)
        return fail(ctx, MLFFD_EINVAL, "workspace request exceeds 32-bit indexing");
    DeviceGuard device_guard(ctx->device);
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    const int64_t N = max_atoms, E = max_edges, P = max_edges / 2 + 1;
    const int H = ctx->H, L = ctx->L;
    size_t cub_bytes = 0;
    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (int*)nullptr, (int*)nullptr,
                                                (int)(N + 1)));
    ArenaPlan plan;
    const size_t o_atom_struct = plan.take(sizeof(int) * N);
    const size_t o_deg = plan.take(sizeof(int) * (N + 1));
    const size_t o_deg_low = plan.take(sizeof(int) * (N + 1));
    const size_t o_rowptr = plan.take(sizeof(int) * (N + 1));
    const size_t o_lowptr = plan.take(sizeof(int) * (N + 1));
    const size_t o_col = plan.take(sizeof(int) * E);
    const size_t o_edge_dst = plan.take(sizeof(int) * E);
    const size_t o_rev = plan.take(sizeof(int) * E);
    const size_t o_pair = plan.take(sizeof(int) * E);
    const size_t o_geo = plan.take(sizeof(float4) * E);
    // per-layer (spline mode: per-layer, per-slice) slabs + debug sum
    const size_t o_adj = plan.take(sizeof(float4) * E * ((size_t)L * ctx->adj_slabs_per_layer + 1));
    const size_t o_erec = plan.take(ctx->spline ? sizeof(float4) * 4 * E : 0);
    const bool skin_on = ctx->skin > 0.f;
    const size_t o_skin = plan.take(skin_on ? sizeof(SkinState) : 0);
    const size_t o_cand_deg = plan.take(skin_on ? sizeof(int) * (N + 1) : 0);
    const size_t o_cand_rowptr = plan.take(skin_on ? sizeof(int) * (N + 1) : 0);
    const size_t o_cand_col = plan.take(skin_on ? sizeof(int) * E : 0);
    const size_t o_pos_ref = plan.take(skin_on ? sizeof(float) * 3 * N : 0);
    const int64_t PT = ctx->spline ? 0 : P;   // the per-step filter tables exist in MLFFD_FILTER_TABLE mode only
    const size_t o_pdist = plan.take(sizeof(float) * P);
    const size_t o_eps = plan.take(sizeof(float) * N);
    const size_t o_virial = plan.take(sizeof(double) * 9 * max_structures);
    const int64_t C = 2 * N + 8 * max_structures + 2;   // cell capacity (grid_setup caps cells per structure)
    {
        size_t cub_cells = 0;
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, cub_cells, (int*)nullptr, (int*)nullptr, (int)(C + 1)));
        cub_bytes = std::max(cub_bytes, cub_cells);
    }
    const size_t o_cub = plan.take(cub_bytes);
    const size_t o_grids = plan.take(sizeof(GridInfo) * max_structures);
    const size_t o_ncells = plan.take(sizeof(int) * (max_structures + 1));
    const size_t o_atom_cell = plan.take(sizeof(int) * N);
    const size_t o_cell_count = plan.take(sizeof(int) * (C + 1));
    const size_t o_cell_start = plan.take(sizeof(int) * (C + 1));
    const size_t o_cell_cursor = plan.take(sizeof(int) * (C + 1));
    const size_t o_cell_atoms = plan.take(sizeof(int) * N);
    size_t o_filt[kMaxLayers], o_dfilt[kMaxLayers], o_s_in[kMaxLayers + 1], o_v_in[kMaxLayers],
        o_s_msg[kMaxLayers], o_v_msg[kMaxLayers], o_y1[kMaxLayers], o_gates[kMaxLayers],
        o_sbar[kMaxLayers], o_vbar[kMaxLayers];
    const int adj_sets = ctx->debug_keep ? L : std::min(L, 2);
    for (int l = 0; l < L; ++l) {
        o_filt[l] = plan.take(sizeof(float) * PT * 3 * H);
        o_dfilt[l] = plan.take(sizeof(float) * PT * 3 * H);
        o_s_in[l] = plan.take(sizeof(float) * N * H);
        o_v_in[l] = (l > 0) ? plan.take(sizeof(float) * N * 3 * H) : 0;
        o_s_msg[l] = plan.take(sizeof(float) * N * H);
        o_v_msg[l] = plan.take(sizeof(float) * N * 3 * H);
        o_y1[l] = plan.take(sizeof(float) * N * H);
        o_gates[l] = (l < L - 1) ? plan.take(sizeof(float) * N * 2 * H) : 0;
        if (l < adj_sets) {
            o_sbar[l] = plan.take(sizeof(float) * N * H);
            o_vbar[l] = plan.take(sizeof(float) * N * 3 * H);
        }
    }
    o_s_in[L] = plan.take(sizeof(float) * N * H);
    if (ws.arena) { cudaFree(ws.arena); ws = Workspace(); }
    ctx->last_adj_slabs = 0;
    void* arena = nullptr;
    cudaError_t e = cudaMalloc(&arena, plan.total);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MLFFD_ENOMEM, "cudaMalloc of " + std::to_string(plan.total) +
                                           " bytes failed: " + cudaGetErrorString(e));
    }
    char* base = (char*)arena;
    ws.arena = arena; ws.arena_bytes = plan.total;
    ws.cap_atoms = N; ws.cap_edges = E; ws.cap_structs = max_structures; ws.cap_pairs = P;
    ws.req_edges = requested_edges;
    ws.atom_struct = (int*)(base + o_atom_struct);
    ws.deg = (int*)(base + o_deg); ws.deg_low = (int*)(base + o_deg_low);
    ws.rowptr = (int*)(base + o_rowptr); ws.lowptr = (int*)(base + o_lowptr);
    ws.col = (int*)(base + o_col); ws.edge_dst = (int*)(base + o_edge_dst);
    ws.rev = (int*)(base + o_rev); ws.pair = (int*)(base + o_pair);
    ws.geo = (float4*)(base + o_geo); ws.edge_adj = (float4*)(base + o_adj);
    ws.erec = (float4*)(base + o_erec);
    if (skin_on) {
        ws.skin = (SkinState*)(base + o_skin);
        ws.cand_deg = (int*)(base + o_cand_deg); ws.cand_rowptr = (int*)(base + o_cand_rowptr);
        ws.cand_col = (int*)(base + o_cand_col); ws.pos_ref = (float*)(base + o_pos_ref);
        CUDA_TRY(ctx, cudaMemset(ws.skin, 0, sizeof(SkinState)));          // no valid candidate list yet
        CUDA_TRY(ctx, cudaMemset(ws.cand_deg, 0, sizeof(int) * (N + 1)));
        ctx->skin_atoms = -1;
    }
    ws.pair_dist = (float*)(base + o_pdist);
    ws.eps = (float*)(base + o_eps);
    ws.virial64 = (double*)(base + o_virial);
    ws.cub_temp = base + o_cub; ws.cub_bytes = cub_bytes;
    ws.grids = (GridInfo*)(base + o_grids);
    ws.ncells = (int*)(base + o_ncells); ws.total_cells = ws.ncells + max_structures;
    ws.atom_cell = (int*)(base + o_atom_cell);
    ws.cell_count = (int*)(base + o_cell_count); ws.cell_start = (int*)(base + o_cell_start);
    ws.cell_cursor = (int*)(base + o_cell_cursor); ws.cell_atoms = (int*)(base + o_cell_atoms);
    ws.cap_cells = C;
    for (int l = 0; l < L; ++l) {
        ws.filt[l] = (float*)(base + o_filt[l]); ws.dfilt[l] = (float*)(base + o_dfilt[l]);
        ws.s_in[l] = (float*)(base + o_s_in[l]);
        ws.v_in[l] = (l > 0) ? (float*)(base + o_v_in[l]) : nullptr;
        ws.s_msg[l] = (float*)(base + o_s_msg[l]); ws.v_msg[l] = (float*)(base + o_v_msg[l]);
        ws.y1[l] = (float*)(base + o_y1[l]);
        ws.gates[l] = (l < L - 1) ? (float*)(base + o_gates[l]) : nullptr;
        if (l < adj_sets) {
            ws.sbar[l] = (float*)(base + o_sbar[l]);
            ws.vbar[l] = (float*)(base + o_vbar[l]);
        }
    }
    ws.s_in[L] = (float*)(base + o_s_in[L]);
    return MLFFD_OK;
}

extern "C" int mlffd_neighbor_list(mlffd_ctx* ctx, const float* pos_d, const int32_t* offsets_d,
                                   int32_t num_structures, int64_t num_atoms, const float* cells_d,
                                   const uint8_t* pbc_d, void* stream) {
    if (!ctx) return MLFFD_EINVAL;
    if (!pos_d || !offsets_d || num_structures < 1 || num_atoms < 1)
        return fail(ctx, MLFFD_EINVAL, "mlffd_neighbor_list: bad argument");
    DeviceGuard device_guard(ctx->device);
    return build_neighbors(ctx, pos_d, offsets_d, num_structures, num_atoms, cells_d, pbc_d,
                           (cudaStream_t)stream);
}

extern "C" int mlffd_export_edges(mlffd_ctx* ctx, int64_t* edge_index_d, int64_t capacity_edges,
                                  int64_t* num_edges_out, void* stream) {
    if (!ctx || !edge_index_d) return MLFFD_EINVAL;
    mlffd_status s;
    int rc = mlffd_get_status(ctx, &s);
    if (rc) return rc;
    if (num_edges_out) *num_edges_out = s.num_edges;
    if (s.overflow) return fail(ctx, MLFFD_ECAPACITY, "edge capacity exceeded in the last build");
    if (s.num_edges > capacity_edges)
        return fail(ctx, MLFFD_ECAPACITY, "edge_index buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (s.num_edges > 0) {
        export_edges_kernel<<<clamp_grid(ceil_div(s.num_edges, 256), kNumSMs * 8), 256, 0, st>>>(
            ctx->ws.col, ctx->ws.edge_dst, (int)s.num_edges, (long long)capacity_edges,
            (long long*)edge_index_d);
        LAUNCH_CHECK(ctx, "export_edges_kernel");
    }
    return MLFFD_OK;
}

extern "C" int mlffd_energy_forces(mlffd_ctx* ctx, const int32_t* z_d, const float* pos_d,
                                   const int32_t* offsets_d, int32_t num_structures,
                                   int64_t num_atoms, const float* cells_d, const uint8_t* pbc_d,
                                   float* energy_d, float* forces_d, void* stream) {
    if (!ctx) return MLFFD_EINVAL;
    if (!z_d || !pos_d || !offsets_d || !energy_d || num_structures < 1 || num_atoms < 1)
        return fail(ctx, MLFFD_EINVAL, "mlffd_energy_forces: bad argument");
    DeviceGuard device_guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = build_neighbors(ctx, pos_d, offsets_d, num_structures, num_atoms, cells_d, pbc_d, st);
    if (rc) return rc;
    ctx->last_had_forces = forces_d != nullptr;
    switch (ctx->H) {
        case 128: return run_model<128>(ctx, z_d, num_atoms, offsets_d, num_structures, energy_d, forces_d, st);
        case 64:  return run_model<64>(ctx, z_d, num_atoms, offsets_d, num_structures, energy_d, forces_d, st);
        default:  return run_model<32>(ctx, z_d, num_atoms, offsets_d, num_structures, energy_d, forces_d, st);
    }
}

extern "C" int mlffd_get_status(mlffd_ctx* ctx, mlffd_status* out) {
    if (!ctx || !out) return MLFFD_EINVAL;
    DeviceGuard device_guard(ctx->device);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->last_stream));
    DeviceStatus h{};
    CUDA_TRY(ctx, cudaMemcpy(&h, ctx->status_d, sizeof(h), cudaMemcpyDeviceToHost));
    out->num_atoms = ctx->last_atoms;
    out->num_edges = h.num_edges;
    out->num_pairs = h.num_pairs;
    out->edge_capacity = ctx->ws.cap_edges;
    out->overflow = h.overflow;
    out->max_degree = h.max_degree;
    out->overflow_events = h.overflow_events;
    out->tc_saturated = h.tc_saturated;
    out->skin_rebuilds = 0;
    if (ctx->ws.skin != nullptr) {
        SkinState sk{};
        CUDA_TRY(ctx, cudaMemcpy(&sk, ctx->ws.skin, sizeof(sk), cudaMemcpyDeviceToHost));
        out->skin_rebuilds = sk.rebuilds;
    }
    return MLFFD_OK;
}

extern "C" int mlffd_status_async(mlffd_ctx* ctx, int32_t* status_out, void* stream) {
    if (!ctx || !status_out) return MLFFD_EINVAL;
    static_assert(sizeof(DeviceStatus) == 6 * sizeof(int32_t), "mlffd_status_async copies six int32 words");
    DeviceGuard device_guard(ctx->device);
    CUDA_TRY(ctx, cudaMemcpyAsync(status_out, ctx->status_d, sizeof(DeviceStatus), cudaMemcpyDefault,
                                  (cudaStream_t)stream));
    return MLFFD_OK;
}

extern "C" int mlffd_filter_table(mlffd_ctx* ctx, int32_t layer, const float* dist_d,
                                  int64_t num_pairs, float* filter_d, float* dfilter_d,
                                  void* stream) {
    if (!ctx) return MLFFD_EINVAL;
    if (!dist_d || !filter_d || !dfilter_d || layer < 0 || layer >= ctx->L || num_pairs < 0 ||
        num_pairs >= (1ll << 31) / (3 * ctx->H))
        return fail(ctx, MLFFD_EINVAL, "mlffd_filter_table: bad argument");
    if (num_pairs == 0) return MLFFD_OK;
    DeviceGuard device_guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    switch (ctx->H) {
        case 128: return launch_filter<128>(ctx, layer, dist_d, nullptr, (int)num_pairs, nullptr, filter_d, dfilter_d, num_pairs, st);
        case 64:  return launch_filter<64>(ctx, layer, dist_d, nullptr, (int)num_pairs, nullptr, filter_d, dfilter_d, num_pairs, st);
        default:  return launch_filter<32>(ctx, layer, dist_d, nullptr, (int)num_pairs, nullptr, filter_d, dfilter_d, num_pairs, st);
    }
}

extern "C" int mlffd_edge_features(const float* pos_d, const int64_t* edge_index_d, int64_t num_edges, float eps,
                                   int32_t eps_mode, float* edge_vec_d, float* dist_d, float* unit_d, void* stream) {
    if (num_edges == 0) return MLFFD_OK;
    if (!pos_d || !edge_index_d || !edge_vec_d || !dist_d || !unit_d || num_edges < 0 || (eps_mode != 0 && eps_mode != 1))
        return MLFFD_EINVAL;
    edge_features_kernel<<<clamp_grid(ceil_div(num_edges, kEdgeFeatThreads), kNumSMs * 8), kEdgeFeatThreads, 0, (cudaStream_t)stream>>>(
        pos_d, (const long long*)edge_index_d, (const long long*)edge_index_d + num_edges, (long long)num_edges, eps, eps_mode,
        edge_vec_d, dist_d, unit_d);
    return cudaGetLastError() == cudaSuccess ? MLFFD_OK : MLFFD_ECUDA;
}

extern "C" int mlffd_rbf_cutoff(const float* dist_d, int64_t num_edges, const float* centers_d, int32_t num_rbf,
                                float gamma, float cutoff, float* out_d, void* stream) {
    if (num_edges == 0) return MLFFD_OK;
    if (!dist_d || !centers_d || !out_d || num_edges < 0 || num_rbf < 1 || !(cutoff > 0.f)) return MLFFD_EINVAL;
    rbf_cutoff_kernel<<<clamp_grid(ceil_div(num_edges * num_rbf, 256), kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(
        dist_d, (long long)num_edges, centers_d, num_rbf, gamma, cutoff, out_d);
    return cudaGetLastError() == cudaSuccess ? MLFFD_OK : MLFFD_ECUDA;
}

extern "C" int mlffd_filter_spline(mlffd_ctx* ctx, int32_t layer, const float* dist_d, int64_t num_pairs,
                                   float* filter_d, float* dfilter_d, void* stream) {
    if (!ctx) return MLFFD_EINVAL;
    if (!dist_d || !filter_d || !dfilter_d || layer < 0 || layer >= ctx->L || num_pairs < 0)
        return fail(ctx, MLFFD_EINVAL, "mlffd_filter_spline: bad argument");
    if (num_pairs == 0) return MLFFD_OK;
    DeviceGuard device_guard(ctx->device);
    spline_filter_eval_kernel<<<clamp_grid(ceil_div(num_pairs * 3 * ctx->H, 256), kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(
        spline_layer_table(ctx, layer), ctx->H, ctx->spline_inv_h, dist_d, (long long)num_pairs, filter_d, dfilter_d);
    LAUNCH_CHECK(ctx, "spline_filter_eval_kernel");
    return MLFFD_OK;
}

extern "C" int mlffd_virial(mlffd_ctx* ctx, const int32_t* offsets_d, int32_t num_structures,
                            float* virial_d, void* stream) {
    if (!ctx) return MLFFD_EINVAL;
    if (!offsets_d || !virial_d || num_structures < 1)
        return fail(ctx, MLFFD_EINVAL, "mlffd_virial: bad argument");
    if (ctx->last_adj_slabs == 0 || num_structures != ctx->last_structs)
        return fail(ctx, MLFFD_EINVAL, "mlffd_virial: call mlffd_energy_forces with forces first (same structures)");
    DeviceGuard device_guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    Workspace& ws = ctx->ws;
    // few structures with many edges each (a periodic box) -> several chunk blocks per structure
    const int64_t per_struct = std::max<int64_t>(1, ws.cap_edges / std::max<int64_t>(1, ctx->last_structs));
    const int chunks = (int)std::min<int64_t>(64, std::max<int64_t>(1, per_struct / kVirialChunk));
    if (chunks > 1) CUDA_TRY(ctx, cudaMemsetAsync(ws.virial64, 0, sizeof(double) * 9 * num_structures, st));
    virial_kernel<<<dim3(num_structures, chunks), 256, 0, st>>>(offsets_d, num_structures, ws.rowptr, ws.rev, ws.geo, ws.edge_adj,
                                                                ctx->last_adj_slabs, ctx->last_adj_pairs, ctx->last_adj_direct, (size_t)ws.cap_edges,
                                                                ws.virial64, ctx->status_d);
    virial_finalize_kernel<<<ceil_div(9 * num_structures, 256), 256, 0, st>>>(ws.virial64, 9 * num_structures,
                                                                               virial_d, ctx->status_d);
    CUDA_TRY(ctx, cudaGetLastError());
    return MLFFD_OK;
}

extern "C" int mlffd_debug_buffer(mlffd_ctx* ctx, const char* name, int32_t layer, void** ptr_out,
                                  int64_t* count_out, int32_t* elem_size_out) {
    if (!ctx || !name || !ptr_out || !count_out || !elem_size_out) return MLFFD_EINVAL;
    mlffd_status s;
    int rc = mlffd_get_status(ctx, &s);
    if (rc) return rc;
    const Workspace& ws = ctx->ws;
    const int64_t N = ctx->last_atoms, E = s.overflow ? 0 : s.num_edges, P = s.overflow ? 0 : s.num_pairs;
    const int H = ctx->H, L = ctx->L;
    const std::string n(name);
    const bool layer_ok = layer >= 0 && layer < L;
    void* p = nullptr; int64_t cnt = 0; int es = 4;
    auto adj = [&](int l) { return ctx->debug_keep ? l : (l & 1); };
    if (n == "rowptr") { p = ws.rowptr; cnt = N + 1; }
    else if (n == "col") { p = ws.col; cnt = E; }
    else if (n == "rev") { p = ws.rev; cnt = E; }
    else if (n == "pair") { p = ws.pair; cnt = E; }
    else if (n == "edge_dst") { p = ws.edge_dst; cnt = E; }
    else if (n == "geo") { p = ws.geo; cnt = E; es = 16; }
    else if (n == "edge_adj") { p = ws.edge_adj + (size_t)ctx->L * ctx->adj_slabs_per_layer * ws.cap_edges; cnt = E; es = 16; }
    else if (n == "pair_dist") { p = ws.pair_dist; cnt = P; }
    else if (n == "atom_energy") { p = ws.eps; cnt = N; }
    else if (n == "s_out") { p = ws.s_in[L]; cnt = N * H; }
    else if (!layer_ok) return fail(ctx, MLFFD_EINVAL, "mlffd_debug_buffer: bad layer");
    else if ((n == "filter" || n == "dfilter") && ctx->spline)
        return fail(ctx, MLFFD_EINVAL, "mlffd_debug_buffer: no filter table in MLFFD_FILTER_SPLINE mode");
    else if (n == "filter") { p = ws.filt[layer]; cnt = P * 3 * H; }
    else if (n == "dfilter") { p = ws.dfilt[layer]; cnt = P * 3 * H; }
    else if (n == "s_in") { p = ws.s_in[layer]; cnt = N * H; }
    else if (n == "v_in") { p = ws.v_in[layer]; cnt = ws.v_in[layer] ? N * 3 * H : 0; }
    else if (n == "s_msg") { p = ws.s_msg[layer]; cnt = N * H; }
    else if (n == "v_msg") { p = ws.v_msg[layer]; cnt = N * 3 * H; }
    else if (n == "y1") { p = ws.y1[layer]; cnt = N * H; }
    else if (n == "gates") { p = ws.gates[layer]; cnt = ws.gates[layer] ? N * 2 * H : 0; }
    else if (n == "sbar") { p = ws.sbar[adj(layer)]; cnt = N * H; }
    else if (n == "vbar") { p = ws.vbar[adj(layer)]; cnt = N * 3 * H; }
    else return fail(ctx, MLFFD_EINVAL, "mlffd_debug_buffer: unknown buffer '" + n + "'");
    *ptr_out = p; *count_out = cnt; *elem_size_out = es;
    return MLFFD_OK;
}

extern "C" int mlffd_profile_enable(mlffd_ctx* ctx, int32_t enable) {
    if (!ctx) return MLFFD_EINVAL;
    DeviceGuard device_guard(ctx->device);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->last_stream));
    ctx->marks.clear();
    ctx->events_used = 0;
    ctx->profiling = enable != 0;
    ctx->launches = 0;
    for (int i = 0; i < MLFFD_NUM_STAGES; ++i) { ctx->stage_ms[i] = 0.0; ctx->stage_launches[i] = 0; }
    return MLFFD_OK;
}

extern "C" int mlffd_profile_read(mlffd_ctx* ctx, mlffd_profile* out) {
    if (!ctx || !out) return MLFFD_EINVAL;
    DeviceGuard device_guard(ctx->device);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->last_stream));
    drain_marks(ctx);
    out->launches = ctx->launches;
    for (int i = 0; i < MLFFD_NUM_STAGES; ++i) {
        out->stage_ms[i] = ctx->stage_ms[i];
        out->stage_launches[i] = ctx->stage_launches[i];
    }
    return MLFFD_OK;
}

extern "C" const char* mlffd_stage_name(int32_t stage) {
    static const char* names[MLFFD_NUM_STAGES] = {
        "neighbor", "embedding", "filter", "message_fwd", "update_fwd", "readout", "energy_sum",
        "update_bwd", "message_bwd", "force"};
    return (stage >= 0 && stage < MLFFD_NUM_STAGES) ? names[stage] : "";
}

// ---- on-device velocity Verlet (stateless helpers; any stream) -------------------------------
extern "C" int mlffd_md_kick_drift(const mlffd_ctx* guard, int64_t num_atoms, double* pos_d, double* vel_d,
                                   const float* forces_d, const double* inv_mass_d, double dt,
                                   float* pos32_d, void* stream) {
    if (num_atoms < 1 || !pos_d || !vel_d || !forces_d || !inv_mass_d || !pos32_d) return MLFFD_EINVAL;
    const long long n3 = 3 * (long long)num_atoms;
    md_kick_drift_kernel<<<clamp_grid(ceil_div(n3, 256), kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(
        n3, pos_d, vel_d, forces_d, inv_mass_d, dt, pos32_d, guard ? guard->status_d : nullptr);
    return cudaGetLastError() == cudaSuccess ? MLFFD_OK : MLFFD_ECUDA;
}

extern "C" int mlffd_md_kick_energy(const mlffd_ctx* guard, int64_t num_atoms, double* vel_d, const float* forces_d,
                                    const double* inv_mass_d, double dt, const float* energy_d,
                                    int32_t num_structures, double* series_d, int32_t* counter_d,
                                    int32_t capacity, void* stream) {
    if (num_atoms < 1 || !vel_d || !forces_d || !inv_mass_d || !energy_d || !series_d || !counter_d)
        return MLFFD_EINVAL;
    const long long n3 = 3 * (long long)num_atoms;
    cudaStream_t st = (cudaStream_t)stream;
    const DeviceStatus* g = guard ? guard->status_d : nullptr;
    if (n3 <= 8192) {   // small system: kick and energy bookkeeping in one single-block launch
        md_energy_kernel<true><<<1, 1024, 0, st>>>(n3, vel_d, forces_d, inv_mass_d, dt, energy_d, num_structures,
                                                   series_d, counter_d, capacity, g);
    } else {
        md_kick_kernel<<<clamp_grid(ceil_div(n3, 256), kNumSMs * 8), 256, 0, st>>>(n3, vel_d, forces_d, inv_mass_d, dt, g);
        md_energy_kernel<false><<<1, 1024, 0, st>>>(n3, vel_d, forces_d, inv_mass_d, dt, energy_d, num_structures,
                                                    series_d, counter_d, capacity, g);
    }
    return cudaGetLastError() == cudaSuccess ? MLFFD_OK : MLFFD_ECUDA;
}

extern "C" int mlffd_set_skin(mlffd_ctx* ctx, float skin) {
    if (!ctx || !(skin >= 0.f) || skin > 4.0f * ctx->cfg.cutoff) return ctx ? fail(ctx, MLFFD_EINVAL, "mlffd_set_skin: bad skin") : MLFFD_EINVAL;
    if (skin == ctx->skin) return MLFFD_OK;
    DeviceGuard device_guard(ctx->device);
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    ctx->skin = skin;
    ctx->skin_atoms = -1;
    if (ctx->ws.arena) { cudaFree(ctx->ws.arena); ctx->ws = Workspace(); }   // the next mlffd_workspace_reserve sizes the candidate buffers
    ctx->last_adj_slabs = 0;
    return MLFFD_OK;
}

extern "C" int mlffd_set_dense_fallback(mlffd_ctx* ctx, int32_t enable) {
    if (!ctx) return MLFFD_EINVAL;
    ctx->dense_fallback = enable != 0;
    return MLFFD_OK;
}
