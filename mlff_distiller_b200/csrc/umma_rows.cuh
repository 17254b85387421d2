// Generic "rows x dense layer" kernel on tcgen05 tensor cores (H = 128), used by the PaiNN update
// block forward and reverse.  Same arithmetic contract as filter_umma.cuh: both operands are split
// into two FP16 terms, three products (W_hi x_hi + W_hi x_lo + W_lo x_hi) accumulate in FP32 in
// tensor memory -> FP32-equivalent results.
//
//   D[channel][row] = sum_k W[channel][k] * x[row][k]      (transposed product: a TMEM lane is an
//                                                           output channel, a column is a row)
// Rows come in tiles of 128 (atoms); K is walked in super-blocks of 128 (KS of them) and the
// output in chunks of 128 channels (NC of them), all NC accumulators (<= 384 columns) stay in
// TMEM across the K loop.  An `Op` supplies
//   produce(row, n, ks, k0) -> float4 : 4 consecutive inputs of super-block ks for one row
//   store(chunk, lane_ch, row0, r, prev) : epilogue for 32 rows x 1 channel per thread
// so norms, SiLU, gates, the 3x3 spatial mixing and residuals are fused around the GEMM and the
// only HBM traffic is the per-atom feature rows.
//
// Roles: 16 compute/epilogue warps + 1 issuer warp (24 MMAs per weight image) + 1 loader warp
// (bulk-copies the pre-swizzled 64 KB weight images [ks][chunk] through a 2-deep ring; image g is
// requested the moment the MMAs of image g-2 have completed -- the issuer itself blocks in
// tcgen05.mma issue while the tensor-core queue is full, so a refill issued from it starts late).
#pragma once
#include "filter_umma.cuh"

namespace mlffd {

struct UmmaRowsSmem {
    static constexpr uint32_t A_HI = 0;
    static constexpr uint32_t A_LO = 32768;
    static constexpr uint32_t B0 = 65536;
    static constexpr uint32_t B1 = 131072;
    static constexpr uint32_t BARS = 196608;   // a_full, act_free, b_full[2], b_free[2], d_full[3], tmem base
    static constexpr uint32_t TOTAL = BARS + 128 + 1024 /*align*/;
};

template <class Op, int MODE = kTcSplit>
__global__ void __launch_bounds__(kFilterUmmaThreads, 1)
umma_rows_kernel(Op op, int num_rows, const uint8_t* __restrict__ images,
                 const DeviceStatus* __restrict__ status) {
    constexpr int KS = Op::KS, NC = Op::NC;
    constexpr bool single = MODE != kTcSplit;        // one product of the `hi` halves (filter_umma.cuh)
    constexpr bool SPLIT = KS > 1 && !single;        // separate TMEM accumulator for the correction products
    constexpr int kImageKBlocks = single ? 2 : 4;    // K blocks of a weight image in use: hi (| lo)
    static_assert(!SPLIT || 2 * NC * 128 <= (int)kTmemCols, "not enough tensor memory columns");
    if (status != nullptr && status->overflow) return;
    const int num_tiles = (num_rows + 127) / 128;
    if ((int)blockIdx.x >= num_tiles) return;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = (uint64_t*)(smem + UmmaRowsSmem::BARS);
    uint64_t* bar_a_full = bars;
    uint64_t* bar_act_free = bars + 1;
    uint64_t* bar_b_full = bars + 2;
    uint64_t* bar_b_free = bars + 4;
    uint64_t* bar_d_full = bars + 6;
    uint32_t* tmem_base_s = (uint32_t*)(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    constexpr int PER_TILE = KS * NC;
    if (warp == kUmmaComputeWarps) {
        if (lane == 0) {
            for (int i = 0; i < 9; ++i) mbar_init(bars + i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // start streaming the first two weight images before anything else (latency path)
            const int total0 = my_tiles * PER_TILE;
            for (int g = 0; g < 2 && g < total0; ++g) {
                mbar_expect_tx(&bar_b_full[g], kImageKBlocks * kKBlockBytes);
                const uint8_t* src = images + (size_t)(g % PER_TILE) * kChunkImageBytes;
                const uint32_t dst = smem_u32(smem + (g ? UmmaRowsSmem::B1 : UmmaRowsSmem::B0));
                for (int qq = 0; qq < kImageKBlocks; ++qq)
                    bulk_g2s(dst + qq * kKBlockBytes, src + qq * kKBlockBytes, kKBlockBytes, &bar_b_full[g]);
            }
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == kUmmaComputeWarps) {
        // =============================== issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_128x128(MODE);
            const uint32_t b_buf[2] = {smem_u32(smem + UmmaRowsSmem::B0), smem_u32(smem + UmmaRowsSmem::B1)};
            const uint64_t act_desc_hi = umma_desc_sw128(smem_u32(smem + UmmaRowsSmem::A_HI));
            const uint64_t act_desc_lo = umma_desc_sw128(smem_u32(smem + UmmaRowsSmem::A_LO));
            const uint64_t w_desc0 = umma_desc_sw128(b_buf[0]);
            int g = 0;   // images 0 and 1 were requested during set-up
            for (int it = 0; it < my_tiles; ++it) {
                for (int ks = 0; ks < KS; ++ks) {
                    const int pc = it * KS + ks;
                    mbar_wait(bar_a_full, pc & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int c = 0; c < NC; ++c, ++g) {
                        const int buf = g & 1;
                        mbar_wait(&bar_b_full[buf], (g >> 1) & 1);
                        const uint32_t d_main = tmem_base + (uint32_t)c * 128;
                        // Truncating accumulation biases long sums into a large accumulator, so
                        // the small correction products (W_hi x_lo, W_lo x_hi) are kept apart:
                        // issued first when there is one K super-block, or accumulated in their
                        // own TMEM columns (added in the epilogue) when K spans several.
                        const uint32_t d_corr = SPLIT ? tmem_base + (uint32_t)(NC + c) * 128 : d_main;
                        const uint64_t w_desc = w_desc0 + (uint64_t)(buf ? (kChunkImageBytes >> 4) : 0);
                        uint32_t acc_corr = (ks > 0) ? 1u : 0u;
#pragma unroll
                        for (int pass = single ? 2 : 0; pass < 2; ++pass) {   // W_hi*x_lo, W_lo*x_hi
                            const uint64_t act_base = (pass == 0) ? act_desc_lo : act_desc_hi;
                            const uint64_t w_base = w_desc + (uint64_t)((pass == 1) ? ((2 * kKBlockBytes) >> 4) : 0);
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t off = (uint64_t)((kb * kKBlockBytes + k * 32) >> 4);
                                    umma_f16(d_corr, w_base + off, act_base + off, idesc, acc_corr);
                                    acc_corr = 1;
                                }
                        }
                        uint32_t acc_main = (SPLIT || single) ? ((ks > 0) ? 1u : 0u) : 1u;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {       // W_hi*x_hi
                                const uint64_t off = (uint64_t)((kb * kKBlockBytes + k * 32) >> 4);
                                umma_f16(d_main, w_desc + off, act_desc_hi + off, idesc, acc_main);
                                acc_main = 1;
                            }
                        umma_commit(&bar_b_free[buf]);
                        if (ks == KS - 1) umma_commit(&bar_d_full[c]);
                    }
                    umma_commit(bar_act_free);
                }
            }
        }
    } else if (warp == kUmmaComputeWarps + 1) {
        // ============================ weight loader ============================
        if (lane == 0) {
            const uint32_t b_buf[2] = {smem_u32(smem + UmmaRowsSmem::B0), smem_u32(smem + UmmaRowsSmem::B1)};
            const int total = my_tiles * PER_TILE;
            for (int g = 2; g < total; ++g) {   // images 0 and 1 were requested during set-up
                const int buf = g & 1;
                mbar_wait(&bar_b_free[buf], ((g - 2) >> 1) & 1);   // MMAs of image g-2 done with the slot
                mbar_expect_tx(&bar_b_full[buf], kImageKBlocks * kKBlockBytes);
                const uint8_t* src = images + (size_t)(g % PER_TILE) * kChunkImageBytes;
#pragma unroll
                for (int qq = 0; qq < kImageKBlocks; ++qq)
                    bulk_g2s(b_buf[buf] + qq * kKBlockBytes, src + qq * kKBlockBytes, kKBlockBytes, &bar_b_full[buf]);
            }
        }
    } else {
        // ========================= compute / epilogue =========================
        // Producer mapping: a warp owns 8 rows of the tile, a lane owns 4 consecutive K values, so
        // every global read of a feature row is one coalesced 512-byte request.
        const int q = warp & 3, cs = warp >> 2;               // channel quarter (lanes), row segment
        const int kb = lane >> 4;                             // K block (64 values) of this lane
        const int chunk = (lane & 15) >> 1, half8 = (lane & 1) * 8;
        bool sat = false;   // an operand left the FP16 range (filter_umma.cuh:split_saturates)
        for (int it = 0; it < my_tiles; ++it) {
            const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128;
            for (int ks = 0; ks < KS; ++ks) {
                const int pc = it * KS + ks;
                float4 x[8];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) x[rr] = op.produce(row0 + warp * 8 + rr, num_rows, ks, 4 * lane);
                if (pc > 0) mbar_wait(bar_act_free, (pc - 1) & 1);   // previous MMAs finished reading the tile
                uint8_t* a_hi = smem + UmmaRowsSmem::A_HI + kb * kKBlockBytes;
                uint8_t* a_lo = smem + UmmaRowsSmem::A_LO + kb * kKBlockBytes;
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    uint2 hi, lo;
                    split4(x[rr], MODE, hi, lo, sat);
                    const uint32_t o = sw128_offset(warp * 8 + rr, chunk) + half8;
                    *reinterpret_cast<uint2*>(a_hi + o) = hi;
                    if (!single) *reinterpret_cast<uint2*>(a_lo + o) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("bar.sync 1, 512;" ::: "memory");
                if (tid == 0) mbar_arrive(bar_a_full);
            }
            // ---- epilogue: D[chunk][channel = lane][row = column] ----
            uint32_t prev[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) prev[j] = 0u;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                mbar_wait(&bar_d_full[c], it & 1);
                __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 128 + cs * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if constexpr (SPLIT) {
                    uint32_t rc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((NC + c) * 128 + cs * 32), rc);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rc[j]));
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * kAccUnscale);
                op.template store<32>(c, q * 32 + lane, row0 + cs * 32, num_rows, r, prev);
#pragma unroll
                for (int j = 0; j < 32; ++j) prev[j] = r[j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        if (sat && status != nullptr) const_cast<DeviceStatus*>(status)->tc_saturated = 1;
    }
    __syncthreads();
    if (warp == kUmmaComputeWarps) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ================================ small systems: FFMA, many blocks ================================
// Same Ops, latency-optimised for a few hundred rows (single-trajectory MD: BASELINE configs C1 /
// C3).  The tensor-core kernel above handles a 128-row tile per CTA and streams every 64 KB weight
// image through one SM: with 3 - 300 rows that is one to three CTAs each walking a serial
// load -> MMA -> epilogue chain (13 - 20 us per launch, measured).  Here a block owns 16 rows x 32
// output channels (of every 128-channel chunk), its 8 warps split K, and the weight matrix is read
// by (rows / 16) x 4 blocks in parallel straight from L2 in the [in][out] layout (coalesced over
// channels), so one launch is a few microseconds.  FP32 FFMA: same accuracy class as the split
// tensor-core products.
constexpr int kSkinnyThreads = 256;   // TR (8 | 16) rows x 32 channels per block, TR / 8 rows per epilogue warp
template <class Op, int TR>
constexpr size_t ffma_rows_smem_bytes() {
    return sizeof(float) * (size_t)(TR * (Op::KS * 128 + 4) + 8 * TR * 32);
}

// wt: the op's weight matrix as [in][out] with row stride ld (UpdateWeights: M1t / M2t / M2 / M1)
template <class Op, int TR>
__global__ void __launch_bounds__(kSkinnyThreads)
ffma_rows_kernel(Op op, int num_rows, const float* __restrict__ wt, int ld,
                 const DeviceStatus* __restrict__ status) {
    constexpr int KS = Op::KS, NC = Op::NC, K = KS * 128, XS = K + 4, KW = K / 8;
    constexpr int WR = TR / 8;   // rows per epilogue warp
    if (status != nullptr && status->overflow) return;
    extern __shared__ __align__(16) float skinny_smem[];
    float* x_s = skinny_smem;             // [TR rows][XS]
    float* red = skinny_smem + TR * XS;   // [8 warps][TR rows][32 channels]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = (int)blockIdx.x * TR, cg = (int)blockIdx.y;
    if (row0 >= num_rows) return;
    // every weight this thread needs for the first chunk, requested before anything else so the L2
    // latency overlaps the produce phase: the kernel is a latency chain, not a throughput problem
    // (ncu: 2 warps per scheduler, ~7 cycles between issues, 5 % of the SM's throughput)
    const int kbeg = warp * KW;
    const float* wp = wt + (size_t)kbeg * ld + cg * 32 + lane;
    float wcur[KW], wnxt[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) { wcur[k] = __ldg(wp + (size_t)k * ld); wnxt[k] = 0.f; }
    const int rows_here = min(TR, num_rows - row0);
    for (int idx = tid; idx < rows_here * (K / 4); idx += kSkinnyThreads) {
        const int r = idx / (K / 4), k4 = idx - r * (K / 4);
        const int ks = k4 >> 5, k0 = (k4 & 31) * 4;
        *reinterpret_cast<float4*>(x_s + r * XS + ks * 128 + k0) = op.produce(row0 + r, num_rows, ks, k0);
    }
    __syncthreads();
    uint32_t prev[WR];
#pragma unroll
    for (int j = 0; j < WR; ++j) prev[j] = 0u;
#pragma unroll 1   // straight-line code that runs once is instruction-fetch bound: keep the body small
    for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC) {   // next chunk's weights travel while this chunk is computed
#pragma unroll
            for (int k = 0; k < KW; ++k) wnxt[k] = __ldg(wp + (size_t)k * ld + (c + 1) * 128);
        }
        float acc[TR];
#pragma unroll
        for (int j = 0; j < TR; ++j) acc[j] = 0.f;
#pragma unroll
        for (int g = 0; g < TR / 8; ++g) {   // rows in groups of 8; groups beyond the tile's rows are skipped
            if (g * 8 < rows_here) {
#pragma unroll
                for (int k = 0; k < KW; k += 4) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = g * 8 + jj;
                        const float4 xv = *reinterpret_cast<const float4*>(x_s + j * XS + kbeg + k);
                        acc[j] = fmaf(xv.x, wcur[k], fmaf(xv.y, wcur[k + 1], fmaf(xv.z, wcur[k + 2], fmaf(xv.w, wcur[k + 3], acc[j]))));
                    }
                }
            }
        }
        // combine the 8 K-slices and run the op's epilogue: warp w owns rows WR w .. WR w + WR - 1
#pragma unroll
        for (int j = 0; j < TR; ++j)
            if (j < rows_here) red[(warp * TR + j) * 32 + lane] = acc[j];
        __syncthreads();
        if (warp * WR < rows_here) {
            uint32_t r[WR];
#pragma unroll
            for (int jj = 0; jj < WR; ++jj) {
                const int j = warp * WR + jj;
                float sum = red[j * 32 + lane];
#pragma unroll
                for (int w = 1; w < 8; ++w) sum += red[(w * TR + j) * 32 + lane];
                r[jj] = __float_as_uint(sum);
            }
            op.template store<WR>(c, cg * 32 + lane, row0 + warp * WR, num_rows, r, prev);
#pragma unroll
            for (int j = 0; j < WR; ++j) prev[j] = r[j];
        }
        if (c + 1 < NC) __syncthreads();
#pragma unroll
        for (int k = 0; k < KW; ++k) wcur[k] = wnxt[k];
    }
}

// =================================== update block ops =========================================
// Restating PaiNNUpdate.forward (reference src/mlff_distiller/models/student_model.py:434-470)
// and its reverse, H = 128.  Weight "images" are [ks][chunk] 64 KB blocks of
// the [out][in] matrix named at each op.

// y1 = M1 [s'; |v'|] + m1          W = update_mlp.0.weight [H][2H]: KS = 2, NC = 1
struct UpdateFwd1Op {
    static constexpr int KS = 2, NC = 1;
    const float* s_msg; const float* v_msg; const float* m1; float* y1;
    __device__ __forceinline__ float4 produce(int row, int n, int ks, int k0) const {
        if (row >= n) return make4(0.f);
        if (ks == 0) return ldg4(s_msg + (size_t)row * 128 + k0);
        const float* vp = v_msg + (size_t)row * 384 + k0;
        const float4 vx = ldg4(vp), vy = ldg4(vp + 128), vz = ldg4(vp + 256);
        return make_float4(sqrtf(vx.x * vx.x + vy.x * vy.x + vz.x * vz.x), sqrtf(vx.y * vx.y + vy.y * vy.y + vz.y * vz.y),
                           sqrtf(vx.z * vx.z + vy.z * vy.z + vz.z * vz.z), sqrtf(vx.w * vx.w + vy.w * vy.w + vz.w * vz.w));
    }
    template <int R>
    __device__ __forceinline__ void store(int, int ch, int row0, int n, const uint32_t (&r)[R], const uint32_t (&)[R]) const {
        [[maybe_unused]] constexpr int B8 = R < 8 ? R : 8, B4 = R < 4 ? R : 4;   // rows whose loads are issued together
        const float b = __ldg(m1 + ch);
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (row0 + j < n) y1[(size_t)(row0 + j) * 128 + ch] = __uint_as_float(r[j]) + b;
    }
};

// (ds | g1 | g2) = M2 SiLU(y1) + m2 ; s'' = s' + ds ; v'' = v' g1 + (U v') g2
//                                   W = update_mlp.2.weight [3H][H]: KS = 1, NC = 3 (1 when LAST)
template <bool LAST>
struct UpdateFwd2Op {
    static constexpr int KS = 1, NC = LAST ? 1 : 3;
    const float* y1; const float* s_msg; const float* v_msg; const float* m2; const float* U;
    float* s_out; float* v_out; float* gates;
    __device__ __forceinline__ float4 produce(int row, int n, int, int k0) const {
        if (row >= n) return make4(0.f);
        const float4 y = ldg4(y1 + (size_t)row * 128 + k0);
        return make_float4(siluf_(y.x), siluf_(y.y), siluf_(y.z), siluf_(y.w));
    }
    template <int R>
    __device__ __forceinline__ void store(int c, int ch, int row0, int n, const uint32_t (&r)[R], const uint32_t (&prev)[R]) const {
        [[maybe_unused]] constexpr int B8 = R < 8 ? R : 8, B4 = R < 4 ? R : 4;   // rows whose loads are issued together
        if (c == 0) {
            const float b = __ldg(m2 + ch);
            // loads of a batch of rows are issued together, then the stores: a load placed after a
            // store through another pointer cannot be hoisted by the compiler (possible alias)
#pragma unroll
            for (int jb = 0; jb < R; jb += B8) {
                float sv[8];
#pragma unroll
                for (int j = 0; j < B8; ++j)
                    sv[j] = (row0 + jb + j < n) ? __ldg(s_msg + (size_t)(row0 + jb + j) * 128 + ch) : 0.f;
#pragma unroll
                for (int j = 0; j < B8; ++j)
                    if (row0 + jb + j < n)
                        s_out[(size_t)(row0 + jb + j) * 128 + ch] = sv[j] + __uint_as_float(r[jb + j]) + b;
            }
        } else if (c == 2) {   // prev = g1 accumulators (chunk 1), r = g2
            const float b1 = __ldg(m2 + 128 + ch), b2 = __ldg(m2 + 256 + ch);
            float u[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) u[k] = __ldg(U + k);
#pragma unroll
            for (int jb = 0; jb < R; jb += B8) {
                float vx[8], vy[8], vz[8];
#pragma unroll
                for (int j = 0; j < B8; ++j) {
                    const bool ok = row0 + jb + j < n;
                    const float* vp = v_msg + (size_t)(row0 + jb + j) * 384 + ch;
                    vx[j] = ok ? __ldg(vp) : 0.f; vy[j] = ok ? __ldg(vp + 128) : 0.f; vz[j] = ok ? __ldg(vp + 256) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < B8; ++j)
                    if (row0 + jb + j < n) {
                        const size_t row = (size_t)(row0 + jb + j);
                        const float g1 = __uint_as_float(prev[jb + j]) + b1, g2 = __uint_as_float(r[jb + j]) + b2;
                        gates[row * 256 + ch] = g1;
                        gates[row * 256 + 128 + ch] = g2;
                        float* vo = v_out + row * 384 + ch;
                        vo[0] = vx[j] * g1 + (u[0] * vx[j] + u[1] * vy[j] + u[2] * vz[j]) * g2;
                        vo[128] = vy[j] * g1 + (u[3] * vx[j] + u[4] * vy[j] + u[5] * vz[j]) * g2;
                        vo[256] = vz[j] * g1 + (u[6] * vx[j] + u[7] * vy[j] + u[8] * vz[j]) * g2;
                    }
            }
        }
    }
};

// hid_bar = M2^T [s_bar; g1_bar; g2_bar] ; y1_bar = hid_bar * SiLU'(y1)   (written over y1)
//                                   W = M2^T [H][3H]: KS = 3 (1 when LAST: g_bar = 0), NC = 1
template <bool LAST>
struct UpdateBwd1Op {
    static constexpr int KS = LAST ? 1 : 3, NC = 1;
    const float* sbar; const float* vbar; const float* v_msg; const float* U; float* y1;
    __device__ __forceinline__ float4 produce(int row, int n, int ks, int k0) const {
        if (row >= n) return make4(0.f);
        if (ks == 0) return ldg4(sbar + (size_t)row * 128 + k0);
        const float* vp = v_msg + (size_t)row * 384 + k0;
        const float* bp = vbar + (size_t)row * 384 + k0;
        const float4 vx = ldg4(vp), vy = ldg4(vp + 128), vz = ldg4(vp + 256);
        const float4 bx = ldg4(bp), by = ldg4(bp + 128), bz = ldg4(bp + 256);
        if (ks == 1)
            return make_float4(bx.x * vx.x + by.x * vy.x + bz.x * vz.x, bx.y * vx.y + by.y * vy.y + bz.y * vz.y,
                               bx.z * vx.z + by.z * vy.z + bz.z * vz.z, bx.w * vx.w + by.w * vy.w + bz.w * vz.w);
        float u[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) u[k] = __ldg(U + k);
        auto mix = [&](float b0, float b1, float b2, float v0, float v1, float v2) {
            return b0 * (u[0] * v0 + u[1] * v1 + u[2] * v2) + b1 * (u[3] * v0 + u[4] * v1 + u[5] * v2) +
                   b2 * (u[6] * v0 + u[7] * v1 + u[8] * v2);
        };
        return make_float4(mix(bx.x, by.x, bz.x, vx.x, vy.x, vz.x), mix(bx.y, by.y, bz.y, vx.y, vy.y, vz.y),
                           mix(bx.z, by.z, bz.z, vx.z, vy.z, vz.z), mix(bx.w, by.w, bz.w, vx.w, vy.w, vz.w));
    }
    template <int R>
    __device__ __forceinline__ void store(int, int ch, int row0, int n, const uint32_t (&r)[R], const uint32_t (&)[R]) const {
        [[maybe_unused]] constexpr int B8 = R < 8 ? R : 8, B4 = R < 4 ? R : 4;   // rows whose loads are issued together
#pragma unroll
        for (int jb = 0; jb < R; jb += B8) {
            float yv[8];
#pragma unroll
            for (int j = 0; j < B8; ++j) yv[j] = (row0 + jb + j < n) ? y1[(size_t)(row0 + jb + j) * 128 + ch] : 0.f;
#pragma unroll
            for (int j = 0; j < B8; ++j)
                if (row0 + jb + j < n)
                    y1[(size_t)(row0 + jb + j) * 128 + ch] = __uint_as_float(r[jb + j]) * silu_gradf_(yv[j]);
        }
    }
};

// [ps_bar | n_bar] = M1^T y1_bar ; s_bar += ps_bar ;
// v_bar <- v_bar g1 + U^T (v_bar g2) + n_bar v'/|v'|      W = M1^T [2H][H]: KS = 1, NC = 2
template <bool LAST>
struct UpdateBwd2Op {
    static constexpr int KS = 1, NC = 2;
    const float* ybar; const float* v_msg; const float* gates; const float* U; float* sbar; float* vbar;
    __device__ __forceinline__ float4 produce(int row, int n, int, int k0) const {
        if (row >= n) return make4(0.f);
        return ldg4(ybar + (size_t)row * 128 + k0);
    }
    template <int R>
    __device__ __forceinline__ void store(int c, int ch, int row0, int n, const uint32_t (&r)[R], const uint32_t (&)[R]) const {
        [[maybe_unused]] constexpr int B8 = R < 8 ? R : 8, B4 = R < 4 ? R : 4;   // rows whose loads are issued together
        if (c == 0) {
#pragma unroll
            for (int jb = 0; jb < R; jb += B8) {
                float sv[8];
#pragma unroll
                for (int j = 0; j < B8; ++j) sv[j] = (row0 + jb + j < n) ? sbar[(size_t)(row0 + jb + j) * 128 + ch] : 0.f;
#pragma unroll
                for (int j = 0; j < B8; ++j)
                    if (row0 + jb + j < n) sbar[(size_t)(row0 + jb + j) * 128 + ch] = sv[j] + __uint_as_float(r[jb + j]);
            }
            return;
        }
        float u[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) u[k] = __ldg(U + k);
#pragma unroll
        for (int jb = 0; jb < R; jb += B4) {
            float vx[4], vy[4], vz[4], bx[4], by[4], bz[4], g1[4], g2[4];
#pragma unroll
            for (int j = 0; j < B4; ++j) {
                const bool ok = row0 + jb + j < n;
                const size_t row = (size_t)(row0 + jb + j);
                const float* vp = v_msg + row * 384 + ch;
                vx[j] = ok ? __ldg(vp) : 0.f; vy[j] = ok ? __ldg(vp + 128) : 0.f; vz[j] = ok ? __ldg(vp + 256) : 0.f;
                bx[j] = by[j] = bz[j] = g1[j] = g2[j] = 0.f;
                if (!LAST && ok) {
                    const float* bp = vbar + row * 384 + ch;
                    bx[j] = bp[0]; by[j] = bp[128]; bz[j] = bp[256];
                    g1[j] = __ldg(gates + row * 256 + ch); g2[j] = __ldg(gates + row * 256 + 128 + ch);
                }
            }
#pragma unroll
            for (int j = 0; j < B4; ++j)
                if (row0 + jb + j < n) {
                    const float nrm = sqrtf(vx[j] * vx[j] + vy[j] * vy[j] + vz[j] * vz[j]);
                    const float sc = (nrm > 0.f) ? __uint_as_float(r[jb + j]) / nrm : 0.f;
                    float ox = sc * vx[j], oy = sc * vy[j], oz = sc * vz[j];
                    if (!LAST) {
                        const float gx = bx[j] * g2[j], gy = by[j] * g2[j], gz = bz[j] * g2[j];
                        ox += bx[j] * g1[j] + (u[0] * gx + u[3] * gy + u[6] * gz);
                        oy += by[j] * g1[j] + (u[1] * gx + u[4] * gy + u[7] * gz);
                        oz += bz[j] * g1[j] + (u[2] * gx + u[5] * gy + u[8] * gz);
                    }
                    float* bp = vbar + (size_t)(row0 + jb + j) * 384 + ch;
                    bp[0] = ox; bp[128] = oy; bp[256] = oz;
                }
        }
    }
};

}  // namespace mlffd
