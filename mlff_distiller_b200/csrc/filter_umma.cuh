// Filter table on the 5th-generation tensor cores (tcgen05 / TMEM), H = 128 / 64 / 32.
//
// Same contract as filter_table_kernel (filter.cuh): for a tile of 64 undirected pairs compute
// f(d) and f'(d) in R^{3H}.  BOTH dense layers run as tcgen05.mma kind::f16 with FP32
// accumulation in tensor memory:
//   layer 1:  D1[channel][row] = W1 (H x 32, K = num_rbf padded)  x  [phi~ ; phi~']^T   (128 rows)
//   layer 2:  D2[channel][row] = W2 chunk (128 x H)               x  [h ; t]^T          (128 rows)
// with h = SiLU(y), t = SiLU'(y) * z the value / tangent of the hidden layer (y, z = the two row
// halves of D1).  To keep FP32-grade accuracy (north-star: 1e-5 eV/atom, 1e-4 eV/A) both operands
// of every product are split into two FP16 terms, x = x_hi + x_lo with x_lo = fp16(x - x_hi)
// (22 significand bits), and three products are accumulated: lo*hi + hi*lo + hi*hi (the dropped
// lo*lo term is ~2^-22 relative).  Each FP16 product is exact in the FP32 accumulator, so the
// result differs from an FP32 FFMA GEMM only by accumulation-order rounding.
//
// Both GEMMs are issued TRANSPOSED (the weight image is the A operand, the activation tile the B
// operand), so a TMEM lane is a channel: the first-layer epilogue applies bias / SiLU per lane,
// and the second-layer epilogue stores 128 contiguous bytes of one filter row per warp
// instruction.
//
// Roles (576 threads, one persistent CTA per SM):
//   warps 0-15 compute: RBF x cutoff (and derivative) -> split -> B tile of layer 1;  D1 -> SiLU /
//              tangent -> split -> B tile of layer 2 (same shared memory: K block 0 of the
//              activation tile is free between the layer-2 MMAs of two tiles);  D2 -> + bias ->
//              global.  The RBF math of tile t+1 runs while the tensor cores work on chunk 0 of
//              tile t.
//   warp 16    issuer (one lane): 6 MMAs of layer 1, then the 24 MMAs of each 128-channel chunk of
//              layer 2, committing to an mbarrier per chunk so the epilogue of chunk c overlaps
//              the MMAs of chunk c+1.
//   warp 17    weight loader (one lane): streams the pre-swizzled W2 chunk images (64 KB each:
//              hi/lo x 2 K-blocks) from L2 with cp.async.bulk into a 2-deep ring; chunk g is
//              requested the moment the MMAs of chunk g-2 have completed.  (When the issuer did
//              this itself every chunk exposed ~2.1k cycles of load latency: tcgen05.mma issue
//              blocks while the tensor-core queue is full, so its refill was always late.)
//   The layer-1 weight image (32 KB) stays resident in shared memory.
// Measured per 64-pair tile on a B200 (C2, -DMLFFD_FILTER_TIMING): epilogue stores 7.8k cycles
// (196 KB at ~25 B/clk/SM: the SM's store path), publish 3.6k, RBF + chunk-0 MMA wait 2.5k,
// layer-1 round trip 0.7k; the tensor pipe is busy ~40 % of the time.
// Operand layout: K-major, SWIZZLE_128B (8-row x 128-byte atoms, 16-byte chunk c of row r stored
// at c ^ (r % 8)), descriptor SBO = 1024 B, version 1; instruction descriptor M=128, N=128, F16
// inputs, F32 accumulate.  Encodings pinned by tools/umma_test.cu on a B200.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "filter.cuh"

namespace mlffd {

constexpr int kUmmaPairs = 64;                 // pairs per tile -> 128 GEMM rows (64 h + 64 t)
constexpr int kUmmaComputeWarps = 16;
constexpr int kUmmaComputeThreads = 32 * kUmmaComputeWarps;   // compute / epilogue warps
constexpr int kUmmaThreads = kUmmaComputeThreads + 32;        // + 1 issuer warp
constexpr int kFilterUmmaThreads = kUmmaComputeThreads + 64;  // filter table: + issuer warp + weight-loader warp
constexpr uint32_t kKBlockBytes = 16384;       // 128 rows x 64 halves
constexpr uint32_t kChunkImageBytes = 65536;   // H = 128 image: hi kb0 | hi kb1 | lo kb0 | lo kb1
constexpr uint32_t kTmemCols = 512;

// Geometry of the tensor-core filter kernel for hidden size H (128, 64 or 32).
//   KBLK   64-wide K blocks of the activation / weight tiles (H = 32 is zero-padded to 64)
//   KSTEPS MMA K-steps (16 each) issued per K block
//   NCH    128-channel output chunks covering the 3H filter channels (padded with zero rows)
// Shared memory: activation tile (hi | lo), two-deep ring of second-layer weight chunks, the
// resident first-layer weight image (hi | lo, one K block: num_rbf <= 32 padded with zeros), tail
// (mbarriers, TMEM base, RBF centres / gammas).  The RBF tile of the first-layer GEMM lives in K
// block 0 of the activation tile (free between the second-layer MMAs of two tiles).
constexpr int kUmmaMaxRbf = 32;
template <int H>
struct UmmaGeom {
    static constexpr int KBLK = (H > 64) ? H / 64 : 1;
    static constexpr int KSTEPS = (H >= 64) ? 4 : H / 16;
    static constexpr int NCH = (3 * H + 127) / 128;
    static constexpr uint32_t TILE = KBLK * kKBlockBytes;          // one FP16 term of a 128-row tile
    static constexpr uint32_t IMAGE = 2 * TILE;                    // hi | lo
    static constexpr uint32_t D1_COL = NCH * 128;                  // first-layer accumulator columns
    static constexpr uint32_t TMEM_COLS = (D1_COL + 128 > 256) ? 512 : 256;
    static constexpr uint32_t A_HI = 0;
    static constexpr uint32_t A_LO = TILE;
    static constexpr uint32_t B0 = 2 * TILE;
    static constexpr uint32_t B1 = B0 + IMAGE;
    static constexpr uint32_t W1_HI = B1 + IMAGE;
    static constexpr uint32_t W1_LO = W1_HI + kKBlockBytes;
    static constexpr uint32_t TAIL = W1_LO + kKBlockBytes;         // 10 mbarriers, tmem base, centres, gammas
    static constexpr uint32_t total() { return TAIL + 128 + 2 * kUmmaMaxRbf * 4 + 1024 /*align*/; }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address
    d |= (uint64_t)(1024u >> 4) << 32;          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// Power-of-two pre-scaling keeps the low FP16 term out of the subnormal range (an unscaled term
// below 6e-5 would be quantised to 2^-24 absolute): activations are split as 2^3 x, weights as
// 2^8 w, and the epilogue multiplies the FP32 accumulator by 2^-11 (all exact).
constexpr float kActScale = 8.0f;
constexpr float kWeightScale = 256.0f;
constexpr float kAccUnscale = 1.0f / (kActScale * kWeightScale);

// Tensor-core arithmetic modes (mlffd_config.precision -> kernel argument `tc_mode`):
//   kTcSplit  two-term FP16 split of both operands, three products: FP32-equivalent (default)
//   kTcF16    one product of FP16-rounded operands (11-bit significands, the precision class of
//             TF32): a third of the MMAs, half the weight bytes, no low-term work
//   kTcBF16   one product of BF16-rounded operands (8-bit significands)
// The single-pass modes only ever touch the `hi` halves of the operand tiles / weight images.
constexpr int kTcSplit = 0, kTcF16 = 1, kTcBF16 = 2;

// instruction descriptor: M = 128, N = 128, FP32 accumulate, A/B formats F16 (0) or BF16 (1)
__device__ __forceinline__ uint32_t umma_idesc_128x128(int tc_mode) {
    const uint32_t fmt = (tc_mode == kTcBF16) ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// 16-bit storage of one pre-scaled value: high term (FP16 or BF16 by mode) and FP16 low term
// (meaningful in kTcSplit only)
// FP16 range guard: a pre-scaled operand at or beyond this magnitude (or a NaN) would round to inf.
// The producers OR the test into a per-thread flag and raise DeviceStatus::tc_saturated once per kernel;
// the host then repeats the step on the FP32 FFMA kernels (mlffd_set_dense_fallback).
constexpr float kSplitLimit = 65000.0f;
__device__ __forceinline__ bool split_saturates(float x) { return !(fabsf(x) < kSplitLimit); }

__device__ __forceinline__ void split_scaled(float x, int tc_mode, uint32_t& hi, uint32_t& lo) {
    if (tc_mode == kTcBF16) {
        hi = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
        lo = 0u;
    } else {
        const __half h = __float2half_rn(x);
        hi = (uint32_t)__half_as_ushort(h);
        lo = (uint32_t)__half_as_ushort(__float2half_rn(x - __half2float(h)));
    }
}

// split of 8 consecutive channels, packed for one 16-byte swizzle chunk
__device__ __forceinline__ void split8(const float (&xin)[8], int tc_mode, uint4& hi, uint4& lo, bool& sat) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) sat |= (tc_mode != kTcBF16) && split_saturates(xin[i] * kActScale);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t h0, l0, h1, l1;
        split_scaled(xin[2 * i] * kActScale, tc_mode, h0, l0);
        split_scaled(xin[2 * i + 1] * kActScale, tc_mode, h1, l1);
        h[i] = h0 | (h1 << 16);
        l[i] = l0 | (l1 << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// same split for 4 consecutive channels (half a swizzle chunk, 8 bytes per term)
__device__ __forceinline__ void split4(const float4& xin, int tc_mode, uint2& hi, uint2& lo, bool& sat) {
    const float x[4] = {xin.x * kActScale, xin.y * kActScale, xin.z * kActScale, xin.w * kActScale};
    sat |= (tc_mode != kTcBF16) && (split_saturates(x[0]) || split_saturates(x[1]) || split_saturates(x[2]) || split_saturates(x[3]));
    uint32_t h[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint32_t h0, l0, h1, l1;
        split_scaled(x[2 * i], tc_mode, h0, l0);
        split_scaled(x[2 * i + 1], tc_mode, h1, l1);
        h[i] = h0 | (h1 << 16);
        l[i] = l0 | (l1 << 16);
    }
    hi = make_uint2(h[0], h[1]);
    lo = make_uint2(l[0], l[1]);
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `r` inside a K-block image
__device__ __forceinline__ uint32_t sw128_offset(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

// Phase timing for tuning (nvcc -DMLFFD_FILTER_TIMING): warp 0 of block 0 prints the cycles it
// spent in each phase of the compute / epilogue role.
#ifdef MLFFD_FILTER_TIMING
#define FT_DECL long long ft_acc[6] = {0, 0, 0, 0, 0, 0}; long long ft_t = clock64();
#define FT_MARK(i) { const long long ft_n = clock64(); ft_acc[i] += ft_n - ft_t; ft_t = ft_n; }
#define FT_PRINT if (blockIdx.x == 0 && tid == 0) printf("filter phases (cycles, %d tiles): wait_d %lld  epilogue %lld  rbf %lld  wait_d1 %lld  publish %lld  sync %lld\n", my_tiles, ft_acc[0], ft_acc[1], ft_acc[2], ft_acc[3], ft_acc[4], ft_acc[5]);
#else
#define FT_DECL
#define FT_MARK(i)
#define FT_PRINT
#endif

// split of one value, pre-scaled like split8 (16-bit patterns)
__device__ __forceinline__ void split1(float x, int tc_mode, uint16_t& hi, uint16_t& lo, bool& sat) {
    uint32_t h, l;
    sat |= (tc_mode != kTcBF16) && split_saturates(x * kActScale);
    split_scaled(x * kActScale, tc_mode, h, l);
    hi = (uint16_t)h;
    lo = (uint16_t)l;
}

// The filter of every layer depends on the pair distances only, so ONE launch evaluates the tables
// of all layers: the grid is split into per-layer groups of persistent CTAs (sized by the layers'
// chunk counts so the groups finish together).  For a small system that is one kernel latency per
// step instead of L; for a large batch nothing changes but the launch count.
constexpr int kFilterMaxLayers = 8;
struct FilterLayerArgs {
    FilterWeights w;
    const uint8_t* w1_image;   // first-layer weight image (hi | lo)
    const uint8_t* w2_images;  // second-layer chunk images
    float* filt;               // [P, 3H]
    float* dfilt;              // [P, 3H]
    int skip_vector_gate;      // layer 0 of the model: the b gate is never read
    int cta_begin;             // first CTA of this layer's group (groups are consecutive)
};
struct FilterBatchArgs {
    FilterLayerArgs layer[kFilterMaxLayers];
    int num_layers;
};

template <int H, int MODE = kTcSplit>
__global__ void __launch_bounds__(kFilterUmmaThreads, 1)
filter_table_umma_kernel(const float* __restrict__ pair_dist, const int* __restrict__ num_pairs_ptr,
                         int num_pairs_arg, const DeviceStatus* __restrict__ status,
                         const float* __restrict__ centers, const float* __restrict__ gammas, int K,
                         float rc, const __grid_constant__ FilterBatchArgs batch) {
    constexpr int tc_mode = MODE;
    using G = UmmaGeom<H>;
    if (status != nullptr && status->overflow) return;
    const int P = (num_pairs_ptr != nullptr) ? *num_pairs_ptr : num_pairs_arg;
    const int num_tiles = (P + kUmmaPairs - 1) / kUmmaPairs;
    // this CTA's layer and its position inside the layer's group
    int li = 0;
    while (li + 1 < batch.num_layers && (int)blockIdx.x >= batch.layer[li + 1].cta_begin) ++li;
    const int bx = (int)blockIdx.x - batch.layer[li].cta_begin;
    const int gx = ((li + 1 < batch.num_layers) ? batch.layer[li + 1].cta_begin : (int)gridDim.x) - batch.layer[li].cta_begin;
    if (bx >= num_tiles) return;
    const FilterWeights w = batch.layer[li].w;
    const uint8_t* __restrict__ w1_image = batch.layer[li].w1_image;
    const uint8_t* __restrict__ w2_images = batch.layer[li].w2_images;
    const int skip_vector_gate = batch.layer[li].skip_vector_gate;
    float* __restrict__ filt = batch.layer[li].filt;
    float* __restrict__ dfilt = batch.layer[li].dfilt;

    // Dynamic shared memory only (no static __shared__), so the 1024-byte alignment pad is
    // computed in the shared address space and every access below stays an LDS/STS.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = (uint64_t*)(smem + G::TAIL);
    uint64_t* bar_a_full = bars;          // second-layer activation tile written
    uint64_t* bar_b_full = bars + 1;      // [2] weight chunk landed
    uint64_t* bar_d_full = bars + 3;      // [3] second-layer accumulator chunk complete
    uint64_t* bar_phi_full = bars + 6;    // RBF tile written
    uint64_t* bar_d1_full = bars + 7;     // first-layer accumulator complete
    uint64_t* bar_w1_full = bars + 8;     // first-layer weight image landed (once)
    uint32_t* tmem_base_s = (uint32_t*)(bars + 10);
    float* cen_s = (float*)(smem + G::TAIL + 128);
    float* gam_s = cen_s + kUmmaMaxRbf;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // layer 0 never reads the b gate (v_in = 0): for H = 128 that is exactly chunk 1, skipped
    // entirely; for smaller H the chunks mix gates, so only the stores are skipped.
    const bool skip_chunk1 = skip_vector_gate && H == 128;
    const int nchunks = skip_chunk1 ? 2 : G::NCH;
    const int my_tiles = (num_tiles - bx + gx - 1) / gx;
    constexpr bool single = tc_mode != kTcSplit;                   // one product, `hi` halves only
    constexpr uint32_t w2_bytes = single ? G::TILE : G::IMAGE;     // bytes of a weight chunk in use
    constexpr int first_pass = single ? 2 : 0;                     // passes: hi*lo, lo*hi, hi*hi

    // ---- one-time setup ----
    if (warp == kUmmaComputeWarps) {
        if (lane == 0) {
            for (int i = 0; i < 9; ++i) mbar_init(bars + i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // start streaming the first-layer image and the first two weight chunks (latency path)
            mbar_expect_tx(bar_w1_full, single ? kKBlockBytes : 2 * kKBlockBytes);
            bulk_g2s(smem_u32(smem + G::W1_HI), w1_image, kKBlockBytes, bar_w1_full);
            if (!single) bulk_g2s(smem_u32(smem + G::W1_LO), w1_image + kKBlockBytes, kKBlockBytes, bar_w1_full);
            const int total0 = my_tiles * nchunks;
            for (int g = 0; g < 2 && g < total0; ++g) {
                mbar_expect_tx(&bar_b_full[g], w2_bytes);
                const int cid = skip_chunk1 ? (g % nchunks) * 2 : (g % nchunks);
                const uint8_t* src = w2_images + (size_t)cid * G::IMAGE;
                const uint32_t dst = smem_u32(smem + (g ? G::B1 : G::B0));
                for (uint32_t off = 0; off < w2_bytes; off += kKBlockBytes)
                    bulk_g2s(dst + off, src + off, kKBlockBytes, &bar_b_full[g]);
            }
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else if (tid < kUmmaMaxRbf) {
        cen_s[tid] = (tid < K) ? __ldg(centers + tid) : 0.f;
        gam_s[tid] = (tid < K) ? __ldg(gammas + tid) : 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == kUmmaComputeWarps) {
        // =============================== issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_128x128(tc_mode);
            const uint32_t b_buf[2] = {smem_u32(smem + G::B0), smem_u32(smem + G::B1)};
            // every operand descriptor is a constant plus a small offset in the 16-byte address field
            const uint64_t act_desc_hi = umma_desc_sw128(smem_u32(smem + G::A_HI));
            const uint64_t act_desc_lo = umma_desc_sw128(smem_u32(smem + G::A_LO));
            const uint64_t w1_desc_hi = umma_desc_sw128(smem_u32(smem + G::W1_HI));
            const uint64_t w1_desc_lo = umma_desc_sw128(smem_u32(smem + G::W1_LO));
            const uint64_t w_desc0 = umma_desc_sw128(b_buf[0]);
            const int total_chunks = my_tiles * nchunks;
            auto chunk_id = [&](int g) { const int ci = g % nchunks; return skip_chunk1 ? ci * 2 : ci; };
            mbar_wait(bar_w1_full, 0);
#ifdef MLFFD_FILTER_TIMING
            long long it_acc[5] = {0, 0, 0, 0, 0}, it_t = clock64();
#define IT_MARK(i) { const long long n_ = clock64(); it_acc[i] += n_ - it_t; it_t = n_; }
#else
#define IT_MARK(i)
#endif
            for (int g = 0; g < total_chunks; ++g) {   // chunks 0 and 1 were requested during set-up
                const int it = g / nchunks, ci = g - it * nchunks, nc = chunk_id(g), buf = g & 1;
                if (ci == 0) {
                    // first layer of tile `it`:  D1[channel][row] = W1 x [phi~ ; phi~']^T, K = 32.
                    // The tensor core accumulates with truncation, a bias that grows with the number
                    // of additions into a LARGE accumulator: the two small correction products go
                    // first (accumulator still ~2^-11 of its final size), the main MMAs last.
                    IT_MARK(4)
                    mbar_wait(bar_phi_full, it & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    IT_MARK(0)
                    uint32_t acc1 = 0;
#pragma unroll
                    for (int pass = first_pass; pass < 3; ++pass) {   // W_hi*phi_lo, W_lo*phi_hi, W_hi*phi_hi
                        const uint64_t phi_base = (pass == 0) ? act_desc_lo : act_desc_hi;
                        const uint64_t w_base = (pass == 1) ? w1_desc_lo : w1_desc_hi;
#pragma unroll
                        for (int k = 0; k < kUmmaMaxRbf / 16; ++k) {
                            umma_f16(tmem_base + G::D1_COL, w_base + (uint64_t)(k * 2), phi_base + (uint64_t)(k * 2), idesc, acc1);
                            acc1 = 1;
                        }
                    }
                    umma_commit(bar_d1_full);
                    // activation tile `it` written (and with it the first-layer accumulator drained)
                    IT_MARK(4)
                    mbar_wait(bar_a_full, it & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    IT_MARK(1)
                }
                mbar_wait(&bar_b_full[buf], (g >> 1) & 1);
                IT_MARK(2)
                const uint32_t d_tmem = tmem_base + (uint32_t)nc * 128;
                const uint64_t w_desc = w_desc0 + (uint64_t)(buf ? (G::IMAGE >> 4) : 0);
                uint32_t acc = 0;
#pragma unroll
                for (int pass = first_pass; pass < 3; ++pass) {   // W_hi*act_lo, W_lo*act_hi, W_hi*act_hi
                    const uint64_t act_base = (pass == 0) ? act_desc_lo : act_desc_hi;
                    const uint64_t w_base = w_desc + (uint64_t)((pass == 1) ? (G::TILE >> 4) : 0);
#pragma unroll
                    for (int kb = 0; kb < G::KBLK; ++kb)
#pragma unroll
                        for (int k = 0; k < G::KSTEPS; ++k) {
                            const uint64_t off = (uint64_t)((kb * kKBlockBytes + k * 32) >> 4);
                            // transposed product: A operand = weight chunk (M = 128 channels),
                            // B operand = activation tile (N = 128 rows)
                            umma_f16(d_tmem, w_base + off, act_base + off, idesc, acc);
                            acc = 1;
                        }
                }
                umma_commit(&bar_d_full[nc]);
            }
#ifdef MLFFD_FILTER_TIMING
            if (blockIdx.x == 0) printf("issuer waits (cycles): phi %lld  act %lld  weights %lld  issue %lld\n", it_acc[0], it_acc[1], it_acc[2], it_acc[4]);
#endif
        }
    } else if (warp == kUmmaComputeWarps + 1) {
        // ============================ weight loader ============================
        // Streams second-layer weight chunk g into ring slot g % 2 as soon as the MMAs of chunk
        // g - 2 (the slot's previous tenant) have completed.  A separate warp, because the issuer
        // blocks in tcgen05.mma issue while the tensor-core queue is full: a refill issued from the
        // issuer thread only starts when the NEXT chunk's MMAs are almost done (measured: 2.1k
        // cycles of exposed load latency per chunk).
        if (lane == 0) {
            const uint32_t b_buf[2] = {smem_u32(smem + G::B0), smem_u32(smem + G::B1)};
            const int total_chunks = my_tiles * nchunks;
            for (int g = 2; g < total_chunks; ++g) {   // chunks 0 and 1 were requested during set-up
                const int gp = g - 2, itp = gp / nchunks, cp = gp - itp * nchunks;
                mbar_wait(&bar_d_full[skip_chunk1 ? cp * 2 : cp], itp & 1);
                const int ci = g % nchunks, buf = g & 1;
                mbar_expect_tx(&bar_b_full[buf], w2_bytes);
                const uint8_t* src = w2_images + (size_t)(skip_chunk1 ? ci * 2 : ci) * G::IMAGE;
                for (uint32_t off = 0; off < w2_bytes; off += kKBlockBytes)
                    bulk_g2s(b_buf[buf] + off, src + off, kKBlockBytes, &bar_b_full[buf]);
            }
        }
    } else {
        // ========================= compute / epilogue =========================
        const int q = warp & 3, cs = warp >> 2;         // TMEM lane quarter (channels), column segment
        const int ch = q * 32 + lane;                   // hidden channel of this lane (first layer)
        const float b1v = (ch < H) ? __ldg(w.b1 + ch) : 0.f;
        float bias[G::NCH];
#pragma unroll
        for (int c = 0; c < G::NCH; ++c) {
            const int gch = c * 128 + ch;
            bias[c] = (gch < 3 * H) ? __ldg(w.b2 + gch) : 0.f;
        }

        // RBF x cutoff and its derivative for the 64 pairs of a tile (same products as rbf_cutoff,
        // filter.cuh), split and written as the B operand of the first-layer GEMM: rows 0..63 =
        // phi~, rows 64..127 = phi~'.  Thread = (pair, 8-wide K chunk); K is padded to 32.
        FT_DECL
        bool sat = false;   // an operand left the FP16 range (split_saturates)
        uint4 phi_hi, phi_lo, dphi_hi, dphi_lo;   // this thread's chunk of the next RBF tile
        auto rbf_compute = [&](int it) {   // pure math (overlaps the tensor cores)
            if (tid < 4 * kUmmaPairs) {
                const int p = tid >> 2, chunk = tid & 3;
                const int gp = (bx + it * gx) * kUmmaPairs + p;
                const float d = (gp < P) ? __ldg(pair_dist + gp) : rc;
                const float kPi = 3.14159274101257324f;
                const float arg = (kPi * d) / rc;
                const bool inside = d < rc;
                const float cut = inside ? 0.5f * (cosf(arg) + 1.0f) : 0.0f;
                const float dcut = inside ? -0.5f * (kPi / rc) * sinf(arg) : 0.0f;
                float v8[8], t8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = chunk * 8 + i;
                    const float diff = d - cen_s[k], gamma = gam_s[k];
                    const float phi = expf(-gamma * (diff * diff));
                    const float dphi = -2.0f * gamma * diff * phi;
                    v8[i] = (k < K) ? phi * cut : 0.f;
                    t8[i] = (k < K) ? dphi * cut + phi * dcut : 0.f;
                }
                split8(v8, tc_mode, phi_hi, phi_lo, sat);
                split8(t8, tc_mode, dphi_hi, dphi_lo, sat);
            }
        };
        auto rbf_store = [&]() {           // needs the activation tile free (second-layer MMAs done)
            FT_MARK(5)
            if (tid < 4 * kUmmaPairs) {
                const int p = tid >> 2, chunk = tid & 3;
                const uint32_t ov = sw128_offset(p, chunk), ot = sw128_offset(kUmmaPairs + p, chunk);
                *reinterpret_cast<uint4*>(smem + G::A_HI + ov) = phi_hi;
                *reinterpret_cast<uint4*>(smem + G::A_HI + ot) = dphi_hi;
                if (!single) {
                    *reinterpret_cast<uint4*>(smem + G::A_LO + ov) = phi_lo;
                    *reinterpret_cast<uint4*>(smem + G::A_LO + ot) = dphi_lo;
                }
            }
            FT_MARK(2)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tid == 0) mbar_arrive(bar_phi_full);
            FT_MARK(5)
        };
        // First-layer accumulator -> SiLU / tangent -> two-term split -> swizzled activation tile.
        // Lane = hidden channel; this warp handles pairs cs*16 .. cs*16+15 (y) and their tangents.
        auto publish_tile = [&](int it) {
            mbar_wait(bar_d1_full, it & 1);
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            FT_MARK(3)
            if (ch < H) {
                uint32_t ry[16], rz[16];
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + G::D1_COL + (uint32_t)(cs * 16);
                tmem_ld16(t0, ry);
                tmem_ld16(t0 + kUmmaPairs, rz);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int kb = ch >> 6;
                const uint32_t in_row = (uint32_t)(ch & 7) * 2;      // byte inside the 16-byte chunk
                const int chunk = (ch & 63) >> 3;
                uint8_t* a_hi = smem + G::A_HI + kb * kKBlockBytes + in_row;
                uint8_t* a_lo = smem + G::A_LO + kb * kKBlockBytes + in_row;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int p = cs * 16 + j;
                    const float yy = fmaf(__uint_as_float(ry[j]), kAccUnscale, b1v);
                    const float zz = __uint_as_float(rz[j]) * kAccUnscale;
                    const float sg = sigmoidf_(yy);
                    const float hv = yy * sg;
                    const float tv = sg * (1.0f + yy * (1.0f - sg)) * zz;
                    uint16_t hh, hl, th, tl;
                    split1(hv, tc_mode, hh, hl, sat);
                    split1(tv, tc_mode, th, tl, sat);
                    const uint32_t oh = sw128_offset(p, chunk), ot = sw128_offset(kUmmaPairs + p, chunk);
                    *reinterpret_cast<uint16_t*>(a_hi + oh) = hh;
                    *reinterpret_cast<uint16_t*>(a_hi + ot) = th;
                    if (!single) {
                        *reinterpret_cast<uint16_t*>(a_lo + oh) = hl;
                        *reinterpret_cast<uint16_t*>(a_lo + ot) = tl;
                    }
                }
            }
            FT_MARK(4)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tid == 0) mbar_arrive(bar_a_full);
            FT_MARK(5)
        };

        rbf_compute(0);
        rbf_store();
        publish_tile(0);
        for (int it = 0; it < my_tiles; ++it) {
            const int p0 = (bx + it * gx) * kUmmaPairs;
            if (it + 1 < my_tiles) rbf_compute(it + 1);        // while the tensor cores work on chunk 0
            FT_MARK(2)
            // ---- epilogue of tile `it`: D[channel = lane][row = column] -> global ----
            const bool tangent = cs >= 2;                      // rows 64..127 hold f'
            const int row0 = p0 + (cs & 1) * 32;               // first pair of this warp's 32 rows
            float* out = (tangent ? dfilt : filt) + (size_t)row0 * (3 * H);
            const bool full = row0 + 32 <= P;
#pragma unroll
            for (int nc = 0; nc < G::NCH; ++nc) {
                if (skip_chunk1 && nc == 1) continue;
                mbar_wait(&bar_d_full[nc], it & 1);
                __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                FT_MARK(0)
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nc * 128 + cs * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int gch = nc * 128 + ch;                  // filter channel of this lane
                bool store = gch < 3 * H;
                if (skip_vector_gate && gch >= H && gch < 2 * H) store = false;   // unused b gate of layer 0
                if (store) {
                    const float b = tangent ? 0.f : bias[nc];
                    float* o = out + gch;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) __stcs(o + (size_t)j * (3 * H), fmaf(__uint_as_float(r[j]), kAccUnscale, b));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (row0 + j < P) __stcs(o + (size_t)j * (3 * H), fmaf(__uint_as_float(r[j]), kAccUnscale, b));
                    }
                }
                FT_MARK(1)
            }
            // all second-layer MMAs of tile `it` are complete (last chunk waited): the activation
            // tile is free for the RBF tile and then the activations of tile it + 1
            if (it + 1 < my_tiles) {
                rbf_store();
                publish_tile(it + 1);
            }
        }
        FT_PRINT
        if (sat && status != nullptr) const_cast<DeviceStatus*>(status)->tc_saturated = 1;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == kUmmaComputeWarps) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::TMEM_COLS));
    }
}

}  // namespace mlffd
