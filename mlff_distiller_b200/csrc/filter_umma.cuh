// Filter table on the 5th-generation tensor cores (tcgen05 / TMEM), H = 128.
//
// Same contract as filter_table_kernel (filter.cuh): for a tile of 64 undirected pairs compute
// f(d) and f'(d) in R^{3H}.  The second dense layer  [h; t] (128 rows x K=128)  x  W2^T (K x 384)
// runs as tcgen05.mma kind::f16 with FP32 accumulation in tensor memory.  To keep FP32-grade
// accuracy (north-star: 1e-5 eV/atom, 1e-4 eV/A) both operands are split into two FP16 terms,
// x = x_hi + x_lo with x_lo = fp16(x - x_hi)  (22 significand bits), and three products are
// accumulated:  hi*hi + lo*hi + hi*lo  (the dropped lo*lo term is ~2^-22 relative).  Each FP16
// product is exact in the FP32 accumulator, so the result differs from an FP32 FFMA GEMM only by
// accumulation-order rounding.  Cost: 3 MMAs at the FP16 rate = 1.5 TF32-rate passes, 4 bytes per
// operand element (same footprint as one TF32 operand).
//
// Roles (544 threads, one persistent CTA per SM):
//   warps 0-15 compute: RBF + first layer (FFMA, registers) -> split -> swizzled activation tile
//              in smem; then epilogue: tcgen05.ld accumulator -> + bias -> global.  The GEMM is
//              issued TRANSPOSED (D[channel][row] = W2_chunk x act^T: the weight image is the A
//              operand, the activation tile the B operand), so a TMEM lane is an output channel
//              and the 32 lanes of a warp store 128 contiguous bytes of one filter row.
//   warp 16    issuer (one lane): streams the pre-swizzled W2 chunk images (64 KB each: hi/lo x 2
//              K-blocks) from L2 with cp.async.bulk into a 2-deep ring, issues the 24 MMAs of a
//              128-channel chunk, commits to an mbarrier per chunk so the epilogue of chunk c
//              overlaps the MMAs of chunk c+1.  The first layer of tile t+1 overlaps the MMAs of
//              tile t (software pipeline in the compute warps).
// Operand layout: K-major, SWIZZLE_128B (8-row x 128-byte atoms, 16-byte chunk c of row r stored
// at c ^ (r % 8)), descriptor SBO = 1024 B, version 1; instruction descriptor M=128, N=128, F16
// inputs, F32 accumulate.  Encodings pinned by tools/umma_test.cu on a B200.
#pragma once
#include <cuda_fp16.h>

#include "filter.cuh"

namespace mlffd {

constexpr int kUmmaPairs = 64;                 // pairs per tile -> 128 GEMM rows (64 h + 64 t)
constexpr int kUmmaComputeWarps = 16;
constexpr int kUmmaComputeThreads = 32 * kUmmaComputeWarps;   // compute / epilogue warps
constexpr int kUmmaThreads = kUmmaComputeThreads + 32;        // + 1 issuer warp
constexpr uint32_t kKBlockBytes = 16384;       // 128 rows x 64 halves
constexpr uint32_t kChunkImageBytes = 65536;   // H = 128 image: hi kb0 | hi kb1 | lo kb0 | lo kb1
constexpr uint32_t kTmemCols = 512;

// Geometry of the tensor-core filter kernel for hidden size H (128, 64 or 32).
//   KBLK   64-wide K blocks of the activation / weight tiles (H = 32 is zero-padded to 64)
//   KSTEPS MMA K-steps (16 each) issued per K block
//   NCH    128-channel output chunks covering the 3H filter channels (padded with zero rows)
template <int H>
struct UmmaGeom {
    static constexpr int KBLK = (H > 64) ? H / 64 : 1;
    static constexpr int KSTEPS = (H >= 64) ? 4 : H / 16;
    static constexpr int NCH = (3 * H + 127) / 128;
    static constexpr uint32_t TILE = KBLK * kKBlockBytes;          // one FP16 term of a 128-row tile
    static constexpr uint32_t IMAGE = 2 * TILE;                    // hi | lo
    static constexpr uint32_t TMEM_COLS = (NCH * 128 > 256) ? 512 : (NCH * 128 > 128 ? 256 : 128);
    static constexpr uint32_t A_HI = 0;
    static constexpr uint32_t A_LO = TILE;
    static constexpr uint32_t B0 = 2 * TILE;
    static constexpr uint32_t B1 = B0 + IMAGE;
    static constexpr uint32_t PHI = B1 + IMAGE;                                // [64][33] f32
    static constexpr uint32_t DPHI = PHI + kUmmaPairs * kPhiStride * 4;        // [64][33] f32
    static constexpr uint32_t W1T = DPHI + kUmmaPairs * kPhiStride * 4;        // [K][H] f32
    static constexpr uint32_t total(int K) {
        return W1T + (uint32_t)K * H * 4 + H * 4 /*b1*/ + 64 /*mbarriers, tmem base*/ + 512 /*cutoff per pair*/ + 1024 /*align*/;
    }
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

// Power-of-two pre-scaling keeps the low FP16 term out of the subnormal range (an unscaled term
// below 6e-5 would be quantised to 2^-24 absolute): activations are split as 2^3 x, weights as
// 2^8 w, and the epilogue multiplies the FP32 accumulator by 2^-11 (all exact).
constexpr float kActScale = 8.0f;
constexpr float kWeightScale = 256.0f;
constexpr float kAccUnscale = 1.0f / (kActScale * kWeightScale);

// fp16 two-term split of 8 consecutive channels, packed for one 16-byte swizzle chunk
__device__ __forceinline__ void split8(const float (&xin)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = xin[i] * kActScale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half h0 = __float2half_rn(x[2 * i]), h1 = __float2half_rn(x[2 * i + 1]);
        const __half l0 = __float2half_rn(x[2 * i] - __half2float(h0));
        const __half l1 = __float2half_rn(x[2 * i + 1] - __half2float(h1));
        h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// same split for 4 consecutive channels (half a swizzle chunk, 8 bytes per term)
__device__ __forceinline__ void split4(const float4& xin, uint2& hi, uint2& lo) {
    const float x[4] = {xin.x * kActScale, xin.y * kActScale, xin.z * kActScale, xin.w * kActScale};
    uint32_t h[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const __half h0 = __float2half_rn(x[2 * i]), h1 = __float2half_rn(x[2 * i + 1]);
        const __half l0 = __float2half_rn(x[2 * i] - __half2float(h0));
        const __half l1 = __float2half_rn(x[2 * i + 1] - __half2float(h1));
        h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    hi = make_uint2(h[0], h[1]);
    lo = make_uint2(l[0], l[1]);
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `r` inside a K-block image
__device__ __forceinline__ uint32_t sw128_offset(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

template <int H>
__global__ void __launch_bounds__(kUmmaThreads, 1)
filter_table_umma_kernel(const float* __restrict__ pair_dist, const int* __restrict__ num_pairs_ptr,
                         int num_pairs_arg, const DeviceStatus* __restrict__ status,
                         const float* __restrict__ centers, const float* __restrict__ gammas, int K,
                         float rc, FilterWeights w, const uint8_t* __restrict__ w2_images,
                         int skip_vector_gate, float* __restrict__ filt, float* __restrict__ dfilt) {
    using G = UmmaGeom<H>;
    constexpr int CPT = H / (kUmmaComputeThreads / kUmmaPairs);   // channels per thread (16 / 8 / 4)
    if (status != nullptr && status->overflow) return;
    const int P = (num_pairs_ptr != nullptr) ? *num_pairs_ptr : num_pairs_arg;
    const int num_tiles = (P + kUmmaPairs - 1) / kUmmaPairs;
    if ((int)blockIdx.x >= num_tiles) return;

    // Dynamic shared memory only (no static __shared__), so the 1024-byte alignment pad is
    // computed in the shared address space and every access below stays an LDS/STS.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* phi_s = (float*)(smem + G::PHI);
    float* dphi_s = (float*)(smem + G::DPHI);
    float* w1t_s = (float*)(smem + G::W1T);
    float* b1_s = w1t_s + K * H;
    uint64_t* bars = (uint64_t*)(b1_s + H);          // a_full, b_full[2], d_full[3]
    uint64_t* bar_a_full = bars;
    uint64_t* bar_b_full = bars + 1;
    uint64_t* bar_d_full = bars + 3;
    uint32_t* tmem_base_s = (uint32_t*)(bars + 6);
    float2* cut_s = (float2*)(bars + 8);             // [64] cutoff value and derivative per pair

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // layer 0 never reads the b gate (v_in = 0): for H = 128 that is exactly chunk 1, skipped
    // entirely; for smaller H the chunks mix gates, so only the stores are skipped.
    const bool skip_chunk1 = skip_vector_gate && H == 128;
    const int nchunks = skip_chunk1 ? 2 : G::NCH;
    const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // ---- one-time setup ----
    if (warp == kUmmaComputeWarps) {
        if (lane == 0) {
            for (int i = 0; i < 6; ++i) mbar_init(bars + i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // start streaming the first two weight chunks before anything else (latency path)
            const int total0 = my_tiles * nchunks;
            for (int g = 0; g < 2 && g < total0; ++g) {
                mbar_expect_tx(&bar_b_full[g], G::IMAGE);
                const int cid = skip_chunk1 ? (g % nchunks) * 2 : (g % nchunks);
                const uint8_t* src = w2_images + (size_t)cid * G::IMAGE;
                const uint32_t dst = smem_u32(smem + (g ? G::B1 : G::B0));
                for (uint32_t off = 0; off < G::IMAGE; off += kKBlockBytes)
                    bulk_g2s(dst + off, src + off, kKBlockBytes, &bar_b_full[g]);
            }
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
        for (int idx = tid; idx < K * H / 4; idx += kUmmaComputeThreads) st4(w1t_s + 4 * idx, ldg4(w.W1t + 4 * idx));
        for (int idx = tid; idx < H; idx += kUmmaComputeThreads) b1_s[idx] = __ldg(w.b1 + idx);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == kUmmaComputeWarps) {
        // =============================== issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t b_buf[2] = {smem_u32(smem + G::B0), smem_u32(smem + G::B1)};
            // every operand descriptor is a constant plus a small offset in the 16-byte address field
            const uint64_t act_desc_hi = umma_desc_sw128(smem_u32(smem + G::A_HI));
            const uint64_t act_desc_lo = umma_desc_sw128(smem_u32(smem + G::A_LO));
            const uint64_t w_desc0 = umma_desc_sw128(b_buf[0]);
            const int total_chunks = my_tiles * nchunks;
            auto chunk_id = [&](int g) { const int ci = g % nchunks; return skip_chunk1 ? ci * 2 : ci; };
            auto issue_load = [&](int g) {
                const int buf = g & 1;
                mbar_expect_tx(&bar_b_full[buf], G::IMAGE);
                const uint8_t* src = w2_images + (size_t)chunk_id(g) * G::IMAGE;
                for (uint32_t off = 0; off < G::IMAGE; off += kKBlockBytes)
                    bulk_g2s(b_buf[buf] + off, src + off, kKBlockBytes, &bar_b_full[buf]);
            };
            for (int g = 0; g < total_chunks; ++g) {   // chunks 0 and 1 were requested during set-up
                const int it = g / nchunks, ci = g - it * nchunks, nc = chunk_id(g), buf = g & 1;
                if (ci == 0) {  // activation tile `it` written; previous accumulators drained
                    mbar_wait(bar_a_full, it & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(&bar_b_full[buf], (g >> 1) & 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)nc * 128;
                const uint64_t w_desc = w_desc0 + (uint64_t)(buf ? (G::IMAGE >> 4) : 0);
                uint32_t acc = 0;
                // The tensor core accumulates with truncation, a bias that grows with the number of
                // additions into a LARGE accumulator: the two small correction products go first
                // (accumulator still ~2^-11 of its final size), the main MMAs last.
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {   // W_hi*act_lo, W_lo*act_hi, W_hi*act_hi
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
                // ring refill: chunk g-1 must have finished reading its buffer before chunk g+1 lands in it
                if (g >= 1 && g + 1 < total_chunks) {
                    const int gp = g - 1, itp = gp / nchunks;
                    mbar_wait(&bar_d_full[chunk_id(gp)], itp & 1);
                    issue_load(g + 1);
                }
            }
        }
    } else {
        // ========================= compute / epilogue =========================
        // Software pipeline: the FFMA first layer of tile it+1 runs while the tensor cores work on
        // tile it; only the final split + store into the (single) activation tile waits for them.
        const int p = tid & 63, cg = tid >> 6;          // pair within tile, channel group (0..7)
        const int q = warp & 3, cs = warp >> 2;         // TMEM lane quarter (channels), row segment
        float y[CPT], z[CPT];
        float bias[G::NCH];
#pragma unroll
        for (int c = 0; c < G::NCH; ++c) {
            const int gch = c * 128 + q * 32 + lane;
            bias[c] = (gch < 3 * H) ? __ldg(w.b2 + gch) : 0.f;
        }

        auto first_layer = [&](int it) {
            const int p0 = ((int)blockIdx.x + it * (int)gridDim.x) * kUmmaPairs;
            // cutoff value/derivative once per pair (not per basis function)
            for (int idx = tid; idx < kUmmaPairs * K; idx += kUmmaComputeThreads) {
                const int pp = idx / K, k = idx - pp * K;
                const float d = (p0 + pp < P) ? __ldg(pair_dist + p0 + pp) : rc;
                if (k == 0) {
                    const float kPi = 3.14159274101257324f;
                    const float arg = (kPi * d) / rc;
                    const bool inside = d < rc;
                    cut_s[pp] = make_float2(inside ? 0.5f * (cosf(arg) + 1.0f) : 0.0f,
                                            inside ? -0.5f * (kPi / rc) * sinf(arg) : 0.0f);
                }
                const float diff = d - __ldg(centers + k), gamma = __ldg(gammas + k);
                const float phi = expf(-gamma * (diff * diff));
                phi_s[pp * kPhiStride + k] = phi;                            // plain Gaussian for now
                dphi_s[pp * kPhiStride + k] = -2.0f * gamma * diff * phi;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            for (int idx = tid; idx < kUmmaPairs * K; idx += kUmmaComputeThreads) {
                const int pp = idx / K, k = idx - pp * K;
                const float2 c = cut_s[pp];
                const float phi = phi_s[pp * kPhiStride + k], dphi = dphi_s[pp * kPhiStride + k];
                phi_s[pp * kPhiStride + k] = phi * c.x;                      // same products as rbf_cutoff
                dphi_s[pp * kPhiStride + k] = dphi * c.x + phi * c.y;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
            for (int c = 0; c < CPT; ++c) { y[c] = b1_s[cg * CPT + c]; z[c] = 0.f; }
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const float ph = phi_s[p * kPhiStride + k], dph = dphi_s[p * kPhiStride + k];
                const float4* wrow = reinterpret_cast<const float4*>(w1t_s + k * H + cg * CPT);
#pragma unroll
                for (int c4 = 0; c4 < CPT / 4; ++c4) {
                    const float4 wv = wrow[c4];
                    y[4 * c4 + 0] = fmaf(ph, wv.x, y[4 * c4 + 0]); z[4 * c4 + 0] = fmaf(dph, wv.x, z[4 * c4 + 0]);
                    y[4 * c4 + 1] = fmaf(ph, wv.y, y[4 * c4 + 1]); z[4 * c4 + 1] = fmaf(dph, wv.y, z[4 * c4 + 1]);
                    y[4 * c4 + 2] = fmaf(ph, wv.z, y[4 * c4 + 2]); z[4 * c4 + 2] = fmaf(dph, wv.z, z[4 * c4 + 2]);
                    y[4 * c4 + 3] = fmaf(ph, wv.w, y[4 * c4 + 3]); z[4 * c4 + 3] = fmaf(dph, wv.w, z[4 * c4 + 3]);
                }
            }
        };
        auto publish_tile = [&]() {   // SiLU / tangent, two-term split, swizzled store, hand-off
            const int ch0 = cg * CPT;                  // first channel (= K index) of this thread
            const int kb = ch0 >> 6;                   // 64-wide K block
            uint8_t* a_hi = smem + G::A_HI + kb * kKBlockBytes;
            uint8_t* a_lo = smem + G::A_LO + kb * kKBlockBytes;
            float hv[CPT], tv[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const float yy = y[i];
                const float sg = sigmoidf_(yy);
                hv[i] = yy * sg;
                tv[i] = sg * (1.0f + yy * (1.0f - sg)) * z[i];
            }
            if constexpr (CPT >= 8) {
#pragma unroll
                for (int j = 0; j < CPT / 8; ++j) {
                    float h8[8], t8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { h8[i] = hv[8 * j + i]; t8[i] = tv[8 * j + i]; }
                    const int chunk = ((ch0 & 63) >> 3) + j;
                    uint4 hi, lo;
                    split8(h8, hi, lo);
                    const uint32_t oh = sw128_offset(p, chunk), ot = sw128_offset(kUmmaPairs + p, chunk);
                    *reinterpret_cast<uint4*>(a_hi + oh) = hi;
                    *reinterpret_cast<uint4*>(a_lo + oh) = lo;
                    split8(t8, hi, lo);
                    *reinterpret_cast<uint4*>(a_hi + ot) = hi;
                    *reinterpret_cast<uint4*>(a_lo + ot) = lo;
                }
            } else {   // 4 channels per thread: half of a 16-byte chunk
                const int chunk = (ch0 & 63) >> 3;
                const uint32_t half8 = (uint32_t)(ch0 & 4) * 2;
                uint2 hi, lo;
                split4(make_float4(hv[0], hv[1], hv[2], hv[3]), hi, lo);
                const uint32_t oh = sw128_offset(p, chunk) + half8, ot = sw128_offset(kUmmaPairs + p, chunk) + half8;
                *reinterpret_cast<uint2*>(a_hi + oh) = hi;
                *reinterpret_cast<uint2*>(a_lo + oh) = lo;
                split4(make_float4(tv[0], tv[1], tv[2], tv[3]), hi, lo);
                *reinterpret_cast<uint2*>(a_hi + ot) = hi;
                *reinterpret_cast<uint2*>(a_lo + ot) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tid == 0) mbar_arrive(bar_a_full);
        };

        first_layer(0);
        publish_tile();
        for (int it = 0; it < my_tiles; ++it) {
            const int p0 = ((int)blockIdx.x + it * (int)gridDim.x) * kUmmaPairs;
            if (it + 1 < my_tiles) first_layer(it + 1);        // overlaps the MMAs of tile `it`
            // ---- epilogue of tile `it`: D[channel = lane][row = column] -> global ----
            const bool tangent = cs >= 2;                      // rows 64..127 hold f'
            const int row0 = p0 + (cs & 1) * 32;               // first pair of this warp's 32 rows
            float* out = (tangent ? dfilt : filt) + (size_t)row0 * (3 * H);
            for (int ci = 0; ci < nchunks; ++ci) {
                const int nc = skip_chunk1 ? ci * 2 : ci;
                mbar_wait(&bar_d_full[nc], it & 1);
                __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nc * 128 + cs * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int gch = nc * 128 + q * 32 + lane;       // filter channel of this lane
                bool store = gch < 3 * H;
                if (skip_vector_gate && gch >= H && gch < 2 * H) store = false;   // unused b gate of layer 0
                if (store) {
                    const float b = tangent ? 0.f : bias[nc];
                    float* o = out + gch;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (row0 + j < P) __stcs(o + (size_t)j * (3 * H), fmaf(__uint_as_float(r[j]), kAccUnscale, b));
                }
            }
            // all MMAs of tile `it` are complete (last chunk waited): the activation tile is free
            if (it + 1 < my_tiles) publish_tile();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == kUmmaComputeWarps) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G::TMEM_COLS));
    }
}

}  // namespace mlffd
