// FP32 register-tiled GEMM building block shared by the filter-table and update kernels.
//
// A block of 256 threads owns a tile of TM = 64 rows.  The activation tile lives in shared
// memory k-major (A_s[k][row], row stride AS = 68 floats so float4 row reads stay aligned), the
// weights are streamed from L2 in [KC = H][H] sub-blocks (W_s[k][n]).  Thread (ty, tx) =
// (tid / 16, tid % 16) accumulates rows ty*4..ty*4+3 and H/16 columns in registers:
//   H = 128: 8 columns  {tx*4..+3} and {64 + tx*4..+3}   (two float4 groups, conflict-free)
//   H =  64: 4 columns  {tx*4..+3}
//   H =  32: 2 columns  {tx*2..+1}
// This is the FFMA reference path for every dense layer; the tensor-core variants replace the
// inner product only and keep this tiling contract (row tile in, H-wide column chunk out).
#pragma once
#include "common.cuh"

namespace mlffd {

constexpr int kTileRows = 64;
constexpr int kAStride = kTileRows + 4;
constexpr int kGemmThreads = 256;

template <int H>
struct TileTraits {
    static constexpr int RN = H / 16;               // columns per thread
    static constexpr int VW = (RN >= 4) ? 4 : RN;   // vector width of a column group
    static constexpr int NG = RN / VW;              // column groups per thread
    static constexpr int GROUP_STRIDE = 16 * VW;    // column distance between groups
    static_assert(H == 32 || H == 64 || H == 128, "hidden_dim must be 32, 64 or 128");
};

// column index (within the H-wide chunk) of accumulator c of thread tx
template <int H>
__device__ __forceinline__ int tile_col(int tx, int c) {
    using T = TileTraits<H>;
    return (c / T::VW) * T::GROUP_STRIDE + tx * T::VW + (c % T::VW);
}

// Copy W[k0 .. k0+kc)[n0 .. n0+H) of a row-major [*, ldw] matrix into W_s[kc][H].
template <int H>
__device__ __forceinline__ void load_weight_chunk(float* __restrict__ W_s,
                                                  const float* __restrict__ W, int ldw, int k0,
                                                  int n0, int kc) {
    constexpr int V = H / 4;  // float4 per row
    for (int idx = threadIdx.x; idx < kc * V; idx += kGemmThreads) {
        const int r = idx / V, c4 = idx - r * V;
        const float4 v = ldg4(W + (size_t)(k0 + r) * ldw + n0 + 4 * c4);
        st4(W_s + r * H + 4 * c4, v);
    }
}

// acc[r][c] += sum_{k < kc} A_s[(k)*AS + ty*4 + r] * W_s[k*H + col(c)]
template <int H>
__device__ __forceinline__ void tile_fma(float (&acc)[4][TileTraits<H>::RN],
                                         const float* __restrict__ A_s,
                                         const float* __restrict__ W_s, int kc, int ty, int tx) {
    using T = TileTraits<H>;
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(A_s + k * kAStride + ty * 4);
        float w[T::RN];
#pragma unroll
        for (int g = 0; g < T::NG; ++g) {
            const float* wp = W_s + k * H + g * T::GROUP_STRIDE + tx * T::VW;
            if constexpr (T::VW == 4) {
                const float4 t = *reinterpret_cast<const float4*>(wp);
                w[g * 4 + 0] = t.x; w[g * 4 + 1] = t.y; w[g * 4 + 2] = t.z; w[g * 4 + 3] = t.w;
            } else {
                const float2 t = *reinterpret_cast<const float2*>(wp);
                w[g * 2 + 0] = t.x; w[g * 2 + 1] = t.y;
            }
        }
#pragma unroll
        for (int c = 0; c < T::RN; ++c) {
            acc[0][c] = fmaf(a.x, w[c], acc[0][c]);
            acc[1][c] = fmaf(a.y, w[c], acc[1][c]);
            acc[2][c] = fmaf(a.z, w[c], acc[2][c]);
            acc[3][c] = fmaf(a.w, w[c], acc[3][c]);
        }
    }
}

// vector load/store of one column group (VW = 4 or 2 consecutive floats)
template <int VW>
__device__ __forceinline__ void ldv(const float* p, float (&v)[VW]) {
    if constexpr (VW == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
}
template <int VW>
__device__ __forceinline__ void stv(float* p, const float (&v)[VW]) {
    if constexpr (VW == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
}

template <int H>
__device__ __forceinline__ void tile_zero(float (&acc)[4][TileTraits<H>::RN]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < TileTraits<H>::RN; ++c) acc[r][c] = 0.f;
}

}  // namespace mlffd
