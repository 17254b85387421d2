// Shared device helpers for the PaiNN-student kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlffd {

constexpr int kNumSMs = 148;       // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kMaxRbf = 32;
constexpr int kMaxLayers = 8;
constexpr float kUnitEps = 1e-8f;  // r / (d + 1e-8): student_model.py:715

// Device-resident counters written by the neighbour build and read by every later kernel, so
// the host never has to synchronise in the middle of a step.
struct DeviceStatus {
    int num_edges;
    int num_pairs;
    int overflow;
    int max_degree;
    int overflow_events;   // sticky: number of neighbour builds that overflowed since creation
    int tc_saturated;      // a tensor-core operand left the FP16 range in this step: dense-layer outputs invalid
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// SiLU as torch evaluates it: x * sigmoid(x).
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }

// d/dx [x sigmoid(x)] = s (1 + x (1 - s))
__device__ __forceinline__ float silu_gradf_(float x) {
    const float s = sigmoidf_(x);
    return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// Streaming (evict-first) 16-byte load/store for data that is touched once per kernel -- the
// filter-table rows -- so it does not push the gathered per-atom feature rows out of L2.
__device__ __forceinline__ float4 ldcs4(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void stcs4(float* p, const float4& v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ void st4(float* p, const float4& v) {
    *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ float4 make4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float4 fma4(const float4& a, const float4& b, const float4& c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z),
                       fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 fma4s(float a, const float4& b, const float4& c) {
    return make_float4(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y), fmaf(a, b.z, c.z), fmaf(a, b.w, c.w));
}
__device__ __forceinline__ float4 mul4(const float4& a, const float4& b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float hsum4(const float4& a) { return (a.x + a.y) + (a.z + a.w); }

// Sum over the `width` consecutive lanes of a sub-warp group (width = 8, 16 or 32).
template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Pair geometry with a FIXED FP32 operation order (no FMA contraction) so the edge decision is
// reproducible bit-for-bit by the oracle (oracle/painn_oracle.py:neighbor_list).
// cell18 = row-major 3x3 cell followed by its row-major inverse.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void min_image(float& dx, float& dy, float& dz, const float* cell18,
                                          unsigned pbc_mask) {
    const float* c = cell18;
    const float* inv = cell18 + 9;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (pbc_mask & 1u)
        n0 = rintf(__fadd_rn(__fadd_rn(__fmul_rn(dx, inv[0]), __fmul_rn(dy, inv[3])),
                             __fmul_rn(dz, inv[6])));
    if (pbc_mask & 2u)
        n1 = rintf(__fadd_rn(__fadd_rn(__fmul_rn(dx, inv[1]), __fmul_rn(dy, inv[4])),
                             __fmul_rn(dz, inv[7])));
    if (pbc_mask & 4u)
        n2 = rintf(__fadd_rn(__fadd_rn(__fmul_rn(dx, inv[2]), __fmul_rn(dy, inv[5])),
                             __fmul_rn(dz, inv[8])));
    const float sx = __fadd_rn(__fadd_rn(__fmul_rn(n0, c[0]), __fmul_rn(n1, c[3])), __fmul_rn(n2, c[6]));
    const float sy = __fadd_rn(__fadd_rn(__fmul_rn(n0, c[1]), __fmul_rn(n1, c[4])), __fmul_rn(n2, c[7]));
    const float sz = __fadd_rn(__fadd_rn(__fmul_rn(n0, c[2]), __fmul_rn(n1, c[5])), __fmul_rn(n2, c[8]));
    dx = __fsub_rn(dx, sx);
    dy = __fsub_rn(dy, sy);
    dz = __fsub_rn(dz, sz);
}

__device__ __forceinline__ float pair_distance(float dx, float dy, float dz) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

}  // namespace mlffd
