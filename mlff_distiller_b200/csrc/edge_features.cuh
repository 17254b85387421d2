// Operator-level stage kernels for the edge features: drop-ins for the reference's two Triton kernels
// (kernels/fused_edge_features.py:25-96, kernels/fused_rbf_cutoff.py:26-88).  On the product path the same
// arithmetic is fused into the neighbour fill pass (neighbor.cuh) and the filter evaluation; these entry
// points exist for callers of the reference's L1 kernel API (StudentForceFieldOptimized.forward,
// student_model_optimized.py:117-140) and for stage-level parity tests.
//
// The Triton kernels launch one program per EDGE with scalar loads (grid (E,), 12 scalar stores per edge);
// here a thread owns an edge, index loads are coalesced 8-byte loads, and the three [E,3] outputs leave
// through shared memory so that every global store of a warp is one contiguous 128-byte line.
#pragma once
#include "common.cuh"

namespace mlffd {

constexpr int kEdgeFeatThreads = 256;

// eps_mode 0: the model's placement (student_model.py:712-715)  d = |r|,  u = r / (d + eps)
// eps_mode 1: the Triton kernel's placement (fused_edge_features.py:77-82)  d = sqrt(|r|^2 + eps),  u = r / d
__global__ void __launch_bounds__(kEdgeFeatThreads)
edge_features_kernel(const float* __restrict__ pos, const long long* __restrict__ src_idx,
                     const long long* __restrict__ dst_idx, long long num_edges, float eps, int eps_mode,
                     float* __restrict__ edge_vec, float* __restrict__ dist, float* __restrict__ unit) {
    __shared__ float stage[2][3 * kEdgeFeatThreads];
    for (long long base = (long long)blockIdx.x * kEdgeFeatThreads; base < num_edges;
         base += (long long)gridDim.x * kEdgeFeatThreads) {
        const long long e = base + threadIdx.x;
        if (e < num_edges) {
            const long long s = __ldg(src_idx + e), t = __ldg(dst_idx + e);
            const float rx = __ldg(pos + 3 * s) - __ldg(pos + 3 * t);
            const float ry = __ldg(pos + 3 * s + 1) - __ldg(pos + 3 * t + 1);
            const float rz = __ldg(pos + 3 * s + 2) - __ldg(pos + 3 * t + 2);
            float d, inv;
            if (eps_mode == 0) {
                d = pair_distance(rx, ry, rz);     // the fixed-order FP32 distance of the neighbour list
                inv = 1.0f / (d + eps);
            } else {
                d = sqrtf(rx * rx + ry * ry + rz * rz + eps);
                inv = 1.0f / d;
            }
            dist[e] = d;
            float* v = stage[0] + 3 * threadIdx.x;
            float* u = stage[1] + 3 * threadIdx.x;
            v[0] = rx; v[1] = ry; v[2] = rz;
            if (eps_mode == 0) { u[0] = rx / (d + eps); u[1] = ry / (d + eps); u[2] = rz / (d + eps); }
            else { u[0] = rx * inv; u[1] = ry * inv; u[2] = rz * inv; }
        }
        __syncthreads();
        const long long count = min((long long)kEdgeFeatThreads, num_edges - base) * 3;
        for (int i = threadIdx.x; i < count; i += kEdgeFeatThreads) {
            edge_vec[3 * base + i] = stage[0][i];
            unit[3 * base + i] = stage[1][i];
        }
        __syncthreads();
    }
}

// out[e][k] = exp(-gamma (d_e - mu_k)^2) * 0.5 (cos(pi d_e / rc) + 1) [d_e < rc]
// (GaussianRBF x CosineCutoff, student_model.py:249-255, 285-292; op order as filter.cuh:rbf_cutoff).
// One thread per output element: consecutive threads write consecutive floats.
__global__ void __launch_bounds__(256)
rbf_cutoff_kernel(const float* __restrict__ dist, long long num_edges, const float* __restrict__ centers,
                  int num_rbf, float gamma, float rc, float* __restrict__ out) {
    const float kPi = 3.14159274101257324f;
    const long long total = num_edges * num_rbf;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx / num_rbf;
        const int k = (int)(idx - e * num_rbf);
        const float d = __ldg(dist + e);
        const float diff = d - __ldg(centers + k);
        const float fc = (d < rc) ? 0.5f * (cosf((kPi * d) / rc) + 1.0f) : 0.0f;
        out[idx] = expf(-gamma * (diff * diff)) * fc;
    }
}

}  // namespace mlffd
