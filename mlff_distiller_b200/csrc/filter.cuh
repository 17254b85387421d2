// Radial filter table: for every undirected pair distance d, the per-layer PaiNN filter
//   f(d)  = W2 SiLU(W1 phi~(d) + b1) + b2                      in R^{3H}
//   f'(d) = W2 ( SiLU'(W1 phi~(d) + b1) * (W1 phi~'(d)) )      (forward-mode d/dd)
// with phi~_k(d) = exp(-gamma_k (d - mu_k)^2) * 0.5 (cos(pi d / rc) + 1) [d < rc].
//
// Restates GaussianRBF.forward, CosineCutoff.forward and PaiNNMessage.rbf_to_scalar of the
// reference (src/mlff_distiller/models/student_model.py:249-255, 285-292, 318-322, 350) and the
// closed-form derivative of src/mlff_distiller/models/analytical_gradients.py:249-262.
//
// Why a table: the filter depends on the scalar distance only (no atom features), and the edge
// set is symmetric, so it is evaluated ONCE per undirected pair and both directed edges, the
// forward message pass and the reverse pass all read it.  Carrying f' next to f turns the whole
// backward through the filter MLP into a 3H-long dot product per edge (d_bar = g_bar . f'), so
// no reverse GEMM and no per-edge adjoint tensor is ever materialised.
//
// Tiling: 32 pairs per tile -> 64 GEMM rows (32 x h, 32 x t) against W2^T in three H-wide
// column chunks (tile_gemm.cuh).
#pragma once
#include "tile_gemm.cuh"

namespace mlffd {

constexpr int kFilterPairs = 32;
constexpr int kPhiStride = kMaxRbf + 1;

struct FilterWeights {
    const float* W1t;  // [K][H]   (transposed rbf_to_scalar.0.weight)
    const float* b1;   // [H]
    const float* W2t;  // [H][3H]  (transposed rbf_to_scalar.2.weight)
    const float* b2;   // [3H]
};

template <int H>
constexpr size_t filter_smem_bytes() {
    return sizeof(float) * (size_t)(H * kAStride + H * H + 2 * kFilterPairs * kPhiStride);
}

// phi~ and d phi~ / dd for one (distance, basis) pair; op order of the forward value follows the
// reference: exp((-gamma) * (diff*diff)), 0.5 * (cos((pi*d)/rc) + 1) * [d < rc].
__device__ __forceinline__ void rbf_cutoff(float d, float mu, float gamma, float rc, float& val,
                                           float& dval) {
    const float kPi = 3.14159274101257324f;  // float32(np.pi)
    const float diff = d - mu;
    const float phi = expf(-gamma * (diff * diff));
    const float arg = (kPi * d) / rc;
    const bool inside = d < rc;
    const float fc = inside ? 0.5f * (cosf(arg) + 1.0f) : 0.0f;
    const float dfc = inside ? -0.5f * (kPi / rc) * sinf(arg) : 0.0f;
    val = phi * fc;
    dval = (-2.0f * gamma * diff * phi) * fc + phi * dfc;
}

template <int H>
__global__ void __launch_bounds__(kGemmThreads)
filter_table_kernel(const float* __restrict__ pair_dist, const int* __restrict__ num_pairs_ptr,
                    int num_pairs_arg, const DeviceStatus* __restrict__ status,
                    const float* __restrict__ centers, const float* __restrict__ gammas, int K,
                    float rc, FilterWeights w, int skip_vector_gate,
                    float* __restrict__ filt, float* __restrict__ dfilt) {
    using T = TileTraits<H>;
    if (status != nullptr && status->overflow) return;
    const int P = (num_pairs_ptr != nullptr) ? *num_pairs_ptr : num_pairs_arg;
    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;                               // [H][AS]   h rows 0..31, t rows 32..63
    float* W_s = A_s + H * kAStride;                 // [H][H]    (first holds W1t [K][H])
    float* phi_s = W_s + H * H;                      // [32][33]
    float* dphi_s = phi_s + kFilterPairs * kPhiStride;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int num_tiles = (P + kFilterPairs - 1) / kFilterPairs;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int p0 = tile * kFilterPairs;
        // ---- phase 0: radial basis and its derivative; stage W1^T ----
        for (int idx = tid; idx < kFilterPairs * K; idx += kGemmThreads) {
            const int p = idx / K, k = idx - p * K;
            const float d = (p0 + p < P) ? pair_dist[p0 + p] : rc;
            float v, dv;
            rbf_cutoff(d, __ldg(centers + k), __ldg(gammas + k), rc, v, dv);
            phi_s[p * kPhiStride + k] = v;
            dphi_s[p * kPhiStride + k] = dv;
        }
        for (int idx = tid; idx < K * H / 4; idx += kGemmThreads)
            st4(W_s + 4 * idx, ldg4(w.W1t + 4 * idx));
        __syncthreads();
        // ---- phase A: hidden layer, value and tangent ----
        {
            constexpr int CPT = H / 8;  // channels per thread: 8 channel groups x 32 pairs
            const int p = tid & 31, c0 = (tid >> 5) * CPT;
            float y[CPT], z[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) { y[c] = __ldg(w.b1 + c0 + c); z[c] = 0.f; }
            for (int k = 0; k < K; ++k) {
                const float ph = phi_s[p * kPhiStride + k], dph = dphi_s[p * kPhiStride + k];
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const float wv = W_s[k * H + c0 + c];
                    y[c] = fmaf(ph, wv, y[c]);
                    z[c] = fmaf(dph, wv, z[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const float s = sigmoidf_(y[c]);
                A_s[(c0 + c) * kAStride + p] = y[c] * s;
                A_s[(c0 + c) * kAStride + kFilterPairs + p] = s * (1.0f + y[c] * (1.0f - s)) * z[c];
            }
        }
        __syncthreads();
        // ---- phase B: [h; t] x W2^T, three H-wide column chunks (a | b | c) ----
        for (int nc = 0; nc < 3; ++nc) {
            if (nc == 1 && skip_vector_gate) continue;  // layer 0: v_in == 0, b is never read
            load_weight_chunk<H>(W_s, w.W2t, 3 * H, 0, nc * H, H);
            __syncthreads();
            float acc[4][T::RN];
            tile_zero<H>(acc);
            tile_fma<H>(acc, A_s, W_s, H, ty, tx);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = ty * 4 + r;
                const bool tangent = m >= kFilterPairs;
                const int p = p0 + (tangent ? m - kFilterPairs : m);
                if (p >= P) continue;
                float* out = (tangent ? dfilt : filt) + (size_t)p * 3 * H + nc * H;
#pragma unroll
                for (int g = 0; g < T::NG; ++g) {
                    const int col = g * T::GROUP_STRIDE + tx * T::VW;
                    float v[T::VW];
#pragma unroll
                    for (int q = 0; q < T::VW; ++q)
                        v[q] = acc[r][g * T::VW + q] + (tangent ? 0.f : __ldg(w.b2 + nc * H + col + q));
                    if constexpr (T::VW == 4)
                        stcs4(out + col, make_float4(v[0], v[1], v[2], v[3]));
                    else
                        __stcs(reinterpret_cast<float2*>(out + col), make_float2(v[0], v[1]));
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace mlffd
