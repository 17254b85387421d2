// Embedding lookup, energy read-out (forward and reverse fused), per-structure energy sums and
// the final edge-geometry reverse + force gather.
//
// Read-out restates energy_head (reference src/mlff_distiller/models/student_model.py:599-605,
// 736-755): eps_j = A3 SiLU(A2 SiLU(A1 s_j + a1) + a2) + a3, E_b = sum_{j in b} eps_j.  Since
// dE/d eps_j = 1 for every atom, the reverse of the head is computed in the same kernel and
// seeds the adjoint s_bar of the last layer.
//
// Forces restate what autograd does through student_model.py:709-715 (r = x_src - x_dst,
// d = |r|, u = r / (d + 1e-8)):  r_bar = (d_bar - (u_bar . u) / q) * (r / d) + u_bar / q,
// q = d + 1e-8;  F_j = sum_{e -> j} (r_bar_e - r_bar_rev(e))  -- a CSR gather through the
// reverse-edge index instead of the +/- atomic scatter of
// analytical_gradients.accumulate_forces_from_edges (:325-339).
#pragma once
#include "common.cuh"
#include "tile_gemm.cuh"

namespace mlffd {

struct HeadWeights {
    const float* A1t;  // [H][H/2]   transposed energy_head.0.weight
    const float* a1;   // [H/2]
    const float* A2t;  // [H/2][H/4] transposed energy_head.2.weight
    const float* a2;   // [H/4]
    const float* A3;   // [H/4]      energy_head.4.weight
    const float* a3;   // [1]
    const float* A1;   // [H/2][H]   original layout (reverse)
    const float* A2;   // [H/4][H/2] original layout (reverse)
};

template <int H>
__global__ void __launch_bounds__(256)
embedding_kernel(const int* __restrict__ z, const float* __restrict__ emb, int max_z,
                 float* __restrict__ s0, int num_atoms) {
    constexpr int V = H / 4;
    const size_t total = (size_t)num_atoms * V;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int atom = (int)(idx / V), c4 = (int)(idx % V) * 4;
        int zi = __ldg(z + atom);
        zi = min(max(zi, 0), max_z);
        st4(s0 + (size_t)atom * H + c4, ldg4(emb + (size_t)zi * H + c4));
    }
}

// A warp evaluates the head for AT = 4 atoms at a time (grid-stride), so every weight fetched
// from L1 feeds four FMAs: with one atom per warp the kernel was bound by L1 bandwidth (one
// 4-byte weight load per FMA).  Writes eps[atom] and, if sbar != nullptr, d eps / d s.
template <int H>
__global__ void __launch_bounds__(256)
readout_kernel(const float* __restrict__ s, HeadWeights w, float* __restrict__ eps,
               float* __restrict__ sbar, int num_atoms, const DeviceStatus* __restrict__ status) {
    constexpr int H2 = H / 2, H4 = H / 4;
    constexpr int WARPS = 8, AT = 4;
    if (status->overflow) return;
    __shared__ float s_sh[WARPS][AT][H];
    __shared__ float y1_sh[WARPS][AT][H2];   // pre-activations
    __shared__ float h1_sh[WARPS][AT][H2];   // activations, then adjoints of y1
    __shared__ float y2_sh[WARPS][AT][H4];   // adjoints of the layer-2 pre-activations
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * WARPS + wib;
    const int num_warps = gridDim.x * WARPS;
    for (int atom0 = warp * AT; atom0 < num_atoms; atom0 += num_warps * AT) {
#pragma unroll
        for (int a = 0; a < AT; ++a)
            for (int c = lane; c < H; c += 32)
                s_sh[wib][a][c] = (atom0 + a < num_atoms) ? __ldg(s + (size_t)(atom0 + a) * H + c) : 0.f;
        __syncwarp();
        // layer 1: H -> H/2
        for (int o = lane; o < H2; o += 32) {
            float y[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) y[a] = __ldg(w.a1 + o);
#pragma unroll 4
            for (int k = 0; k < H; ++k) {
                const float wv = __ldg(w.A1t + k * H2 + o);
#pragma unroll
                for (int a = 0; a < AT; ++a) y[a] = fmaf(s_sh[wib][a][k], wv, y[a]);
            }
#pragma unroll
            for (int a = 0; a < AT; ++a) { y1_sh[wib][a][o] = y[a]; h1_sh[wib][a][o] = siluf_(y[a]); }
        }
        __syncwarp();
        // layer 2: H/2 -> H/4, layer 3: H/4 -> 1
        float part[AT];
#pragma unroll
        for (int a = 0; a < AT; ++a) part[a] = 0.f;
        for (int o = lane; o < H4; o += 32) {
            float y[AT];
#pragma unroll
            for (int a = 0; a < AT; ++a) y[a] = __ldg(w.a2 + o);
#pragma unroll 4
            for (int k = 0; k < H2; ++k) {
                const float wv = __ldg(w.A2t + k * H4 + o);
#pragma unroll
                for (int a = 0; a < AT; ++a) y[a] = fmaf(h1_sh[wib][a][k], wv, y[a]);
            }
            const float a3 = __ldg(w.A3 + o);
#pragma unroll
            for (int a = 0; a < AT; ++a) {
                part[a] = fmaf(siluf_(y[a]), a3, part[a]);
                y2_sh[wib][a][o] = a3 * silu_gradf_(y[a]);   // adjoint of the layer-2 pre-activation
            }
        }
#pragma unroll
        for (int a = 0; a < AT; ++a) {
            part[a] = group_sum<32>(part[a]);
            if (lane == 0 && atom0 + a < num_atoms) eps[atom0 + a] = part[a] + __ldg(w.a3);
        }
        __syncwarp();
        if (sbar != nullptr) {
            // y1_bar = (A2^T y2_bar) * SiLU'(y1)
            for (int k = lane; k < H2; k += 32) {
                float hb[AT];
#pragma unroll
                for (int a = 0; a < AT; ++a) hb[a] = 0.f;
#pragma unroll 4
                for (int o = 0; o < H4; ++o) {
                    const float wv = __ldg(w.A2 + o * H2 + k);
#pragma unroll
                    for (int a = 0; a < AT; ++a) hb[a] = fmaf(y2_sh[wib][a][o], wv, hb[a]);
                }
#pragma unroll
                for (int a = 0; a < AT; ++a) h1_sh[wib][a][k] = hb[a] * silu_gradf_(y1_sh[wib][a][k]);
            }
            __syncwarp();
            for (int c = lane; c < H; c += 32) {
                float sb[AT];
#pragma unroll
                for (int a = 0; a < AT; ++a) sb[a] = 0.f;
#pragma unroll 4
                for (int k = 0; k < H2; ++k) {
                    const float wv = __ldg(w.A1 + k * H + c);
#pragma unroll
                    for (int a = 0; a < AT; ++a) sb[a] = fmaf(h1_sh[wib][a][k], wv, sb[a]);
                }
#pragma unroll
                for (int a = 0; a < AT; ++a)
                    if (atom0 + a < num_atoms) sbar[(size_t)(atom0 + a) * H + c] = sb[a];
            }
        }
        __syncwarp();
    }
}

// Latency variant for small systems (single-trajectory MD): one BLOCK per 4 atoms, every small
// GEMV of the head split over all 256 threads (output x K-slice, partial sums combined through
// shared memory), so the dependent chain is <= H/8 FMAs per layer instead of H with one warp per 4
// atoms (readout_kernel: 43 us for a 3-atom system, measured).  Same math, same outputs.
template <int H>
__global__ void __launch_bounds__(256)
readout_block_kernel(const float* __restrict__ s, HeadWeights w, float* __restrict__ eps,
                     float* __restrict__ sbar, int num_atoms, const DeviceStatus* __restrict__ status) {
    constexpr int H2 = H / 2, H4 = H / 4, AT = 4, T = 256;
    if (status->overflow) return;
    __shared__ float s_sh[AT][H];
    __shared__ float y1_sh[AT][H2];
    __shared__ float h1_sh[AT][H2];    // activations, then adjoints of y1
    __shared__ float y2b_sh[AT][H4];   // adjoints of the layer-2 pre-activations
    __shared__ float part[T * AT];     // [slice][atom][output]
    const int t = threadIdx.x;
    const int atom0 = (int)blockIdx.x * AT;
    if (atom0 >= num_atoms) return;
    for (int idx = t; idx < AT * H; idx += T) {
        const int a = idx / H, c = idx - a * H;
        s_sh[a][c] = (atom0 + a < num_atoms) ? __ldg(s + (size_t)(atom0 + a) * H + c) : 0.f;
    }
    __syncthreads();
    {   // layer 1: H -> H/2
        constexpr int NSL = (T / H2 < H) ? T / H2 : H, KPER = H / NSL;
        const int o = t % H2, sl = t / H2;
        float y[AT] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < (sl < NSL ? KPER : 0); ++kk) {
            const int k = sl * KPER + kk;
            const float wv = __ldg(w.A1t + k * H2 + o);
#pragma unroll
            for (int a = 0; a < AT; ++a) y[a] = fmaf(s_sh[a][k], wv, y[a]);
        }
#pragma unroll
        for (int a = 0; a < AT; ++a) if (sl < NSL) part[(sl * AT + a) * H2 + o] = y[a];
        __syncthreads();
        if (t < AT * H2) {
            const int a = t / H2, oo = t - a * H2;
            float yy = __ldg(w.a1 + oo);
#pragma unroll
            for (int q = 0; q < NSL; ++q) yy += part[(q * AT + a) * H2 + oo];
            y1_sh[a][oo] = yy;
            h1_sh[a][oo] = siluf_(yy);
        }
        __syncthreads();
    }
    {   // layer 2: H/2 -> H/4, layer 3: H/4 -> 1
        constexpr int NSL = (T / H4 < H2) ? T / H4 : H2, KPER = H2 / NSL;
        const int o = t % H4, sl = t / H4;
        float y[AT] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < (sl < NSL ? KPER : 0); ++kk) {
            const int k = sl * KPER + kk;
            const float wv = __ldg(w.A2t + k * H4 + o);
#pragma unroll
            for (int a = 0; a < AT; ++a) y[a] = fmaf(h1_sh[a][k], wv, y[a]);
        }
#pragma unroll
        for (int a = 0; a < AT; ++a) if (sl < NSL) part[(sl * AT + a) * H4 + o] = y[a];
        __syncthreads();
        if (t < AT * H4) {   // H4 consecutive threads of one warp per atom
            const int a = t / H4, oo = t - a * H4;
            float yy = __ldg(w.a2 + oo);
#pragma unroll
            for (int q = 0; q < NSL; ++q) yy += part[(q * AT + a) * H4 + oo];
            const float a3 = __ldg(w.A3 + oo);
            y2b_sh[a][oo] = a3 * silu_gradf_(yy);
            const float e = group_sum<H4>(siluf_(yy) * a3);
            if (oo == 0 && atom0 + a < num_atoms) eps[atom0 + a] = e + __ldg(w.a3);
        }
        __syncthreads();
    }
    if (sbar == nullptr) return;
    {   // y1_bar = (A2^T y2_bar) * SiLU'(y1)
        constexpr int NSL = (T / H2 < H4) ? T / H2 : H4, OPER = H4 / NSL;
        const int k = t % H2, sl = t / H2;
        float hb[AT] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int oo = 0; oo < (sl < NSL ? OPER : 0); ++oo) {
            const int o = sl * OPER + oo;
            const float wv = __ldg(w.A2 + o * H2 + k);
#pragma unroll
            for (int a = 0; a < AT; ++a) hb[a] = fmaf(y2b_sh[a][o], wv, hb[a]);
        }
#pragma unroll
        for (int a = 0; a < AT; ++a) if (sl < NSL) part[(sl * AT + a) * H2 + k] = hb[a];
        __syncthreads();
        if (t < AT * H2) {
            const int a = t / H2, kk = t - a * H2;
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < NSL; ++q) v += part[(q * AT + a) * H2 + kk];
            h1_sh[a][kk] = v * silu_gradf_(y1_sh[a][kk]);
        }
        __syncthreads();
    }
    {   // s_bar = A1^T y1_bar
        constexpr int NSL = (T / H < H2) ? T / H : H2, KPER = H2 / NSL;
        const int c = t % H, sl = t / H;
        float sb[AT] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < (sl < NSL ? KPER : 0); ++kk) {
            const int k = sl * KPER + kk;
            const float wv = __ldg(w.A1 + k * H + c);
#pragma unroll
            for (int a = 0; a < AT; ++a) sb[a] = fmaf(h1_sh[a][k], wv, sb[a]);
        }
#pragma unroll
        for (int a = 0; a < AT; ++a) if (sl < NSL) part[(sl * AT + a) * H + c] = sb[a];
        __syncthreads();
        for (int idx = t; idx < AT * H; idx += T) {
            const int a = idx / H, cc = idx - a * H;
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < NSL; ++q) v += part[(q * AT + a) * H + cc];
            if (atom0 + a < num_atoms) sbar[(size_t)(atom0 + a) * H + cc] = v;
        }
    }
}

// Tiled variant of readout_kernel for H = 128: 64 atoms per block, the four small GEMMs (head
// forward H -> H/2 -> H/4, reverse H/4 -> H/2 -> H) on the register-tiled FFMA block of
// tile_gemm.cuh with weights staged in shared memory, so every weight fetched feeds 4 x RN FMAs
// instead of 4 (the warp-per-4-atoms kernel above is bound by L1 weight loads).  Pre-activations
// stay in registers between the forward and the reverse half (same thread <-> same rows/columns).
constexpr size_t readout_tile_smem_bytes() { return sizeof(float) * (size_t)(128 * kAStride + 128 * 64); }

__global__ void __launch_bounds__(kGemmThreads)
readout_tile_kernel(const float* __restrict__ s, HeadWeights w, float* __restrict__ eps,
                    float* __restrict__ sbar, int num_atoms, const DeviceStatus* __restrict__ status) {
    constexpr int H = 128, H2 = 64, H4 = 32;
    if (status->overflow) return;
    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;                      // [K <= 128][AS]  activations, k-major
    float* W_s = A_s + H * kAStride;        // [K][N] weight chunk, K*N <= 8192
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int num_tiles = (num_atoms + kTileRows - 1) / kTileRows;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int a0 = tile * kTileRows;
        __syncthreads();   // previous tile done with A_s / W_s
        // stage s rows k-major (row fastest across threads: conflict-free smem stores)
        for (int idx = tid; idx < kTileRows * (H / 4); idx += kGemmThreads) {
            const int m = idx % kTileRows, c4 = idx / kTileRows;
            const int atom = a0 + m;
            const float4 v = (atom < num_atoms) ? ldg4(s + (size_t)atom * H + 4 * c4) : make4(0.f);
            A_s[(4 * c4 + 0) * kAStride + m] = v.x;
            A_s[(4 * c4 + 1) * kAStride + m] = v.y;
            A_s[(4 * c4 + 2) * kAStride + m] = v.z;
            A_s[(4 * c4 + 3) * kAStride + m] = v.w;
        }
        load_weight_chunk<H2>(W_s, w.A1t, H2, 0, 0, H);
        __syncthreads();
        // ---- y1 = s A1^T + a1  (K = 128, N = 64) ----
        float y1[4][TileTraits<H2>::RN];
        tile_zero<H2>(y1);
        tile_fma<H2>(y1, A_s, W_s, H, ty, tx);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < TileTraits<H2>::RN; ++c) {
            const int n = tile_col<H2>(tx, c);
            const float b = __ldg(w.a1 + n);
            float h[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { y1[r][c] += b; h[r] = siluf_(y1[r][c]); }
            st4(A_s + n * kAStride + ty * 4, make_float4(h[0], h[1], h[2], h[3]));
        }
        load_weight_chunk<H4>(W_s, w.A2t, H4, 0, 0, H2);
        __syncthreads();
        // ---- y2 = h1 A2^T + a2  (K = 64, N = 32);  eps = A3 . SiLU(y2) + a3 ----
        float y2[4][TileTraits<H4>::RN];
        tile_zero<H4>(y2);
        tile_fma<H4>(y2, A_s, W_s, H2, ty, tx);
        __syncthreads();
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < TileTraits<H4>::RN; ++c) {
            const int n = tile_col<H4>(tx, c);
            const float b = __ldg(w.a2 + n), a3 = __ldg(w.A3 + n);
            float yb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float y = y2[r][c] + b;
                part[r] = fmaf(siluf_(y), a3, part[r]);
                yb[r] = a3 * silu_gradf_(y);          // adjoint of the layer-2 pre-activation
            }
            st4(A_s + n * kAStride + ty * 4, make_float4(yb[0], yb[1], yb[2], yb[3]));
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);   // over tx
            const int atom = a0 + ty * 4 + r;
            if (tx == 0 && atom < num_atoms) eps[atom] = part[r] + __ldg(w.a3);
        }
        if (sbar == nullptr) continue;
        load_weight_chunk<H2>(W_s, w.A2, H2, 0, 0, H4);
        __syncthreads();
        // ---- y1_bar = (y2_bar A2) * SiLU'(y1)  (K = 32, N = 64) ----
        float hb[4][TileTraits<H2>::RN];
        tile_zero<H2>(hb);
        tile_fma<H2>(hb, A_s, W_s, H4, ty, tx);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < TileTraits<H2>::RN; ++c) {
            const int n = tile_col<H2>(tx, c);
            float v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = hb[r][c] * silu_gradf_(y1[r][c]);
            st4(A_s + n * kAStride + ty * 4, make_float4(v[0], v[1], v[2], v[3]));
        }
        load_weight_chunk<H>(W_s, w.A1, H, 0, 0, H2);
        __syncthreads();
        // ---- s_bar = y1_bar A1  (K = 64, N = 128) ----
        float sb[4][TileTraits<H>::RN];
        tile_zero<H>(sb);
        tile_fma<H>(sb, A_s, W_s, H2, ty, tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int atom = a0 + ty * 4 + r;
            if (atom >= num_atoms) continue;
#pragma unroll
            for (int g = 0; g < TileTraits<H>::NG; ++g) {
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = sb[r][g * 4 + q];
                stv<4>(sbar + (size_t)atom * H + g * TileTraits<H>::GROUP_STRIDE + tx * 4, v);
            }
        }
    }
}

// One warp per structure: E_b = sum of eps over the structure's atoms (FP64 accumulation, fixed
// order -> deterministic).
__global__ void __launch_bounds__(256)
structure_energy_kernel(const float* __restrict__ eps, const int* __restrict__ offsets,
                        int num_structures, float* __restrict__ energy,
                        const DeviceStatus* __restrict__ status) {
    if (status->overflow) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int num_warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < num_structures; b += num_warps) {
        const int lo = __ldg(offsets + b), hi = __ldg(offsets + b + 1);
        double acc = 0.0;
        for (int i = lo + lane; i < hi; i += 32) acc += (double)__ldg(eps + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) energy[b] = (float)acc;
    }
}

__device__ __forceinline__ float3 edge_position_adjoint(const float4 g, const float4 adj) {
    // g = (u_x, u_y, u_z, d), adj = (u_bar_x, u_bar_y, u_bar_z, d_bar)
    const float q = g.w + kUnitEps;
    const float udot = g.x * adj.x + g.y * adj.y + g.z * adj.z;
    const float inv_q = 1.0f / q;
    // r / d = u * q / d ; torch's norm backward gives 0 at d == 0
    const float scale = (g.w > 0.f) ? (adj.w - udot * inv_q) * (q / g.w) : 0.f;
    return make_float3(fmaf(scale, g.x, adj.x * inv_q), fmaf(scale, g.y, adj.y * inv_q),
                       fmaf(scale, g.z, adj.z * inv_q));
}

// Edge adjoint summed over the slabs [lo, hi), last first (the order the single-buffer accumulation of the
// reverse pass uses); zero for an empty range.
// Three kinds of slab, in this order (message_spline.cuh says who writes which):
//   [0, pairs)       at the UPPER entry e of a pair (rev(e) > e) the sum over both directions, (u_bar_e - u_bar_r,
//                    d_bar_e + d_bar_r); the lower entry is not written.  r_bar_e - r_bar_r is the position
//                    adjoint of that sum taken with e's geometry.
//   [pairs, direct)  at index e the adjoint (u_bar, d_bar) of edge e
//   [direct, slabs)  at index e the adjoint of rev(e)
__device__ __forceinline__ float4 edge_adjoint_sum(const float4* __restrict__ edge_adj, int e, int lo, int hi,
                                                   size_t slab_stride) {
    if (hi <= lo) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 t = __ldg(edge_adj + (size_t)(hi - 1) * slab_stride + e);
    for (int l = hi - 2; l >= lo; --l) t = add4(__ldg(edge_adj + (size_t)l * slab_stride + e), t);
    return t;
}

// One warp per atom j: F_j = sum_{e in row j} (r_bar_e - r_bar_rev(e)).  edge_adj holds `slabs`
// per-layer slabs of slab_stride entries (1 when the reverse pass accumulated in place);
// adj_sum_out (debug only) receives the summed adjoint of every edge.
__global__ void __launch_bounds__(256)
force_kernel(const int* __restrict__ rowptr, const int* __restrict__ rev,
             const float4* __restrict__ geo, const float4* __restrict__ edge_adj, int slabs, int pairs, int direct,
             size_t slab_stride, float4* __restrict__ adj_sum_out, float* __restrict__ forces,
             int num_atoms, const DeviceStatus* __restrict__ status) {
    if (status->overflow) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int num_warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < num_atoms; j += num_warps) {
        const int e0 = rowptr[j], e1 = rowptr[j + 1];
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int e = e0 + lane; e < e1; e += 32) {
            const int r = __ldg(rev + e);
            float4 adj_e = edge_adjoint_sum(edge_adj, e, pairs, direct, slab_stride);
            float4 adj_r = edge_adjoint_sum(edge_adj, r, pairs, direct, slab_stride);
            if (pairs > 0) {   // the pair's sum lives at its upper entry: r_bar_e - r_bar_r = +- its position adjoint
                if (r > e) adj_e = add4(adj_e, edge_adjoint_sum(edge_adj, e, 0, pairs, slab_stride));
                else adj_r = add4(adj_r, edge_adjoint_sum(edge_adj, r, 0, pairs, slab_stride));
            }
            if (direct < slabs) {   // swapped slabs: what is stored at e belongs to r and vice versa
                const float4 at_e = edge_adjoint_sum(edge_adj, e, direct, slabs, slab_stride);
                const float4 at_r = edge_adjoint_sum(edge_adj, r, direct, slabs, slab_stride);
                adj_e = add4(adj_e, at_r);
                adj_r = add4(adj_r, at_e);
            }
            if (adj_sum_out != nullptr) adj_sum_out[e] = adj_e;
            const float3 a = edge_position_adjoint(__ldg(geo + e), adj_e);
            const float3 b = edge_position_adjoint(__ldg(geo + r), adj_r);
            fx += a.x - b.x; fy += a.y - b.y; fz += a.z - b.z;
        }
        fx = group_sum<32>(fx); fy = group_sum<32>(fy); fz = group_sum<32>(fz);
        if (lane == 0) {
            forces[3 * (size_t)j] = fx;
            forces[3 * (size_t)j + 1] = fy;
            forces[3 * (size_t)j + 2] = fz;
        }
    }
}

// Strain derivative of the energy (virial), per structure:  W_ab = dE/d eps_ab = sum_e r_bar_{e,a}
// r_{e,b} over all directed edges of the structure, with r_e = x_src - x_dst (minimum image) the
// edge vector the model saw and r_bar_e its adjoint (same expression as the force kernel).  The
// reference has no working equivalent: inference/ase_calculator.py:521-588 differentiates with
// respect to a `cell` tensor the model never uses and falls back to zeros.
// A structure's edges are one contiguous CSR range; a block reduces one chunk of kVirialChunk edges
// in FP64 and adds its 9 partial sums with FP64 atomics (sum order varies, the FP32 result does not).
constexpr int kVirialChunk = 8192;

__global__ void __launch_bounds__(256)
virial_kernel(const int* __restrict__ offsets, int num_structures, const int* __restrict__ rowptr,
              const int* __restrict__ rev, const float4* __restrict__ geo, const float4* __restrict__ edge_adj,
              int slabs, int pairs, int direct, size_t slab_stride, double* __restrict__ virial, const DeviceStatus* __restrict__ status) {
    if (status->overflow) return;
    __shared__ double part[8][9];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // blockIdx.x walks (structure, chunk) pairs; chunks per structure are not known on the host
    // (edge counts live on the device), so every structure gets gridDim.y chunk slots
    const int b = blockIdx.x;
    if (b >= num_structures) return;
    const int e_lo = __ldg(rowptr + __ldg(offsets + b)), e_hi = __ldg(rowptr + __ldg(offsets + b + 1));
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int c0 = e_lo + (int)blockIdx.y * kVirialChunk; c0 < e_hi; c0 += (int)gridDim.y * kVirialChunk) {
        const int c1 = min(c0 + kVirialChunk, e_hi);
        for (int e = c0 + (int)threadIdx.x; e < c1; e += (int)blockDim.x) {
            const float4 g = __ldg(geo + e);
            float4 adj = edge_adjoint_sum(edge_adj, e, pairs, direct, slab_stride);
            // pair slabs: (r_bar_e - r_bar_r) (x) r_e = r_bar_e (x) r_e + r_bar_r (x) r_r, counted at the upper entry
            if (pairs > 0 && __ldg(rev + e) > e) adj = add4(adj, edge_adjoint_sum(edge_adj, e, 0, pairs, slab_stride));
            float3 rb = edge_position_adjoint(g, adj);
            if (direct < slabs) {
                // a swapped slab entry is the adjoint of rev(e): unit vector -u, edge vector -r.  The sum over all
                // e of r_bar_rev(e) (x) r_rev(e) is the same total, so it is folded in here with both signs flipped.
                const float3 rs = edge_position_adjoint(make_float4(-g.x, -g.y, -g.z, g.w),
                                                        edge_adjoint_sum(edge_adj, e, direct, slabs, slab_stride));
                rb.x -= rs.x; rb.y -= rs.y; rb.z -= rs.z;
            }
            const float q = g.w + kUnitEps;
            const float r[3] = {g.x * q, g.y * q, g.z * q};
            const float a[3] = {rb.x, rb.y, rb.z};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[3 * i + j] += (double)a[i] * (double)r[j];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) part[wib][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += part[w][threadIdx.x];
        if (gridDim.y == 1) virial[(size_t)b * 9 + threadIdx.x] = v;
        else atomicAdd(virial + (size_t)b * 9 + threadIdx.x, v);
    }
}

__global__ void __launch_bounds__(256)
virial_finalize_kernel(const double* __restrict__ virial, int count, float* __restrict__ out,
                       const DeviceStatus* __restrict__ status) {
    if (status->overflow) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (float)virial[i];
}

}  // namespace mlffd
