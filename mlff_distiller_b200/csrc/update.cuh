// PaiNN update block, forward and reverse, as per-atom-tile fused kernels.
//
// Forward restates PaiNNUpdate.forward (reference src/mlff_distiller/models/student_model.py:
// 434-470): n = |v'| over xyz; (ds, g1, g2) = M2 SiLU(M1 [s'; n] + m1) + m2; s'' = s' + ds;
// v''[a] = v'[a] g1 + (sum_b U[a][b] v'[b]) g2.  The 3x3 mixing_matrix acts on the SPATIAL axis
// (einsum 'ij,njk->nik', :458-462) -- reproduced as is, the model is not rotation-equivariant.
// The last layer's vector update never reaches the energy (:736) and is skipped (LAST == true).
//
// Reverse follows SURVEY App. A.3 (update-bwd); norm backward at |v'| = 0 yields 0 like
// torch.linalg.vector_norm.
//
// One block = 64 atoms; both dense layers run on the shared register-tiled GEMM
// (tile_gemm.cuh); norms, gates, spatial mixing and residuals are fused around them so the
// only HBM traffic is the feature rows themselves.
#pragma once
#include "tile_gemm.cuh"

namespace mlffd {

struct UpdateWeights {
    const float* M1t;  // [2H][H]  transposed update_mlp.0.weight (forward)
    const float* m1;   // [H]
    const float* M2t;  // [H][3H]  transposed update_mlp.2.weight (forward)
    const float* m2;   // [3H]
    const float* M1;   // [H][2H]  original layout (reverse: y1_bar @ M1)
    const float* M2;   // [3H][H]  original layout (reverse: out_bar @ M2)
    const float* U;    // [3][3]
};

template <int H>
constexpr size_t update_fwd_smem_bytes() { return sizeof(float) * (size_t)(2 * H * kAStride + H * H); }
template <int H>
constexpr size_t update_bwd_smem_bytes() { return sizeof(float) * (size_t)(3 * H * kAStride + H * H); }

template <int H, bool LAST>
__global__ void __launch_bounds__(kGemmThreads)
update_forward_kernel(const float* __restrict__ s_msg, const float* __restrict__ v_msg,
                      UpdateWeights w, float* __restrict__ s_out, float* __restrict__ v_out,
                      float* __restrict__ y1_save, float* __restrict__ gates_save, int num_atoms,
                      const DeviceStatus* __restrict__ status) {
    using T = TileTraits<H>;
    constexpr int VW = T::VW;
    if (status->overflow) return;
    extern __shared__ __align__(16) float smem[];
    float* A_s = smem;                     // [2H][AS]: rows 0..H-1 = s', rows H..2H-1 = |v'|
    float* W_s = A_s + 2 * H * kAStride;   // [H][H]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int num_tiles = (num_atoms + kTileRows - 1) / kTileRows;
    float U[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) U[q] = __ldg(w.U + q);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int a0 = tile * kTileRows;
        // ---- stage [s' ; |v'|] k-major ----
        for (int idx = tid; idx < kTileRows * H; idx += kGemmThreads) {
            const int m = idx / H, c = idx - m * H;
            const int atom = a0 + m;
            float sv = 0.f, nv = 0.f;
            if (atom < num_atoms) {
                sv = __ldg(s_msg + (size_t)atom * H + c);
                const float* vp = v_msg + (size_t)atom * 3 * H + c;
                const float x = __ldg(vp), y = __ldg(vp + H), z = __ldg(vp + 2 * H);
                nv = sqrtf(x * x + y * y + z * z);
            }
            A_s[c * kAStride + m] = sv;
            A_s[(H + c) * kAStride + m] = nv;
        }
        // ---- GEMM 1: y1 = [s'; n] M1^T + m1 ----
        float acc[4][T::RN];
        tile_zero<H>(acc);
        for (int kc = 0; kc < 2; ++kc) {
            __syncthreads();
            load_weight_chunk<H>(W_s, w.M1t, H, kc * H, 0, H);
            __syncthreads();
            tile_fma<H>(acc, A_s + kc * H * kAStride, W_s, H, ty, tx);
        }
        __syncthreads();  // everyone done reading A_s before it is overwritten with SiLU(y1)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int m = ty * 4 + r, atom = a0 + m;
#pragma unroll
            for (int g = 0; g < T::NG; ++g) {
                const int n0 = g * T::GROUP_STRIDE + tx * VW;
                float y[VW];
#pragma unroll
                for (int q = 0; q < VW; ++q) {
                    y[q] = acc[r][g * VW + q] + __ldg(w.m1 + n0 + q);
                    A_s[(n0 + q) * kAStride + m] = siluf_(y[q]);
                }
                if (atom < num_atoms) stv<VW>(y1_save + (size_t)atom * H + n0, y);
            }
        }
        // ---- GEMM 2: (ds | g1 | g2) = SiLU(y1) M2^T + m2 ----
        float g1[4][T::RN], g2[4][T::RN];
#pragma unroll
        for (int nc = 0; nc < (LAST ? 1 : 3); ++nc) {
            __syncthreads();
            load_weight_chunk<H>(W_s, w.M2t, 3 * H, 0, nc * H, H);
            __syncthreads();
            tile_zero<H>(acc);
            tile_fma<H>(acc, A_s, W_s, H, ty, tx);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int atom = a0 + ty * 4 + r;
#pragma unroll
                for (int g = 0; g < T::NG; ++g) {
                    const int n0 = g * T::GROUP_STRIDE + tx * VW;
                    float o[VW];
#pragma unroll
                    for (int q = 0; q < VW; ++q)
                        o[q] = acc[r][g * VW + q] + __ldg(w.m2 + nc * H + n0 + q);
                    if (nc == 0) {
                        if (atom < num_atoms) {
                            float sv[VW];
                            ldv<VW>(s_msg + (size_t)atom * H + n0, sv);
#pragma unroll
                            for (int q = 0; q < VW; ++q) sv[q] += o[q];
                            stv<VW>(s_out + (size_t)atom * H + n0, sv);
                        }
                    } else if (nc == 1) {
#pragma unroll
                        for (int q = 0; q < VW; ++q) g1[r][g * VW + q] = o[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < VW; ++q) g2[r][g * VW + q] = o[q];
                    }
                }
            }
        }
        if (!LAST) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int atom = a0 + ty * 4 + r;
                if (atom >= num_atoms) continue;
#pragma unroll
                for (int g = 0; g < T::NG; ++g) {
                    const int n0 = g * T::GROUP_STRIDE + tx * VW;
                    const float* vp = v_msg + (size_t)atom * 3 * H + n0;
                    float vx[VW], vy[VW], vz[VW], a[VW], b[VW], ox[VW], oy[VW], oz[VW];
                    ldv<VW>(vp, vx); ldv<VW>(vp + H, vy); ldv<VW>(vp + 2 * H, vz);
#pragma unroll
                    for (int q = 0; q < VW; ++q) {
                        a[q] = g1[r][g * VW + q];
                        b[q] = g2[r][g * VW + q];
                        ox[q] = vx[q] * a[q] + (U[0] * vx[q] + U[1] * vy[q] + U[2] * vz[q]) * b[q];
                        oy[q] = vy[q] * a[q] + (U[3] * vx[q] + U[4] * vy[q] + U[5] * vz[q]) * b[q];
                        oz[q] = vz[q] * a[q] + (U[6] * vx[q] + U[7] * vy[q] + U[8] * vz[q]) * b[q];
                    }
                    stv<VW>(gates_save + (size_t)atom * 2 * H + n0, a);
                    stv<VW>(gates_save + (size_t)atom * 2 * H + H + n0, b);
                    float* vo = v_out + (size_t)atom * 3 * H + n0;
                    stv<VW>(vo, ox); stv<VW>(vo + H, oy); stv<VW>(vo + 2 * H, oz);
                }
            }
        }
        __syncthreads();  // A_s / W_s reused by the next tile
    }
}

// Reverse of the update block, in place on the adjoint buffers:
//   in : sbar (adjoint of s''), vbar (adjoint of v''; ignored when LAST: it is zero)
//   out: sbar <- adjoint of s', vbar <- adjoint of v'
template <int H, bool LAST>
__global__ void __launch_bounds__(kGemmThreads)
update_backward_kernel(const float* __restrict__ v_msg, const float* __restrict__ y1_save,
                       const float* __restrict__ gates_save, UpdateWeights w,
                       float* sbar, float* vbar, int num_atoms,
                       const DeviceStatus* __restrict__ status) {
    using T = TileTraits<H>;
    constexpr int VW = T::VW;
    if (status->overflow) return;
    extern __shared__ __align__(16) float smem[];
    constexpr int KB = LAST ? H : 3 * H;   // rows of the output adjoint [ds_bar; g1_bar; g2_bar]
    float* B_s = smem;                     // [3H][AS]
    float* W_s = B_s + 3 * H * kAStride;   // [H][H]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int num_tiles = (num_atoms + kTileRows - 1) / kTileRows;
    float U[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) U[q] = __ldg(w.U + q);

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int a0 = tile * kTileRows;
        // ---- stage [s_bar ; g1_bar ; g2_bar] k-major ----
        for (int idx = tid; idx < kTileRows * H; idx += kGemmThreads) {
            const int m = idx / H, c = idx - m * H;
            const int atom = a0 + m;
            float sv = 0.f, ga = 0.f, gb = 0.f;
            if (atom < num_atoms) {
                sv = sbar[(size_t)atom * H + c];
                if (!LAST) {
                    const float* vp = v_msg + (size_t)atom * 3 * H + c;
                    const float* bp = vbar + (size_t)atom * 3 * H + c;
                    const float vx = __ldg(vp), vy = __ldg(vp + H), vz = __ldg(vp + 2 * H);
                    const float bx = bp[0], by = bp[H], bz = bp[2 * H];
                    ga = bx * vx + by * vy + bz * vz;
                    gb = bx * (U[0] * vx + U[1] * vy + U[2] * vz) +
                         by * (U[3] * vx + U[4] * vy + U[5] * vz) +
                         bz * (U[6] * vx + U[7] * vy + U[8] * vz);
                }
            }
            B_s[c * kAStride + m] = sv;
            if (!LAST) {
                B_s[(H + c) * kAStride + m] = ga;
                B_s[(2 * H + c) * kAStride + m] = gb;
            }
        }
        // ---- GEMM 1: hid_bar = [ds_bar; g1_bar; g2_bar] M2 ; y1_bar = hid_bar * SiLU'(y1) ----
        float acc[4][T::RN];
        tile_zero<H>(acc);
        for (int kc = 0; kc < KB / H; ++kc) {
            __syncthreads();
            load_weight_chunk<H>(W_s, w.M2, H, kc * H, 0, H);
            __syncthreads();
            tile_fma<H>(acc, B_s + kc * H * kAStride, W_s, H, ty, tx);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int m = ty * 4 + r, atom = a0 + m;
#pragma unroll
            for (int g = 0; g < T::NG; ++g) {
                const int n0 = g * T::GROUP_STRIDE + tx * VW;
                float y[VW];
#pragma unroll
                for (int q = 0; q < VW; ++q) y[q] = 0.f;
                if (atom < num_atoms) ldv<VW>(y1_save + (size_t)atom * H + n0, y);
#pragma unroll
                for (int q = 0; q < VW; ++q)
                    B_s[(n0 + q) * kAStride + m] = acc[r][g * VW + q] * silu_gradf_(y[q]);
            }
        }
        // ---- GEMM 2: [ps_bar | n_bar] = y1_bar M1 ----
#pragma unroll
        for (int nc = 0; nc < 2; ++nc) {
            __syncthreads();
            load_weight_chunk<H>(W_s, w.M1, 2 * H, 0, nc * H, H);
            __syncthreads();
            tile_zero<H>(acc);
            tile_fma<H>(acc, B_s, W_s, H, ty, tx);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int atom = a0 + ty * 4 + r;
                if (atom >= num_atoms) continue;
#pragma unroll
                for (int g = 0; g < T::NG; ++g) {
                    const int n0 = g * T::GROUP_STRIDE + tx * VW;
                    if (nc == 0) {
                        float sv[VW];
                        ldv<VW>(sbar + (size_t)atom * H + n0, sv);
#pragma unroll
                        for (int q = 0; q < VW; ++q) sv[q] += acc[r][g * VW + q];
                        stv<VW>(sbar + (size_t)atom * H + n0, sv);
                    } else {
                        const float* vp = v_msg + (size_t)atom * 3 * H + n0;
                        float* bp = vbar + (size_t)atom * 3 * H + n0;
                        float vx[VW], vy[VW], vz[VW], ox[VW], oy[VW], oz[VW];
                        ldv<VW>(vp, vx); ldv<VW>(vp + H, vy); ldv<VW>(vp + 2 * H, vz);
#pragma unroll
                        for (int q = 0; q < VW; ++q) {
                            const float nrm = sqrtf(vx[q] * vx[q] + vy[q] * vy[q] + vz[q] * vz[q]);
                            const float sc = (nrm > 0.f) ? acc[r][g * VW + q] / nrm : 0.f;
                            ox[q] = sc * vx[q]; oy[q] = sc * vy[q]; oz[q] = sc * vz[q];
                        }
                        if (!LAST) {
                            float a[VW], b[VW], bx[VW], by[VW], bz[VW];
                            ldv<VW>(gates_save + (size_t)atom * 2 * H + n0, a);
                            ldv<VW>(gates_save + (size_t)atom * 2 * H + H + n0, b);
                            ldv<VW>(bp, bx); ldv<VW>(bp + H, by); ldv<VW>(bp + 2 * H, bz);
#pragma unroll
                            for (int q = 0; q < VW; ++q) {
                                const float gx = bx[q] * b[q], gy = by[q] * b[q], gz = bz[q] * b[q];
                                ox[q] += bx[q] * a[q] + (U[0] * gx + U[3] * gy + U[6] * gz);
                                oy[q] += by[q] * a[q] + (U[1] * gx + U[4] * gy + U[7] * gz);
                                oz[q] += bz[q] * a[q] + (U[2] * gx + U[5] * gy + U[8] * gz);
                            }
                        }
                        stv<VW>(bp, ox); stv<VW>(bp + H, oy); stv<VW>(bp + 2 * H, oz);
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace mlffd
