// Message block for SMALL systems (single-trajectory MD): a team of T warps shares every CSR row.
//
// With a few hundred atoms the chip is almost empty and the row-per-warp kernels (message.cuh,
// message_pipe.cuh) are a serial chain: a warp walks its ~18 edges one L2 round trip at a time
// (10 - 15 us per launch at 300 atoms, whatever the ring depth or prefetch distance: measured).
// Here a block is one team: warp t handles edges t, t + T, ... of the row, the T partial sums meet
// in shared memory and are added in fixed order (t = 0 .. T-1), so results are deterministic; they
// differ from the row-per-warp kernels by summation order only (a few ulp, inside the tolerance;
// tests compare the two paths).  The per-pair edge adjoints of the reverse pass have exactly one
// writer, as in message_backward_pairs_kernel, whose arithmetic this file repeats.
#pragma once
#include "message.cuh"

namespace mlffd {

template <int T>
__device__ __forceinline__ void team_reduce4(float4 (&part)[T][4][32], int t, int lane, float4& a0, float4& a1,
                                             float4& a2, float4& a3) {
    part[t][0][lane] = a0; part[t][1][lane] = a1; part[t][2][lane] = a2; part[t][3][lane] = a3;
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int w = 1; w < T; ++w) {
            a0 = add4(a0, part[w][0][lane]); a1 = add4(a1, part[w][1][lane]);
            a2 = add4(a2, part[w][2][lane]); a3 = add4(a3, part[w][3][lane]);
        }
    }
    __syncthreads();
}

// Forward; contract of message_forward_kernel<H, LAYER0>.
template <int H, bool LAYER0, int T>
__global__ void __launch_bounds__(32 * T)
message_forward_team_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                            const int* __restrict__ pair, const float4* __restrict__ geo,
                            const float* __restrict__ filt, const float* __restrict__ s_in,
                            const float* __restrict__ v_in, float* __restrict__ s_msg,
                            float* __restrict__ v_msg, int num_atoms,
                            const DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    __shared__ float4 part[T][4][32];
    const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
    const int sub = lane / M::LPA;
    const int c4 = (lane % M::LPA) * 4;
    for (int j0 = (int)blockIdx.x * M::APW; j0 < num_atoms; j0 += (int)gridDim.x * M::APW) {
        const int j = j0 + sub;
        const bool valid = j < num_atoms;
        const int e0 = valid ? rowptr[j] : 0;
        const int e1 = valid ? rowptr[j + 1] : 0;
        float4 acc_s = make4(0.f), acc_x = make4(0.f), acc_y = make4(0.f), acc_z = make4(0.f);
#pragma unroll 2
        for (int e = e0 + t; e < e1; e += T) {
            const int i = __ldg(col + e);
            const float4 g = __ldg(geo + e);
            const float* f = filt + (size_t)__ldg(pair + e) * (3 * H) + c4;
            const float4 fa = ldg4(f);
            const float4 fc = ldg4(f + 2 * H);
            const float4 si = ldg4(s_in + (size_t)i * H + c4);
            acc_s = fma4(si, fa, acc_s);
            if (!LAYER0) {
                const float4 fb = ldg4(f + H);
                const float* vi = v_in + (size_t)i * 3 * H + c4;
                acc_x = fma4(ldg4(vi), fb, acc_x);
                acc_y = fma4(ldg4(vi + H), fb, acc_y);
                acc_z = fma4(ldg4(vi + 2 * H), fb, acc_z);
            }
            acc_x = fma4s(g.x, fc, acc_x);
            acc_y = fma4s(g.y, fc, acc_y);
            acc_z = fma4s(g.z, fc, acc_z);
        }
        team_reduce4<T>(part, t, lane, acc_s, acc_x, acc_y, acc_z);
        if (t == 0 && valid) {
            st4(s_msg + (size_t)j * H + c4, add4(ldg4(s_in + (size_t)j * H + c4), acc_s));
            float* vo = v_msg + (size_t)j * 3 * H + c4;
            if (LAYER0) {
                st4(vo, acc_x); st4(vo + H, acc_y); st4(vo + 2 * H, acc_z);
            } else {
                const float* vj = v_in + (size_t)j * 3 * H + c4;
                st4(vo, add4(ldg4(vj), acc_x));
                st4(vo + H, add4(ldg4(vj + H), acc_y));
                st4(vo + 2 * H, add4(ldg4(vj + 2 * H), acc_z));
            }
        }
    }
}

// Reverse, every undirected pair once; contract and arithmetic of
// message_backward_pairs_kernel<H, LAYER0, false> (per-layer adjoint slab, no accumulation).
template <int H, bool LAYER0, int T>
__global__ void __launch_bounds__(32 * T)
message_backward_pairs_team_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                   const int* __restrict__ pair, const int* __restrict__ rev,
                                   const float4* __restrict__ geo, const float* __restrict__ filt,
                                   const float* __restrict__ dfilt, const float* __restrict__ s_in,
                                   const float* __restrict__ v_in, const float* __restrict__ sbar_m,
                                   const float* __restrict__ vbar_m, float* __restrict__ sbar_in,
                                   float* __restrict__ vbar_in, float4* __restrict__ edge_adj,
                                   int num_atoms, const DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    __shared__ float4 part_s[T][4][32];
    const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
    const int sub = lane / M::LPA;
    const int gl = lane % M::LPA;
    const int c4 = gl * 4;
    const int held = ((gl & (M::LPA / 2)) ? 4 : 0) + ((gl & (M::LPA / 4)) ? 2 : 0) + ((gl & (M::LPA / 8)) ? 1 : 0);
    const bool holder = (gl % (M::LPA / 8)) == 0;
    for (int i0 = (int)blockIdx.x * M::APW; i0 < num_atoms; i0 += (int)gridDim.x * M::APW) {
        const int i = i0 + sub;
        const bool valid = i < num_atoms;
        const int e0 = valid ? rowptr[i] : 0;
        const int deg = valid ? rowptr[i + 1] - e0 : 0;
        const int max_deg = (M::APW > 1) ? group_max_int<32>(deg) : deg;
        float4 sb = make4(0.f), vbx = make4(0.f), vby = make4(0.f), vbz = make4(0.f);
        float4 si = make4(0.f), vix = make4(0.f), viy = make4(0.f), viz = make4(0.f);
        if (valid) {
            sb = ldg4(sbar_m + (size_t)i * H + c4);
            const float* vb = vbar_m + (size_t)i * 3 * H + c4;
            vbx = ldg4(vb); vby = ldg4(vb + H); vbz = ldg4(vb + 2 * H);
            si = ldg4(s_in + (size_t)i * H + c4);
            if (!LAYER0) {
                const float* vi = v_in + (size_t)i * 3 * H + c4;
                vix = ldg4(vi); viy = ldg4(vi + H); viz = ldg4(vi + 2 * H);
            }
        }
        // the residual path enters once, with warp 0's partial sum
        float4 acc_s = t ? make4(0.f) : sb, acc_x = t ? make4(0.f) : vbx, acc_y = t ? make4(0.f) : vby,
               acc_z = t ? make4(0.f) : vbz;
        for (int k = t; k < max_deg; k += T) {
            const bool active = k < deg;
            const int e = e0 + k;
            const int j = active ? __ldg(col + e) : i;
            const bool upper = active && j > i;
            float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int r = 0;
            if (active && (upper || !LAYER0)) {
                const size_t prow = (size_t)__ldg(pair + e) * (3 * H) + c4;
                const float4 sbj = ldg4(sbar_m + (size_t)j * H + c4);
                const float* vbj_p = vbar_m + (size_t)j * 3 * H + c4;
                const float4 vbjx = ldg4(vbj_p), vbjy = ldg4(vbj_p + H), vbjz = ldg4(vbj_p + 2 * H);
                if (!LAYER0) {
                    const float4 fa = ldg4(filt + prow);
                    const float4 fb = ldg4(filt + prow + H);
                    acc_s = fma4(fa, sbj, acc_s);
                    acc_x = fma4(fb, vbjx, acc_x);
                    acc_y = fma4(fb, vbjy, acc_y);
                    acc_z = fma4(fb, vbjz, acc_z);
                }
                if (upper) {
                    r = __ldg(rev + e);
                    const float4 g = __ldg(geo + e);                        // unit vector of (j -> i)
                    const float4 gr = make_float4(-g.x, -g.y, -g.z, g.w);   // (i -> j): exact negation
                    const float4 fc = ldg4(filt + prow + 2 * H);
                    const float4 dfa = ldg4(dfilt + prow);
                    const float4 dfc = ldg4(dfilt + prow + 2 * H);
                    const float4 sj = ldg4(s_in + (size_t)j * H + c4);
                    const float4 abar = fma4(sj, sb, mul4(si, sbj));
                    float4 cbar = fma4s(g.x, vbx, fma4s(g.y, vby, fma4s(g.z, vbz, make4(0.f))));
                    cbar = fma4s(gr.x, vbjx, fma4s(gr.y, vbjy, fma4s(gr.z, vbjz, cbar)));
                    float d = dot4(abar, dfa) + dot4(cbar, dfc);
                    if (!LAYER0) {
                        const float4 dfb = ldg4(dfilt + prow + H);
                        const float* vj = v_in + (size_t)j * 3 * H + c4;
                        float4 bbar = fma4(ldg4(vj), vbx, fma4(ldg4(vj + H), vby, mul4(ldg4(vj + 2 * H), vbz)));
                        bbar = fma4(vix, vbjx, fma4(viy, vbjy, fma4(viz, vbjz, bbar)));
                        d += dot4(bbar, dfb);
                    }
                    part[0] = dot4(fc, vbx); part[1] = dot4(fc, vby); part[2] = dot4(fc, vbz);
                    part[3] = d;
                    part[4] = dot4(fc, vbjx); part[5] = dot4(fc, vbjy); part[6] = dot4(fc, vbjz);
                }
            }
            if (__any_sync(0xffffffffu, upper)) {
                const float total = group_sum8<M::LPA>(part, gl);
                if (upper && holder)
                    reinterpret_cast<float*>(edge_adj + (held < 4 ? e : r))[held & 3] = total;
            }
        }
        team_reduce4<T>(part_s, t, lane, acc_s, acc_x, acc_y, acc_z);
        if (!LAYER0 && valid && t == 0) {
            st4(sbar_in + (size_t)i * H + c4, acc_s);
            float* vo = vbar_in + (size_t)i * 3 * H + c4;
            st4(vo, acc_x); st4(vo + H, acc_y); st4(vo + 2 * H, acc_z);
        }
    }
}

}  // namespace mlffd
