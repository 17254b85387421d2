// PaiNN message block as gather-filter-scatter over destination-sorted CSR, forward and reverse.
//
// Forward restates PaiNNMessage.forward (reference src/mlff_distiller/models/student_model.py:
// 346-387):  s'_j = s_j + sum_{e=(i->j)} s_i * a_e ;  v'_j = v_j + sum_e (v_i * b_e + u_e c_e),
// (a,b,c) = the filter-table row of the edge's pair.  index_add_ becomes a sequential
// accumulation over the row in source-ascending order -- the same order the reference's CPU
// index_add_ applies -- in registers of the sub-warp group that owns the destination atom: no
// atomics, deterministic.
//
// Reverse follows SURVEY App. A.3.  The scatter to SOURCE atoms (s_bar_i += a_e s_bar'_j) is
// turned into a gather over the same CSR row through the symmetry of the edge set: edge (i->j)
// and its reverse (j->i) share the filter row, so atom i sums a_e * s_bar'_j over its own row.
// The filter MLP is never back-propagated: d_bar_e = [a_bar; b_bar; c_bar] . f'(d_e) with f' from
// the table (filter.cuh).
//
// Thread mapping: H/4 lanes per atom (32 / 16 / 8), one float4 of channels per lane, so every
// feature row is read with 16-byte loads, fully coalesced (H*4 bytes contiguous per group).
#pragma once
#include "common.cuh"

namespace mlffd {

template <int H>
struct MsgTraits {
    static constexpr int LPA = H / 4;           // lanes per atom
    static constexpr int APW = 32 / LPA;        // atoms per warp
};

template <int WIDTH>
__device__ __forceinline__ int group_max_int(int v) {
    // max over the whole warp (all sub-groups) so loop trip counts stay warp-uniform
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// LAYER0: v_in == 0 (not stored), so the v gather and the b filter are skipped.
template <int H, bool LAYER0>
__global__ void __launch_bounds__(256)
message_forward_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                       const int* __restrict__ pair, const float4* __restrict__ geo,
                       const float* __restrict__ filt, const float* __restrict__ s_in,
                       const float* __restrict__ v_in, float* __restrict__ s_msg,
                       float* __restrict__ v_msg, int num_atoms,
                       const DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane / M::LPA;               // which atom of the warp
    const int c4 = (lane % M::LPA) * 4;          // first channel of this lane
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int num_warps = (gridDim.x * blockDim.x) >> 5;
    for (int j0 = warp * M::APW; j0 < num_atoms; j0 += num_warps * M::APW) {
        const int j = j0 + sub;
        const bool valid = j < num_atoms;
        const int e0 = valid ? rowptr[j] : 0;
        const int e1 = valid ? rowptr[j + 1] : 0;
        float4 acc_s = make4(0.f), acc_x = make4(0.f), acc_y = make4(0.f), acc_z = make4(0.f);
#pragma unroll 2
        for (int e = e0; e < e1; ++e) {
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
        if (valid) {
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

// Reverse of the message block for layer l.
//   in : sbar_m / vbar_m   adjoints of (s_msg, v_msg)                      [N,H] / [N,3,H]
//        s_in / v_in       the layer's input features (saved by the forward)
//   out: sbar_in / vbar_in adjoints of the layer inputs (not needed for LAYER0: the embedding
//        does not depend on positions)
//        edge_adj[e] = (dE/du_x, dE/du_y, dE/du_z, dE/dd through the filter), summed over layers
//        (ACCUMULATE == false for the first layer processed, i.e. l = L-1).
template <int H, bool LAYER0, bool ACCUMULATE>
__global__ void __launch_bounds__(256)
message_backward_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                        const int* __restrict__ pair, const float4* __restrict__ geo,
                        const float* __restrict__ filt, const float* __restrict__ dfilt,
                        const float* __restrict__ s_in, const float* __restrict__ v_in,
                        const float* __restrict__ sbar_m, const float* __restrict__ vbar_m,
                        float* __restrict__ sbar_in, float* __restrict__ vbar_in,
                        float4* __restrict__ edge_adj, int num_atoms,
                        const DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane / M::LPA;
    const int gl = lane % M::LPA;
    const int c4 = gl * 4;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int num_warps = (gridDim.x * blockDim.x) >> 5;
    for (int i0 = warp * M::APW; i0 < num_atoms; i0 += num_warps * M::APW) {
        const int i = i0 + sub;
        const bool valid = i < num_atoms;
        const int e0 = valid ? rowptr[i] : 0;
        const int deg = valid ? rowptr[i + 1] - e0 : 0;
        const int max_deg = (M::APW > 1) ? group_max_int<32>(deg) : deg;
        float4 sb = make4(0.f), vbx = make4(0.f), vby = make4(0.f), vbz = make4(0.f);
        if (valid) {
            sb = ldg4(sbar_m + (size_t)i * H + c4);
            const float* vb = vbar_m + (size_t)i * 3 * H + c4;
            vbx = ldg4(vb); vby = ldg4(vb + H); vbz = ldg4(vb + 2 * H);
        }
        float4 acc_s = sb, acc_x = vbx, acc_y = vby, acc_z = vbz;  // residual path
        for (int k = 0; k < max_deg; ++k) {
            const bool active = k < deg;
            float d_part = 0.f, ux_part = 0.f, uy_part = 0.f, uz_part = 0.f;
            const int e = e0 + k;
            if (active) {
                const int j = __ldg(col + e);
                const float4 g = __ldg(geo + e);
                const size_t prow = (size_t)__ldg(pair + e) * (3 * H) + c4;
                const float4 fa = ldg4(filt + prow);
                const float4 fc = ldg4(filt + prow + 2 * H);
                const float4 dfa = ldg4(dfilt + prow);
                const float4 dfc = ldg4(dfilt + prow + 2 * H);
                const float4 sj = ldg4(s_in + (size_t)j * H + c4);
                // as the edge (j -> i): adjoints of its filter row
                const float4 abar = mul4(sj, sb);
                const float4 cbar = fma4s(g.x, vbx, fma4s(g.y, vby, fma4s(g.z, vbz, make4(0.f))));
                d_part = dot4(abar, dfa) + dot4(cbar, dfc);
                ux_part = dot4(fc, vbx);
                uy_part = dot4(fc, vby);
                uz_part = dot4(fc, vbz);
                if (!LAYER0) {
                    const float4 fb = ldg4(filt + prow + H);
                    const float4 dfb = ldg4(dfilt + prow + H);
                    const float* vj = v_in + (size_t)j * 3 * H + c4;
                    const float4 bbar = fma4(ldg4(vj), vbx, fma4(ldg4(vj + H), vby,
                                             mul4(ldg4(vj + 2 * H), vbz)));
                    d_part += dot4(bbar, dfb);
                    // as the reverse edge (i -> j): same filter row, adjoints of neighbour j
                    acc_s = fma4(fa, ldg4(sbar_m + (size_t)j * H + c4), acc_s);
                    const float* vbj = vbar_m + (size_t)j * 3 * H + c4;
                    acc_x = fma4(fb, ldg4(vbj), acc_x);
                    acc_y = fma4(fb, ldg4(vbj + H), acc_y);
                    acc_z = fma4(fb, ldg4(vbj + 2 * H), acc_z);
                }
            }
            d_part = group_sum<M::LPA>(d_part);
            ux_part = group_sum<M::LPA>(ux_part);
            uy_part = group_sum<M::LPA>(uy_part);
            uz_part = group_sum<M::LPA>(uz_part);
            if (active && gl == 0) {
                float4 out = make_float4(ux_part, uy_part, uz_part, d_part);
                if (ACCUMULATE) {
                    const float4 old = edge_adj[e];
                    out = add4(out, old);
                }
                edge_adj[e] = out;
            }
        }
        if (!LAYER0 && valid) {
            st4(sbar_in + (size_t)i * H + c4, acc_s);
            float* vo = vbar_in + (size_t)i * 3 * H + c4;
            st4(vo, acc_x); st4(vo + H, acc_y); st4(vo + 2 * H, acc_z);
        }
    }
}

// Sum of 8 per-lane values over the LPA lanes of a sub-warp group with 4 + 2 + 1 + log2(LPA / 8)
// shuffles (a butterfly that halves the number of live values at each of the first three steps)
// instead of 8 * log2(LPA).  Afterwards lane gl holds the total of value
// idx = 4 * bit(gl, LPA/2) + 2 * bit(gl, LPA/4) + bit(gl, LPA/8).
template <int LPA>
__device__ __forceinline__ float group_sum8(const float (&v)[8], int gl) {
    static_assert(LPA >= 8, "needs at least 8 lanes per group");
    const unsigned full = 0xffffffffu;
    const bool b0 = (gl & (LPA / 2)) != 0;
    float w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float send = b0 ? v[t] : v[t + 4], keep = b0 ? v[t + 4] : v[t];
        w[t] = keep + __shfl_xor_sync(full, send, LPA / 2);
    }
    const bool b1 = (gl & (LPA / 4)) != 0;
    float x[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float send = b1 ? w[t] : w[t + 2], keep = b1 ? w[t + 2] : w[t];
        x[t] = keep + __shfl_xor_sync(full, send, LPA / 4);
    }
    const bool b2 = (gl & (LPA / 8)) != 0;
    float y = (b2 ? x[1] : x[0]) + __shfl_xor_sync(full, b2 ? x[0] : x[1], LPA / 8);
#pragma unroll
    for (int o = LPA / 16; o > 0; o >>= 1) y += __shfl_xor_sync(full, y, o);
    return y;
}

// Reverse of the message block, every undirected pair handled ONCE.
//
// Row i gathers, for each neighbour j, the neighbour's adjoints (s_bar'_j, v_bar'_j) and, for
// j > i, also its features (s_j, v_j): together with the row's own (s_i, v_i, s_bar'_i,
// v_bar'_i) that is everything BOTH directed edges of the pair {i, j} need, and both share the
// filter row.  So the derivative row f' (half of the table bytes the reverse pass streams) is
// read once per pair instead of once per directed edge:
//     d_bar_pair = (a_bar + a_bar') . a' + (b_bar + b_bar') . b' + (c_bar + c_bar') . c'
// with (a_bar, b_bar, c_bar) the gate adjoints of edge (j -> i) and the primed ones those of
// (i -> j).  Only the SUM d_bar_e + d_bar_rev(e) enters the forces (r_rev = -r_e), so the pair
// value is stored with edge e and 0 with rev(e); u_bar is stored for both edges.  For j < i the
// row only needs (a, b) and the neighbour's adjoints for its own s_bar_i / v_bar_i sums.
// Same contract as message_backward_kernel; every edge_adj entry has exactly one writer.
template <int H, bool LAYER0, bool ACCUMULATE>
__global__ void __launch_bounds__(256, 3)
message_backward_pairs_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                              const int* __restrict__ pair, const int* __restrict__ rev,
                              const float4* __restrict__ geo, const float* __restrict__ filt,
                              const float* __restrict__ dfilt, const float* __restrict__ s_in,
                              const float* __restrict__ v_in, const float* __restrict__ sbar_m,
                              const float* __restrict__ vbar_m, float* __restrict__ sbar_in,
                              float* __restrict__ vbar_in, float4* __restrict__ edge_adj,
                              int num_atoms, const DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane / M::LPA;
    const int gl = lane % M::LPA;
    const int c4 = gl * 4;
    // lane that ends up holding reduced value t (see group_sum8) and the value this lane holds
    const int held = ((gl & (M::LPA / 2)) ? 4 : 0) + ((gl & (M::LPA / 4)) ? 2 : 0) + ((gl & (M::LPA / 8)) ? 1 : 0);
    const bool holder = (gl % (M::LPA / 8)) == 0;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int num_warps = (gridDim.x * blockDim.x) >> 5;
    for (int i0 = warp * M::APW; i0 < num_atoms; i0 += num_warps * M::APW) {
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
        float4 acc_s = sb, acc_x = vbx, acc_y = vby, acc_z = vbz;  // residual path
        for (int k = 0; k < max_deg; ++k) {
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
                    // as the source of edge (i -> j): same filter row, adjoints of neighbour j
                    const float4 fa = ldg4(filt + prow);
                    const float4 fb = ldg4(filt + prow + H);
                    acc_s = fma4(fa, sbj, acc_s);
                    acc_x = fma4(fb, vbjx, acc_x);
                    acc_y = fma4(fb, vbjy, acc_y);
                    acc_z = fma4(fb, vbjz, acc_z);
                }
                if (upper) {
                    r = __ldg(rev + e);
                    const float4 g = __ldg(geo + e);    // unit vector of (j -> i)
                    const float4 gr = make_float4(-g.x, -g.y, -g.z, g.w);   // unit vector of (i -> j) = exact negation of (j -> i)
                    const float4 fc = ldg4(filt + prow + 2 * H);
                    const float4 dfa = ldg4(dfilt + prow);
                    const float4 dfc = ldg4(dfilt + prow + 2 * H);
                    const float4 sj = ldg4(s_in + (size_t)j * H + c4);
                    // gate adjoints of (j -> i) plus those of (i -> j)
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
                float total = group_sum8<M::LPA>(part, gl);
                if (upper && holder) {
                    // values 0..3 -> edge_adj[e] = (u_bar, d_bar_pair); 4..7 -> edge_adj[rev] = (u_bar', 0)
                    float* dst = reinterpret_cast<float*>(edge_adj + (held < 4 ? e : r)) + (held & 3);
                    if (ACCUMULATE) total += *dst;
                    *dst = total;
                }
            }
        }
        if (!LAYER0 && valid) {
            st4(sbar_in + (size_t)i * H + c4, acc_s);
            float* vo = vbar_in + (size_t)i * 3 * H + c4;
            st4(vo, acc_x); st4(vo + H, acc_y); st4(vo + 2 * H, acc_z);
        }
    }
}

}  // namespace mlffd
