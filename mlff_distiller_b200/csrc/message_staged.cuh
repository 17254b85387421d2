// Structure-blocked message kernels with shared-memory-staged per-atom features.
//
// Same arithmetic, same CSR order and therefore bit-identical results as message.cuh, for batches
// of small structures (BASELINE configs C2 / C5: ~50 atoms).  One thread block owns one structure
// at a time: the structure's feature rows are contiguous in memory (atoms of a structure are
// contiguous), so staging them is a straight coalesced copy of n x 16H bytes, after which every
// neighbour gather -- 16H bytes per edge, ~20x reuse of each row -- is served from shared memory
// instead of L2.  The only streaming traffic left is the filter-table rows; both directed edges
// of a pair are processed by the same block microseconds apart, so the second read is an L2 hit.
//
// The reverse pass is split in two so each half stages only one pair of arrays (16H bytes/atom):
//   edge kernel : stages (s_in, v_in)      -> edge_adj (d_bar, u_bar per edge)   reads c, a', b', c'
//   atom kernel : stages (s_bar', v_bar')  -> s_bar_in, v_bar_in                 reads a, b
// together exactly the bytes message_backward_kernel reads.
//
// Rows are handed to warps through a shared-memory counter (dynamic balance: a 50-atom structure
// does not divide evenly among 8 warps).  Requires every structure to have at most
// `max_atoms` atoms (host hint, mlffd_set_structure_hint); a larger structure raises
// DeviceStatus::hint_violation and the host reruns the step on the generic kernels.
#pragma once
#include "message.cuh"

namespace mlffd {

constexpr int kStagedThreads = 512;

template <int H>
constexpr size_t staged_smem_bytes(int max_atoms) { return (size_t)max_atoms * 4 * H * sizeof(float) + 16; }

// copy n rows of `width` floats (contiguous) global -> shared
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int count4) {
    for (int i = threadIdx.x; i < count4; i += blockDim.x)
        reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
}

template <int H, bool LAYER0>
__global__ void __launch_bounds__(kStagedThreads)
message_forward_staged_kernel(const int* __restrict__ offsets, int num_structures, int max_atoms,
                              const int* __restrict__ rowptr, const int* __restrict__ col,
                              const int* __restrict__ pair, const float4* __restrict__ geo,
                              const float* __restrict__ filt, const float* __restrict__ s_in,
                              const float* __restrict__ v_in, float* __restrict__ s_msg,
                              float* __restrict__ v_msg, DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    extern __shared__ __align__(16) float stage[];
    float* s_s = stage;                              // [n][H]
    float* v_s = stage + (size_t)max_atoms * H;      // [n][3H]
    int* next_row = reinterpret_cast<int*>(stage + (size_t)max_atoms * 4 * H);
    const int lane = threadIdx.x & 31, sub = lane / M::LPA, c4 = (lane % M::LPA) * 4;
    for (int b = blockIdx.x; b < num_structures; b += gridDim.x) {
        const int lo = __ldg(offsets + b), n = __ldg(offsets + b + 1) - lo;
        if (n > max_atoms) { if (threadIdx.x == 0) status->hint_violation = 1; continue; }
        __syncthreads();   // previous structure fully consumed
        stage_rows(s_s, s_in + (size_t)lo * H, n * H / 4);
        if (!LAYER0) stage_rows(v_s, v_in + (size_t)lo * 3 * H, n * 3 * H / 4);
        if (threadIdx.x == 0) *next_row = 0;
        __syncthreads();
        for (;;) {
            int r0 = 0;
            if (lane == 0) r0 = atomicAdd(next_row, M::APW);
            r0 = __shfl_sync(0xffffffffu, r0, 0);
            if (r0 >= n) break;
            const int rl = r0 + sub;                 // local row
            const bool valid = rl < n;
            const int j = lo + rl;
            const int e0 = valid ? rowptr[j] : 0, e1 = valid ? rowptr[j + 1] : 0;
            float4 acc_s = make4(0.f), acc_x = make4(0.f), acc_y = make4(0.f), acc_z = make4(0.f);
#pragma unroll 4
            for (int e = e0; e < e1; ++e) {
                const int il = __ldg(col + e) - lo;
                const float4 g = __ldg(geo + e);
                const float* f = filt + (size_t)__ldg(pair + e) * (3 * H) + c4;
                const float4 fa = ldg4(f);
                const float4 fc = ldg4(f + 2 * H);
                acc_s = fma4(*reinterpret_cast<const float4*>(s_s + il * H + c4), fa, acc_s);
                if (!LAYER0) {
                    const float4 fb = ldg4(f + H);
                    const float* vi = v_s + il * 3 * H + c4;
                    acc_x = fma4(*reinterpret_cast<const float4*>(vi), fb, acc_x);
                    acc_y = fma4(*reinterpret_cast<const float4*>(vi + H), fb, acc_y);
                    acc_z = fma4(*reinterpret_cast<const float4*>(vi + 2 * H), fb, acc_z);
                }
                acc_x = fma4s(g.x, fc, acc_x);
                acc_y = fma4s(g.y, fc, acc_y);
                acc_z = fma4s(g.z, fc, acc_z);
            }
            if (valid) {
                st4(s_msg + (size_t)j * H + c4, add4(*reinterpret_cast<const float4*>(s_s + rl * H + c4), acc_s));
                float* vo = v_msg + (size_t)j * 3 * H + c4;
                if (LAYER0) {
                    st4(vo, acc_x); st4(vo + H, acc_y); st4(vo + 2 * H, acc_z);
                } else {
                    const float* vj = v_s + rl * 3 * H + c4;
                    st4(vo, add4(*reinterpret_cast<const float4*>(vj), acc_x));
                    st4(vo + H, add4(*reinterpret_cast<const float4*>(vj + H), acc_y));
                    st4(vo + 2 * H, add4(*reinterpret_cast<const float4*>(vj + 2 * H), acc_z));
                }
            }
        }
    }
}

// Reverse, part 1: per-edge adjoints (d_bar through the filters, u_bar).  Stages (s_in, v_in).
template <int H, bool LAYER0, bool ACCUMULATE>
__global__ void __launch_bounds__(kStagedThreads)
message_backward_edges_staged_kernel(const int* __restrict__ offsets, int num_structures, int max_atoms,
                                     const int* __restrict__ rowptr, const int* __restrict__ col,
                                     const int* __restrict__ pair, const float4* __restrict__ geo,
                                     const float* __restrict__ filt, const float* __restrict__ dfilt,
                                     const float* __restrict__ s_in, const float* __restrict__ v_in,
                                     const float* __restrict__ sbar_m, const float* __restrict__ vbar_m,
                                     float4* __restrict__ edge_adj, DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    extern __shared__ __align__(16) float stage[];
    float* s_s = stage;
    float* v_s = stage + (size_t)max_atoms * H;
    int* next_row = reinterpret_cast<int*>(stage + (size_t)max_atoms * 4 * H);
    const int lane = threadIdx.x & 31, sub = lane / M::LPA, gl = lane % M::LPA, c4 = gl * 4;
    for (int b = blockIdx.x; b < num_structures; b += gridDim.x) {
        const int lo = __ldg(offsets + b), n = __ldg(offsets + b + 1) - lo;
        if (n > max_atoms) { if (threadIdx.x == 0) status->hint_violation = 1; continue; }
        __syncthreads();
        stage_rows(s_s, s_in + (size_t)lo * H, n * H / 4);
        if (!LAYER0) stage_rows(v_s, v_in + (size_t)lo * 3 * H, n * 3 * H / 4);
        if (threadIdx.x == 0) *next_row = 0;
        __syncthreads();
        for (;;) {
            int r0 = 0;
            if (lane == 0) r0 = atomicAdd(next_row, M::APW);
            r0 = __shfl_sync(0xffffffffu, r0, 0);
            if (r0 >= n) break;
            const int rl = r0 + sub;
            const bool valid = rl < n;
            const int i = lo + rl;
            const int e0 = valid ? rowptr[i] : 0;
            const int deg = valid ? rowptr[i + 1] - e0 : 0;
            const int max_deg = (M::APW > 1) ? group_max_int<32>(deg) : deg;
            float4 sb = make4(0.f), vbx = make4(0.f), vby = make4(0.f), vbz = make4(0.f);
            if (valid) {
                sb = ldg4(sbar_m + (size_t)i * H + c4);
                const float* vb = vbar_m + (size_t)i * 3 * H + c4;
                vbx = ldg4(vb); vby = ldg4(vb + H); vbz = ldg4(vb + 2 * H);
            }
            for (int k = 0; k < max_deg; ++k) {
                const bool active = k < deg;
                float d_part = 0.f, ux_part = 0.f, uy_part = 0.f, uz_part = 0.f;
                const int e = e0 + k;
                if (active) {
                    const int jl = __ldg(col + e) - lo;
                    const float4 g = __ldg(geo + e);
                    const size_t prow = (size_t)__ldg(pair + e) * (3 * H) + c4;
                    const float4 fc = ldg4(filt + prow + 2 * H);
                    const float4 dfa = ldg4(dfilt + prow);
                    const float4 dfc = ldg4(dfilt + prow + 2 * H);
                    const float4 sj = *reinterpret_cast<const float4*>(s_s + jl * H + c4);
                    const float4 abar = mul4(sj, sb);
                    const float4 cbar = fma4s(g.x, vbx, fma4s(g.y, vby, fma4s(g.z, vbz, make4(0.f))));
                    d_part = dot4(abar, dfa) + dot4(cbar, dfc);
                    ux_part = dot4(fc, vbx);
                    uy_part = dot4(fc, vby);
                    uz_part = dot4(fc, vbz);
                    if (!LAYER0) {
                        const float4 dfb = ldg4(dfilt + prow + H);
                        const float* vj = v_s + jl * 3 * H + c4;
                        const float4 bbar = fma4(*reinterpret_cast<const float4*>(vj), vbx,
                                                 fma4(*reinterpret_cast<const float4*>(vj + H), vby,
                                                      mul4(*reinterpret_cast<const float4*>(vj + 2 * H), vbz)));
                        d_part += dot4(bbar, dfb);
                    }
                }
                d_part = group_sum<M::LPA>(d_part);
                ux_part = group_sum<M::LPA>(ux_part);
                uy_part = group_sum<M::LPA>(uy_part);
                uz_part = group_sum<M::LPA>(uz_part);
                if (active && gl == 0) {
                    float4 out = make_float4(ux_part, uy_part, uz_part, d_part);
                    if (ACCUMULATE) out = add4(out, edge_adj[e]);
                    edge_adj[e] = out;
                }
            }
        }
    }
}

// Reverse, part 2: adjoints of the layer inputs.  Stages (s_bar', v_bar').  Not used for layer 0.
template <int H>
__global__ void __launch_bounds__(kStagedThreads)
message_backward_atoms_staged_kernel(const int* __restrict__ offsets, int num_structures, int max_atoms,
                                     const int* __restrict__ rowptr, const int* __restrict__ col,
                                     const int* __restrict__ pair, const float* __restrict__ filt,
                                     const float* __restrict__ sbar_m, const float* __restrict__ vbar_m,
                                     float* __restrict__ sbar_in, float* __restrict__ vbar_in,
                                     DeviceStatus* __restrict__ status) {
    using M = MsgTraits<H>;
    if (status->overflow) return;
    extern __shared__ __align__(16) float stage[];
    float* s_s = stage;
    float* v_s = stage + (size_t)max_atoms * H;
    int* next_row = reinterpret_cast<int*>(stage + (size_t)max_atoms * 4 * H);
    const int lane = threadIdx.x & 31, sub = lane / M::LPA, c4 = (lane % M::LPA) * 4;
    for (int b = blockIdx.x; b < num_structures; b += gridDim.x) {
        const int lo = __ldg(offsets + b), n = __ldg(offsets + b + 1) - lo;
        if (n > max_atoms) { if (threadIdx.x == 0) status->hint_violation = 1; continue; }
        __syncthreads();
        stage_rows(s_s, sbar_m + (size_t)lo * H, n * H / 4);
        stage_rows(v_s, vbar_m + (size_t)lo * 3 * H, n * 3 * H / 4);
        if (threadIdx.x == 0) *next_row = 0;
        __syncthreads();
        for (;;) {
            int r0 = 0;
            if (lane == 0) r0 = atomicAdd(next_row, M::APW);
            r0 = __shfl_sync(0xffffffffu, r0, 0);
            if (r0 >= n) break;
            const int rl = r0 + sub;
            const bool valid = rl < n;
            const int i = lo + rl;
            const int e0 = valid ? rowptr[i] : 0, e1 = valid ? rowptr[i + 1] : 0;
            float4 acc_s = make4(0.f), acc_x = make4(0.f), acc_y = make4(0.f), acc_z = make4(0.f);
            if (valid) {   // residual path
                acc_s = *reinterpret_cast<const float4*>(s_s + rl * H + c4);
                const float* vb = v_s + rl * 3 * H + c4;
                acc_x = *reinterpret_cast<const float4*>(vb);
                acc_y = *reinterpret_cast<const float4*>(vb + H);
                acc_z = *reinterpret_cast<const float4*>(vb + 2 * H);
            }
#pragma unroll 4
            for (int e = e0; e < e1; ++e) {
                const int jl = __ldg(col + e) - lo;
                const float* f = filt + (size_t)__ldg(pair + e) * (3 * H) + c4;
                const float4 fa = ldg4(f);
                const float4 fb = ldg4(f + H);
                acc_s = fma4(fa, *reinterpret_cast<const float4*>(s_s + jl * H + c4), acc_s);
                const float* vbj = v_s + jl * 3 * H + c4;
                acc_x = fma4(fb, *reinterpret_cast<const float4*>(vbj), acc_x);
                acc_y = fma4(fb, *reinterpret_cast<const float4*>(vbj + H), acc_y);
                acc_z = fma4(fb, *reinterpret_cast<const float4*>(vbj + 2 * H), acc_z);
            }
            if (valid) {
                st4(sbar_in + (size_t)i * H + c4, acc_s);
                float* vo = vbar_in + (size_t)i * 3 * H + c4;
                st4(vo, acc_x); st4(vo + H, acc_y); st4(vo + 2 * H, acc_z);
            }
        }
    }
}

}  // namespace mlffd
