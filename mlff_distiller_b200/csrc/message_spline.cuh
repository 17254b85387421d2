// PaiNN message block with the radial filter evaluated in-kernel from a shared-memory spline table:
// no per-step filter table, no filter kernel, no [E,3H] tensor anywhere in HBM.
//
// What it restates: PaiNNMessage.forward (reference src/mlff_distiller/models/student_model.py:
// 346-387) and its reverse, with (a_e, b_e, c_e) = f_l(d_e) read from the per-model
// quintic B-spline of spline_table.h instead of the two dense layers of rbf_to_scalar.
//
// Decomposition: the H channels are cut into slices of 32.  A CTA owns ONE slice of ONE layer: its
// (192 + 5) x 3 x 32 coefficient block (75 648 B) sits in shared memory for the whole launch.  Eight
// lanes own a CSR row (one float4 of channels per lane, 4 rows per warp) and accumulate it in
// registers in source-ascending order -- the order the reference's CPU index_add_ applies: no atomics,
// bit-reproducible.  The four row groups of a warp advance through their own row sequences
// independently (a group that finishes a row writes it out and takes its next one while the others
// keep going), so ragged degrees do not idle lanes.  Per directed edge and lane the kernel reads
// 6 x 3 float4 coefficients from shared memory (conflict-free: a quarter-warp reads 128 contiguous
// bytes), gathers the neighbour's feature slice (4 x 128 B segments) from L2 and issues packed
// two-wide FP32 FMAs (fma.rn.f32x2, SASS FFMA2).  The six basis weights of an edge (and their
// d-derivatives) depend on the distance only; spline_basis_kernel computes them once per step and all
// layers, both directions and all slices read them (2 x 32 B per edge).
//
// Bound: the shared-memory pipe (128 B/clk/SM): 18 LDS.128 per directed edge and slice.  HBM traffic
// is the compulsory feature bytes only (2 N 16H + E 20 per layer).
#pragma once
#include "common.cuh"
#include "spline_table.h"

namespace mlffd {

constexpr int kSplineRowFloat4 = 3 * kSliceChannels / 4;          // float4 per table row: 24
constexpr int kSplineSliceFloat4 = kSplineRows * kSplineRowFloat4;
constexpr size_t kSplineSmemBytes = (size_t)kSplineSliceFloat4 * sizeof(float4);   // 75 648 B

// ---- packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): four floats = two 64-bit registers ----
typedef ulonglong2 pk4;
__device__ __forceinline__ unsigned long long pk_dup(float s) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(s));
    return r;
}
__device__ __forceinline__ pk4 pk_zero() { return make_ulonglong2(0ull, 0ull); }
__device__ __forceinline__ pk4 pk_fma(pk4 a, pk4 b, pk4 c) {
    pk4 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.x) : "l"(a.x), "l"(b.x), "l"(c.x));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.y) : "l"(a.y), "l"(b.y), "l"(c.y));
    return d;
}
__device__ __forceinline__ pk4 pk_fma_s(float s, pk4 b, pk4 c) {   // s * b + c, s broadcast by the instruction
    const unsigned long long ss = pk_dup(s);
    pk4 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.x) : "l"(ss), "l"(b.x), "l"(c.x));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.y) : "l"(ss), "l"(b.y), "l"(c.y));
    return d;
}
__device__ __forceinline__ pk4 pk_mul(pk4 a, pk4 b) {
    pk4 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.x) : "l"(a.x), "l"(b.x));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.y) : "l"(a.y), "l"(b.y));
    return d;
}
__device__ __forceinline__ pk4 pk_mul_s(float s, pk4 b) {
    const unsigned long long ss = pk_dup(s);
    pk4 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.x) : "l"(ss), "l"(b.x));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.y) : "l"(ss), "l"(b.y));
    return d;
}
__device__ __forceinline__ pk4 pk_add(pk4 a, pk4 b) {
    pk4 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.x) : "l"(a.x), "l"(b.x));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.y) : "l"(a.y), "l"(b.y));
    return d;
}
__device__ __forceinline__ float pk_hsum(pk4 a) {   // (x + z) + (y + w)
    unsigned long long t;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(a.x), "l"(a.y));
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(t));
    return x + y;
}
__device__ __forceinline__ pk4 pk_ldg(const float* p) { return __ldg(reinterpret_cast<const pk4*>(p)); }
__device__ __forceinline__ void pk_st(float* p, pk4 v) { *reinterpret_cast<pk4*>(p) = v; }

// ---- per-edge basis weights ------------------------------------------------------------------
// Uniform quintic B-spline basis at u in [0,1] (see spline_table.h) and its u-derivative.
__device__ __forceinline__ void quintic_basis(float u, float (&b)[6], float (&db)[6]) {
    const float v = 1.0f - u;
    const float b2_0 = 0.5f * v * v, b2_2 = 0.5f * u * u, b2_1 = 1.0f - b2_0 - b2_2;
    const float k3 = 1.0f / 3.0f;
    const float b3_0 = k3 * v * b2_0;
    const float b3_1 = k3 * ((u + 2.0f) * b2_0 + (2.0f - u) * b2_1);
    const float b3_2 = k3 * ((u + 1.0f) * b2_1 + (3.0f - u) * b2_2);
    const float b3_3 = k3 * u * b2_2;
    const float b4_0 = 0.25f * v * b3_0;
    const float b4_1 = 0.25f * ((u + 3.0f) * b3_0 + (2.0f - u) * b3_1);
    const float b4_2 = 0.25f * ((u + 2.0f) * b3_1 + (3.0f - u) * b3_2);
    const float b4_3 = 0.25f * ((u + 1.0f) * b3_2 + (4.0f - u) * b3_3);
    const float b4_4 = 0.25f * u * b3_3;
    b[0] = 0.2f * v * b4_0;
    b[1] = 0.2f * ((u + 4.0f) * b4_0 + (2.0f - u) * b4_1);
    b[2] = 0.2f * ((u + 3.0f) * b4_1 + (3.0f - u) * b4_2);
    b[3] = 0.2f * ((u + 2.0f) * b4_2 + (4.0f - u) * b4_3);
    b[4] = 0.2f * ((u + 1.0f) * b4_3 + (5.0f - u) * b4_4);
    b[5] = 0.2f * u * b4_4;
    db[0] = -b4_0; db[1] = b4_0 - b4_1; db[2] = b4_1 - b4_2;
    db[3] = b4_2 - b4_3; db[4] = b4_3 - b4_4; db[5] = b4_4;
}

// segment and local parameter of distance d (x = d / h); d == rc maps to the right end of the last segment
__device__ __forceinline__ int spline_segment(float d, float inv_h, float& u) {
    const int seg = min((int)(d * inv_h), kSplineIntervals - 1);
    u = fmaf(d, inv_h, -(float)seg);   // one rounding: the product is not rounded before the segment is removed
    return seg;
}

// Per-edge record of the spline message kernels, written once per step after the neighbour list
// (every layer, both directions and all channel slices read it):
//   erec[4e]     = (source atom as int bits, u_x, u_y, u_z)      unit vector of the edge
//   erec[4e + 1] = (b0, b1, b2, b3)                              basis weights at the edge's distance
//   erec[4e + 2] = (b4, b5, segment as int bits, b4' / h)
//   erec[4e + 3] = (b0', b1', b2', b3') / h                      d/dd of the weights; b5' = -(b0' + .. + b4')
// The forward pass reads the first three entries, the reverse pass all four.
__global__ void __launch_bounds__(256)
spline_basis_kernel(const int* __restrict__ colidx, const float4* __restrict__ geo, float inv_h,
                    float4* __restrict__ erec, const DeviceStatus* __restrict__ status) {
    if (status->overflow) return;
    const int E = status->num_edges;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        float u, b[6], db[6];
        const float4 g = __ldg(geo + e);
        const int seg = spline_segment(g.w, inv_h, u);
        quintic_basis(u, b, db);
        erec[4 * (size_t)e] = make_float4(__int_as_float(__ldg(colidx + e)), g.x, g.y, g.z);
        erec[4 * (size_t)e + 1] = make_float4(b[0], b[1], b[2], b[3]);
        erec[4 * (size_t)e + 2] = make_float4(b[4], b[5], __int_as_float(seg), db[4] * inv_h);
        erec[4 * (size_t)e + 3] = make_float4(db[0] * inv_h, db[1] * inv_h, db[2] * inv_h, db[3] * inv_h);
    }
}

__device__ __forceinline__ void load_spline_slice(float4* tab, const float* __restrict__ table, int slice) {
    const float4* src = reinterpret_cast<const float4*>(table) + (size_t)slice * kSplineSliceFloat4;
    for (int idx = threadIdx.x; idx < kSplineSliceFloat4; idx += blockDim.x) tab[idx] = __ldg(src + idx);
    __syncthreads();
}

// value of one component (comp: 0 = a, 1 = b, 2 = c) for this lane's four channels
__device__ __forceinline__ pk4 spline_value(const pk4* row, int comp, const float (&b)[6]) {
    pk4 acc = pk_mul_s(b[0], row[comp * 8]);
#pragma unroll
    for (int j = 1; j < 6; ++j) acc = pk_fma_s(b[j], row[j * kSplineRowFloat4 + comp * 8], acc);
    return acc;
}
__device__ __forceinline__ void spline_value_deriv(const pk4* row, int comp, const float (&b)[6],
                                                   const float (&db)[6], pk4& val, pk4& der) {
    const pk4 c0 = row[comp * 8];
    val = pk_mul_s(b[0], c0);
    der = pk_mul_s(db[0], c0);
#pragma unroll
    for (int j = 1; j < 6; ++j) {
        const pk4 c = row[j * kSplineRowFloat4 + comp * 8];
        val = pk_fma_s(b[j], c, val);
        der = pk_fma_s(db[j], c, der);
    }
}
__device__ __forceinline__ pk4 spline_deriv(const pk4* row, int comp, const float (&db)[6]) {
    pk4 acc = pk_mul_s(db[0], row[comp * 8]);
#pragma unroll
    for (int j = 1; j < 6; ++j) acc = pk_fma_s(db[j], row[j * kSplineRowFloat4 + comp * 8], acc);
    return acc;
}

// ---- a row group's walk through its rows -------------------------------------------------------
// Group g of CTA partition p owns rows p * R + g, (p + parts) * R + g, ... (R = row groups per CTA).
// The next row's edge range is fetched one row ahead, and the source index of the next edge one edge
// ahead (across row boundaries), so the only exposed global-memory latency per edge is the gather.
// With `lowptr` (prefix sums of the number of entries with source < row) a row's walk starts at its first
// entry with source > row: the upper half of every undirected pair (pair-once kernels).
struct RowCursor {
    int row, e, e_end, nrow, ne0, ne1, step;
    float4 head;   // first record entry (source, unit vector) of the edge to process next
    __device__ __forceinline__ static void row_range(const int* __restrict__ rowptr, const int* __restrict__ lowptr,
                                                     int r, int& first, int& last) {
        first = __ldg(rowptr + r); last = __ldg(rowptr + r + 1);
        if (lowptr != nullptr) first += __ldg(lowptr + r + 1) - __ldg(lowptr + r);
    }
    __device__ __forceinline__ void fetch_next_range(const int* __restrict__ rowptr, int num_atoms,
                                                     const int* __restrict__ lowptr) {
        ne0 = ne1 = 0;
        if (nrow < num_atoms) row_range(rowptr, lowptr, nrow, ne0, ne1);
    }
    __device__ __forceinline__ void init(int first_row, int row_step, const int* __restrict__ rowptr,
                                         const float4* __restrict__ erec, int num_atoms,
                                         const int* __restrict__ lowptr = nullptr) {
        row = first_row; step = row_step; e = e_end = 0; head = make4(0.f);
        if (row < num_atoms) row_range(rowptr, lowptr, row, e, e_end);
        nrow = row + step;
        fetch_next_range(rowptr, num_atoms, lowptr);
        if (e < e_end) head = __ldg(erec + 4 * (size_t)e);
    }
    __device__ __forceinline__ void advance_row(const int* __restrict__ rowptr, const float4* __restrict__ erec,
                                                int num_atoms, bool had_edges,
                                                const int* __restrict__ lowptr = nullptr) {
        row = nrow; e = ne0; e_end = ne1; nrow += step;
        fetch_next_range(rowptr, num_atoms, lowptr);
        // the head of the new row's first edge was prefetched by the last edge of the previous row
        if (!had_edges && e < e_end) head = __ldg(erec + 4 * (size_t)e);
    }
    // called while edge e is being processed: returns its head and prefetches the next edge's
    __device__ __forceinline__ float4 take_head(const float4* __restrict__ erec) {
        const float4 cur = head;
        const int nxt = (e + 1 < e_end) ? e + 1 : ((ne0 < ne1) ? ne0 : -1);
        if (nxt >= 0) head = __ldg(erec + 4 * (size_t)nxt);
        return cur;
    }
};

// ---------------------------------------------------------------------------------------------
// forward:  s'_j = s_j + sum_{e=(i->j)} s_i * a(d_e) ;  v'_j = v_j + sum_e (v_i * b(d_e) + u_e c(d_e))
// grid = slices x partitions.  LAYER0: v_in == 0 (not stored): the v gather and the b filter are skipped.
// ---------------------------------------------------------------------------------------------
template <bool LAYER0, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
spline_message_forward_kernel(const float* __restrict__ table, int H,
                              const int* __restrict__ rowptr, const float4* __restrict__ erec,
                              const float* __restrict__ s_in, const float* __restrict__ v_in,
                              float* __restrict__ s_msg, float* __restrict__ v_msg, int num_atoms,
                              const DeviceStatus* __restrict__ status) {
    extern __shared__ float4 spline_tab[];
    if (status->overflow) return;
    constexpr int kGroups = THREADS / 8;
    const int slices = H / kSliceChannels;
    const int slice = blockIdx.x % slices, part = blockIdx.x / slices, parts = gridDim.x / slices;
    load_spline_slice(spline_tab, table, slice);
    const int sub = threadIdx.x >> 3, l8 = threadIdx.x & 7;
    const int ch = slice * kSliceChannels + l8 * 4;
    const pk4* tab = reinterpret_cast<const pk4*>(spline_tab) + l8;
    RowCursor c;
    c.init(part * kGroups + sub, parts * kGroups, rowptr, erec, num_atoms);
    pk4 acc_s = pk_zero(), acc_x = pk_zero(), acc_y = pk_zero(), acc_z = pk_zero();
    bool had_edges = false;
    while (__any_sync(0xffffffffu, c.row < num_atoms)) {
        if (c.row < num_atoms && c.e == c.e_end) {   // row complete (or empty): write it out, take the next one
            const int j = c.row;
            pk_st(s_msg + (size_t)j * H + ch, pk_add(pk_ldg(s_in + (size_t)j * H + ch), acc_s));
            float* vo = v_msg + (size_t)j * 3 * H + ch;
            if (LAYER0) {
                pk_st(vo, acc_x); pk_st(vo + H, acc_y); pk_st(vo + 2 * H, acc_z);
            } else {
                const float* vj = v_in + (size_t)j * 3 * H + ch;
                pk_st(vo, pk_add(pk_ldg(vj), acc_x));
                pk_st(vo + H, pk_add(pk_ldg(vj + H), acc_y));
                pk_st(vo + 2 * H, pk_add(pk_ldg(vj + 2 * H), acc_z));
            }
            acc_s = acc_x = acc_y = acc_z = pk_zero();
            c.advance_row(rowptr, erec, num_atoms, had_edges);
            had_edges = false;
        }
        if (c.row < num_atoms && c.e < c.e_end) {
            const int e = c.e;
            const float4 g = c.take_head(erec);   // (source, unit vector)
            const int i = __float_as_int(g.x);
            const pk4 si = pk_ldg(s_in + (size_t)i * H + ch);
            pk4 vix, viy, viz;
            if (!LAYER0) {
                const float* vi = v_in + (size_t)i * 3 * H + ch;
                vix = pk_ldg(vi); viy = pk_ldg(vi + H); viz = pk_ldg(vi + 2 * H);
            }
            const float4 w0 = __ldg(erec + 4 * (size_t)e + 1), w1 = __ldg(erec + 4 * (size_t)e + 2);
            const float b[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
            const pk4* row = tab + __float_as_int(w1.z) * kSplineRowFloat4;
            acc_s = pk_fma(si, spline_value(row, 0, b), acc_s);
            if (!LAYER0) {
                const pk4 fb = spline_value(row, 1, b);
                acc_x = pk_fma(vix, fb, acc_x);
                acc_y = pk_fma(viy, fb, acc_y);
                acc_z = pk_fma(viz, fb, acc_z);
            }
            const pk4 fc = spline_value(row, 2, b);
            acc_x = pk_fma_s(g.y, fc, acc_x);
            acc_y = pk_fma_s(g.z, fc, acc_y);
            acc_z = pk_fma_s(g.w, fc, acc_z);
            c.e = e + 1;
            had_edges = true;
        }
    }
}

// Sum of 4 per-lane values over the 8 lanes of a row group with 2 + 1 + 1 shuffles; afterwards lane l8
// holds the total of value 2 * bit(l8, 4) + bit(l8, 2) (lanes that differ in bit 1 hold duplicates).
// `full` = the lanes of the warp that execute the call (whole row groups).
__device__ __forceinline__ float group8_sum4(const float (&v)[4], int l8, unsigned full) {
    const bool b0 = (l8 & 4) != 0;
    float w[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float send = b0 ? v[t] : v[t + 2], keep = b0 ? v[t + 2] : v[t];
        w[t] = keep + __shfl_xor_sync(full, send, 4);
    }
    const bool b1 = (l8 & 2) != 0;
    float y = (b1 ? w[1] : w[0]) + __shfl_xor_sync(full, b1 ? w[0] : w[1], 2);
    y += __shfl_xor_sync(full, y, 1);
    return y;
}

// ---------------------------------------------------------------------------------------------
// reverse of the message block for layer l, one slice per CTA (same decomposition as the forward).
//   in : sbar_m / vbar_m   adjoints of (s_msg, v_msg)        [N,H] / [N,3,H]
//        s_in / v_in       the layer's input features
//   out: sbar_in / vbar_in adjoints of the layer inputs (not for LAYER0: the embedding does not
//        depend on positions)
//        edge_adj (this layer's, this slice's slab): (dE/du_x, dE/du_y, dE/du_z, dE/dd through the
//        filter) of every directed edge restricted to the slice's channels; the force kernel sums the
//        L x slices slabs.
// Row i walks its CSR entries e = (j -> i) one by one (uniform work, no pairing).  Edge e and its reverse
// r = (i -> j) have the same distance, hence the same filter, and u_r = -u_e.
// LAYER0 (v_in == 0, no input adjoints, nothing to accumulate per row): every undirected PAIR once.  Row i
// walks only its entries with j > i (RowCursor with lowptr) and forms the sum over both directions -- only
// u_bar_e - u_bar_r and d_bar_e + d_bar_r enter the forces (r_bar_e - r_bar_r, readout.cuh):
//   d_bar = (s_j s_bar'_i + s_i s_bar'_j) . a'(d) + (sum_x u_e[x] (v_bar'_i - v_bar'_j)[x]) . c'(d)
//   u_bar[x] = c(d) . (v_bar'_i - v_bar'_j)[x]                         -> slab[e], e upper   ("pair" slab)
// 12 coefficient reads per pair instead of 24.
// other layers: everything row i computes comes from the REVERSE edge r, whose source is i itself, so one
// gather of the neighbour's adjoints (s_bar'_j, v_bar'_j: 4 x 128 B) serves both the scatter-turned-gather
//   s_bar_i += a(d) * s_bar'_j,  v_bar_i[x] += b(d) * v_bar'_j[x]
// and the edge adjoint of r, with the row's own FEATURES (registers) in place of gathered ones:
//   a_bar = s_i * s_bar'_j,  b_bar = sum_x v_i[x] * v_bar'_j[x],  c_bar = -sum_x u_e[x] v_bar'_j[x],
//   d_bar_r = a_bar . a' + b_bar . b' + c_bar . c',  u_bar_r[x] = c(d) . v_bar'_j[x]
//                                                                     -> slab[e] = adjoint of rev(e) ("swapped" slab)
// The force / virial kernels (readout.cuh) read both e and rev(e) anyway and know which slabs are swapped.
// Half the gathers of a formulation that computes the adjoint of e itself (features AND adjoints of j).
// The filter MLP is never back-propagated: f' is the exact derivative of the spline the forward used.
// ---------------------------------------------------------------------------------------------
template <bool LAYER0, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
spline_message_backward_kernel(const float* __restrict__ table, int H,
                               const int* __restrict__ rowptr, const float4* __restrict__ erec,
                               const float* __restrict__ s_in, const float* __restrict__ v_in,
                               const float* __restrict__ sbar_m, const float* __restrict__ vbar_m,
                               float* __restrict__ sbar_in, float* __restrict__ vbar_in,
                               float4* __restrict__ edge_adj, size_t slab_stride, int num_atoms,
                               const int* __restrict__ lowptr, const DeviceStatus* __restrict__ status) {
    extern __shared__ float4 spline_tab[];
    if (status->overflow) return;
    constexpr int kGroups = THREADS / 8;
    const int slices = H / kSliceChannels;
    const int slice = blockIdx.x % slices, part = blockIdx.x / slices, parts = gridDim.x / slices;
    load_spline_slice(spline_tab, table, slice);
    float* adj_out = reinterpret_cast<float*>(edge_adj + (size_t)slice * slab_stride);
    const int sub = threadIdx.x >> 3, l8 = threadIdx.x & 7;
    const int ch = slice * kSliceChannels + l8 * 4;
    const pk4* tab = reinterpret_cast<const pk4*>(spline_tab) + l8;
    const int held = ((l8 & 4) ? 2 : 0) + ((l8 & 2) ? 1 : 0);   // which reduced value this lane ends up with
    const int* upper = LAYER0 ? lowptr : nullptr;   // LAYER0: pairs once, rows start at their first j > i
    RowCursor c;
    c.init(part * kGroups + sub, parts * kGroups, rowptr, erec, num_atoms, upper);
    // LAYER0: the row's adjoints (sb, vb*) and own_s;  other layers: the row's features (own_s, own_*)
    pk4 sb = pk_zero(), vbx = pk_zero(), vby = pk_zero(), vbz = pk_zero();
    pk4 own_s = pk_zero(), own_x = pk_zero(), own_y = pk_zero(), own_z = pk_zero();
    pk4 acc_s = pk_zero(), acc_x = pk_zero(), acc_y = pk_zero(), acc_z = pk_zero();
    bool had_edges = false, loaded = false;   // loaded: the row's own vectors are in registers
    while (__any_sync(0xffffffffu, c.row < num_atoms)) {
        if (c.row < num_atoms && loaded && c.e == c.e_end) {   // row complete: write it out, take the next one
            if (!LAYER0) {
                const int i = c.row;
                pk_st(sbar_in + (size_t)i * H + ch, acc_s);
                float* vo = vbar_in + (size_t)i * 3 * H + ch;
                pk_st(vo, acc_x); pk_st(vo + H, acc_y); pk_st(vo + 2 * H, acc_z);
            }
            c.advance_row(rowptr, erec, num_atoms, had_edges, upper);
            had_edges = false;
            loaded = false;
        }
        if (c.row < num_atoms && !loaded) {
            const int i = c.row;
            const float* vb = vbar_m + (size_t)i * 3 * H + ch;
            if (LAYER0) {
                if (c.e < c.e_end) {   // a row without upper entries has nothing to do
                    sb = pk_ldg(sbar_m + (size_t)i * H + ch);
                    vbx = pk_ldg(vb); vby = pk_ldg(vb + H); vbz = pk_ldg(vb + 2 * H);
                    own_s = pk_ldg(s_in + (size_t)i * H + ch);
                }
            } else {   // the row's adjoints start the residual path; its features stay for the edge adjoints
                acc_s = pk_ldg(sbar_m + (size_t)i * H + ch);
                acc_x = pk_ldg(vb); acc_y = pk_ldg(vb + H); acc_z = pk_ldg(vb + 2 * H);
                own_s = pk_ldg(s_in + (size_t)i * H + ch);
                const float* vi = v_in + (size_t)i * 3 * H + ch;
                own_x = pk_ldg(vi); own_y = pk_ldg(vi + H); own_z = pk_ldg(vi + 2 * H);
            }
            loaded = true;
        }
        const bool active = c.row < num_atoms && c.e < c.e_end;
        const unsigned active_lanes = __ballot_sync(0xffffffffu, active);   // whole row groups: the edge reduction below shuffles among them only
        if (active) {
            const int e = c.e;
            const float4 g = c.take_head(erec);   // (source, unit vector)
            const int j = __float_as_int(g.x);
            // the neighbour's adjoints s_bar'_j, v_bar'_j; LAYER0 also its features s_j
            const pk4 gs = pk_ldg(sbar_m + (size_t)j * H + ch);
            const float* vbj = vbar_m + (size_t)j * 3 * H + ch;
            const pk4 gx = pk_ldg(vbj), gy = pk_ldg(vbj + H), gz = pk_ldg(vbj + 2 * H);
            pk4 gf = pk_zero();
            if (LAYER0) gf = pk_ldg(s_in + (size_t)j * H + ch);
            const float4 w0 = __ldg(erec + 4 * (size_t)e + 1), w1 = __ldg(erec + 4 * (size_t)e + 2);
            const float4 d0 = __ldg(erec + 4 * (size_t)e + 3);
            const float b[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
            const float db[6] = {d0.x, d0.y, d0.z, d0.w, w1.w, -((d0.x + d0.y) + (d0.z + d0.w) + w1.w)};
            const pk4* row = tab + __float_as_int(w1.z) * kSplineRowFloat4;
            pk4 dv;   // per-channel products of gate adjoints and filter derivatives, summed at the end
            pk4 fc, dfc;
            float part4[4];
            if (LAYER0) {
                dv = pk_mul(pk_fma(gf, sb, pk_mul(own_s, gs)), spline_deriv(row, 0, db));
                spline_value_deriv(row, 2, b, db, fc, dfc);
                const pk4 dx = pk_fma_s(-1.0f, gx, vbx), dy = pk_fma_s(-1.0f, gy, vby), dz = pk_fma_s(-1.0f, gz, vbz);
                const pk4 cbar = pk_fma_s(g.y, dx, pk_fma_s(g.z, dy, pk_mul_s(g.w, dz)));
                dv = pk_fma(cbar, dfc, dv);
                part4[0] = pk_hsum(pk_mul(fc, dx)); part4[1] = pk_hsum(pk_mul(fc, dy));
                part4[2] = pk_hsum(pk_mul(fc, dz));
            } else {
                pk4 fa, dfa, fb, dfb;
                spline_value_deriv(row, 0, b, db, fa, dfa);
                dv = pk_mul(pk_mul(own_s, gs), dfa);
                acc_s = pk_fma(fa, gs, acc_s);
                spline_value_deriv(row, 1, b, db, fb, dfb);
                const pk4 bbar = pk_fma(own_x, gx, pk_fma(own_y, gy, pk_mul(own_z, gz)));
                dv = pk_fma(bbar, dfb, dv);
                acc_x = pk_fma(fb, gx, acc_x);
                acc_y = pk_fma(fb, gy, acc_y);
                acc_z = pk_fma(fb, gz, acc_z);
                spline_value_deriv(row, 2, b, db, fc, dfc);
                const pk4 cbar = pk_fma_s(-g.y, gx, pk_fma_s(-g.z, gy, pk_mul_s(-g.w, gz)));   // u_r = -u_e
                dv = pk_fma(cbar, dfc, dv);
                part4[0] = pk_hsum(pk_mul(fc, gx)); part4[1] = pk_hsum(pk_mul(fc, gy));
                part4[2] = pk_hsum(pk_mul(fc, gz));
            }
            part4[3] = pk_hsum(dv);
            const float total = group8_sum4(part4, l8, active_lanes);
            if ((l8 & 1) == 0) adj_out[4 * (size_t)e + held] = total;
            c.e = e + 1;
            had_edges = true;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Small systems (single-trajectory MD: N <= MLFFD_SMALL_ROWS): with a few hundred rows the kernels above
// leave most SMs idle while every row group walks its ~20 edges one L2 round trip at a time.  Here the
// four row groups of a WARP share one row (group t takes edges t, t + 4, ...), a CTA is 8 warps = 8 rows,
// so a 300-atom system spreads over ~150 CTAs and a row costs ~5 dependent round trips.  The four partial
// sums meet through shuffles in a fixed order ((g0 + g1) + (g2 + g3)): deterministic; differs from the
// row-per-group kernels by summation order only.
// ---------------------------------------------------------------------------------------------
constexpr int kSplineTeamThreads = 256;

__device__ __forceinline__ pk4 pk_team_sum(pk4 v) {   // sum over the 4 row groups of a warp (lanes l8, l8 + 8, ...)
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        pk4 w;
        w.x = __shfl_xor_sync(full, v.x, o);
        w.y = __shfl_xor_sync(full, v.y, o);
        v = pk_add(v, w);
    }
    return v;
}

template <bool LAYER0>
__global__ void __launch_bounds__(kSplineTeamThreads)
spline_message_forward_team_kernel(const float* __restrict__ table, int H,
                                   const int* __restrict__ rowptr, const float4* __restrict__ erec,
                                   const float* __restrict__ s_in, const float* __restrict__ v_in,
                                   float* __restrict__ s_msg, float* __restrict__ v_msg, int num_atoms,
                                   const DeviceStatus* __restrict__ status) {
    extern __shared__ float4 spline_tab[];
    if (status->overflow) return;
    constexpr int kWarps = kSplineTeamThreads / 32;
    const int slices = H / kSliceChannels;
    const int slice = blockIdx.x % slices, part = blockIdx.x / slices, parts = gridDim.x / slices;
    load_spline_slice(spline_tab, table, slice);
    const int warp = threadIdx.x >> 5, t = (threadIdx.x >> 3) & 3, l8 = threadIdx.x & 7;
    const int ch = slice * kSliceChannels + l8 * 4;
    const pk4* tab = reinterpret_cast<const pk4*>(spline_tab) + l8;
    for (int j = part * kWarps + warp; j < num_atoms; j += parts * kWarps) {
        const int e0 = __ldg(rowptr + j), e1 = __ldg(rowptr + j + 1);
        pk4 acc_s = pk_zero(), acc_x = pk_zero(), acc_y = pk_zero(), acc_z = pk_zero();
        float4 head = make4(0.f);
        if (e0 + t < e1) head = __ldg(erec + 4 * (size_t)(e0 + t));
        for (int e = e0 + t; e < e1; e += 4) {
            const float4 g = head;   // (source, unit vector)
            if (e + 4 < e1) head = __ldg(erec + 4 * (size_t)(e + 4));
            const int i = __float_as_int(g.x);
            const pk4 si = pk_ldg(s_in + (size_t)i * H + ch);
            pk4 vix, viy, viz;
            if (!LAYER0) {
                const float* vi = v_in + (size_t)i * 3 * H + ch;
                vix = pk_ldg(vi); viy = pk_ldg(vi + H); viz = pk_ldg(vi + 2 * H);
            }
            const float4 w0 = __ldg(erec + 4 * (size_t)e + 1), w1 = __ldg(erec + 4 * (size_t)e + 2);
            const float b[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
            const pk4* row = tab + __float_as_int(w1.z) * kSplineRowFloat4;
            acc_s = pk_fma(si, spline_value(row, 0, b), acc_s);
            if (!LAYER0) {
                const pk4 fb = spline_value(row, 1, b);
                acc_x = pk_fma(vix, fb, acc_x);
                acc_y = pk_fma(viy, fb, acc_y);
                acc_z = pk_fma(viz, fb, acc_z);
            }
            const pk4 fc = spline_value(row, 2, b);
            acc_x = pk_fma_s(g.y, fc, acc_x);
            acc_y = pk_fma_s(g.z, fc, acc_y);
            acc_z = pk_fma_s(g.w, fc, acc_z);
        }
        acc_s = pk_team_sum(acc_s); acc_x = pk_team_sum(acc_x); acc_y = pk_team_sum(acc_y); acc_z = pk_team_sum(acc_z);
        if (t == 0) {
            pk_st(s_msg + (size_t)j * H + ch, pk_add(pk_ldg(s_in + (size_t)j * H + ch), acc_s));
            float* vo = v_msg + (size_t)j * 3 * H + ch;
            if (LAYER0) {
                pk_st(vo, acc_x); pk_st(vo + H, acc_y); pk_st(vo + 2 * H, acc_z);
            } else {
                const float* vj = v_in + (size_t)j * 3 * H + ch;
                pk_st(vo, pk_add(pk_ldg(vj), acc_x));
                pk_st(vo + H, pk_add(pk_ldg(vj + H), acc_y));
                pk_st(vo + 2 * H, pk_add(pk_ldg(vj + 2 * H), acc_z));
            }
        }
    }
}

template <bool LAYER0>
__global__ void __launch_bounds__(kSplineTeamThreads)
spline_message_backward_team_kernel(const float* __restrict__ table, int H,
                                    const int* __restrict__ rowptr, const float4* __restrict__ erec,
                                    const float* __restrict__ s_in, const float* __restrict__ v_in,
                                    const float* __restrict__ sbar_m, const float* __restrict__ vbar_m,
                                    float* __restrict__ sbar_in, float* __restrict__ vbar_in,
                                    float4* __restrict__ edge_adj, size_t slab_stride, int num_atoms,
                                    const DeviceStatus* __restrict__ status) {
    extern __shared__ float4 spline_tab[];
    if (status->overflow) return;
    constexpr int kWarps = kSplineTeamThreads / 32;
    const int slices = H / kSliceChannels;
    const int slice = blockIdx.x % slices, part = blockIdx.x / slices, parts = gridDim.x / slices;
    load_spline_slice(spline_tab, table, slice);
    float* adj_out = reinterpret_cast<float*>(edge_adj + (size_t)slice * slab_stride);
    const int warp = threadIdx.x >> 5, t = (threadIdx.x >> 3) & 3, l8 = threadIdx.x & 7;
    const int ch = slice * kSliceChannels + l8 * 4;
    const pk4* tab = reinterpret_cast<const pk4*>(spline_tab) + l8;
    const int held = ((l8 & 4) ? 2 : 0) + ((l8 & 2) ? 1 : 0);
    for (int i = part * kWarps + warp; i < num_atoms; i += parts * kWarps) {
        const int e0 = __ldg(rowptr + i), e1 = __ldg(rowptr + i + 1);
        // the row's adjoints: LAYER0 operands / residual path of the other layers (requested up front: a load
        // at the end of the row would add an exposed L2 round trip to a kernel that is one latency chain)
        const pk4 sb = pk_ldg(sbar_m + (size_t)i * H + ch);
        const float* vb = vbar_m + (size_t)i * 3 * H + ch;
        const pk4 vbx = pk_ldg(vb), vby = pk_ldg(vb + H), vbz = pk_ldg(vb + 2 * H);
        pk4 own_s = pk_zero(), own_x = pk_zero(), own_y = pk_zero(), own_z = pk_zero();   // the row's features
        if (!LAYER0) {
            own_s = pk_ldg(s_in + (size_t)i * H + ch);
            const float* vi = v_in + (size_t)i * 3 * H + ch;
            own_x = pk_ldg(vi); own_y = pk_ldg(vi + H); own_z = pk_ldg(vi + 2 * H);
        }
        pk4 acc_s = pk_zero(), acc_x = pk_zero(), acc_y = pk_zero(), acc_z = pk_zero();
        float4 head = make4(0.f);
        if (e0 + t < e1) head = __ldg(erec + 4 * (size_t)(e0 + t));
        const int trips = (e1 - e0 + 3) >> 2;
        for (int k = 0; k < trips; ++k) {
            const int e = e0 + 4 * k + t;
            const bool active = e < e1;
            const unsigned active_lanes = __ballot_sync(0xffffffffu, active);
            if (active) {
                const float4 g = head;
                if (e + 4 < e1) head = __ldg(erec + 4 * (size_t)(e + 4));
                const int j = __float_as_int(g.x);
                pk4 gs, gx, gy, gz;   // LAYER0: s_j;  other layers: s_bar'_j, v_bar'_j (see the row kernel)
                if (LAYER0) {
                    gs = pk_ldg(s_in + (size_t)j * H + ch);
                } else {
                    gs = pk_ldg(sbar_m + (size_t)j * H + ch);
                    const float* vbj = vbar_m + (size_t)j * 3 * H + ch;
                    gx = pk_ldg(vbj); gy = pk_ldg(vbj + H); gz = pk_ldg(vbj + 2 * H);
                }
                const float4 w0 = __ldg(erec + 4 * (size_t)e + 1), w1 = __ldg(erec + 4 * (size_t)e + 2);
                const float4 d0 = __ldg(erec + 4 * (size_t)e + 3);
                const float b[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
                const float db[6] = {d0.x, d0.y, d0.z, d0.w, w1.w, -((d0.x + d0.y) + (d0.z + d0.w) + w1.w)};
                const pk4* row = tab + __float_as_int(w1.z) * kSplineRowFloat4;
                pk4 dv, fc, dfc;
                float part4[4];
                if (LAYER0) {
                    dv = pk_mul(pk_mul(gs, sb), spline_deriv(row, 0, db));
                    spline_value_deriv(row, 2, b, db, fc, dfc);
                    const pk4 cbar = pk_fma_s(g.y, vbx, pk_fma_s(g.z, vby, pk_mul_s(g.w, vbz)));
                    dv = pk_fma(cbar, dfc, dv);
                    part4[0] = pk_hsum(pk_mul(fc, vbx)); part4[1] = pk_hsum(pk_mul(fc, vby));
                    part4[2] = pk_hsum(pk_mul(fc, vbz));
                } else {
                    pk4 fa, dfa, fb, dfb;
                    spline_value_deriv(row, 0, b, db, fa, dfa);
                    dv = pk_mul(pk_mul(own_s, gs), dfa);
                    acc_s = pk_fma(fa, gs, acc_s);
                    spline_value_deriv(row, 1, b, db, fb, dfb);
                    const pk4 bbar = pk_fma(own_x, gx, pk_fma(own_y, gy, pk_mul(own_z, gz)));
                    dv = pk_fma(bbar, dfb, dv);
                    acc_x = pk_fma(fb, gx, acc_x);
                    acc_y = pk_fma(fb, gy, acc_y);
                    acc_z = pk_fma(fb, gz, acc_z);
                    spline_value_deriv(row, 2, b, db, fc, dfc);
                    const pk4 cbar = pk_fma_s(-g.y, gx, pk_fma_s(-g.z, gy, pk_mul_s(-g.w, gz)));   // u_r = -u_e
                    dv = pk_fma(cbar, dfc, dv);
                    part4[0] = pk_hsum(pk_mul(fc, gx)); part4[1] = pk_hsum(pk_mul(fc, gy));
                    part4[2] = pk_hsum(pk_mul(fc, gz));
                }
                part4[3] = pk_hsum(dv);
                const float total = group8_sum4(part4, l8, active_lanes);
                if ((l8 & 1) == 0) adj_out[4 * (size_t)e + held] = total;
            }
        }
        if (!LAYER0) {
            acc_s = pk_team_sum(acc_s); acc_x = pk_team_sum(acc_x); acc_y = pk_team_sum(acc_y); acc_z = pk_team_sum(acc_z);
            if (t == 0) {   // residual path + the gathered sums
                pk_st(sbar_in + (size_t)i * H + ch, pk_add(sb, acc_s));
                float* vo = vbar_in + (size_t)i * 3 * H + ch;
                pk_st(vo, pk_add(vbx, acc_x)); pk_st(vo + H, pk_add(vby, acc_y)); pk_st(vo + 2 * H, pk_add(vbz, acc_z));
            }
        }
    }
}

// Stage entry point (parity tests): value and d-derivative of the layer's filter spline at `num`
// distances, all 3H channels: filt / dfilt [num][3H] in the (a | b | c) column order of the filter table.
__global__ void __launch_bounds__(256)
spline_filter_eval_kernel(const float* __restrict__ table, int H, float inv_h, const float* __restrict__ dist,
                          long long num, float* __restrict__ filt, float* __restrict__ dfilt) {
    const long long total = num * 3 * H;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / (3 * H);
        const int c = (int)(idx - p * 3 * H), comp = c / H, chan = c % H;
        const int slice = chan / kSliceChannels, cc = chan % kSliceChannels;
        float u, b[6], db[6];
        const int seg = spline_segment(__ldg(dist + p), inv_h, u);
        quintic_basis(u, b, db);
        const float* base = table + (((size_t)slice * kSplineRows + seg) * 3 + comp) * kSliceChannels + cc;
        float val = 0.f, der = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float coef = __ldg(base + (size_t)j * 3 * kSliceChannels);
            val = fmaf(b[j], coef, val);
            der = fmaf(db[j], coef, der);
        }
        filt[idx] = val;
        dfilt[idx] = der * inv_h;
    }
}

}  // namespace mlffd
