// Radius-graph neighbour list: destination-sorted CSR with sources ascending inside a row.
//
// Replaces radius_graph_native (reference src/mlff_distiller/models/student_model.py:63-109),
// whose dense [N,N,3] difference tensor is O(N^2) memory.  Semantics kept: ordered pairs i != j
// of the same structure with d <= cutoff (inclusive), FP32.  Because the edge set is symmetric,
// CSR order (dst row, src ascending) enumerates the same sequence of index pairs as the
// reference's lexicographic (src, dst) order with the two labels swapped; mlffd_export_edges uses
// that to hand back the reference's edge_index without a sort.
//
// Two candidate generators feed the same warp-cooperative pair test:
//   * per-structure sweep  (small structures: every atom of the structure is a candidate)
//   * cell list            (large structures: candidates from the 27 surrounding cells)
// A warp owns one destination atom; lanes test 32 candidates at a time, and a ballot/popc
// compaction keeps sources in ascending order without atomics or a sort.
#pragma once
#include "common.cuh"

namespace mlffd {

// structure id of every atom from the offsets array (binary search; B is small or N is large)
__global__ void atom_structure_kernel(const int* __restrict__ offsets, int num_structures,
                                      int num_atoms, int* __restrict__ atom_struct) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_atoms; i += gridDim.x * blockDim.x) {
        int lo = 0, hi = num_structures;  // invariant: offsets[lo] <= i < offsets[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid;
        }
        atom_struct[i] = lo;
    }
}

// One warp per destination atom j; candidates = all atoms of j's structure.
// FILL == false: deg[j] = #neighbours, deg_low[j] = #neighbours with index < j.
// FILL == true : writes col/edge_dst/geo at rowptr[j]... (sources ascending).
template <bool FILL>
__global__ void __launch_bounds__(256)
neighbor_sweep_kernel(const float* __restrict__ pos, const int* __restrict__ offsets,
                      const int* __restrict__ atom_struct, const float* __restrict__ cells,
                      const uint8_t* __restrict__ pbc, int num_atoms, float cutoff,
                      int* __restrict__ deg, int* __restrict__ deg_low,
                      const int* __restrict__ rowptr, int* __restrict__ col,
                      int* __restrict__ edge_dst, float4* __restrict__ geo,
                      DeviceStatus* __restrict__ status, const int* __restrict__ gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;   // skin list: the candidate build is not due this step
    if (FILL && status->overflow) return;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int warp_stride = gridDim.x * warps_per_block;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int j = warp0; j < num_atoms; j += warp_stride) {
        const int b = atom_struct[j];
        const int lo = __ldg(offsets + b), hi = __ldg(offsets + b + 1);
        const float xj = __ldg(pos + 3 * j), yj = __ldg(pos + 3 * j + 1), zj = __ldg(pos + 3 * j + 2);
        unsigned pmask = 0;
        const float* cell18 = nullptr;
        if (pbc != nullptr) {
            pmask = (pbc[3 * b] ? 1u : 0u) | (pbc[3 * b + 1] ? 2u : 0u) | (pbc[3 * b + 2] ? 4u : 0u);
            cell18 = cells + 18 * b;
        }
        int count = 0, count_low = 0;
        const int base = FILL ? rowptr[j] : 0;
        for (int i0 = lo; i0 < hi; i0 += 32) {
            const int i = i0 + lane;
            bool ok = false;
            float dx = 0.f, dy = 0.f, dz = 0.f, d = 0.f;
            if (i < hi && i != j) {
                dx = __fsub_rn(__ldg(pos + 3 * i), xj);      // x_src - x_dst
                dy = __fsub_rn(__ldg(pos + 3 * i + 1), yj);
                dz = __fsub_rn(__ldg(pos + 3 * i + 2), zj);
                if (pmask) min_image(dx, dy, dz, cell18, pmask);
                d = pair_distance(dx, dy, dz);
                ok = d <= cutoff;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (FILL) {
                if (ok) {
                    const int e = base + count + __popc(m & lt_mask);
                    const float q = __fadd_rn(d, kUnitEps);
                    col[e] = i;
                    edge_dst[e] = j;
                    geo[e] = make_float4(__fdiv_rn(dx, q), __fdiv_rn(dy, q), __fdiv_rn(dz, q), d);
                }
            } else {
                count_low += __popc(__ballot_sync(0xffffffffu, ok && i < j));
            }
            count += __popc(m);
        }
        if (!FILL && lane == 0) {
            deg[j] = count;
            deg_low[j] = count_low;
        }
    }
}

// After the two exclusive scans: publish E and P, flag overflow.
__global__ void neighbor_finalize_kernel(const int* __restrict__ rowptr,
                                         const int* __restrict__ lowptr, int num_atoms,
                                         int edge_capacity, DeviceStatus* __restrict__ status) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int e = rowptr[num_atoms];
        status->num_edges = e;
        status->num_pairs = lowptr[num_atoms];
        status->overflow = (e > edge_capacity) ? 1 : 0;
        if (e > edge_capacity) status->overflow_events += 1;
        status->max_degree = 0;
        status->tc_saturated = 0;
    }
}

// Small systems (latency path): both exclusive scans and the finalize step in ONE single-block
// launch instead of five (2 x CUB init + scan, finalize).  Each thread scans a contiguous chunk.
__global__ void __launch_bounds__(1024)
neighbor_scan_small_kernel(const int* __restrict__ deg, const int* __restrict__ deg_low,
                           int* __restrict__ rowptr, int* __restrict__ lowptr, int num_atoms,
                           int edge_capacity, DeviceStatus* __restrict__ status /* may be NULL: scan only */) {
    __shared__ int warp_a[32], warp_b[32];
    const int n = num_atoms + 1;   // deg[num_atoms] == 0 closes the CSR
    const int chunk = (n + 1023) / 1024;
    const int lo = min(threadIdx.x * chunk, n), hi = min(lo + chunk, n);
    int sa = 0, sb = 0;
    for (int i = lo; i < hi; ++i) { sa += deg[i]; sb += deg_low[i]; }
    // exclusive scan of the 1024 per-thread sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ia = sa, ib = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { warp_a[warp] = ia; warp_b[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        int wa = warp_a[lane], wb = warp_b[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, wa, o), tb = __shfl_up_sync(0xffffffffu, wb, o);
            if (lane >= o) { wa += ta; wb += tb; }
        }
        warp_a[lane] = wa; warp_b[lane] = wb;   // inclusive over warps
    }
    __syncthreads();
    int pa = ia - sa + (warp ? warp_a[warp - 1] : 0);   // exclusive prefix of this thread's chunk
    int pb = ib - sb + (warp ? warp_b[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) {
        rowptr[i] = pa; lowptr[i] = pb;
        pa += deg[i]; pb += deg_low[i];
    }
    if (threadIdx.x == 1023 && status != nullptr) {
        const int e = warp_a[31];
        status->num_edges = e;
        status->num_pairs = warp_b[31];
        status->overflow = (e > edge_capacity) ? 1 : 0;
        if (e > edge_capacity) status->overflow_events += 1;
        status->max_degree = 0;
        status->tc_saturated = 0;
    }
}

// Per edge e = (i -> j): rev[e] = position of (j -> i) (binary search in row i), pair id, and the
// compact per-pair distance list the filter-table kernel consumes.  Lower edges (i < j) come
// first in a row because sources ascend, so pair ids need no extra sort.
__global__ void reverse_pair_kernel(const int* __restrict__ rowptr, const int* __restrict__ lowptr,
                                    const int* __restrict__ col, const int* __restrict__ edge_dst,
                                    const float4* __restrict__ geo, int* __restrict__ rev,
                                    int* __restrict__ pair, float* __restrict__ pair_dist,
                                    DeviceStatus* __restrict__ status) {
    if (status->overflow) return;
    const int num_edges = status->num_edges;
    int local_max = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += gridDim.x * blockDim.x) {
        const int i = col[e], j = edge_dst[e];
        int lo = rowptr[i], hi = rowptr[i + 1] - 1;
        local_max = max(local_max, rowptr[j + 1] - rowptr[j]);
        while (lo < hi) {  // sources ascending -> lower_bound of j in row i
            const int mid = (lo + hi) >> 1;
            if (col[mid] < j) lo = mid + 1; else hi = mid;
        }
        rev[e] = lo;
        int p;
        if (i < j) {
            p = lowptr[j] + (e - rowptr[j]);
            pair_dist[p] = geo[e].w;
        } else {
            p = lowptr[i] + (lo - rowptr[i]);
        }
        pair[e] = p;
    }
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 16));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 8));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 4));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 2));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 1));
    if ((threadIdx.x & 31) == 0 && local_max > 0) atomicMax(&status->max_degree, local_max);
}

// Small systems (single-trajectory MD, N <= kSmallNeighborAtoms): the whole list in ONE single-block
// launch -- structure ids, count pass, both scans, status, fill pass, reverse / pair indices --
// instead of five kernels and two memsets (each graph node costs 2 - 3 us even when empty).  Same
// pair test, same order-preserving compaction, therefore the same bits as the multi-kernel path.
constexpr int kSmallNeighborAtoms = 64;

__global__ void __launch_bounds__(1024)
neighbor_small_kernel(const float* __restrict__ pos, const int* __restrict__ offsets, int num_structures,
                      const float* __restrict__ cells, const uint8_t* __restrict__ pbc, int num_atoms,
                      float cutoff, int edge_capacity, int* __restrict__ atom_struct,
                      int* __restrict__ rowptr, int* __restrict__ lowptr, int* __restrict__ col,
                      int* __restrict__ edge_dst, float4* __restrict__ geo, int* __restrict__ rev,
                      int* __restrict__ pair, float* __restrict__ pair_dist,
                      DeviceStatus* __restrict__ status) {
    __shared__ int struct_s[kSmallNeighborAtoms];
    __shared__ int row_s[kSmallNeighborAtoms + 1], low_s[kSmallNeighborAtoms + 1];   // degrees, then exclusive prefixes
    __shared__ int warp_a[32], warp_b[32], warp_m[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned full = 0xffffffffu;
    __shared__ float x_s[kSmallNeighborAtoms], y_s[kSmallNeighborAtoms], z_s[kSmallNeighborAtoms];
    if (tid < num_atoms) {   // positions once into shared memory: the pair loops below are latency chains
        x_s[tid] = __ldg(pos + 3 * tid); y_s[tid] = __ldg(pos + 3 * tid + 1); z_s[tid] = __ldg(pos + 3 * tid + 2);
    }
    if (tid < num_atoms) {   // structure of every atom (binary search over offsets)
        int lo = 0, hi = num_structures;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(offsets + mid) <= tid) lo = mid; else hi = mid;
        }
        struct_s[tid] = lo;
        atom_struct[tid] = lo;
    }
    __syncthreads();
    // pass == 0 counts, pass == 1 fills; a warp owns destination rows warp, warp + 32, ...
    for (int pass = 0; pass < 2; ++pass) {
        for (int j = warp; j < num_atoms; j += 32) {
            const int b = struct_s[j];
            const int lo = __ldg(offsets + b), hi = __ldg(offsets + b + 1);
            const float xj = x_s[j], yj = y_s[j], zj = z_s[j];
            unsigned pmask = 0;
            const float* cell18 = nullptr;
            if (pbc != nullptr) {
                pmask = (pbc[3 * b] ? 1u : 0u) | (pbc[3 * b + 1] ? 2u : 0u) | (pbc[3 * b + 2] ? 4u : 0u);
                cell18 = cells + 18 * b;
            }
            int count = 0, count_low = 0;
            const int base = pass ? row_s[j] : 0;
            for (int i0 = lo; i0 < hi; i0 += 32) {
                const int i = i0 + lane;
                bool ok = false;
                float dx = 0.f, dy = 0.f, dz = 0.f, d = 0.f;
                if (i < hi && i != j) {
                    dx = __fsub_rn(x_s[i], xj);      // x_src - x_dst
                    dy = __fsub_rn(y_s[i], yj);
                    dz = __fsub_rn(z_s[i], zj);
                    if (pmask) min_image(dx, dy, dz, cell18, pmask);
                    d = pair_distance(dx, dy, dz);
                    ok = d <= cutoff;
                }
                const unsigned m = __ballot_sync(full, ok);
                if (pass) {
                    if (ok) {
                        const int e = base + count + __popc(m & lt_mask);
                        const float q = __fadd_rn(d, kUnitEps);
                        col[e] = i;
                        edge_dst[e] = j;
                        geo[e] = make_float4(__fdiv_rn(dx, q), __fdiv_rn(dy, q), __fdiv_rn(dz, q), d);
                    }
                } else {
                    count_low += __popc(__ballot_sync(full, ok && i < j));
                }
                count += __popc(m);
            }
            if (!pass && lane == 0) { row_s[j] = count; low_s[j] = count_low; }
        }
        if (pass) break;
        __syncthreads();
        // exclusive scans of the degrees over num_atoms + 1 entries (one per thread), max degree
        const int n = num_atoms + 1;
        const int da = (tid < num_atoms) ? row_s[tid] : 0, db = (tid < num_atoms) ? low_s[tid] : 0;
        int ia = da, ib = db, mx = da;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(full, ia, o), tb = __shfl_up_sync(full, ib, o);
            if (lane >= o) { ia += ta; ib += tb; }
            mx = max(mx, __shfl_xor_sync(full, mx, o));
        }
        if (lane == 31) { warp_a[warp] = ia; warp_b[warp] = ib; warp_m[warp] = mx; }
        __syncthreads();
        if (warp == 0) {
            int wa = warp_a[lane], wb = warp_b[lane], wm = warp_m[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_up_sync(full, wa, o), tb = __shfl_up_sync(full, wb, o);
                if (lane >= o) { wa += ta; wb += tb; }
                wm = max(wm, __shfl_xor_sync(full, wm, o));
            }
            warp_a[lane] = wa; warp_b[lane] = wb; warp_m[lane] = wm;   // inclusive over warps
        }
        __syncthreads();
        const int pa = ia - da + (warp ? warp_a[warp - 1] : 0), pb = ib - db + (warp ? warp_b[warp - 1] : 0);
        const int num_edges = warp_a[31];
        __syncthreads();   // everyone has read the degrees before they become prefixes
        if (tid < n) {
            row_s[tid] = pa; low_s[tid] = pb;
            rowptr[tid] = pa; lowptr[tid] = pb;
        }
        if (tid == 0) {
            status->num_edges = num_edges;
            status->num_pairs = warp_b[31];
            status->overflow = (num_edges > edge_capacity) ? 1 : 0;
            if (num_edges > edge_capacity) status->overflow_events += 1;
            status->max_degree = warp_m[0];
            status->tc_saturated = 0;
        }
        __syncthreads();
        if (num_edges > edge_capacity) return;   // uniform: outputs invalid, the host grows and re-runs
    }
    __syncthreads();   // the block's own global writes (col, geo) are visible to all its threads
    const int num_edges = row_s[num_atoms];
    for (int e = tid; e < num_edges; e += 1024) {
        const int i = col[e], j = edge_dst[e];
        int lo = row_s[i], hi = row_s[i + 1] - 1;
        while (lo < hi) {  // sources ascending -> lower_bound of j in row i
            const int mid = (lo + hi) >> 1;
            if (col[mid] < j) lo = mid + 1; else hi = mid;
        }
        rev[e] = lo;
        int p;
        if (i < j) {
            p = low_s[j] + (e - row_s[j]);
            pair_dist[p] = geo[e].w;
        } else {
            p = low_s[i] + (lo - row_s[i]);
        }
        pair[e] = p;
    }
}

// edge_index [2, cap] int64 in the reference's order: row 0 = src, row 1 = dst, lexicographic.
// By symmetry the k-th CSR entry (row j, col i) is the k-th lexicographic pair (src=j, dst=i).
__global__ void export_edges_kernel(const int* __restrict__ col, const int* __restrict__ edge_dst,
                                    int num_edges, long long capacity,
                                    long long* __restrict__ edge_index) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += gridDim.x * blockDim.x) {
        edge_index[e] = edge_dst[e];
        edge_index[capacity + e] = col[e];
    }
}

}  // namespace mlffd
