// Cell-list candidate generator for large structures (periodic or open), feeding the same FP32
// pair test and producing the same destination-sorted CSR as neighbor_sweep_kernel.
//
// The reference has no such path: radius_graph_native builds a dense [N,N,3] tensor
// (src/mlff_distiller/models/student_model.py:89-99, 1.2 GB at 10 k atoms) and ignores cell/pbc
// (:694-703).  Semantics are unchanged -- an edge exists iff pair_distance(min_image(x_i - x_j))
// <= cutoff, evaluated with the identical device functions -- the grid only prunes candidates.
//
// Per structure a grid is laid over "grid coordinates" s: fractional coordinates s = x * cell^-1
// when the structure has any periodic axis, Cartesian coordinates otherwise.  Along axis k the
// bin thickness (perpendicular distance between bin planes) is >= cutoff, so two atoms within the
// cutoff (in any periodic image) sit in the same or adjacent bins and the 27-cell stencil is a
// superset of the true neighbours.  Periodic axes wrap; an axis with fewer than 3 bins is
// collapsed to one bin so the stencil never visits a cell twice.
#pragma once
#include "common.cuh"

namespace mlffd {

struct GridInfo {
    float origin[3];    // grid coordinate of the lower corner (0 for periodic axes)
    float inv_width[3]; // bins per unit of grid coordinate
    float frame[9];     // x -> s matrix (row-major, s_k = sum_r x_r * frame[r*3+k]); identity if open
    int nb[3];          // bins per axis
    int cell_offset;    // first global cell id of this structure
    unsigned pbc_mask;  // periodic axes
};

__device__ __forceinline__ void grid_coords(const GridInfo& g, float x, float y, float z, float (&s)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = x * g.frame[k] + y * g.frame[3 + k] + z * g.frame[6 + k];
}

__device__ __forceinline__ int grid_bin(const GridInfo& g, int k, float s) {
    float t = s;
    if (g.pbc_mask & (1u << k)) t = s - floorf(s);  // wrap into [0,1)
    int b = (int)floorf((t - g.origin[k]) * g.inv_width[k]);
    return min(max(b, 0), g.nb[k] - 1);
}

// One block per structure: bounding box in grid coordinates, bin counts.
__global__ void __launch_bounds__(256)
grid_setup_kernel(const float* __restrict__ pos, const int* __restrict__ offsets,
                  const float* __restrict__ cells, const uint8_t* __restrict__ pbc, float cutoff,
                  int max_cells_per_struct, GridInfo* __restrict__ grids, int* __restrict__ ncells,
                  const int* __restrict__ gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;
    const int b = blockIdx.x;
    const int lo = offsets[b], hi = offsets[b + 1];
    __shared__ GridInfo g;
    __shared__ float red_min[3][8], red_max[3][8];
    unsigned pmask = 0;
    if (pbc != nullptr)
        pmask = (pbc[3 * b] ? 1u : 0u) | (pbc[3 * b + 1] ? 2u : 0u) | (pbc[3 * b + 2] ? 4u : 0u);
    if (threadIdx.x == 0) {
        g.pbc_mask = pmask;
        for (int q = 0; q < 9; ++q) g.frame[q] = (q % 4 == 0) ? 1.f : 0.f;
        if (pmask) for (int q = 0; q < 9; ++q) g.frame[q] = cells[18 * b + 9 + q];
    }
    __syncthreads();
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float s[3];
        grid_coords(g, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], s);
#pragma unroll
        for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], s[k]); mx[k] = fmaxf(mx[k], s[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if ((threadIdx.x & 31) == 0) { red_min[k][threadIdx.x >> 5] = mn[k]; red_max[k][threadIdx.x >> 5] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float height[3] = {1.f, 1.f, 1.f};  // perpendicular thickness of one unit of s_k
        if (pmask) {
            // |row k of cell^-T|^-1 = distance between the planes s_k = const one unit apart
            for (int k = 0; k < 3; ++k) {
                const float a = g.frame[k], bb = g.frame[3 + k], c = g.frame[6 + k];
                height[k] = rsqrtf(a * a + bb * bb + c * c);
            }
        }
        long long total = 1;
        for (int k = 0; k < 3; ++k) {
            float lo_k = 3.4e38f, hi_k = -3.4e38f;
            for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) { lo_k = fminf(lo_k, red_min[k][wv]); hi_k = fmaxf(hi_k, red_max[k][wv]); }
            if (hi <= lo) { lo_k = 0.f; hi_k = 0.f; }
            float extent;
            if (pmask & (1u << k)) { g.origin[k] = 0.f; extent = 1.f; }
            else { g.origin[k] = lo_k; extent = fmaxf(hi_k - lo_k, 0.f); }
            // 1 % slack on the bin thickness absorbs the FP32 rounding of the bin assignment
            int nb = (int)floorf(extent * height[k] / (cutoff * 1.01f));
            if (pmask & (1u << k)) { if (nb < 3) nb = 1; }
            else nb = max(nb, 1);
            g.nb[k] = nb;
            total *= nb;
        }
        // cap the number of cells (sparse open systems): coarsen uniformly
        while (total > (long long)max_cells_per_struct) {
            total = 1;
            for (int k = 0; k < 3; ++k) {
                int nb = max(g.nb[k] / 2, 1);
                if ((pmask & (1u << k)) && nb < 3) nb = 1;
                g.nb[k] = nb;
                total *= nb;
            }
        }
        for (int k = 0; k < 3; ++k) {
            float extent = (pmask & (1u << k)) ? 1.f : 0.f;
            if (!(pmask & (1u << k))) {
                float lo_k = 3.4e38f, hi_k = -3.4e38f;
                for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) { lo_k = fminf(lo_k, red_min[k][wv]); hi_k = fmaxf(hi_k, red_max[k][wv]); }
                extent = (hi > lo) ? fmaxf(hi_k - lo_k, 0.f) : 0.f;
            }
            g.inv_width[k] = (extent > 0.f) ? (float)g.nb[k] / extent : 0.f;
        }
        g.cell_offset = 0;
        grids[b] = g;
        ncells[b] = (int)total;
    }
}

// serial prefix over structures (this path is taken for few, large structures)
__global__ void grid_offsets_kernel(GridInfo* __restrict__ grids, const int* __restrict__ ncells,
                                    int num_structures, int* __restrict__ total_cells,
                                    const int* __restrict__ gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < num_structures; ++b) { grids[b].cell_offset = acc; acc += ncells[b]; }
        *total_cells = acc;
    }
}

__device__ __forceinline__ int atom_cell_id(const GridInfo& g, float x, float y, float z, int (&bin)[3]) {
    float s[3];
    grid_coords(g, x, y, z, s);
#pragma unroll
    for (int k = 0; k < 3; ++k) bin[k] = grid_bin(g, k, s[k]);
    return g.cell_offset + (bin[0] * g.nb[1] + bin[1]) * g.nb[2] + bin[2];
}

__global__ void __launch_bounds__(256)
cell_count_kernel(const float* __restrict__ pos, const int* __restrict__ atom_struct,
                  const GridInfo* __restrict__ grids, int num_atoms, int* __restrict__ atom_cell,
                  int* __restrict__ cell_count, const int* __restrict__ gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_atoms; i += gridDim.x * blockDim.x) {
        const GridInfo g = grids[atom_struct[i]];
        int bin[3];
        const int c = atom_cell_id(g, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], bin);
        atom_cell[i] = c;
        atomicAdd(cell_count + c, 1);
    }
}

__global__ void __launch_bounds__(256)
cell_fill_kernel(const int* __restrict__ atom_cell, const int* __restrict__ cell_start,
                 int* __restrict__ cell_cursor, int num_atoms, int* __restrict__ cell_atoms,
                 const int* __restrict__ gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_atoms; i += gridDim.x * blockDim.x) {
        const int c = atom_cell[i];
        cell_atoms[cell_start[c] + atomicAdd(cell_cursor + c, 1)] = i;  // order inside a cell is irrelevant:
    }                                                                  // rows are sorted by source below
}

// One warp per destination atom j; candidates from the 27 surrounding cells.
// FILL == false: deg / deg_low.  FILL == true: unsorted (src, geo) into tmp_col / tmp_geo at
// rowptr[j].., then a rank sort by source index into col / geo (sources ascending).
template <bool FILL>
__global__ void __launch_bounds__(256)
neighbor_cells_kernel(const float* __restrict__ pos, const int* __restrict__ atom_struct,
                      const float* __restrict__ cells, const GridInfo* __restrict__ grids,
                      const int* __restrict__ cell_start, const int* __restrict__ cell_atoms,
                      int num_atoms, float cutoff, int* __restrict__ deg, int* __restrict__ deg_low,
                      const int* __restrict__ rowptr, int* __restrict__ tmp_col,
                      float4* __restrict__ tmp_geo, int* __restrict__ col,
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
        const GridInfo g = grids[b];
        const float xj = __ldg(pos + 3 * j), yj = __ldg(pos + 3 * j + 1), zj = __ldg(pos + 3 * j + 2);
        const float* cell18 = g.pbc_mask ? cells + 18 * b : nullptr;
        int bin[3];
        atom_cell_id(g, xj, yj, zj, bin);
        int count = 0, count_low = 0;
        const int base = FILL ? rowptr[j] : 0;
        for (int d0 = -1; d0 <= 1; ++d0) {
            int c0 = bin[0] + d0;
            if (g.nb[0] == 1) { if (d0 != 0) continue; c0 = 0; }
            else if (g.pbc_mask & 1u) c0 = (c0 + g.nb[0]) % g.nb[0];
            else if (c0 < 0 || c0 >= g.nb[0]) continue;
            for (int d1 = -1; d1 <= 1; ++d1) {
                int c1 = bin[1] + d1;
                if (g.nb[1] == 1) { if (d1 != 0) continue; c1 = 0; }
                else if (g.pbc_mask & 2u) c1 = (c1 + g.nb[1]) % g.nb[1];
                else if (c1 < 0 || c1 >= g.nb[1]) continue;
                for (int d2 = -1; d2 <= 1; ++d2) {
                    int c2 = bin[2] + d2;
                    if (g.nb[2] == 1) { if (d2 != 0) continue; c2 = 0; }
                    else if (g.pbc_mask & 4u) c2 = (c2 + g.nb[2]) % g.nb[2];
                    else if (c2 < 0 || c2 >= g.nb[2]) continue;
                    const int cell = g.cell_offset + (c0 * g.nb[1] + c1) * g.nb[2] + c2;
                    const int a0 = cell_start[cell], a1 = cell_start[cell + 1];
                    for (int t0 = a0; t0 < a1; t0 += 32) {
                        const int t = t0 + lane;
                        bool ok = false;
                        int i = -1;
                        float dx = 0.f, dy = 0.f, dz = 0.f, d = 0.f;
                        if (t < a1) {
                            i = cell_atoms[t];
                            if (i != j) {
                                dx = __fsub_rn(__ldg(pos + 3 * i), xj);
                                dy = __fsub_rn(__ldg(pos + 3 * i + 1), yj);
                                dz = __fsub_rn(__ldg(pos + 3 * i + 2), zj);
                                if (g.pbc_mask) min_image(dx, dy, dz, cell18, g.pbc_mask);
                                d = pair_distance(dx, dy, dz);
                                ok = d <= cutoff;
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, ok);
                        if (FILL) {
                            if (ok) {
                                const int e = base + count + __popc(m & lt_mask);
                                const float qn = __fadd_rn(d, kUnitEps);
                                tmp_col[e] = i;
                                tmp_geo[e] = make_float4(__fdiv_rn(dx, qn), __fdiv_rn(dy, qn), __fdiv_rn(dz, qn), d);
                            }
                        } else {
                            count_low += __popc(__ballot_sync(0xffffffffu, ok && i < j));
                        }
                        count += __popc(m);
                    }
                }
            }
        }
        if (!FILL) {
            if (lane == 0) { deg[j] = count; deg_low[j] = count_low; }
        } else {
            __syncwarp();
            // rank sort of the row by source index (sources are unique within a row)
            for (int e = lane; e < count; e += 32) {
                const int ci = tmp_col[base + e];
                int rank = 0;
                for (int f = 0; f < count; ++f) rank += (tmp_col[base + f] < ci) ? 1 : 0;
                col[base + rank] = ci;
                edge_dst[base + rank] = j;
                geo[base + rank] = tmp_geo[base + e];
            }
        }
    }
}

}  // namespace mlffd
