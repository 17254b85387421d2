// Verlet-skin neighbour list, an OPTION (mlffd_set_skin, off by default).
//
// The reference rebuilds its O(N^2) list on every forward (src/mlff_distiller/models/student_model.py:
// 694-703).  With a skin s > 0 a CANDIDATE list of all pairs within cutoff + s is kept together with the
// positions it was built from; it stays complete as long as no atom has moved further than s / 2 since
// then.  Every step (i) measures the largest displacement and decides ON THE DEVICE whether the candidate
// list has to be rebuilt (no host synchronisation, so a captured CUDA graph replays unchanged), (ii) runs
// the count / scan / fill passes of the exact list over the candidates instead of over the whole structure
// or the 27 surrounding cells.  The pair test is the same code on the same raw difference vector
// (common.cuh:min_image, pair_distance), the candidates ascend in source index like the cell list's rows,
// so the exact list -- edges, order, geometry -- is bit-identical to the one the full build produces.
#pragma once
#include "common.cuh"

namespace mlffd {

struct SkinState {
    int rebuild;        // this step rebuilds the candidate list (gate of the build kernels)
    int valid;          // a candidate list exists for the current system
    int cand_overflow;  // the candidate list did not fit: outputs of the step are invalid
    unsigned acc;       // max squared displacement of this step against the reference positions (float bits);
                        // written by skin_check_kernel, read and cleared by skin_decide_kernel (stream order)
    int rebuilds;       // candidate builds so far
    int cand_edges;     // size of the current candidate list
};

// largest squared displacement against the reference positions -> acc
__global__ void __launch_bounds__(256)
skin_check_kernel(const float* __restrict__ pos, const float* __restrict__ pos_ref, int num_atoms,
                  SkinState* __restrict__ skin) {
    float worst = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_atoms; i += gridDim.x * blockDim.x) {
        const float dx = pos[3 * i] - pos_ref[3 * i], dy = pos[3 * i + 1] - pos_ref[3 * i + 1],
                    dz = pos[3 * i + 2] - pos_ref[3 * i + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        worst = fmaxf(worst, (d2 == d2) ? d2 : 3.0e38f);   // NaN positions force a rebuild
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0 && worst > 0.f) atomicMax(&skin->acc, __float_as_uint(worst));
}

// one thread: rebuild iff there is no valid list or an atom moved further than skin / 2
__global__ void skin_decide_kernel(SkinState* __restrict__ skin, float half_skin_sq, DeviceStatus* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    status->overflow = 0;             // this step's flags start clean (the exact list's finalize sets them)
    const bool rebuild = !skin->valid || __uint_as_float(skin->acc) > half_skin_sq;
    skin->rebuild = rebuild ? 1 : 0;
    skin->acc = 0u;                   // the next step's check kernel starts from zero
    if (rebuild) { skin->rebuilds += 1; skin->cand_overflow = 0; }
}

// after the candidate scan (only when rebuilding): capacity check, bookkeeping
__global__ void skin_cand_finalize_kernel(const int* __restrict__ cand_rowptr, int num_atoms, int capacity,
                                          SkinState* __restrict__ skin, DeviceStatus* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || !skin->rebuild) return;
    const int e = cand_rowptr[num_atoms];
    skin->cand_edges = e;
    skin->cand_overflow = (e > capacity) ? 1 : 0;
    skin->valid = (e > capacity) ? 0 : 1;
    if (e > capacity) status->overflow = 1;   // stops the candidate fill pass; merged again at the end of the build
}

__global__ void __launch_bounds__(256)
skin_copy_ref_kernel(const float* __restrict__ pos, float* __restrict__ pos_ref, int n3,
                     const SkinState* __restrict__ skin) {
    if (!skin->rebuild) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += gridDim.x * blockDim.x) pos_ref[i] = pos[i];
}

// Exact list from the candidates: neighbor_sweep_kernel (neighbor.cuh) with the candidate row of atom j as
// the source range.  FILL == false: deg / deg_low.  FILL == true: col / edge_dst / geo at rowptr[j].
template <bool FILL>
__global__ void __launch_bounds__(256)
neighbor_cand_kernel(const float* __restrict__ pos, const int* __restrict__ atom_struct,
                     const float* __restrict__ cells, const uint8_t* __restrict__ pbc, int num_atoms, float cutoff,
                     const int* __restrict__ cand_rowptr, const int* __restrict__ cand_col,
                     int* __restrict__ deg, int* __restrict__ deg_low, const int* __restrict__ rowptr,
                     int* __restrict__ col, int* __restrict__ edge_dst, float4* __restrict__ geo,
                     const SkinState* __restrict__ skin, DeviceStatus* __restrict__ status) {
    if (FILL && status->overflow) return;
    const bool broken = skin->cand_overflow != 0;   // no usable candidates: empty rows, the step is flagged invalid
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int warp_stride = gridDim.x * warps_per_block;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int j = warp0; j < num_atoms; j += warp_stride) {
        if (broken) {
            if (!FILL && lane == 0) { deg[j] = 0; deg_low[j] = 0; }
            continue;
        }
        const int b = atom_struct[j];
        const float xj = __ldg(pos + 3 * j), yj = __ldg(pos + 3 * j + 1), zj = __ldg(pos + 3 * j + 2);
        unsigned pmask = 0;
        const float* cell18 = nullptr;
        if (pbc != nullptr) {
            pmask = (pbc[3 * b] ? 1u : 0u) | (pbc[3 * b + 1] ? 2u : 0u) | (pbc[3 * b + 2] ? 4u : 0u);
            cell18 = cells + 18 * b;
        }
        int count = 0, count_low = 0;
        const int base = FILL ? rowptr[j] : 0;
        const int c_lo = __ldg(cand_rowptr + j), c_hi = __ldg(cand_rowptr + j + 1);
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            const int c = c0 + lane;
            bool ok = false;
            int i = -1;
            float dx = 0.f, dy = 0.f, dz = 0.f, d = 0.f;
            if (c < c_hi) {
                i = __ldg(cand_col + c);
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

// a candidate-list overflow invalidates the step like an edge overflow does
__global__ void skin_merge_status_kernel(const SkinState* __restrict__ skin, DeviceStatus* __restrict__ status) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && skin->cand_overflow) {
        status->overflow = 1;
        status->overflow_events += 1;
        status->num_edges = skin->cand_edges;   // what the caller has to make room for (the exact list is a subset)
    }
}

}  // namespace mlffd
