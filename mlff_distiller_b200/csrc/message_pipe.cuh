// Message block with the filter-table rows streamed through an asynchronous shared-memory ring
// (cp.async), H = 128.
//
// Same math, same accumulation order and therefore bit-identical results to message.cuh; what
// changes is how the bytes arrive.  The generic kernels are latency bound: a warp walks its CSR
// row edge by edge and every edge costs an index load (col / pair / geo) followed by a dependent
// 1.5 - 3 KB filter-row load from HBM, all held in registers, so the bytes in flight per SM are
// capped by registers x occupancy (measured: time ~ 1 / resident warps, 58 - 78 % of HBM peak).
// Here
//   * the indices of a whole row segment (<= 32 edges) are fetched by ONE coalesced load per
//     array (lane k holds edge k) one segment ahead and broadcast with shuffles;
//   * the filter rows of the next D-1 edges are always in flight as cp.async.cg copies into a
//     per-lane private ring (no registers held, no barriers: a lane only ever reads the 16-byte
//     chunks it copied itself), across row boundaries, so the pipeline never drains;
//   * rowptr is fetched two rows ahead.
// Rows stay assigned grid-stride (warp w owns rows w, w + W, ...), so the rows in flight at any
// time are consecutive atoms: both directions of a pair are read within one row time and the
// second read hits the L2.
#pragma once
#include "message.cuh"

namespace mlffd {

__device__ __forceinline__ uint32_t smem_addr32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One row segment: up to 32 consecutive edges of a CSR row, lane k holding edge k's indices.
struct RowSegment {
    int row;      // atom, -1 = past the end
    int e0;       // first edge of the segment
    int n;        // edges in the segment (<= 32)
    int rem;      // edges of the row after this segment (0 = last segment: flush)
    int col;      // per lane: neighbour atom of edge e0 + lane
    int pair;     // per lane: filter-table row
    int aux;      // per lane: reverse edge (reverse pass only)
    float4 geo;   // per lane: (u_x, u_y, u_z, d)
};

// Walks the rows of one warp (grid-stride) as a stream of segments, with rowptr two rows ahead.
template <bool WITH_REV>
struct SegmentStream {
    const int* rowptr; const int* col; const int* pair; const int* rev; const float4* geo;
    int num_atoms, stride, lane;
    int la_row, la_a, la_b;   // look-ahead row and its rowptr pair (loads may still be in flight)

    __device__ __forceinline__ void fetch_lookahead(int row) {
        la_row = (row >= 0 && row < num_atoms) ? row : -1;
        la_a = 0; la_b = 0;
        if (la_row >= 0) { la_a = __ldg(rowptr + row); la_b = __ldg(rowptr + row + 1); }
    }
    __device__ __forceinline__ void load_lanes(RowSegment& s) const {
        s.col = 0; s.pair = 0; s.aux = 0; s.geo = make4(0.f);
        if (lane < s.n) {
            s.col = __ldg(col + s.e0 + lane);
            s.pair = __ldg(pair + s.e0 + lane);
            s.geo = __ldg(geo + s.e0 + lane);
            if (WITH_REV) s.aux = __ldg(rev + s.e0 + lane);
        }
    }
    __device__ __forceinline__ RowSegment from_lookahead() {
        RowSegment s;
        s.row = la_row; s.e0 = la_a;
        const int deg = la_b - la_a;
        s.n = min(deg, 32); s.rem = deg - s.n;
        load_lanes(s);
        fetch_lookahead(la_row >= 0 ? la_row + stride : -1);
        return s;
    }
    __device__ __forceinline__ RowSegment next_of(const RowSegment& c) {
        if (c.row >= 0 && c.rem > 0) {
            RowSegment s;
            s.row = c.row; s.e0 = c.e0 + 32; s.n = min(c.rem, 32); s.rem = c.rem - s.n;
            load_lanes(s);
            return s;
        }
        return from_lookahead();
    }
};

__device__ __forceinline__ float4 shfl4(const float4& v, int src) {
    return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                       __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}


// Ring-slot parts (512 bytes each = one float4 per lane): forward (a, b, c); reverse (a, b, c, a',
// b', c').  Measured on C2: depth 2 is enough (2, 3, 4 within 1 %; deeper rings shrink the L1 and
// get slower); also streaming the gathered neighbour rows through the ring (cp.async.ca) was 10 %
// slower than plain loads, and blocks of 16 - 32 warps were slower than 8.
constexpr int kPipeWarps = 8;
constexpr int kFwdParts = 3, kBwdParts = 6;
template <int D>
constexpr size_t message_forward_pipe_smem() { return (size_t)kPipeWarps * D * kFwdParts * 512; }
template <int D>
constexpr size_t message_backward_pipe_smem() { return (size_t)kPipeWarps * D * kBwdParts * 512; }

// Forward; contract of message_forward_kernel<128, LAYER0>.
// (register cap measured: 64 registers / 4 resident blocks spills 40 bytes and is 19 % slower;
// letting the compiler go above 80 registers / 3 blocks is 5 % slower)
template <bool LAYER0, int D>
__global__ void __launch_bounds__(32 * kPipeWarps)
message_forward_pipe_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                            const int* __restrict__ pair, const float4* __restrict__ geo,
                            const float* __restrict__ filt, const float* __restrict__ s_in,
                            const float* __restrict__ v_in, float* __restrict__ s_msg,
                            float* __restrict__ v_msg, int num_atoms,
                            const DeviceStatus* __restrict__ status) {
    constexpr int H = 128;
    constexpr unsigned full = 0xffffffffu;
    if (status->overflow) return;
    extern __shared__ float4 pipe_ring[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int PARTS = kFwdParts;
    const float4* ring = pipe_ring + (size_t)wib * (D * PARTS * 32) + lane;   // slot s, part p: ring[(s*PARTS+p)*32]
    const uint32_t ring_addr = smem_addr32(ring);
    const int c4 = lane * 4;
    const int warp = blockIdx.x * kPipeWarps + wib, num_warps = gridDim.x * kPipeWarps;

    SegmentStream<false> stream{rowptr, col, pair, nullptr, geo, num_atoms, num_warps, lane, -1, 0, 0};
    stream.fetch_lookahead(warp);
    RowSegment cur = stream.from_lookahead();
    RowSegment nxt = stream.next_of(cur);
    int issued = 0;        // edges of cur ++ nxt already requested (index relative to cur's first edge)
    int drain_until = -1;  // edges up to this index were requested late: wait for everything
    int put = 0, get = 0;  // ring slots
    auto issue_next = [&]() {
        const int idx = issued;
        const int pr = (idx < cur.n) ? __shfl_sync(full, cur.pair, idx) : __shfl_sync(full, nxt.pair, idx - cur.n);
        const float* src = filt + (size_t)pr * (3 * H) + c4;
        const uint32_t dst = ring_addr + (uint32_t)put * (PARTS * 512);
        cp_async16(dst, src);
        if (!LAYER0) cp_async16(dst + 512, src + H);
        cp_async16(dst + 1024, src + 2 * H);
        put = (put + 1 == D) ? 0 : put + 1;
        ++issued;
    };
    auto more_known = [&]() { return issued < cur.n || issued - cur.n < nxt.n; };
#pragma unroll 1
    for (int g = 0; g < D - 1; ++g) {
        if (more_known()) issue_next();
        cp_async_commit();
    }
    float4 acc_s = make4(0.f), acc_x = make4(0.f), acc_y = make4(0.f), acc_z = make4(0.f);
    while (cur.row >= 0) {
#pragma unroll 1
        for (int k = 0; k < cur.n; ++k) {
            while (issued < k + D && more_known()) {
                if (issued < k + D - 1) drain_until = issued;
                issue_next();
            }
            cp_async_commit();
            const float4 g = shfl4(cur.geo, k);
            const int i = __shfl_sync(full, cur.col, k);
            const float4 si = ldg4(s_in + (size_t)i * H + c4);   // neighbour rows: plain loads (L1 / L2 hits)
            float4 vx, vy, vz;
            if (!LAYER0) {
                const float* vi = v_in + (size_t)i * 3 * H + c4;
                vx = ldg4(vi); vy = ldg4(vi + H); vz = ldg4(vi + 2 * H);
            }
            if (k <= drain_until) cp_async_wait<0>(); else cp_async_wait<D - 1>();
            const float4* slot = ring + get * (PARTS * 32);
            get = (get + 1 == D) ? 0 : get + 1;
            const float4 fa = slot[0];
            const float4 fc = slot[64];
            acc_s = fma4(si, fa, acc_s);
            if (!LAYER0) {
                const float4 fb = slot[32];
                acc_x = fma4(vx, fb, acc_x);
                acc_y = fma4(vy, fb, acc_y);
                acc_z = fma4(vz, fb, acc_z);
            }
            acc_x = fma4s(g.x, fc, acc_x);
            acc_y = fma4s(g.y, fc, acc_y);
            acc_z = fma4s(g.z, fc, acc_z);
        }
        if (cur.rem == 0) {   // last segment of the row: residual + store
            const int j = cur.row;
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
            acc_s = make4(0.f); acc_x = make4(0.f); acc_y = make4(0.f); acc_z = make4(0.f);
        }
        issued -= cur.n;
        drain_until -= cur.n;
        cur = nxt;
        nxt = stream.next_of(cur);
    }
    cp_async_wait<0>();
}

// Reverse, every undirected pair handled once; contract and math of
// message_backward_pairs_kernel<128, LAYER0, false> (message.cuh).  Ring slot = the six 512-byte
// parts (a, b, c, a', b', c') of the pair's table rows; edges with j < i only request (a, b), and
// for LAYER0 nothing at all.
template <bool LAYER0, int D>
__global__ void __launch_bounds__(32 * kPipeWarps, 2)
message_backward_pipe_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                             const int* __restrict__ pair, const int* __restrict__ rev,
                             const float4* __restrict__ geo, const float* __restrict__ filt,
                             const float* __restrict__ dfilt, const float* __restrict__ s_in,
                             const float* __restrict__ v_in, const float* __restrict__ sbar_m,
                             const float* __restrict__ vbar_m, float* __restrict__ sbar_in,
                             float* __restrict__ vbar_in, float4* __restrict__ edge_adj,
                             int num_atoms, const DeviceStatus* __restrict__ status) {
    constexpr int H = 128;
    constexpr unsigned full = 0xffffffffu;
    if (status->overflow) return;
    extern __shared__ float4 pipe_ring[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int PARTS = kBwdParts;
    const float4* ring = pipe_ring + (size_t)wib * (D * PARTS * 32) + lane;   // slot s, part p: ring[(s*PARTS+p)*32]
    const uint32_t ring_addr = smem_addr32(ring);
    const int c4 = lane * 4;
    const int held = ((lane & 16) ? 4 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
    const bool holder = (lane & 3) == 0;
    const int warp = blockIdx.x * kPipeWarps + wib, num_warps = gridDim.x * kPipeWarps;

    SegmentStream<true> stream{rowptr, col, pair, rev, geo, num_atoms, num_warps, lane, -1, 0, 0};
    stream.fetch_lookahead(warp);
    RowSegment cur = stream.from_lookahead();
    RowSegment nxt = stream.next_of(cur);
    int issued = 0, drain_until = -1, put = 0, get = 0;
    auto issue_next = [&]() {
        const int idx = issued;
        const bool in_cur = idx < cur.n;
        const int pr = in_cur ? __shfl_sync(full, cur.pair, idx) : __shfl_sync(full, nxt.pair, idx - cur.n);
        const int cj = in_cur ? __shfl_sync(full, cur.col, idx) : __shfl_sync(full, nxt.col, idx - cur.n);
        const bool up = cj > (in_cur ? cur.row : nxt.row);
        const size_t off = (size_t)pr * (3 * H) + c4;
        const uint32_t dst = ring_addr + (uint32_t)put * (PARTS * 512);
        if (!LAYER0) {
            cp_async16(dst, filt + off);
            cp_async16(dst + 512, filt + off + H);
        }
        if (up) {
            cp_async16(dst + 2 * 512, filt + off + 2 * H);
            cp_async16(dst + 3 * 512, dfilt + off);
            if (!LAYER0) cp_async16(dst + 4 * 512, dfilt + off + H);
            cp_async16(dst + 5 * 512, dfilt + off + 2 * H);
        }
        put = (put + 1 == D) ? 0 : put + 1;
        ++issued;
    };
    auto more_known = [&]() { return issued < cur.n || issued - cur.n < nxt.n; };
#pragma unroll 1
    for (int g = 0; g < D - 1; ++g) {
        if (more_known()) issue_next();
        cp_async_commit();
    }
    int row_loaded = -1;
    float4 sb = make4(0.f), vbx = make4(0.f), vby = make4(0.f), vbz = make4(0.f);
    float4 si = make4(0.f), vix = make4(0.f), viy = make4(0.f), viz = make4(0.f);
    float4 acc_s = make4(0.f), acc_x = make4(0.f), acc_y = make4(0.f), acc_z = make4(0.f);
    while (cur.row >= 0) {
        const int i = cur.row;
        if (row_loaded != i) {   // first segment of the row: own adjoints / features, residual path
            row_loaded = i;
            sb = ldg4(sbar_m + (size_t)i * H + c4);
            const float* vb = vbar_m + (size_t)i * 3 * H + c4;
            vbx = ldg4(vb); vby = ldg4(vb + H); vbz = ldg4(vb + 2 * H);
            si = ldg4(s_in + (size_t)i * H + c4);
            if (!LAYER0) {
                const float* vi = v_in + (size_t)i * 3 * H + c4;
                vix = ldg4(vi); viy = ldg4(vi + H); viz = ldg4(vi + 2 * H);
            }
            acc_s = sb; acc_x = vbx; acc_y = vby; acc_z = vbz;
        }
#pragma unroll 1
        for (int k = 0; k < cur.n; ++k) {
            while (issued < k + D && more_known()) {
                if (issued < k + D - 1) drain_until = issued;
                issue_next();
            }
            cp_async_commit();
            const int j = __shfl_sync(full, cur.col, k);
            const bool upper = j > i;
            const float4* slot = ring + get * (PARTS * 32);
            get = (get + 1 == D) ? 0 : get + 1;
            if (LAYER0 && !upper) {   // nothing to do for this edge (its slot is empty)
                if (k <= drain_until) cp_async_wait<0>(); else cp_async_wait<D - 1>();
                continue;
            }
            const float4 sbj = ldg4(sbar_m + (size_t)j * H + c4);
            const float* vbj_p = vbar_m + (size_t)j * 3 * H + c4;
            const float4 vbjx = ldg4(vbj_p), vbjy = ldg4(vbj_p + H), vbjz = ldg4(vbj_p + 2 * H);
            if (!upper) {
                if (k <= drain_until) cp_async_wait<0>(); else cp_async_wait<D - 1>();
                const float4 fa = slot[0], fb = slot[32];
                acc_s = fma4(fa, sbj, acc_s);
                acc_x = fma4(fb, vbjx, acc_x);
                acc_y = fma4(fb, vbjy, acc_y);
                acc_z = fma4(fb, vbjz, acc_z);
                continue;
            }
            const int e = cur.e0 + k;
            const int r = __shfl_sync(full, cur.aux, k);
            const float4 g = shfl4(cur.geo, k);      // unit vector of (j -> i)
            // unit vector of (i -> j): the exact negation of (j -> i) -- subtraction, the rounding of
            // the minimum-image shift and the division are all odd functions -- so no second gather
            const float4 gr = make_float4(-g.x, -g.y, -g.z, g.w);
            const float4 sj = ldg4(s_in + (size_t)j * H + c4);
            float4 vjx, vjy, vjz;
            if (!LAYER0) {
                const float* vj = v_in + (size_t)j * 3 * H + c4;
                vjx = ldg4(vj); vjy = ldg4(vj + H); vjz = ldg4(vj + 2 * H);
            }
            if (k <= drain_until) cp_async_wait<0>(); else cp_async_wait<D - 1>();
            if (!LAYER0) {
                const float4 fa = slot[0], fb = slot[32];
                acc_s = fma4(fa, sbj, acc_s);
                acc_x = fma4(fb, vbjx, acc_x);
                acc_y = fma4(fb, vbjy, acc_y);
                acc_z = fma4(fb, vbjz, acc_z);
            }
            float part[8];
            {
                const float4 abar = fma4(sj, sb, mul4(si, sbj));
                float4 cbar = fma4s(g.x, vbx, fma4s(g.y, vby, fma4s(g.z, vbz, make4(0.f))));
                cbar = fma4s(gr.x, vbjx, fma4s(gr.y, vbjy, fma4s(gr.z, vbjz, cbar)));
                float d = dot4(abar, slot[3 * 32]) + dot4(cbar, slot[5 * 32]);
                if (!LAYER0) {
                    float4 bbar = fma4(vjx, vbx, fma4(vjy, vby, mul4(vjz, vbz)));
                    bbar = fma4(vix, vbjx, fma4(viy, vbjy, fma4(viz, vbjz, bbar)));
                    d += dot4(bbar, slot[4 * 32]);
                }
                const float4 fc = slot[2 * 32];
                part[0] = dot4(fc, vbx); part[1] = dot4(fc, vby); part[2] = dot4(fc, vbz);
                part[3] = d;
                part[4] = dot4(fc, vbjx); part[5] = dot4(fc, vbjy); part[6] = dot4(fc, vbjz);
                part[7] = 0.f;
            }
            const float total = group_sum8<32>(part, lane);
            if (holder) reinterpret_cast<float*>(edge_adj + (held < 4 ? e : r))[held & 3] = total;
        }
        if (cur.rem == 0 && !LAYER0) {
            st4(sbar_in + (size_t)i * H + c4, acc_s);
            float* vo = vbar_in + (size_t)i * 3 * H + c4;
            st4(vo, acc_x); st4(vo + H, acc_y); st4(vo + 2 * H, acc_z);
        }
        issued -= cur.n;
        drain_until -= cur.n;
        cur = nxt;
        nxt = stream.next_of(cur);
    }
    cp_async_wait<0>();
}

}  // namespace mlffd
