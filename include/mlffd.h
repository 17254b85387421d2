/*
 * mlffd.h -- C ABI of the B200-native PaiNN-student energy+force library (libmlffd.so).
 *
 * The reference (atfrank/MLFF-Distiller) is 100% Python and has no FFI boundary for this path;
 * the entry points below are what a binding for its two hot-path classes needs.  Each export
 * cites the reference interface it replaces (paths relative to /root/reference).  All pointers
 * suffixed _d are DEVICE pointers (plain CUDA memory, e.g. torch.Tensor.data_ptr()); `stream`
 * is a cudaStream_t passed as void*.  No torch / C++ types cross the boundary.
 *
 * Conventions: return 0 on success, a negative MLFFD_E* code on failure; never throws.  A
 * context is bound to one device and is NOT thread-safe; distinct contexts are independent
 * (multi-GPU = one context per GPU).  All work is enqueued asynchronously on `stream`; only
 * mlffd_model_create, mlffd_workspace_reserve, mlffd_get_status and mlffd_debug_buffer
 * synchronise.  There is no CPU fallback: every call fails with MLFFD_ECUDA without a GPU.
 */
#ifndef MLFFD_H
#define MLFFD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLFFD_ABI_VERSION 2

enum {
    MLFFD_OK = 0,
    MLFFD_EINVAL = -1,    /* bad argument (null pointer, unsupported hidden_dim, ...) */
    MLFFD_ECUDA = -2,     /* CUDA runtime error; text in mlffd_last_error */
    MLFFD_ECAPACITY = -3, /* workspace too small for the request (reserve more and retry) */
    MLFFD_ENOMEM = -4
};

/* Arithmetic of the dense layers.  FP32 is the parity default (north-star tolerances 1e-5
 * eV/atom, 1e-4 eV/A); the others trade accuracy for tensor-core speed with looser bounds. */
enum {
    MLFFD_PREC_FP32 = 0,      /* FP32 FFMA everywhere */
    MLFFD_PREC_TC_FP16X2 = 1, /* tcgen05 tensor cores, operands split into two FP16 terms, three
                                 products, FP32 accumulation in TMEM: FP32-equivalent (meets the
                                 FP32 bounds).  Filter table for H = 128 / 64 / 32, update block
                                 for H = 128 (FFMA otherwise). */
    MLFFD_PREC_TF32 = 2,      /* reserved (not implemented): MLFFD_PREC_TC_FP16 has the same 11-bit
                                 significands at twice the tensor-core rate */
    MLFFD_PREC_TC_BF16 = 3,   /* single product of BF16-rounded operands (8-bit significands), FP32
                                 accumulation; looser bounds: DESIGN.md section 6 */
    MLFFD_PREC_TC_FP16 = 4    /* single product of FP16-rounded operands (11-bit significands, the
                                 precision class of TF32), FP32 accumulation; looser bounds */
};

/* How the radial filter f_l(d) = W2 SiLU(W1 phi~(d) + b1) + b2 (student_model.py:318-322, 350) reaches
 * the message kernels.  It depends on the distance only, never on the structure. */
enum {
    MLFFD_FILTER_SPLINE = 0, /* default: a quintic B-spline of every component function on [0, cutoff]
                                (192 intervals) built once per model in FP64; the message kernels keep a
                                32-channel slice of it in shared memory and evaluate value and
                                d-derivative per edge.  Nothing per-pair is written to HBM.  Stated bound:
                                |f - spline| <= 2e-7, |f' - spline'| <= 1e-5 / Angstrom as functions (FP64
                                evaluation); <= 2e-6 / 2e-5 evaluated in FP32 on the device. */
    MLFFD_FILTER_TABLE = 1   /* the two dense layers evaluated per undirected pair and step into a
                                [P,3H] table (+ derivative) in HBM that the message kernels stream
                                (tensor cores per `precision`). */
};

/* Hyper-parameters = the `config` dict of StudentForceField.save
 * (src/mlff_distiller/models/student_model.py:1087-1094). */
typedef struct mlffd_config {
    int32_t hidden_dim;       /* H: 128 (Original), 64 (Tiny), 32 (Ultra-tiny) */
    int32_t num_rbf;          /* K <= 32 */
    int32_t num_interactions; /* L <= 8 */
    int32_t max_z;            /* embedding has max_z + 1 rows */
    float cutoff;             /* r_c in Angstrom */
    int32_t precision;        /* MLFFD_PREC_* */
    int32_t filter_mode;      /* MLFFD_FILTER_* */
} mlffd_config;

/* Read back by mlffd_get_status (synchronises the stream of the last call). */
typedef struct mlffd_status {
    int64_t num_atoms;
    int64_t num_edges;   /* directed edges found by the last neighbour build */
    int64_t num_pairs;   /* undirected pairs = num_edges / 2 */
    int64_t edge_capacity;
    int32_t overflow;    /* 1: edges exceeded capacity, outputs invalid -> reserve and retry */
    int32_t max_degree;
    int64_t overflow_events; /* sticky count of overflowed builds since mlffd_model_create (lets a
                                caller that enqueues many steps, e.g. on-device MD, detect one) */
    int32_t tc_saturated;    /* 1: an operand of a tensor-core dense layer left the FP16 range in this step
                                (|activation| >= 8 125): energies / forces invalid -> call
                                mlffd_set_dense_fallback(ctx, 1) and evaluate again */
    int32_t skin_rebuilds;   /* candidate-list builds since the skin list was (re)started (0 without a skin) */
} mlffd_status;

typedef struct mlffd_ctx mlffd_ctx;

int mlffd_version(void);

/* Message of the last failure on `ctx` (or of the last failed create when ctx == NULL). */
const char* mlffd_last_error(const mlffd_ctx* ctx);

/*
 * Replaces StudentForceField.__init__ + load_state_dict (student_model.py:566-616, 1164-1170).
 * `weights_host`: every tensor of the reference state_dict, FP32 little-endian, row-major, torch
 * [out,in] layout for Linear weights, concatenated in this order:
 *   embedding.weight [max_z+1,H]; rbf.centers [K]; rbf.widths [K];
 *   for l in 0..L-1: message.rbf_to_scalar.0.{weight [H,K], bias [H]},
 *                    message.rbf_to_scalar.2.{weight [3H,H], bias [3H]},
 *                    update.update_mlp.0.{weight [H,2H], bias [H]},
 *                    update.update_mlp.2.{weight [3H,H], bias [3H]}, update.mixing_matrix [3,3];
 *   energy_head.0.{weight [H/2,H], bias [H/2]}; energy_head.2.{weight [H/4,H/2], bias [H/4]};
 *   energy_head.4.{weight [1,H/4], bias [1]}.
 * The library copies and re-lays the weights out on `device`; the caller keeps the host blob.
 */
int mlffd_model_create(mlffd_ctx** out, int device, const mlffd_config* config,
                       const float* weights_host, size_t num_floats);
void mlffd_model_destroy(mlffd_ctx* ctx);

/* Pre-size every scratch buffer; nothing is allocated on the hot path afterwards.  Growing is
 * allowed at any time (synchronises).  max_edges counts DIRECTED edges. */
int mlffd_workspace_reserve(mlffd_ctx* ctx, int64_t max_atoms, int64_t max_edges,
                            int64_t max_structures);

/*
 * Replaces radius_graph / radius_graph_native (student_model.py:63-109, 165-191).
 *   pos_d      [N,3] f32 positions
 *   offsets_d  [B+1] i32 first atom of each structure (replaces the `batch` vector)
 *   cells_d    [B,18] f32: row-major 3x3 cell followed by its row-major 3x3 inverse, or NULL
 *   pbc_d      [B,3] u8 periodic flags, or NULL (open boundaries, the reference's behaviour)
 * Builds destination-sorted CSR (sources ascending within a row) inside the context.
 * Edge (i -> j) iff same structure, i != j and |x_i - x_j (minimum image)| <= cutoff in FP32.
 */
int mlffd_neighbor_list(mlffd_ctx* ctx, const float* pos_d, const int32_t* offsets_d,
                        int32_t num_structures, int64_t num_atoms, const float* cells_d,
                        const uint8_t* pbc_d, void* stream);

/* Copy the last neighbour list out as the reference's edge_index [2,E] int64, row 0 = src,
 * row 1 = dst, lexicographic (src,dst) order (student_model.py:106-107).  `capacity_edges` is
 * the room in edge_index_d per row; fails with MLFFD_ECAPACITY if smaller than E (syncs). */
int mlffd_export_edges(mlffd_ctx* ctx, int64_t* edge_index_d, int64_t capacity_edges,
                       int64_t* num_edges_out, void* stream);

/*
 * Replaces StudentForceField.forward + predict_energy_and_forces (student_model.py:634-795) and
 * the batched recipe of StudentForceFieldCalculator._batch_forward
 * (src/mlff_distiller/inference/ase_calculator.py:708-770).
 *   z_d        [N] i32 atomic numbers (0..max_z)
 *   energy_d   [B] f32 out: total energy per structure (eV)
 *   forces_d   [N,3] f32 out: -dE/dx (eV/A); NULL = energy only (no reverse pass)
 * Runs neighbour build, (filter tables,) L x (message, update), readout and the analytical
 * reverse pass, all on `stream`, no host synchronisation.  Check mlffd_get_status().overflow
 * after synchronising: if set, the outputs are invalid; reserve more edges and call again.
 */
int mlffd_energy_forces(mlffd_ctx* ctx, const int32_t* z_d, const float* pos_d,
                        const int32_t* offsets_d, int32_t num_structures, int64_t num_atoms,
                        const float* cells_d, const uint8_t* pbc_d, float* energy_d,
                        float* forces_d, void* stream);

int mlffd_get_status(mlffd_ctx* ctx, mlffd_status* out);

/*
 * Non-blocking companion of mlffd_get_status for callers that keep several steps in flight (the
 * pipelined batched-structure interface, StudentForceFieldCalculator.evaluate_stream): enqueues on
 * `stream` a copy of the device status words of the step enqueued just before it into
 * `status_out`, six int32 in pinned host (or device) memory:
 *   [0] num_edges  [1] num_pairs  [2] overflow  [3] max_degree  [4] overflow_events (sticky)
 *   [5] tc_saturated
 * Valid once the stream has passed this point (record an event after the call).  Never synchronises.
 */
int mlffd_status_async(mlffd_ctx* ctx, int32_t* status_out, void* stream);

/*
 * Verlet-skin neighbour list, an option; default 0 = the exact list is rebuilt
 * from scratch on every call, like the reference does (student_model.py:694-703).  skin > 0 (Angstrom):
 * a candidate list of all pairs within cutoff + skin is kept and rebuilt only when some atom has moved
 * further than skin / 2 since it was built (decided on the device, no host synchronisation); every call
 * still derives the EXACT d <= cutoff list from the candidates with the same pair test, so edges, their
 * order and the geometry are bit-identical to the full build.  Meant for a trajectory of one system: the
 * list is tied to (num_atoms, num_structures) and restarts when they change; call mlffd_set_skin again
 * (or change the skin) after replacing the atoms or the cell of a same-sized system.  Periodic cells need
 * heights >= 2 (cutoff + skin).  Frees the workspace: reserve again afterwards.  Systems of up to 64 atoms
 * keep the one-launch exact kernel.
 */
int mlffd_set_skin(mlffd_ctx* ctx, float skin);

/*
 * Range guard of the tensor-core precisions.  MLFFD_PREC_TC_FP16X2 / _TC_FP16 split operands into FP16
 * terms after a power-of-two pre-scale (activations x 2^3, weights x 2^8): an activation or adjoint with
 * |x| >= 8 125 (unphysically dense inputs; the reference aggregates un-normalised neighbour sums) would
 * round to infinity.  Weights are checked once in mlffd_model_create (a model whose max|w| >= 253 keeps
 * its dense layers on the FP32 FFMA kernels); activations are checked by the kernels that split them,
 * which raise mlffd_status.tc_saturated.  enable = 1 makes every later call run the dense layers on the
 * FP32 FFMA kernels (the MLFFD_PREC_FP32 arithmetic) whatever `precision` says; 0 restores the tensor
 * cores.  The reference has no counterpart (it computes in FP32 throughout).
 */
int mlffd_set_dense_fallback(mlffd_ctx* ctx, int32_t enable);

/*
 * Stage entry point (parity tests, ncu): evaluate the per-layer radial filter and its
 * d-derivative for `num_pairs` distances.  Restates PaiNNMessage.rbf_to_scalar applied to
 * GaussianRBF * CosineCutoff (student_model.py:249-255, 285-292, 318-322, 350).
 *   dist_d [P] f32 -> filter_d [P,3H] f32, dfilter_d [P,3H] f32 (d filter / d distance)
 */
int mlffd_filter_table(mlffd_ctx* ctx, int32_t layer, const float* dist_d, int64_t num_pairs,
                       float* filter_d, float* dfilter_d, void* stream);

/*
 * Operator-level drop-ins for the reference's two Triton kernels (stateless: no context; device pointers).
 *
 * mlffd_edge_features replaces fused_edge_features_triton(positions, edge_index, eps)
 * (kernels/fused_edge_features.py:99-162):  edge_index_d [2,E] i64 (row 0 = src, row 1 = dst) ->
 *   edge_vec_d [E,3] = pos[src] - pos[dst],  dist_d [E],  unit_d [E,3].
 *   eps_mode 0: d = |r|, u = r / (d + eps)          the model's own placement (student_model.py:712-715)
 *   eps_mode 1: d = sqrt(|r|^2 + eps), u = r / d    the Triton kernel's placement (fused_edge_features.py:77-82)
 * mlffd_rbf_cutoff replaces fused_rbf_cutoff_triton(distances, centers, gamma, r_cut)
 * (kernels/fused_rbf_cutoff.py:90-134):  out_d [E,K] = exp(-gamma (d - mu_k)^2) * 0.5 (cos(pi d / rc) + 1) [d < rc].
 * On the product path both are fused into the neighbour fill pass and the filter evaluation; these exports
 * serve callers of the reference's kernel-level API and the stage parity tests.
 */
int mlffd_edge_features(const float* pos_d, const int64_t* edge_index_d, int64_t num_edges, float eps,
                        int32_t eps_mode, float* edge_vec_d, float* dist_d, float* unit_d, void* stream);
int mlffd_rbf_cutoff(const float* dist_d, int64_t num_edges, const float* centers_d, int32_t num_rbf,
                     float gamma, float cutoff, float* out_d, void* stream);

/*
 * Stage entry point (parity tests): the same quantities as mlffd_filter_table, evaluated from the
 * per-model filter spline the MLFFD_FILTER_SPLINE message kernels use (value and exact derivative of
 * the interpolant).  Works in either filter mode.
 */
int mlffd_filter_spline(mlffd_ctx* ctx, int32_t layer, const float* dist_d, int64_t num_pairs,
                        float* filter_d, float* dfilter_d, void* stream);

/*
 * Test hook: device pointer + element count of an internal buffer of the last call
 * (synchronises).  Names: "rowptr","col","rev","pair","edge_dst","geo" (float4 ux,uy,uz,d),
 * "pair_dist", "filter"/"dfilter" (layer; MLFFD_FILTER_TABLE only), "s_in"/"v_in"/"s_msg"/"v_msg"/"y1"/"gates" (layer),
 * "s_out", "atom_energy", "sbar"/"vbar" (adjoints of s_msg / v_msg of `layer`; all layers are
 * kept only when the context was created with MLFFD_DEBUG_KEEP=1 in the environment, otherwise
 * two sets ping-pong), "edge_adj" (float4 dE/du_x, dE/du_y, dE/du_z, dE/dd through the filters).
 * elem_size_out: bytes per element.
 */
int mlffd_debug_buffer(mlffd_ctx* ctx, const char* name, int32_t layer, void** ptr_out,
                       int64_t* count_out, int32_t* elem_size_out);

/*
 * Launch accounting and per-stage device timing (bench.py roofline / gpu_launches).  While
 * enabled, a CUDA event is recorded on the step's stream after every kernel of a stage, so
 * stage_ms is the device time of that stage's kernels inside the live step (the analogue of the
 * reference's CUDATimer, src/mlff_distiller/cuda/benchmark_utils.py:147-193, moved inside the
 * step).  enable resets the counters; read synchronises and accumulates.
 */
enum {
    MLFFD_STAGE_NEIGHBOR = 0,
    MLFFD_STAGE_EMBEDDING = 1,
    MLFFD_STAGE_FILTER = 2,
    MLFFD_STAGE_MESSAGE_FWD = 3,
    MLFFD_STAGE_UPDATE_FWD = 4,
    MLFFD_STAGE_READOUT = 5,
    MLFFD_STAGE_ENERGY_SUM = 6,
    MLFFD_STAGE_UPDATE_BWD = 7,
    MLFFD_STAGE_MESSAGE_BWD = 8,
    MLFFD_STAGE_FORCE = 9,
    MLFFD_NUM_STAGES = 10
};

typedef struct mlffd_profile {
    int64_t launches;                         /* kernels launched since enable */
    int64_t stage_launches[MLFFD_NUM_STAGES];
    double stage_ms[MLFFD_NUM_STAGES];        /* only filled while profiling is enabled */
} mlffd_profile;

int mlffd_profile_enable(mlffd_ctx* ctx, int32_t enable);
int mlffd_profile_read(mlffd_ctx* ctx, mlffd_profile* out);
const char* mlffd_stage_name(int32_t stage);

/*
 * Strain derivative of the energy of the LAST mlffd_energy_forces call (which must have been made
 * with forces_d != NULL, on the same stream, with the same offsets):
 *   virial_d[b*9 + 3*a + c] = dE_b / d eps_ac = sum over the directed edges e of structure b of
 *   (dE/dr_e)_a (r_e)_c,   r_e = x_src - x_dst (minimum image),   x -> (1 + eps) x.
 * stress = virial / volume (eV/A^3; ASE Voigt order xx, yy, zz, yz, xz, xy after symmetrising).
 * The reference's counterpart, inference/ase_calculator.py:521-588 (_compute_stress), differentiates
 * with respect to a `cell` tensor the model never reads and returns zeros; this entry point is the
 * working version and its parity is pinned only by the oracle's autograd
 * strain derivative and by finite differences.
 */
int mlffd_virial(mlffd_ctx* ctx, const int32_t* offsets_d, int32_t num_structures, float* virial_d,
                 void* stream);

/*
 * On-device velocity Verlet (replaces the host-side ASE VelocityVerlet loop the reference drives
 * through src/mlff_distiller/testing/nve_harness.py:214-235; one step = kick_drift,
 * mlffd_energy_forces on pos32_d, kick_energy).  Integrator state is FP64 (as ASE's), the model
 * sees FP32 positions (as inference/ase_calculator.py:497-500).  ASE units: eV, Angstrom, amu;
 * dt in ASE time units (fs * 0.0982269...).  The three calls are graph-capturable.
 *   kick_drift : v += dt/2 F/m ; x += dt v ; pos32 = float(x)
 *   kick_energy: v += dt/2 F/m ; series[*counter] = (sum_b E_b, sum m v^2/2) ; ++*counter
 * `guard` (may be NULL) is the context whose mlffd_energy_forces supplies the forces: when its last force
 * evaluation failed on the device (mlffd_status.overflow or .tc_saturated), both calls do nothing, so a
 * trajectory enqueued many steps ahead freezes at the valid mid-step state (x_k, v_{k-1/2}) of the failing
 * step: reserve more edges / mlffd_set_dense_fallback, evaluate the forces again, call kick_energy, go on.
 */
int mlffd_md_kick_drift(const mlffd_ctx* guard, int64_t num_atoms, double* pos_d, double* vel_d, const float* forces_d,
                        const double* inv_mass_d, double dt, float* pos32_d, void* stream);
int mlffd_md_kick_energy(const mlffd_ctx* guard, int64_t num_atoms, double* vel_d, const float* forces_d,
                         const double* inv_mass_d, double dt, const float* energy_d,
                         int32_t num_structures, double* series_d, int32_t* counter_d,
                         int32_t capacity, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLFFD_H */
