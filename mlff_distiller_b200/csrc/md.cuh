// On-device velocity-Verlet: positions, velocities and forces stay in
// HBM for a whole trajectory; the host only reads the energy series.  Restates what ASE's
// VelocityVerlet does around the reference calculator (src/mlff_distiller/testing/
// nve_harness.py:214-235): half kick, drift, force evaluation, half kick -- integrator state in
// FP64 like ASE's numpy arrays, positions rounded to FP32 only as model input, exactly like
// inference/ase_calculator.py:497-500 does every step.
//
// Guard: when the force evaluation of a step failed on the device (edge-workspace overflow or an FP16
// saturation of a tensor-core operand -- DeviceStatus), every kick / drift / record kernel that follows
// returns without touching the state, so the trajectory FREEZES at the mid-step state (x_k, v_{k-1/2}) of
// the failing step k, which is valid: the host grows the workspace (or switches the dense layers to the
// FP32 kernels), repeats the force evaluation and the second half kick, and carries on (md.DeviceMD.run).
#pragma once
#include "common.cuh"

namespace mlffd {

// v += dt/2 * F/m ; x += dt * v ; x32 = float(x)
__global__ void __launch_bounds__(256)
md_kick_drift_kernel(long long n3, double* __restrict__ pos, double* __restrict__ vel,
                     const float* __restrict__ forces, const double* __restrict__ inv_mass,
                     double dt, float* __restrict__ pos32, const DeviceStatus* __restrict__ guard) {
    if (guard != nullptr && (guard->overflow | guard->tc_saturated)) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n3;
         i += (long long)gridDim.x * blockDim.x) {
        const double v = vel[i] + 0.5 * dt * (double)forces[i] * inv_mass[i / 3];
        const double x = pos[i] + dt * v;
        vel[i] = v;
        pos[i] = x;
        pos32[i] = (float)x;
    }
}

// v += dt/2 * F/m ; then one block reduces KE = sum m v^2 / 2 and PE = sum_b E_b (fixed order),
// appends (PE, KE) to the series at *counter and increments it.
__global__ void __launch_bounds__(256)
md_kick_kernel(long long n3, double* __restrict__ vel, const float* __restrict__ forces,
               const double* __restrict__ inv_mass, double dt, const DeviceStatus* __restrict__ guard) {
    if (guard != nullptr && (guard->overflow | guard->tc_saturated)) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n3;
         i += (long long)gridDim.x * blockDim.x)
        vel[i] += 0.5 * dt * (double)forces[i] * inv_mass[i / 3];
}

// KICK (small systems): the second half kick happens in the same single block, element by element
// before it enters the kinetic energy -- one launch instead of two, same values in the same order.
template <bool KICK>
__global__ void __launch_bounds__(1024)
md_energy_kernel(long long n3, double* __restrict__ vel, const float* __restrict__ forces,
                 const double* __restrict__ inv_mass, double dt,
                 const float* __restrict__ energy, int num_structures, double* __restrict__ series,
                 int* __restrict__ counter, int capacity, const DeviceStatus* __restrict__ guard) {
    if (guard != nullptr && (guard->overflow | guard->tc_saturated)) return;
    __shared__ double red[32];
    double ke = 0.0;
    for (long long i = threadIdx.x; i < n3; i += blockDim.x) {
        double v = vel[i];
        if (KICK) {
            v += 0.5 * dt * (double)forces[i] * inv_mass[i / 3];
            vel[i] = v;
        }
        ke += 0.5 * v * v / inv_mass[i / 3];
    }
    double pe = 0.0;
    for (int b = threadIdx.x; b < num_structures; b += blockDim.x) pe += (double)energy[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ke += __shfl_xor_sync(0xffffffffu, ke, o);
        pe += __shfl_xor_sync(0xffffffffu, pe, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = ke;
    __syncthreads();
    double ke_tot = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) ke_tot += red[w];
    __syncthreads();
    if (lane == 0) red[warp] = pe;
    __syncthreads();
    if (threadIdx.x == 0) {
        double pe_tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) pe_tot += red[w];
        const int k = *counter;
        if (k < capacity) { series[2 * k] = pe_tot; series[2 * k + 1] = ke_tot; }
        *counter = k + 1;
    }
}

}  // namespace mlffd
