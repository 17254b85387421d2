// Micro-benchmark of the SM's L1 data pipe for the access shapes of the spline message kernels (sm_100a):
// cycles per warp-level load instruction at saturation (32 warps per SM, L1-resident 16 KB working set).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_probe tools/l1_probe.cu && ./l1_probe
//   shape 0: LDG.128, the four 8-lane groups of a warp read four different 128-byte lines (gather of the row kernels)
//   shape 1: LDG.32, the 32 lanes read one 128-byte line
//   shape 2: LDG.128, each group reads ONE 16-byte entry (broadcast), four different lines per instruction (records)
//   shape 3: LDG.128, all 32 lanes read the same 16 bytes
//   shape 4: LDS.128, four groups read four different 128-byte rows of shared memory (coefficient rows)
//   shape 5: LDS.32, 32 lanes read one 128-byte row
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kLines = 128;   // 128 lines x 128 B = 16 KB

template <int SHAPE>
__global__ void __launch_bounds__(1024) probe(const float4* __restrict__ buf, float* out, long long* cycles) {
    __shared__ float4 sm[kLines * 8];
    for (int i = threadIdx.x; i < kLines * 8; i += blockDim.x) sm[i] = buf[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 3, l8 = lane & 7, warp = threadIdx.x >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc1 = 0.f;
    unsigned h = 1234567u * (warp + 1) + blockIdx.x;
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < kIters; ++it) {
        h = h * 1664525u + 1013904223u;
        const int line_g = ((h >> 8) + 37 * g) % kLines;    // a different line per group
        const int line_w = (h >> 8) % kLines;               // one line per warp
        if (SHAPE == 0) { const float4 v = __ldg(buf + line_g * 8 + l8); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        if (SHAPE == 1) { acc1 += __ldg(reinterpret_cast<const float*>(buf) + line_w * 32 + lane); }
        if (SHAPE == 2) { const float4 v = __ldg(buf + line_g * 8 + ((h >> 4) & 7)); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        if (SHAPE == 3) { const float4 v = __ldg(buf + line_w * 8 + ((h >> 4) & 7)); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        if (SHAPE == 4) { const float4 v = sm[line_g * 8 + l8]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        if (SHAPE == 5) { acc1 += reinterpret_cast<const float*>(sm)[line_w * 32 + lane]; }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w + acc1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int SHAPE>
void run(const float4* buf, float* out, long long* cyc_d, const char* name) {
    const int blocks = 148, threads = 1024;   // 32 warps per SM
    probe<SHAPE><<<blocks, threads>>>(buf, out, cyc_d);
    probe<SHAPE><<<blocks, threads>>>(buf, out, cyc_d);
    cudaDeviceSynchronize();
    long long c[148];
    cudaMemcpy(c, cyc_d, sizeof(c), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += (double)c[i]; mean /= blocks;
    // per SM: 32 warps x kIters instructions
    printf("%-70s %7.2f cycles per warp instruction (SM-wide), %s\n", name, mean / (32.0 * kIters), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float4* buf; float* out; long long* cyc;
    cudaMalloc(&buf, kLines * 8 * sizeof(float4)); cudaMemset(buf, 0, kLines * 8 * sizeof(float4));
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    run<0>(buf, out, cyc, "LDG.128  4 groups x 128 B, four lines per instruction");
    run<1>(buf, out, cyc, "LDG.32   32 lanes x 4 B, one line per instruction");
    run<2>(buf, out, cyc, "LDG.128  4 groups, one 16 B entry each (broadcast), four lines");
    run<3>(buf, out, cyc, "LDG.128  all lanes the same 16 B");
    run<4>(buf, out, cyc, "LDS.128  4 groups x 128 B, four rows per instruction");
    run<5>(buf, out, cyc, "LDS.32   32 lanes x 4 B, one row per instruction");
    return 0;
}
