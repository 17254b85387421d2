// Standalone tcgen05 probe: D[128x128] = A[128xK] * B[128xK]^T, fp16 inputs, fp32 accumulate in
// TMEM, K-major SWIZZLE_128B shared-memory operands written by ordinary threads.  Sweeps the
// descriptor fields this project relies on and prints the max error of each candidate, so the
// encodings used in csrc/filter_umma.cuh are pinned by measurement, not by memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_test.bin tools/umma_test.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 128, KB = 64;  // one swizzle atom along K = 64 halves = 128 bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
    uint32_t lbo_enc, sbo_enc, layout_type, version, idesc;
    int num_kblocks;      // K = 64 * num_kblocks
    uint32_t kblock_bytes; // distance between K-block images in smem
};

__global__ void __launch_bounds__(128) umma_probe(const uint4* __restrict__ a_img, const uint4* __restrict__ b_img,
                                                  float* __restrict__ d_out, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + p.num_kblocks * p.kblock_bytes;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int img_vec = p.num_kblocks * p.kblock_bytes / 16;
    for (int i = tid; i < img_vec; i += 128) {
        reinterpret_cast<uint4*>(a_s)[i] = a_img[i];
        reinterpret_cast<uint4*>(b_s)[i] = b_img[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        auto make_desc = [&](uint32_t saddr) {
            uint64_t d = 0;
            d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
            d |= (uint64_t)(p.lbo_enc & 0x3FFFu) << 16;
            d |= (uint64_t)(p.sbo_enc & 0x3FFFu) << 32;
            d |= (uint64_t)(p.version & 3u) << 46;
            d |= (uint64_t)(p.layout_type & 7u) << 61;
            return d;
        };
        int first = 1;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
            for (int k = 0; k < KB / 16; ++k) {
                const uint64_t da = make_desc(smem_u32(a_s) + kb * p.kblock_bytes + k * 32);
                const uint64_t db = make_desc(smem_u32(b_s) + kb * p.kblock_bytes + k * 32);
                const uint32_t acc = first ? 0u : 1u;
                first = 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                    "l"(da), "l"(db), "r"(p.idesc), "r"(acc)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&mbar)), "r"(0u)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) d_out[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

// K-major SWIZZLE_128B image of a [rows x 64] fp16 block: 8-row groups of 1024 B, row r at
// (r%8)*128, 16-byte chunk c stored at chunk position c ^ (r%8).
static void make_image(const std::vector<__half>& src, int rows, int ld, int k0, std::vector<__half>& img) {
    img.assign((size_t)rows * 64, __float2half(0.f));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < 64; ++c) {
            const int chunk = c / 8, within = c % 8;
            const size_t off = (size_t)(r / 8) * 512 + (size_t)(r % 8) * 64 + (size_t)((chunk ^ (r % 8)) * 8) + within;
            img[off] = src[(size_t)r * ld + k0 + c];
        }
}

int main() {
    const int nkb = 2, K = 64 * nkb;
    std::vector<__half> A((size_t)M * K), B((size_t)N * K);
    srand(1);
    for (auto& v : A) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    std::vector<float> ref((size_t)M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)__half2float(A[(size_t)m * K + k]) * __half2float(B[(size_t)n * K + k]);
            ref[(size_t)m * N + n] = (float)s;
        }
    std::vector<__half> a_img, b_img, tmp;
    for (int kb = 0; kb < nkb; ++kb) {
        make_image(A, M, K, kb * 64, tmp); a_img.insert(a_img.end(), tmp.begin(), tmp.end());
        make_image(B, N, K, kb * 64, tmp); b_img.insert(b_img.end(), tmp.begin(), tmp.end());
    }
    __half *a_d, *b_d; float* d_d;
    CK(cudaMalloc(&a_d, a_img.size() * 2)); CK(cudaMalloc(&b_d, b_img.size() * 2)); CK(cudaMalloc(&d_d, ref.size() * 4));
    CK(cudaMemcpy(a_d, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b_d, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = 2 * nkb * 16384 + 1024;
    CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto idesc_of = [](int mshift_bit) {
        uint32_t d = 0;
        d |= 1u << 4;                    // c_format = F32
        d |= 0u << 7;                    // a_format = F16
        d |= 0u << 10;                   // b_format = F16
        d |= 0u << 15;                   // a K-major
        d |= 0u << 16;                   // b K-major
        d |= (uint32_t)(N >> 3) << 17;   // n_dim
        d |= (uint32_t)(M >> 4) << mshift_bit;  // m_dim
        return d;
    };
    struct Cand { const char* name; Params p; };
    std::vector<Cand> cands = {
        {"sw128 lbo=0 sbo=64 v=1 m@24", {0, 64, 2, 1, idesc_of(24), nkb, 16384}},
        {"sw128 lbo=1 sbo=64 v=1 m@24", {1, 64, 2, 1, idesc_of(24), nkb, 16384}},
        {"sw128 lbo=64 sbo=64 v=1 m@24", {64, 64, 2, 1, idesc_of(24), nkb, 16384}},
        {"sw128 lbo=0 sbo=64 v=0 m@24", {0, 64, 2, 0, idesc_of(24), nkb, 16384}},
        {"sw128 lbo=0 sbo=64 v=1 m@23", {0, 64, 2, 1, idesc_of(23), nkb, 16384}},
    };
    std::vector<float> out(ref.size());
    for (auto& c : cands) {
        CK(cudaMemset(d_d, 0xff, ref.size() * 4));
        umma_probe<<<1, 128, smem>>>((const uint4*)a_d, (const uint4*)b_d, d_d, c.p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-34s -> CUDA error: %s\n", c.name, cudaGetErrorString(e)); return 2; }
        CK(cudaMemcpy(out.data(), d_d, ref.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (size_t i = 0; i < ref.size(); ++i) {
            const double err = std::fabs((double)out[i] - ref[i]);
            if (!(err <= 1e-3)) ++bad;
            if (err > maxerr || err != err) maxerr = err;
        }
        printf("%-34s -> max err %.3e, bad %d / %zu  (D[0][0]=%f ref %f, D[5][77]=%f ref %f)\n", c.name, maxerr, bad,
               ref.size(), out[0], ref[0], out[5 * N + 77], ref[5 * N + 77]);
    }
    return 0;
}
