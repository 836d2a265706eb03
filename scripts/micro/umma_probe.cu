// Probe: one CTA, D[128 x N] (fp32, TMEM) = A[128 x K] . B[N x K]^T with bf16 operands in the no-swizzle K-major
// canonical layout ([K/8 chunks][rows][8 bf16]: SBO = 128 B between 8-row groups, LBO = rows * 16 B between the two
// 16-byte K chunks of one MMA), staged by one cp.async.bulk each.  Validates the shared-memory / instruction
// descriptors and the TMEM read-back (32x32b) used by the LBS kernel.   nvcc -arch=sm_100a -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    return d;                               // layout type 0 = no swizzle, base offset 0
}

template <int N, int K>
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                             float* __restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_base;
    unsigned char* sA = smem;                                  // [K/8][128][16 B]
    unsigned char* sB = smem + (size_t)K / 8 * 128 * 16;       // [K/8][N][16 B]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t bytesA = K / 8 * 128 * 16, bytesB = K / 8 * N * 16;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"(bytesA + bytesB) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sA)), "l"(A), "r"(bytesA), "r"(smem_u32(&bar_full)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sB)), "l"(B), "r"(bytesB), "r"(smem_u32(&bar_full)) : "memory");
    }
    if (warp == 1) {
        // wait for the operands
        asm volatile("{\n.reg .pred p;\nW0:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D0;\nbra W0;\nD0:\n}\n"
                     ::"r"(smem_u32(&bar_full)), "r"(0) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            // instruction descriptor: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), K-major both, N >> 3 at 17, M >> 4 at 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int s = 0; s < K / 16; ++s) {
                const uint64_t da = make_desc(smem_u32(sA) + s * 2 * 128 * 16, 128 * 16, 128);
                const uint64_t db = make_desc(smem_u32(sB) + s * 2 * N * 16, N * 16, 128);
                const uint32_t acc = s > 0;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
        }
        __syncwarp();
    }
    // all four warps read their lane quadrant back
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n"
                 ::"r"(smem_u32(&bar_done)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int N, int K>
int run() {
    std::vector<float> a(128 * K), b(N * K);
    for (auto& x : a) x = (float)(rand() % 17 - 8) / 8.f;
    for (auto& x : b) x = (float)(rand() % 13 - 6) / 4.f;
    std::vector<__nv_bfloat16> ta(a.size()), tb(b.size());
    for (int k = 0; k < K; ++k) for (int r = 0; r < 128; ++r) ta[((size_t)(k / 8) * 128 + r) * 8 + k % 8] = __float2bfloat16(a[r * K + k]);
    for (int k = 0; k < K; ++k) for (int r = 0; r < N; ++r) tb[((size_t)(k / 8) * N + r) * 8 + k % 8] = __float2bfloat16(b[r * K + k]);
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, ta.size() * 2); cudaMalloc(&dB, tb.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, ta.data(), ta.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, tb.data(), tb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * N * 4);
    const size_t smem = (size_t)K / 8 * (128 + N) * 16;
    cudaFuncSetAttribute(probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<N, K><<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d: CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
    std::vector<float> d(128 * N);
    cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)a[i * K + k] * b[j * K + k];
        worst = fmax(worst, fabs(ref - d[i * N + j]));
    }
    printf("N=%d K=%d: max |D - ref| = %g  (D[0][0]=%g D[5][3]=%g D[127][%d]=%g)\n", N, K, worst, d[0], d[5 * N + 3], N - 1, d[127 * N + N - 1]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return worst < 1e-3 ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run<128, 32>();
    bad += run<64, 64>();
    bad += run<16, 16>();
    bad += run<128, 224>();
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
