// Pose-blend gradient contraction of the LBS backward on the tensor cores (sm_100a, tcgen05):
//
//     g_pf[b][k] = sum_c posedirs[k][c] g_vposed[b][c]          k < 207,  c < 3V = 20,670,  b < B
//
// a [207 x 3V] x [3V x B] GEMM whose K dimension is the long one.  Same recipe as the forward (lbs_tc.cu): every
// fp32 operand is three bf16 terms, the six products a_i b_j with i + j <= 4 are accumulated in fp32 in TMEM, the
// operands sit in global memory in the no-swizzle K-major canonical layout so that a pipeline stage is two plain
// cp.async.bulk copies.  M = 128 rows of k (two M blocks: 207 -> 256, the padding rows are zero), N = up to 128
// bodies, K is cut into FIXED slabs of 176 coordinates (11 k-steps; 118 slabs at SMPL size): a CTA owns one
// (slab, body tile, M block), writes its [128 x N] partial, and lbs_tc_bwd_reduce_kernel adds the slabs in index
// order.  Slab boundaries do not depend on the batch, and a body's column never mixes with another's, so a body's
// gradient is bit-identical whatever batch it is part of.
//   model operand   (tuch_smpl_create, once):  [slab][k-step][M block][term][chunk][128 rows][8]   12,288 B / stage
//   gradient operand (lbs_bwd_vertex_kernel):  [body tile][k-step][term][chunk][128 bodies][8]      12,288 B / stage
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue, warps 2-5 = epilogue.
#include "api_internal.h"
#include "smpl_internal.h"

#include <cuda_bf16.h>
#include <cstring>

namespace tuch {

static inline uint16_t bf16_rn_bits_b(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bf16_bits_to_float_b(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float x;
    memcpy(&x, &u, 4);
    return x;
}

int lbs_tcb_slabs(int V) { return cdiv((long long)V * 3, LBS_TCB_SLAB); }

void lbs_tcb_pack_model(int V, const float* posedirs, std::vector<uint16_t>& blob) {
    const int S = lbs_tcb_slabs(V);
    const size_t V3 = (size_t)V * 3;
    blob.assign((size_t)S * LBS_TCB_KSTEPS * 2 * 3 * 2 * 128 * 8, 0);
    for (int s = 0; s < S; ++s)
        for (int ks = 0; ks < LBS_TCB_KSTEPS; ++ks)
            for (int mb = 0; mb < 2; ++mb)
                for (int ch = 0; ch < 2; ++ch)
                    for (int r = 0; r < 128; ++r) {
                        const int k = mb * 128 + r;
                        if (k >= 207) continue;
                        for (int e = 0; e < 8; ++e) {
                            const size_t c = (size_t)s * LBS_TCB_SLAB + ks * 16 + ch * 8 + e;
                            if (c >= V3) continue;
                            float rem = posedirs[(size_t)k * V3 + c];
                            for (int p = 0; p < 3; ++p) {
                                const uint16_t h = bf16_rn_bits_b(rem);
                                rem -= bf16_bits_to_float_b(h);
                                blob[(((((((size_t)s * LBS_TCB_KSTEPS + ks) * 2 + mb) * 3 + p) * 2 + ch) * 128 + r) * 8) + e] = h;
                            }
                        }
                    }
}

__device__ __forceinline__ uint64_t umma_desc_b(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

constexpr int TCB_THREADS = 192;
constexpr int TCB_STAGES = 4;
constexpr int TCB_A_BYTES = 3 * 2 * 128 * 16;            // 12,288: term x chunk x 128 rows x 16 B (one M block)
constexpr int TCB_B_BYTES = 3 * 2 * LBS_TC_NB * 16;      // 12,288
constexpr int TCB_STAGE_BYTES = TCB_A_BYTES + TCB_B_BYTES;
constexpr int TCB_SMEM_BYTES = TCB_STAGES * TCB_STAGE_BYTES;     // 98,304 -> two CTAs per SM
constexpr int TCB_TMEM_COLS = 128;

__global__ void __launch_bounds__(TCB_THREADS, 2)
lbs_tc_bwd_kernel(const uint16_t* __restrict__ model, const uint16_t* __restrict__ gradop, int B, int n_slabs,
                  float* __restrict__ partial) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_full[TCB_STAGES], bar_empty[TCB_STAGES], bar_acc;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slab = blockIdx.x, mb = blockIdx.y, bt = blockIdx.z;
    const int b0 = bt * LBS_TC_NB;
    const int nb = min(LBS_TC_NB, B - b0);
    const int N = (nb + 15) & ~15;
    const int KS_TOTAL = n_slabs * LBS_TCB_KSTEPS;       // k-steps per body tile in the gradient operand

    if (threadIdx.x == 0) {
        for (int i = 0; i < TCB_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_acc, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TCB_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char* gA = (const unsigned char*)model + ((size_t)slab * LBS_TCB_KSTEPS * 2 + mb) * TCB_A_BYTES;
            const unsigned char* gB = (const unsigned char*)gradop + ((size_t)bt * KS_TOTAL + (size_t)slab * LBS_TCB_KSTEPS) * TCB_B_BYTES;
            for (int s = 0; s < LBS_TCB_KSTEPS; ++s) {
                const int st = s % TCB_STAGES;
                if (s >= TCB_STAGES) mbar_wait(&bar_empty[st], (uint32_t)((s / TCB_STAGES - 1) & 1));
                unsigned char* dst = smem + st * TCB_STAGE_BYTES;
                mbar_expect_tx(&bar_full[st], TCB_STAGE_BYTES);
                tma_load_1d(dst, gA + (size_t)s * 2 * TCB_A_BYTES, TCB_A_BYTES, &bar_full[st]);
                tma_load_1d(dst + TCB_A_BYTES, gB + (size_t)s * TCB_B_BYTES, TCB_B_BYTES, &bar_full[st]);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int s = 0; s < LBS_TCB_KSTEPS; ++s) {
            const int st = s % TCB_STAGES;
            mbar_wait(&bar_full[st], (uint32_t)((s / TCB_STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + st * TCB_STAGE_BYTES), sb = sa + TCB_A_BYTES;
                const int pa[6] = {0, 0, 1, 0, 1, 2}, pb[6] = {0, 1, 0, 2, 1, 0};
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const uint64_t da = umma_desc_b(sa + (uint32_t)(pa[q] * 2 * 128 * 16), 128 * 16, 128);
                    const uint64_t db = umma_desc_b(sb + (uint32_t)(pb[q] * 2 * LBS_TC_NB * 16), LBS_TC_NB * 16, 128);
                    const uint32_t acc = (s > 0 || q > 0) ? 1u : 0u;
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_empty[st])) : "memory");
                if (s == LBS_TCB_KSTEPS - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_acc)) : "memory");
            }
            __syncwarp();
        }
    } else {
        // epilogue: thread = row k of this M block (TMEM lane), loop over the bodies (columns): partial[slab][b][k]
        const int quad = warp & 3;
        const int k = mb * 128 + quad * 32 + lane;
        mbar_wait(&bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int col = 0; col < N; col += 16) {
            uint32_t r[16];
            const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)col;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int b = b0 + col + j;
                if (b < B) partial[((size_t)slab * B + b) * 256 + k] = __uint_as_float(r[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCB_TMEM_COLS));
}

// g_pf[b][k] = sum over the slabs, in index order
__global__ void lbs_tc_bwd_reduce_kernel(const float* __restrict__ partial, int B, int n_slabs, float* __restrict__ g_pf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 207) return;
    const int b = i / 207, k = i - b * 207;
    float x = 0.f;
    for (int s = 0; s < n_slabs; ++s) x += partial[((size_t)s * B + b) * 256 + k];
    g_pf[i] = x;
}

int launch_lbs_tc_bwd(const SmplDev& m, const uint16_t* gradop, int B, float* partial, float* g_pf, cudaStream_t st) {
    static bool attr_set[64] = {false};
    int dev = 0;
    TUCH_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        TUCH_CUDA(cudaFuncSetAttribute(lbs_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM_BYTES));
        attr_set[dev] = true;
    }
    const int S = lbs_tcb_slabs(m.V);
    dim3 grid(S, 2, cdiv(B, LBS_TC_NB));
    lbs_tc_bwd_kernel<<<grid, TCB_THREADS, TCB_SMEM_BYTES, st>>>(m.tcb_model, gradop, B, S, partial);
    TUCH_LAUNCH_CHECK(); count_launch();
    lbs_tc_bwd_reduce_kernel<<<cdiv(B * 207, 256), 256, 0, st>>>(partial, B, S, g_pf);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

}  // namespace tuch
