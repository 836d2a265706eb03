// SMPL blend shapes + pose blend shapes + skinning in ONE tcgen05 kernel (sm_100a).
//
// The two blend-shape contractions of smplx.lbs.lbs -- v_shaped = T + S beta ([B,10] x [10,3V]) and
// v_posed = v_shaped + P^T vec(R_1..23 - I) ([B,207] x [207,3V]) -- are ONE GEMM over the concatenated feature
// vector f_b = [vec(R - I) | beta | 0] (K = 224):  offset[v,c,b] = sum_k M[k,v,c] f[b,k].  It runs on the 5th-gen
// tensor cores with fp32 accuracy by splitting every fp32 operand into three bf16 terms (a = a1 + a2 + a3, 8 + 8
// + 8 mantissa bits) and keeping the six products a_i b_j with i + j <= 4 (what is dropped is below 2^-22 of a
// term, i.e. < 3e-8 m on decimetre-sized blend offsets); accumulation is fp32 in TMEM.
//
// Orientation: the MMA's M dimension is a tile of 128 VERTICES, its N dimension the bodies of the batch, and the
// three coordinates are three accumulators (TMEM columns [0,128) x, [128,256) y, [256,384) z).  So TMEM lane =
// vertex, column = body: an epilogue thread owns one vertex -- its template position and skinning weights live in
// registers -- walks over the bodies, blends the body's joint transforms (staged in shared memory), applies them
// and writes vertices that are CONTIGUOUS across the warp (32 x 12 B per body).
//
// Operands are pre-arranged in global memory in the tensor core's no-swizzle K-major canonical layout
// ([16-byte K chunk][row][8 bf16]: SBO = 128 B between 8-row groups, LBO = rows x 16 B between the two K chunks of
// one K=16 MMA), so a pipeline stage is two plain cp.async.bulk (TMA) copies and no thread touches an operand:
//   model operand  (tuch_smpl_create, once):   [vertex tile][k-step][coord][term][chunk][128 vertices][8]  36,864 B / stage
//   feature operand (lbs_pose_kernel, per call): [body tile][k-step][term][chunk][128 bodies][8]            12,288 B / stage
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue (one elected lane), warps
// 2-9 = epilogue (two warps per TMEM lane quadrant, 64 bodies at a time).  3-stage mbarrier ring.
#include "api_internal.h"
#include "smpl_internal.h"

#include <cuda_bf16.h>

namespace tuch {

// ------------------------------------------------------------------------------------------
// host: fp32 -> three bf16 terms, model operand blob
// ------------------------------------------------------------------------------------------
static inline uint16_t bf16_rn_bits(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);      // inf / nan: truncate
    u += 0x7fffu + ((u >> 16) & 1u);                                         // round to nearest even
    return (uint16_t)(u >> 16);
}
static inline float bf16_bits_to_float(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float x;
    memcpy(&x, &u, 4);
    return x;
}

void lbs_tc_pack_model(int V, int L, const float* shapedirs, const float* posedirs, std::vector<uint16_t>& blob) {
    const int TV = cdiv(V, LBS_TC_M);
    blob.assign((size_t)TV * LBS_TC_KSTEPS * 3 * 3 * 2 * LBS_TC_M * 8, 0);
    const size_t V3 = (size_t)V * 3;
    for (int t = 0; t < TV; ++t)
        for (int s = 0; s < LBS_TC_KSTEPS; ++s)
            for (int c = 0; c < 3; ++c)
                for (int ch = 0; ch < 2; ++ch)
                    for (int r = 0; r < LBS_TC_M; ++r) {
                        const int v = t * LBS_TC_M + r;
                        if (v >= V) continue;
                        for (int e = 0; e < 8; ++e) {
                            const int k = s * 16 + ch * 8 + e;
                            float a = 0.f;
                            if (k < 207) a = posedirs[(size_t)k * V3 + (size_t)v * 3 + c];
                            else if (k < 207 + L) a = shapedirs[((size_t)v * 3 + c) * L + (k - 207)];
                            float rem = a;
                            for (int p = 0; p < 3; ++p) {
                                const uint16_t h = bf16_rn_bits(rem);
                                rem -= bf16_bits_to_float(h);
                                const size_t o = ((((((size_t)t * LBS_TC_KSTEPS + s) * 3 + c) * 3 + p) * 2 + ch) * LBS_TC_M + r) * 8 + e;
                                blob[o] = h;
                            }
                        }
                    }
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                     // descriptor version 1 (sm_100); layout type 0 = no swizzle
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

constexpr int TC_THREADS = 320;                  // 10 warps
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = 3 * 3 * 2 * LBS_TC_M * 16;      // 36,864: coord x term x chunk x 128 rows x 16 B
constexpr int TC_B_BYTES = 3 * 2 * LBS_TC_NB * 16;         // 12,288: term x chunk x 128 rows x 16 B
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;    // 49,152
constexpr int TC_SKIN_BODIES = 64;                         // joint transforms staged per epilogue pass
constexpr int TC_SKIN_BYTES = TC_SKIN_BODIES * 288 * 4;    // 73,728
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_SKIN_BYTES;    // 221,184
constexpr int TC_TMEM_COLS = 512;

__global__ void __launch_bounds__(TC_THREADS, 1)
lbs_skin_tc_kernel(SmplDev m, const uint16_t* __restrict__ featop, const float* __restrict__ A, int B,
                   float* __restrict__ verts, float* __restrict__ v_posed_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_acc;
    __shared__ uint32_t s_tmem;
    unsigned char* s_skin = smem + TC_STAGES * TC_STAGE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x;                    // vertex tile
    const int bt = blockIdx.y;                   // body tile
    const int b0 = bt * LBS_TC_NB;
    const int nb = min(LBS_TC_NB, B - b0);       // live bodies of this CTA
    const int N = (nb + 15) & ~15;               // MMA N (multiple of 16 for M = 128)

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_acc, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ===== TMA producer: stage s = k-step s of the model operand (this vertex tile) + feature operand (this body tile)
        if (lane == 0) {
            const unsigned char* gA = (const unsigned char*)m.tc_model + (size_t)t * LBS_TC_KSTEPS * TC_A_BYTES;
            const unsigned char* gB = (const unsigned char*)featop + (size_t)bt * LBS_TC_KSTEPS * TC_B_BYTES;
            for (int s = 0; s < LBS_TC_KSTEPS; ++s) {
                const int st = s % TC_STAGES;
                if (s >= TC_STAGES) mbar_wait(&bar_empty[st], (uint32_t)((s / TC_STAGES - 1) & 1));
                unsigned char* dst = smem + st * TC_STAGE_BYTES;
                mbar_expect_tx(&bar_full[st], TC_STAGE_BYTES);
                tma_load_1d(dst, gA + (size_t)s * TC_A_BYTES, TC_A_BYTES, &bar_full[st]);
                tma_load_1d(dst + TC_A_BYTES, gB + (size_t)s * TC_B_BYTES, TC_B_BYTES, &bar_full[st]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: 3 coordinates x 6 term products per k-step, fp32 accumulators in TMEM
        // instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(LBS_TC_M >> 4) << 24);
        for (int s = 0; s < LBS_TC_KSTEPS; ++s) {
            const int st = s % TC_STAGES;
            mbar_wait(&bar_full[st], (uint32_t)((s / TC_STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + st * TC_STAGE_BYTES), sb = sa + TC_A_BYTES;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // the six products a_i b_j with i + j <= 4, largest first
                    const int pa[6] = {0, 0, 1, 0, 1, 2}, pb[6] = {0, 1, 0, 2, 1, 0};
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const uint64_t da = umma_desc(sa + (uint32_t)((c * 3 + pa[q]) * 2 * LBS_TC_M * 16), LBS_TC_M * 16, 128);
                        const uint64_t db = umma_desc(sb + (uint32_t)(pb[q] * 2 * LBS_TC_NB * 16), LBS_TC_NB * 16, 128);
                        umma_bf16(tmem + (uint32_t)(c * LBS_TC_NB), da, db, idesc, (s > 0 || q > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&bar_empty[st]);                       // the stage may be refilled once these MMAs retire
                if (s == LBS_TC_KSTEPS - 1) umma_commit(&bar_acc);  // accumulators complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: thread = vertex (TMEM lane), loop over bodies (TMEM columns)
        const int ew = warp - 2;                     // 0..7
        const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
        const int sub = ew >> 2;                     // which 32 of the 64 staged bodies
        const int et = threadIdx.x - 64;             // 0..255 among the epilogue threads
        const int v = t * LBS_TC_M + quad * 32 + lane;
        const bool vlive = v < m.V;
        const int K = m.K;
        float tx = 0.f, ty = 0.f, tz = 0.f;
        float w[LBS_TC_MAXK];
        int ji[LBS_TC_MAXK];
#pragma unroll
        for (int k = 0; k < LBS_TC_MAXK; ++k) { w[k] = 0.f; ji[k] = 0; }
        if (vlive) {
            tx = m.v_template[3 * v]; ty = m.v_template[3 * v + 1]; tz = m.v_template[3 * v + 2];
#pragma unroll
            for (int k = 0; k < LBS_TC_MAXK; ++k)
                if (k < K) { w[k] = m.skin_w[(size_t)v * K + k]; ji[k] = 12 * (int)m.skin_idx[(size_t)v * K + k]; }
        }
        const float4* s_skin4 = reinterpret_cast<const float4*>(s_skin);
        for (int h = 0; h * TC_SKIN_BODIES < nb; ++h) {
            // stage the joint transforms A[b][24][12] of 64 bodies (rows past the batch: zeros)
            if (h > 0) asm volatile("bar.sync 1, 256;" ::: "memory");       // everyone is done with the previous 64
            {
                const int first = b0 + h * TC_SKIN_BODIES;
                const int live = min(TC_SKIN_BODIES, B - first);
                const float4* src = reinterpret_cast<const float4*>(A + (size_t)first * 288);
                float4* dst = reinterpret_cast<float4*>(s_skin);
                for (int i = et; i < TC_SKIN_BODIES * 72; i += 256)
                    dst[i] = (i < live * 72) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (h == 0) {
                mbar_wait(&bar_acc, 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
#pragma unroll 1
            for (int ck = 0; ck < 2; ++ck) {
                const int col = h * TC_SKIN_BODIES + sub * 32 + ck * 16;      // first body (in the tile) of this chunk
                if (col >= N) break;                                          // warp-uniform
                uint32_t rx[16], ry[16], rz[16];
                const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)col;
                tmem_ld16(ta, rx);
                tmem_ld16(ta + LBS_TC_NB, ry);
                tmem_ld16(ta + 2 * LBS_TC_NB, rz);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = b0 + col + j;
                    if (b >= B || !vlive) continue;
                    const float x = tx + __uint_as_float(rx[j]), y = ty + __uint_as_float(ry[j]), z = tz + __uint_as_float(rz[j]);
                    const float4* sa = s_skin4 + (size_t)(sub * 32 + ck * 16 + j) * 72;
                    float4 T0 = make_float4(0.f, 0.f, 0.f, 0.f), T1 = T0, T2 = T0;
#pragma unroll
                    for (int k = 0; k < LBS_TC_MAXK; ++k) {
                        if (k < K) {
                            const float4* a = sa + (ji[k] >> 2);
                            const float4 a0 = a[0], a1 = a[1], a2 = a[2];
                            const float wk = w[k];
                            T0.x = fmaf(wk, a0.x, T0.x); T0.y = fmaf(wk, a0.y, T0.y); T0.z = fmaf(wk, a0.z, T0.z); T0.w = fmaf(wk, a0.w, T0.w);
                            T1.x = fmaf(wk, a1.x, T1.x); T1.y = fmaf(wk, a1.y, T1.y); T1.z = fmaf(wk, a1.z, T1.z); T1.w = fmaf(wk, a1.w, T1.w);
                            T2.x = fmaf(wk, a2.x, T2.x); T2.y = fmaf(wk, a2.y, T2.y); T2.z = fmaf(wk, a2.z, T2.z); T2.w = fmaf(wk, a2.w, T2.w);
                        }
                    }
                    float* o = verts + ((size_t)b * m.V + v) * 3;
                    o[0] = fmaf(T0.z, z, fmaf(T0.y, y, fmaf(T0.x, x, T0.w)));
                    o[1] = fmaf(T1.z, z, fmaf(T1.y, y, fmaf(T1.x, x, T1.w)));
                    o[2] = fmaf(T2.z, z, fmaf(T2.y, y, fmaf(T2.x, x, T2.w)));
                    if (v_posed_out != nullptr) {
                        float* p = v_posed_out + ((size_t)b * m.V + v) * 3;
                        p[0] = x; p[1] = y; p[2] = z;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

int launch_lbs_skin_tc(const SmplDev& m, const uint16_t* featop, const float* A, int B, float* verts, float* v_posed,
                       cudaStream_t st) {
    static bool attr_set[64] = {false};
    int dev = 0;
    TUCH_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        TUCH_CUDA(cudaFuncSetAttribute(lbs_skin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set[dev] = true;
    }
    dim3 grid(cdiv(m.V, LBS_TC_M), cdiv(B, LBS_TC_NB));
    lbs_skin_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(m, featop, A, B, verts, v_posed);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

}  // namespace tuch
