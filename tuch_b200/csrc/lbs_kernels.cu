// SMPL linear blend skinning, forward and backward (sm_100a).
//
// Replaces the op chain the reference obtains from smplx==0.1.13 `lbs()` plus its own wrapper
// tuch/models/smpl.py:44-56 (~60-80 ATen/cuBLAS launches incl. a 23-step Python kinematic chain)
// with three launches forward (pose chain | blend shapes + skinning | joints) and four backward.
//
//   v_shaped = T + S beta            J = J_T + J_S beta   (J_T = Jreg T, J_S = Jreg S precomputed)
//   R_k = rodrigues(theta_k)         pose_feature = vec(R_1..23 - I)
//   v_posed = v_shaped + P^T pose_feature
//   G_k = G_parent(k) [R_k | J_k - J_parent(k)]           A_k = [G3_k | Gt_k - G3_k J_k]
//   v = (sum_k w_vk A_k) [v_posed; 1]
//   joints54 = [Gt_0..23 | v[picked 21] | J_extra v]      joints49 = joints54[joint_map]
#include "api_internal.h"
#include "smpl_internal.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

namespace tuch {


// ------------------------------------------------------------------------------------------
// small 3x3 helpers (row-major)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* c) {   // c = a b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            c[3 * i + j] = fmaf(a[3 * i + 2], b[6 + j], fmaf(a[3 * i + 1], b[3 + j], a[3 * i] * b[j]));
}
__device__ __forceinline__ void mat3_vec(const float* a, const float* v, float* o) {   // o = a v
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = fmaf(a[3 * i + 2], v[2], fmaf(a[3 * i + 1], v[1], a[3 * i] * v[0]));
}
__device__ __forceinline__ void mat3T_vec(const float* a, const float* v, float* o) {  // o = a^T v
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = fmaf(a[6 + i], v[2], fmaf(a[3 + i], v[1], a[i] * v[0]));
}

// smplx.lbs.batch_rodrigues: angle = |r + 1e-8|, axis = r / angle, R = I + sin K + (1 - cos) K K
__device__ __forceinline__ void rodrigues_fwd(const float* r, float* R) {
    const float e0 = r[0] + 1e-8f, e1 = r[1] + 1e-8f, e2 = r[2] + 1e-8f;
    const float th = sqrtf(fmaf(e2, e2, fmaf(e1, e1, e0 * e0)));
    const float x = r[0] / th, y = r[1] / th, z = r[2] / th;
    float s, c;
    sincosf(th, &s, &c);
    const float K[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
    float K2[9];
    mat3_mul(K, K, K2);
    const float oc = 1.f - c;
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = fmaf(oc, K2[i], s * K[i]);
    R[0] += 1.f; R[4] += 1.f; R[8] += 1.f;
}

// gradient of rodrigues_fwd: gR[9] -> gr[3]
__device__ __forceinline__ void rodrigues_bwd(const float* r, const float* gR, float* gr) {
    const float e0 = r[0] + 1e-8f, e1 = r[1] + 1e-8f, e2 = r[2] + 1e-8f;
    const float th = sqrtf(fmaf(e2, e2, fmaf(e1, e1, e0 * e0)));
    const float inv = 1.f / th;
    const float x = r[0] * inv, y = r[1] * inv, z = r[2] * inv;
    float s, c;
    sincosf(th, &s, &c);
    const float K[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
    float K2[9];
    mat3_mul(K, K, K2);
    float gs = 0.f, goc = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { gs = fmaf(gR[i], K[i], gs); goc = fmaf(gR[i], K2[i], goc); }
    float gth = gs * c + goc * s;
    // gK = s gR + (1-c) (gR K^T + K^T gR)
    const float oc = 1.f - c;
    float KT[9] = {K[0], K[3], K[6], K[1], K[4], K[7], K[2], K[5], K[8]};
    float t1[9], t2[9], gK[9];
    mat3_mul(gR, KT, t1);
    mat3_mul(KT, gR, t2);
#pragma unroll
    for (int i = 0; i < 9; ++i) gK[i] = fmaf(oc, t1[i] + t2[i], s * gR[i]);
    const float gx = gK[7] - gK[5], gy = gK[2] - gK[6], gz = gK[3] - gK[1];
    // axis = r / th
    gth -= (gx * r[0] + gy * r[1] + gz * r[2]) * inv * inv;
    gr[0] = fmaf(gth, e0 * inv, gx * inv);
    gr[1] = fmaf(gth, e1 * inv, gy * inv);
    gr[2] = fmaf(gth, e2 * inv, gz * inv);
}

// ------------------------------------------------------------------------------------------
// forward 1/3: pose chain.  One warp per body, lane k < 24 owns joint k.
// ------------------------------------------------------------------------------------------
constexpr int POSE_WARPS = 4;

__global__ void __launch_bounds__(POSE_WARPS * 32)
lbs_pose_kernel(SmplDev m, const float* __restrict__ betas, const float* __restrict__ pose,
                const float* __restrict__ orient, int pose_is_rotmat,
                int B, float* __restrict__ Rout, float* __restrict__ Jrest, float* __restrict__ G,
                float* __restrict__ A, float* __restrict__ pf, uint16_t* __restrict__ featop, int* __restrict__ step_a,
                int* __restrict__ step_b) {
    // `orient` != NULL: split axis-angle pose -- joint 0 from orient[B,3], joints 1..23 from pose[B,69]
    // (the two parameter tensors of SMPLify-DC's stage 2, smplifydc.py:149) instead of one pose[B,72]
    __shared__ float sG[POSE_WARPS][24][12];
    __shared__ float sJ[POSE_WARPS][24][3];
    const int w = threadIdx.x / 32, lane = threadIdx.x & 31;
    const int b = blockIdx.x * POSE_WARPS + w;
    const bool act = (b < B) && lane < 24;
    // the fused Adam step at the end of the iteration (lbs_bwd_chain_kernel) reads the step counters this first
    // kernel of the iteration advances: stream order makes the increment visible to every later kernel
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (step_a != nullptr) *step_a += 1;
        if (step_b != nullptr) *step_b += 1;
    }
    // feature operand of lbs_tc.cu: feature k of body b as three bf16 terms at
    // [b / 128][k / 16][term][(k % 16) / 8][b % 128][k % 8]; rows past the batch and the K padding are zero
    auto put_feature = [&](int k, float a) {
        const size_t base = ((size_t)(b / LBS_TC_NB) * LBS_TC_KSTEPS + k / 16) * 3;
        const size_t tail = ((size_t)((k % 16) / 8) * LBS_TC_NB + b % LBS_TC_NB) * 8 + k % 8;
        float rem = a;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const __nv_bfloat16 h = __float2bfloat16_rn(rem);
            rem -= __bfloat162float(h);
            featop[(base + p) * 2 * LBS_TC_NB * 8 + tail] = __bfloat16_as_ushort(h);
        }
    };
    if (featop != nullptr && b >= B) {                   // padding bodies of the last 128-body tile
        for (int k = lane; k < LBS_TC_K; k += 32) put_feature(k, 0.f);
        return;
    }
    float R[9], J[3] = {0.f, 0.f, 0.f}, Gk[12];
    if (act) {
        if (pose_is_rotmat) {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = pose[((size_t)b * 24 + lane) * 9 + i];
        } else if (orient != nullptr) {
            rodrigues_fwd(lane == 0 ? orient + (size_t)b * 3 : pose + (size_t)b * 69 + (lane - 1) * 3, R);
        } else {
            rodrigues_fwd(pose + ((size_t)b * 24 + lane) * 3, R);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float acc = m.J_template[3 * lane + c];
            const float* js = m.J_shapedirs + (size_t)(3 * lane + c) * m.L;
            for (int l = 0; l < m.L; ++l) acc = fmaf(js[l], betas[(size_t)b * m.L + l], acc);
            J[c] = acc;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) Rout[((size_t)b * 24 + lane) * 9 + i] = R[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) Jrest[((size_t)b * 24 + lane) * 3 + c] = J[c];
        if (lane >= 1) {
            float* o = pf + (size_t)b * 207 + (lane - 1) * 9;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const float f = R[i] - ((i == 0 || i == 4 || i == 8) ? 1.f : 0.f);
                o[i] = f;
                if (featop != nullptr) put_feature((lane - 1) * 9 + i, f);
            }
        }
    }
    if (featop != nullptr && b < B)                      // betas, then the K padding
        for (int k = 207 + lane; k < LBS_TC_K; k += 32) put_feature(k, k - 207 < m.L ? betas[(size_t)b * m.L + (k - 207)] : 0.f);
    const int par = lane < 24 ? m.parents[lane] : -1;
    const int dep = lane < 24 ? m.depth[lane] : -1;
    if (act) { sJ[w][lane][0] = J[0]; sJ[w][lane][1] = J[1]; sJ[w][lane][2] = J[2]; }
    __syncwarp();
    float Jp[3] = {0.f, 0.f, 0.f};
    if (act && par >= 0) { Jp[0] = sJ[w][par][0]; Jp[1] = sJ[w][par][1]; Jp[2] = sJ[w][par][2]; }
    for (int level = 0; level < 24; ++level) {
        if (act && dep == level) {
            const float t[3] = {J[0] - Jp[0], J[1] - Jp[1], J[2] - Jp[2]};
            if (par < 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    Gk[4 * i] = R[3 * i]; Gk[4 * i + 1] = R[3 * i + 1]; Gk[4 * i + 2] = R[3 * i + 2]; Gk[4 * i + 3] = t[i];
                }
            } else {
                const float* Gp = sG[w][par];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        Gk[4 * i + j] = fmaf(Gp[4 * i + 2], R[6 + j], fmaf(Gp[4 * i + 1], R[3 + j], Gp[4 * i] * R[j]));
                    Gk[4 * i + 3] = fmaf(Gp[4 * i + 2], t[2], fmaf(Gp[4 * i + 1], t[1], fmaf(Gp[4 * i], t[0], Gp[4 * i + 3])));
                }
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) sG[w][lane][i] = Gk[i];
        }
        __syncwarp();
    }
    if (act) {
        float* g = G + ((size_t)b * 24 + lane) * 12;
        float* a = A + ((size_t)b * 24 + lane) * 12;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float corr = fmaf(Gk[4 * i + 2], J[2], fmaf(Gk[4 * i + 1], J[1], Gk[4 * i] * J[0]));
#pragma unroll
            for (int j = 0; j < 4; ++j) g[4 * i + j] = Gk[4 * i + j];
            a[4 * i] = Gk[4 * i]; a[4 * i + 1] = Gk[4 * i + 1]; a[4 * i + 2] = Gk[4 * i + 2];
            a[4 * i + 3] = Gk[4 * i + 3] - corr;
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward 2/3: blend shapes + skinning.  Block = LBS_VT vertices x LBS_NB bodies; every thread
// owns one vertex and LBS_NB bodies so each posedirs / shapedirs value read feeds LBS_NB FMAs.
// ------------------------------------------------------------------------------------------
constexpr int LBS_VT = 128;
constexpr int LBS_NB = 8;

__global__ void __launch_bounds__(LBS_VT)
lbs_skin_kernel(SmplDev m, const float* __restrict__ betas, const float* __restrict__ pf,
                const float* __restrict__ A, int B, float* __restrict__ verts, float* __restrict__ v_posed_out) {
    __shared__ float s_pf[207][LBS_NB];
    __shared__ float s_A[LBS_NB][24 * 12];
    __shared__ float s_beta[LBS_NB][SMPL_MAX_BETAS];
    const int b0 = blockIdx.y * LBS_NB;
    for (int i = threadIdx.x; i < 207 * LBS_NB; i += LBS_VT) {
        const int k = i / LBS_NB, bb = i % LBS_NB;
        s_pf[k][bb] = (b0 + bb < B) ? pf[(size_t)(b0 + bb) * 207 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < LBS_NB * 288; i += LBS_VT) {
        const int bb = i / 288, e = i % 288;
        s_A[bb][e] = (b0 + bb < B) ? A[(size_t)(b0 + bb) * 288 + e] : 0.f;
    }
    for (int i = threadIdx.x; i < LBS_NB * m.L; i += LBS_VT) {
        const int bb = i / m.L, l = i % m.L;
        s_beta[bb][l] = (b0 + bb < B) ? betas[(size_t)(b0 + bb) * m.L + l] : 0.f;
    }
    __syncthreads();
    const int v = blockIdx.x * LBS_VT + threadIdx.x;
    if (v >= m.V) return;
    const size_t V3 = (size_t)m.V * 3;
    float acc[LBS_NB][3];
    {
        const float t0 = m.v_template[3 * v], t1 = m.v_template[3 * v + 1], t2 = m.v_template[3 * v + 2];
#pragma unroll
        for (int bb = 0; bb < LBS_NB; ++bb) { acc[bb][0] = t0; acc[bb][1] = t1; acc[bb][2] = t2; }
    }
    for (int l = 0; l < m.L; ++l) {
        const float* row = m.shapedirsT + (size_t)l * V3 + 3 * v;
        const float s0 = __ldg(row), s1 = __ldg(row + 1), s2 = __ldg(row + 2);
#pragma unroll
        for (int bb = 0; bb < LBS_NB; ++bb) {
            const float be = s_beta[bb][l];
            acc[bb][0] = fmaf(s0, be, acc[bb][0]);
            acc[bb][1] = fmaf(s1, be, acc[bb][1]);
            acc[bb][2] = fmaf(s2, be, acc[bb][2]);
        }
    }
#pragma unroll 3
    for (int k = 0; k < 207; ++k) {
        const float* row = m.posedirs + (size_t)k * V3 + 3 * v;
        const float p0 = __ldg(row), p1 = __ldg(row + 1), p2 = __ldg(row + 2);
#pragma unroll
        for (int bb = 0; bb < LBS_NB; ++bb) {
            const float f = s_pf[k][bb];
            acc[bb][0] = fmaf(p0, f, acc[bb][0]);
            acc[bb][1] = fmaf(p1, f, acc[bb][1]);
            acc[bb][2] = fmaf(p2, f, acc[bb][2]);
        }
    }
    const int K = m.K;
    const uint8_t* si = m.skin_idx + (size_t)v * K;
    const float* sw = m.skin_w + (size_t)v * K;
#pragma unroll
    for (int bb = 0; bb < LBS_NB; ++bb) {
        if (b0 + bb >= B) break;
        float T[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = 0.f;
        for (int k = 0; k < K; ++k) {
            const float wk = sw[k];
            const float* a = s_A[bb] + 12 * (int)si[k];
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = fmaf(wk, a[i], T[i]);
        }
        const float x = acc[bb][0], y = acc[bb][1], z = acc[bb][2];
        float* o = verts + ((size_t)(b0 + bb) * m.V + v) * 3;
        o[0] = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
        o[1] = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
        o[2] = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
        if (v_posed_out != nullptr) {
            float* p = v_posed_out + ((size_t)(b0 + bb) * m.V + v) * 3;
            p[0] = x; p[1] = y; p[2] = z;
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward 3/3: 54 joints (24 posed + picked vertices + regressed extras) -> joint_map (smpl.py:47-49)
// One block per body; warp e reduces extra regressor e over its non-zeros.
// ------------------------------------------------------------------------------------------
constexpr int JOINT_THREADS = 256;

__global__ void __launch_bounds__(JOINT_THREADS)
lbs_joints_kernel(SmplDev m, const float* __restrict__ verts, const float* __restrict__ G,
                  float* __restrict__ joints_out) {
    __shared__ float s_j[SMPL_MAX_JOINTS54][3];
    const int b = blockIdx.x;
    const float* vb = verts + (size_t)b * m.V * 3;
    for (int j = threadIdx.x; j < 24 + m.NX; j += JOINT_THREADS) {
        if (j < 24) {
            const float* g = G + ((size_t)b * 24 + j) * 12;
            s_j[j][0] = g[3]; s_j[j][1] = g[7]; s_j[j][2] = g[11];
        } else {
            const int v = m.extra_vertex_ids[j - 24];
            s_j[j][0] = vb[3 * v]; s_j[j][1] = vb[3 * v + 1]; s_j[j][2] = vb[3 * v + 2];
        }
    }
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
    for (int e = warp; e < m.NE; e += JOINT_THREADS / 32) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int i = m.ex_off[e] + lane; i < m.ex_off[e + 1]; i += 32) {
            const int v = m.ex_vert[i];
            const float w = m.ex_w[i];
            a0 = fmaf(w, vb[3 * v], a0); a1 = fmaf(w, vb[3 * v + 1], a1); a2 = fmaf(w, vb[3 * v + 2], a2);
        }
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) { s_j[24 + m.NX + e][0] = a0; s_j[24 + m.NX + e][1] = a1; s_j[24 + m.NX + e][2] = a2; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m.NO * 3; i += JOINT_THREADS) {
        const int o = i / 3, c = i % 3;
        joints_out[((size_t)b * m.NO + o) * 3 + c] = s_j[m.joint_map[o]][c];
    }
}

// ------------------------------------------------------------------------------------------
// backward 1/4: per vertex.  g = gV + sum over the vertex's joint contributions (picked joints,
// extra regressors); g_vposed = T3^T g.  Writes the combined g (for the per-joint pass) and
// g_vposed (for the blend-shape contractions).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LBS_VT)
lbs_bwd_vertex_kernel(SmplDev m, const float* __restrict__ A, const float* __restrict__ gV,
                      const float* __restrict__ gJ49, int B, float* __restrict__ g_comb,
                      float* __restrict__ g_vposed, float* __restrict__ beta_part, uint16_t* __restrict__ gradop,
                      int ks_total) {
    // gradop != NULL: g_vposed goes out as the bf16 x 3 operand of the tensor-core contraction (lbs_tc_bwd.cu),
    // [body tile][c / 16][term][(c % 16) / 8][body % 128][c % 8], instead of the fp32 array
    // beta_part != NULL: also the vertex part of the shape gradient, g_beta[b][l] = sum_{v,c} S[l][v,c] g_vposed[b][v,c],
    // as one partial sum per (vertex tile, body, l) -- beta_part[tile][B][L], summed over the tiles in index order
    // by lbs_bwd_chain_kernel -- instead of a second pass over g_vposed
    __shared__ float s_A[LBS_NB][24 * 12];
    __shared__ float s_gj[LBS_NB][SMPL_MAX_JOINTS54][3];
    __shared__ float s_beta[LBS_VT / 32][LBS_NB][SMPL_MAX_BETAS];
    const int b0 = blockIdx.y * LBS_NB;
    for (int i = threadIdx.x; i < LBS_NB * 288; i += LBS_VT) {
        const int bb = i / 288, e = i % 288;
        s_A[bb][e] = (b0 + bb < B) ? A[(size_t)(b0 + bb) * 288 + e] : 0.f;
    }
    for (int i = threadIdx.x; i < LBS_NB * SMPL_MAX_JOINTS54 * 3; i += LBS_VT) (&s_gj[0][0][0])[i] = 0.f;
    __syncthreads();
    if (gJ49 != nullptr) {
        // scatter-add joints49 -> joints54 (several outputs may alias one source joint)
        if (threadIdx.x < LBS_NB * 3) {
            const int bb = threadIdx.x / 3, c = threadIdx.x % 3;
            if (b0 + bb < B)
                for (int o = 0; o < m.NO; ++o)
                    s_gj[bb][m.joint_map[o]][c] += gJ49[((size_t)(b0 + bb) * m.NO + o) * 3 + c];
        }
        __syncthreads();
    }
    const int v = blockIdx.x * LBS_VT + threadIdx.x;
    const bool live = v < m.V;
    const int K = m.K;
    float gp[LBS_NB][3];
#pragma unroll
    for (int bb = 0; bb < LBS_NB; ++bb) { gp[bb][0] = 0.f; gp[bb][1] = 0.f; gp[bb][2] = 0.f; }
    if (live) {
        const uint8_t* si = m.skin_idx + (size_t)v * K;
        const float* sw = m.skin_w + (size_t)v * K;
        const int c0 = m.vj_off[v], c1 = m.vj_off[v + 1];
#pragma unroll
        for (int bb = 0; bb < LBS_NB; ++bb) {
            if (b0 + bb >= B) break;
            const size_t o = ((size_t)(b0 + bb) * m.V + v) * 3;
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (gV != nullptr) { g0 = gV[o]; g1 = gV[o + 1]; g2 = gV[o + 2]; }
            for (int i = c0; i < c1; ++i) {
                const float w = m.vj_w[i];
                const float* gj = s_gj[bb][m.vj_joint[i]];
                g0 = fmaf(w, gj[0], g0); g1 = fmaf(w, gj[1], g1); g2 = fmaf(w, gj[2], g2);
            }
            g_comb[o] = g0; g_comb[o + 1] = g1; g_comb[o + 2] = g2;
            // most vertices carry no gradient (the contact terms touch the interior / in-contact vertices and their
            // partners, the joints a few hundred vertices): their blended transform is never formed
            if (g0 != 0.f || g1 != 0.f || g2 != 0.f) {
                float T[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) T[i] = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float wk = sw[k];
                    const float* a = s_A[bb] + 12 * (int)si[k];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        T[3 * r] = fmaf(wk, a[4 * r], T[3 * r]);
                        T[3 * r + 1] = fmaf(wk, a[4 * r + 1], T[3 * r + 1]);
                        T[3 * r + 2] = fmaf(wk, a[4 * r + 2], T[3 * r + 2]);
                    }
                }
                gp[bb][0] = fmaf(T[6], g2, fmaf(T[3], g1, T[0] * g0));
                gp[bb][1] = fmaf(T[7], g2, fmaf(T[4], g1, T[1] * g0));
                gp[bb][2] = fmaf(T[8], g2, fmaf(T[5], g1, T[2] * g0));
            }
            if (gradop == nullptr) {
                g_vposed[o] = gp[bb][0]; g_vposed[o + 1] = gp[bb][1]; g_vposed[o + 2] = gp[bb][2];
            }
        }
    }
    if (gradop != nullptr) {
        // the tile's 384 coordinates x 8 bodies as bf16 x 3, staged in shared memory so that they leave as 16-byte
        // rows of the operand layout (one row = 8 consecutive coordinates of one body and term) instead of 72
        // two-byte stores per thread; the tile starts at a multiple of 16 coordinates, so rows never straddle tiles
        __shared__ __align__(16) uint16_t s_op[3][LBS_NB][3 * LBS_VT];
#pragma unroll
        for (int bb = 0; bb < LBS_NB; ++bb)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float rem = gp[bb][i];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const __nv_bfloat16 h = __float2bfloat16_rn(rem);
                    rem -= __bfloat162float(h);
                    s_op[p][bb][3 * threadIdx.x + i] = __bfloat16_as_ushort(h);
                }
            }
        __syncthreads();
        const int c_tile = blockIdx.x * 3 * LBS_VT;                 // first coordinate of the tile
        const int n_rows = 3 * LBS_NB * (3 * LBS_VT / 8);           // (term, body, chunk of 8 coordinates)
        for (int r = threadIdx.x; r < n_rows; r += LBS_VT) {
            const int ch8 = r % (3 * LBS_VT / 8), bb = (r / (3 * LBS_VT / 8)) % LBS_NB, p = r / (3 * LBS_VT / 8 * LBS_NB);
            const int b = b0 + bb, c = c_tile + 8 * ch8;
            if (b >= B || c / 16 >= ks_total) continue;
            const size_t row = ((((size_t)(b / LBS_TC_NB) * ks_total + c / 16) * 3 + p) * 2 + (c % 16) / 8) * LBS_TC_NB + b % LBS_TC_NB;
            *reinterpret_cast<uint4*>(gradop + row * 8) = *reinterpret_cast<const uint4*>(&s_op[p][bb][8 * ch8]);
        }
    }
    if (beta_part == nullptr) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t V3 = (size_t)m.V * 3;
    for (int l = 0; l < m.L; ++l) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (live) {
            const float* row = m.shapedirsT + (size_t)l * V3 + 3 * v;
            s0 = __ldg(row); s1 = __ldg(row + 1); s2 = __ldg(row + 2);
        }
#pragma unroll
        for (int bb = 0; bb < LBS_NB; ++bb) {
            float x = fmaf(s2, gp[bb][2], fmaf(s1, gp[bb][1], s0 * gp[bb][0]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) s_beta[warp][bb][l] = x;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LBS_NB * m.L; i += LBS_VT) {
        const int bb = i / m.L, l = i % m.L;
        if (b0 + bb >= B) continue;
        float x = 0.f;
#pragma unroll
        for (int w = 0; w < LBS_VT / 32; ++w) x += s_beta[w][bb][l];
        beta_part[((size_t)blockIdx.x * B + b0 + bb) * m.L + l] = x;
    }
}

// ------------------------------------------------------------------------------------------
// backward 2/4: the pose-blend contraction over the 3V coordinates,  g_pf[b][r] = sum_c posedirs[r][c] g_vposed[b][c]
// (the shape-basis contraction is folded into the vertex pass above).
// ------------------------------------------------------------------------------------------
// A [B, 3V] x [3V, 207] GEMM, K = 20,670: a CTA owns
// a 64-body x 64-row output tile over one K split, streams 32-wide K slabs of both operands through shared
// memory ([k][row] so that a thread reads its 8 rows / 4 bodies as vectors) and keeps a 4 x 8 register
// tile; the K splits are summed in a fixed order by lbs_contract_reduce_kernel.  Every operand
// element is read 4 times from L2 instead of 32 (M) / 26 (G).
constexpr int GT_B = 64, GT_R = 64, GT_K = 32, GT_THREADS = 128;

__global__ void __launch_bounds__(GT_THREADS, 6)
lbs_contract_tiled_kernel(const float* __restrict__ M, const float* __restrict__ G, int n_rows, int n_coords, int B,
                          int k_per_split, float* __restrict__ partial) {
    __shared__ __align__(16) float sM[GT_K][GT_R + 4];
    __shared__ __align__(16) float sG[GT_K][GT_B + 4];
    const int r0 = blockIdx.x * GT_R, b0 = blockIdx.y * GT_B, split = blockIdx.z;
    const int k0 = split * k_per_split, k1 = min(n_coords, k0 + k_per_split);
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;         // rows tx*8.., bodies ty*4..
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int kk = k0; kk < k1; kk += GT_K) {
#pragma unroll 8
        for (int i = 0; i < GT_R * GT_K / GT_THREADS; ++i) {
            const int e = threadIdx.x + GT_THREADS * i, row = e / GT_K, k = e % GT_K;
            sM[k][row] = (r0 + row < n_rows && kk + k < k1) ? __ldg(M + (size_t)(r0 + row) * n_coords + kk + k) : 0.f;
        }
#pragma unroll 8
        for (int i = 0; i < GT_B * GT_K / GT_THREADS; ++i) {
            const int e = threadIdx.x + GT_THREADS * i, bb = e / GT_K, k = e % GT_K;
            sG[k][bb] = (b0 + bb < B && kk + k < k1) ? G[(size_t)(b0 + bb) * n_coords + kk + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < GT_K; ++k) {
            const float4 m0 = *reinterpret_cast<const float4*>(&sM[k][tx * 8]);
            const float4 m1 = *reinterpret_cast<const float4*>(&sM[k][tx * 8 + 4]);
            const float4 g = *reinterpret_cast<const float4*>(&sG[k][ty * 4]);
            const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(gv[i], mv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int bb = b0 + ty * 4 + i, r = r0 + tx * 8 + j;
            if (bb < B && r < n_rows) partial[((size_t)split * B + bb) * n_rows + r] = acc[i][j];
        }
}

__global__ void lbs_contract_reduce_kernel(const float* __restrict__ partial, int n, int S, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = 0.f;
    for (int s = 0; s < S; ++s) x += partial[(size_t)s * n + i];
    out[i] = x;
}

// ------------------------------------------------------------------------------------------
// backward 3/4: per (body, joint) reduction gA[b][k] = sum_{v in list(k)} w_vk g_v [v_posed; 1]^T.
// One CTA per (body, joint), no atomics (deterministic).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lbs_bwd_joint_kernel(SmplDev m, const float* __restrict__ g_comb, const float* __restrict__ v_posed, int B,
                     float* __restrict__ gA) {
    // one CTA of four warps per (body, joint): the longest lists (torso joints, ~3k entries) were the critical
    // path of the whole backward with one warp each; the four partial sums meet in warp order (deterministic,
    // independent of the batch)
    __shared__ float s_part[4][12];
    const int b = blockIdx.x / 24, k = blockIdx.x % 24;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
    const float* gb = g_comb + (size_t)b * m.V * 3;
    const float* pb = v_posed + (size_t)b * m.V * 3;
    const int i1 = m.jl_off[k + 1];
#pragma unroll 2
    for (int i = m.jl_off[k] + threadIdx.x; i < i1; i += 128) {
        const int v = m.jl_vert[i];
        const float w = m.jl_w[i];
        const float g[3] = {w * gb[3 * v], w * gb[3 * v + 1], w * gb[3 * v + 2]};
        const float p[3] = {pb[3 * v], pb[3 * v + 1], pb[3 * v + 2]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            acc[4 * r] = fmaf(g[r], p[0], acc[4 * r]);
            acc[4 * r + 1] = fmaf(g[r], p[1], acc[4 * r + 1]);
            acc[4 * r + 2] = fmaf(g[r], p[2], acc[4 * r + 2]);
            acc[4 * r + 3] += g[r];
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        if (lane == 0) s_part[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 12)
        gA[((size_t)b * 24 + k) * 12 + threadIdx.x] =
            ((s_part[0][threadIdx.x] + s_part[1][threadIdx.x]) + s_part[2][threadIdx.x]) + s_part[3][threadIdx.x];
}

// ------------------------------------------------------------------------------------------
// backward 4/4: kinematic chain + Rodrigues.  One warp per body, lane k owns joint k; children
// hand their contribution to the parent through shared memory level by level (deepest first).
// Outputs: g_pose [B,72] (axis-angle) or [B,24,9] (rotation matrices), g_betas [B,L] (optional:
// joint-regressor path + the precomputed vertex path g_beta_vert).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(POSE_WARPS * 32)
lbs_bwd_chain_kernel(SmplDev m, const float* __restrict__ pose, int pose_is_rotmat, const float* __restrict__ R,
                     const float* __restrict__ Jrest, const float* __restrict__ G, const float* __restrict__ gA,
                     const float* __restrict__ gJ49, const float* __restrict__ g_pf,
                     const float* __restrict__ g_beta_vert, int beta_tiles, int B, float* __restrict__ g_pose,
                     float* __restrict__ g_betas, LbsAdam ad) {
    __shared__ float s_c3[POSE_WARPS][24][9];     // child -> parent contribution to gG3
    __shared__ float s_ct[POSE_WARPS][24][3];     // child -> parent contribution to gGt
    __shared__ float s_cj[POSE_WARPS][24][3];     // child -> parent contribution to gJ (= -gt_child)
    __shared__ float s_gJ[POSE_WARPS][24][3];
    const int w = threadIdx.x / 32, lane = threadIdx.x & 31;
    const int b = blockIdx.x * POSE_WARPS + w;
    const bool act = (b < B) && lane < 24;
    const int par = lane < 24 ? m.parents[lane] : -1;
    const int dep = lane < 24 ? m.depth[lane] : -1;
    float Rk[9], Jk[3], G3[9], gG3[9], gGt[3], gJ[3] = {0.f, 0.f, 0.f}, gR[9];
    float Jp[3] = {0.f, 0.f, 0.f}, G3p[9];
    if (act) {
        const size_t o = (size_t)b * 24 + lane;
#pragma unroll
        for (int i = 0; i < 9; ++i) Rk[i] = R[o * 9 + i];
#pragma unroll
        for (int c = 0; c < 3; ++c) Jk[c] = Jrest[o * 3 + c];
        const float* g = G + o * 12;
#pragma unroll
        for (int r = 0; r < 3; ++r) { G3[3 * r] = g[4 * r]; G3[3 * r + 1] = g[4 * r + 1]; G3[3 * r + 2] = g[4 * r + 2]; }
        if (par >= 0) {
            const float* gp = G + ((size_t)b * 24 + par) * 12;
#pragma unroll
            for (int r = 0; r < 3; ++r) { G3p[3 * r] = gp[4 * r]; G3p[3 * r + 1] = gp[4 * r + 1]; G3p[3 * r + 2] = gp[4 * r + 2]; }
#pragma unroll
            for (int c = 0; c < 3; ++c) Jp[c] = Jrest[((size_t)b * 24 + par) * 3 + c];
        }
        // A3 = G3, At = Gt - G3 J, posed joint = Gt
        const float* ga = gA + o * 12;
        const float gAt[3] = {ga[3], ga[7], ga[11]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            gG3[3 * r] = ga[4 * r] - gAt[r] * Jk[0];
            gG3[3 * r + 1] = ga[4 * r + 1] - gAt[r] * Jk[1];
            gG3[3 * r + 2] = ga[4 * r + 2] - gAt[r] * Jk[2];
            gGt[r] = gAt[r];
        }
        float t[3];
        mat3T_vec(G3, gAt, t);
        gJ[0] = -t[0]; gJ[1] = -t[1]; gJ[2] = -t[2];
        if (gJ49 != nullptr)
            for (int oj = 0; oj < m.NO; ++oj)
                if (m.joint_map[oj] == lane) {
                    gGt[0] += gJ49[((size_t)b * m.NO + oj) * 3];
                    gGt[1] += gJ49[((size_t)b * m.NO + oj) * 3 + 1];
                    gGt[2] += gJ49[((size_t)b * m.NO + oj) * 3 + 2];
                }
    }
    for (int level = 23; level >= 0; --level) {
        // parents at `level` collect what their (already processed) children left for them
        if (act && dep == level) {
            for (int ch = 0; ch < 24; ++ch)
                if (m.parents[ch] == lane) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) gG3[i] += s_c3[w][ch][i];
#pragma unroll
                    for (int c = 0; c < 3; ++c) { gGt[c] += s_ct[w][ch][c]; gJ[c] += s_cj[w][ch][c]; }
                }
            // own local transform: G3 = G3p R, Gt = G3p t + Gtp, t = J - Jp
            if (par >= 0) {
                const float t[3] = {Jk[0] - Jp[0], Jk[1] - Jp[1], Jk[2] - Jp[2]};
                float RT[9] = {Rk[0], Rk[3], Rk[6], Rk[1], Rk[4], Rk[7], Rk[2], Rk[5], Rk[8]};
                float c3[9];
                mat3_mul(gG3, RT, c3);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    c3[3 * r] = fmaf(gGt[r], t[0], c3[3 * r]);
                    c3[3 * r + 1] = fmaf(gGt[r], t[1], c3[3 * r + 1]);
                    c3[3 * r + 2] = fmaf(gGt[r], t[2], c3[3 * r + 2]);
                }
                float G3pT[9] = {G3p[0], G3p[3], G3p[6], G3p[1], G3p[4], G3p[7], G3p[2], G3p[5], G3p[8]};
                mat3_mul(G3pT, gG3, gR);
                float gt[3];
                mat3T_vec(G3p, gGt, gt);
                gJ[0] += gt[0]; gJ[1] += gt[1]; gJ[2] += gt[2];
#pragma unroll
                for (int i = 0; i < 9; ++i) s_c3[w][lane][i] = c3[i];
#pragma unroll
                for (int c = 0; c < 3; ++c) { s_ct[w][lane][c] = gGt[c]; s_cj[w][lane][c] = -gt[c]; }
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i) gR[i] = gG3[i];
                gJ[0] += gGt[0]; gJ[1] += gGt[1]; gJ[2] += gGt[2];
            }
        }
        __syncwarp();
    }
    if (act) {
        if (lane >= 1 && g_pf != nullptr) {
            const float* gp = g_pf + (size_t)b * 207 + (lane - 1) * 9;
#pragma unroll
            for (int i = 0; i < 9; ++i) gR[i] += gp[i];
        }
        if (g_pose != nullptr) {
            if (pose_is_rotmat) {
#pragma unroll
                for (int i = 0; i < 9; ++i) g_pose[((size_t)b * 24 + lane) * 9 + i] = gR[i];
            } else {
                float gr[3];
                rodrigues_bwd(pose + ((size_t)b * 24 + lane) * 3, gR, gr);
                g_pose[((size_t)b * 24 + lane) * 3] = gr[0];
                g_pose[((size_t)b * 24 + lane) * 3 + 1] = gr[1];
                g_pose[((size_t)b * 24 + lane) * 3 + 2] = gr[2];
            }
        }
        if (ad.body_pose != nullptr) {
            // SMPLify-DC stage 2 (smplifydc.py:149-183): the two parameter tensors are body_pose [B,69] and
            // global_orient [B,3]; this lane owns joint `lane`, i.e. three entries of one of them.  The gradient
            // is the chain's plus the pose-prior term's; torch.optim.Adam follows in place, no g_pose round trip.
            float* prm = lane == 0 ? ad.global_orient + (size_t)b * 3 : ad.body_pose + (size_t)b * 69 + (lane - 1) * 3;
            float* mm = lane == 0 ? ad.m_orient + (size_t)b * 3 : ad.m_pose + (size_t)b * 69 + (lane - 1) * 3;
            float* vv = lane == 0 ? ad.v_orient + (size_t)b * 3 : ad.v_pose + (size_t)b * 69 + (lane - 1) * 3;
            float gr[3];
            rodrigues_bwd(prm, gR, gr);
            if (lane >= 1 && ad.g_extra_pose != nullptr) {
#pragma unroll
                for (int c = 0; c < 3; ++c) gr[c] += ad.g_extra_pose[(size_t)b * 69 + (lane - 1) * 3 + c];
            }
            const int t = lane == 0 ? *ad.step_orient : *ad.step_pose;       // already advanced by lbs_pose_kernel
            const float w1 = (float)(1.0 - ad.beta1), b2 = (float)ad.beta2, w2 = (float)(1.0 - ad.beta2);
            const double bc1 = 1.0 - pow(ad.beta1, (double)t);
            const double bc2 = 1.0 - pow(ad.beta2, (double)t);
            const float sq2 = (float)sqrt(bc2), eps = (float)ad.eps, lr1 = -(float)(ad.lr / bc1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float g = gr[c];
                if (ad.g_out_pose != nullptr) {
                    if (lane == 0) ad.g_out_orient[(size_t)b * 3 + c] = g; else ad.g_out_pose[(size_t)b * 69 + (lane - 1) * 3 + c] = g;
                }
                const float mi = fmaf(w1, g - mm[c], mm[c]);
                const float vi = fmaf(w2, g * g, b2 * vv[c]);
                mm[c] = mi; vv[c] = vi;
                const float denom = sqrtf(vi) / sq2 + eps;
                prm[c] = fmaf(lr1, mi / denom, prm[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) s_gJ[w][lane][c] = gJ[c];
    }
    __syncwarp();
    if (g_betas != nullptr && b < B) {
        for (int l = lane; l < m.L; l += 32) {
            float acc = 0.f;                          // vertex part: the per-tile partials of lbs_bwd_vertex_kernel, in order
            if (g_beta_vert != nullptr)
                for (int t = 0; t < beta_tiles; ++t) acc += g_beta_vert[((size_t)t * B + b) * m.L + l];
            for (int k = 0; k < 24; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) acc = fmaf(m.J_shapedirs[(size_t)(3 * k + c) * m.L + l], s_gJ[w][k][c], acc);
            g_betas[(size_t)b * m.L + l] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
int launch_lbs_forward(const SmplDev& m, const float* betas, const float* pose, int pose_is_rotmat, int B,
                       const LbsBuffers& w, float* verts, float* joints, cudaStream_t st, const float* orient,
                       int* step_a, int* step_b) {
    if (B == 0) return 0;
    KernelTimer timer("lbs_forward_kernels", st);
    // blend shapes + skinning: the tcgen05 kernel (lbs_tc.cu) whenever the model fits its operand layout;
    // TUCH_LBS_FFMA=1 keeps the CUDA-core kernel for A/B measurements
    static const bool force_ffma = getenv("TUCH_LBS_FFMA") != nullptr && atoi(getenv("TUCH_LBS_FFMA")) != 0;
    const bool tc = m.tc_model != nullptr && !force_ffma;
    const int Bp = tc ? cdiv(B, LBS_TC_NB) * LBS_TC_NB : B;           // the pose kernel zero-fills the padding rows
    lbs_pose_kernel<<<cdiv(Bp, POSE_WARPS), POSE_WARPS * 32, 0, st>>>(m, betas, pose, orient, pose_is_rotmat, B, w.R, w.Jrest,
                                                                    w.G, w.A, w.pf, tc ? w.featop : nullptr, step_a, step_b);
    TUCH_LAUNCH_CHECK(); count_launch();
    if (tc) {
        if (int rc = launch_lbs_skin_tc(m, w.featop, w.A, B, verts, w.v_posed, st)) return rc;
    } else {
        dim3 grid(cdiv(m.V, LBS_VT), cdiv(B, LBS_NB));
        lbs_skin_kernel<<<grid, LBS_VT, 0, st>>>(m, betas, w.pf, w.A, B, verts, w.v_posed);
        TUCH_LAUNCH_CHECK(); count_launch();
    }
    if (joints != nullptr) {
        lbs_joints_kernel<<<B, JOINT_THREADS, 0, st>>>(m, verts, w.G, joints);
        TUCH_LAUNCH_CHECK(); count_launch();
    }
    return 0;
}

int launch_lbs_joints(const SmplDev& m, const float* verts, const LbsBuffers& w, int B, float* joints, cudaStream_t st) {
    if (B == 0 || joints == nullptr) return 0;
    lbs_joints_kernel<<<B, JOINT_THREADS, 0, st>>>(m, verts, w.G, joints);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_lbs_backward(const SmplDev& m, const float* pose, int pose_is_rotmat, int B, const LbsBuffers& w,
                        const float* gV, const float* gJ49, float* g_pose, float* g_betas, cudaStream_t st,
                        const LbsAdam* adam, const LbsSide* side) {
    if (B == 0) return 0;
    KernelTimer timer("lbs_backward_kernels", st);
    dim3 grid(cdiv(m.V, LBS_VT), cdiv(B, LBS_NB));
    const bool need_pf = g_pose != nullptr || (adam != nullptr && adam->body_pose != nullptr);
    static const bool force_ffma = getenv("TUCH_LBS_FFMA") != nullptr && atoi(getenv("TUCH_LBS_FFMA")) != 0;
    const bool tc = need_pf && m.tcb_model != nullptr && !force_ffma;        // tensor-core contraction (lbs_tc_bwd.cu)
    const int n_coords = m.V * 3;
    const int per = 11 * GT_K;                    // see below
    const int S = tc ? lbs_tcb_slabs(m.V) : cdiv(n_coords, per);
    Scratch sc;
    const size_t h_part = sc.plan(!need_pf ? 0 : tc ? sizeof(float) * (size_t)S * B * 256 : sizeof(float) * (size_t)S * B * 207);
    const size_t h_beta = sc.plan(g_betas != nullptr ? sizeof(float) * (size_t)grid.x * B * m.L : 0);
    if (int rc = sc.commit_slot(st, 2)) return rc;
    float* beta_part = g_betas != nullptr ? sc.get<float>(h_beta) : nullptr;
    const int ks_total = lbs_tcb_slabs(m.V) * LBS_TCB_KSTEPS;
    if (tc) {
        // the K padding behind the last coordinate must be finite (it multiplies zero rows of the model operand):
        // zero the k-steps from the one that holds coordinate 3V on, per body tile, before the vertex pass fills in
        const int first = n_coords / 16;
        for (int bt = 0; bt < cdiv(B, LBS_TC_NB); ++bt)
            TUCH_CUDA(cudaMemsetAsync(reinterpret_cast<unsigned char*>(w.gradop) + ((size_t)bt * ks_total + first) * LBS_TCB_STEP_BYTES, 0,
                                      (size_t)(ks_total - first) * LBS_TCB_STEP_BYTES, st));
    }
    lbs_bwd_vertex_kernel<<<grid, LBS_VT, 0, st>>>(m, w.A, gV, gJ49, B, w.g_comb, w.g_vposed, beta_part,
                                                   tc ? w.gradop : nullptr, ks_total);
    TUCH_LAUNCH_CHECK(); count_launch();
    // the per-joint reduction runs beside the contractions when the caller lends a second stream
    const cudaStream_t st_j = side != nullptr ? side->stream : st;
    if (side != nullptr) {
        TUCH_CUDA(cudaEventRecord(side->fork, st));
        TUCH_CUDA(cudaStreamWaitEvent(st_j, side->fork, 0));
        lbs_bwd_joint_kernel<<<B * 24, 128, 0, st_j>>>(m, w.g_comb, w.v_posed, B, w.gA);
        TUCH_LAUNCH_CHECK(); count_launch();
        TUCH_CUDA(cudaEventRecord(side->join, st_j));
    }
    // the pose-feature gradient is only needed when the pose is differentiated (not in SMPLify-DC's stage 1, which
    // optimises betas and the camera): skip its [B,3V] x [3V,207] contraction otherwise
    if (tc) {
        if (int rc = launch_lbs_tc_bwd(m, w.gradop, B, sc.get<float>(h_part), w.g_pf, st)) return rc;
    } else if (need_pf) {
        // split K into FIXED slabs of 352 coordinates (59 splits at SMPL size, >= 236 CTAs at any batch): the
        // grouping of the partial sums must not depend on the batch size, or a body fitted in a shard of the batch
        // (BASELINE config 4) would round differently from the same body fitted in the whole batch
        float* partial = sc.get<float>(h_part);
        dim3 g2(cdiv(207, GT_R), cdiv(B, GT_B), S);
        lbs_contract_tiled_kernel<<<g2, GT_THREADS, 0, st>>>(m.posedirs, w.g_vposed, 207, n_coords, B, per, partial);
        TUCH_LAUNCH_CHECK(); count_launch();
        lbs_contract_reduce_kernel<<<cdiv(B * 207, 256), 256, 0, st>>>(partial, B * 207, S, w.g_pf);
        TUCH_LAUNCH_CHECK(); count_launch();
    }
    if (side == nullptr) {
        lbs_bwd_joint_kernel<<<B * 24, 128, 0, st>>>(m, w.g_comb, w.v_posed, B, w.gA);
        TUCH_LAUNCH_CHECK(); count_launch();
    } else {
        TUCH_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    }
    lbs_bwd_chain_kernel<<<cdiv(B, POSE_WARPS), POSE_WARPS * 32, 0, st>>>(
        m, pose, pose_is_rotmat, w.R, w.Jrest, w.G, w.gA, gJ49, need_pf ? w.g_pf : nullptr, beta_part, (int)grid.x, B,
        g_pose, g_betas, adam != nullptr ? *adam : LbsAdam{});
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

}  // namespace tuch
