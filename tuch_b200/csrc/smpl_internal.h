// Internal: device-side view of a SMPL body model and the per-batch LBS workspace.
#pragma once
#include <vector>

#include "kernels.h"

namespace tuch {

constexpr int SMPL_MAX_BETAS = 32;
constexpr int SMPL_MAX_JOINTS54 = 64;   // 24 posed + picked vertices + regressed extras

// tcgen05 blend-shape GEMM + skinning (lbs_tc.cu): tile = 128 vertices x up to 128 bodies, K = 224 = 14 k-steps of 16
// (207 pose features + up to 17 betas, zero padded), every fp32 operand as three bf16 terms
constexpr int LBS_TC_M = 128, LBS_TC_NB = 128, LBS_TC_KSTEPS = 14, LBS_TC_K = 16 * LBS_TC_KSTEPS, LBS_TC_MAXK = 8;
constexpr size_t LBS_TC_FEAT_BYTES_PER_TILE = (size_t)LBS_TC_KSTEPS * 3 * 2 * LBS_TC_NB * 16;   // 172,032 per 128 bodies
// tcgen05 pose-blend gradient contraction (lbs_tc_bwd.cu): K = the 3V coordinates in fixed slabs of 11 k-steps
constexpr int LBS_TCB_KSTEPS = 11, LBS_TCB_SLAB = 16 * LBS_TCB_KSTEPS;
constexpr size_t LBS_TCB_STEP_BYTES = (size_t)3 * 2 * LBS_TC_NB * 16;                           // 12,288 per k-step and body tile

// POD passed by value to the LBS kernels (all pointers are device pointers)
struct SmplDev {
    int V, L, K, NX, NE, NO;
    int8_t parents[24];           // kinematic tree (parents[0] = -1) and node depths
    int8_t depth[24];
    const float* v_template;      // [V][3]
    const float* shapedirsT;      // [L][3V]
    const float* posedirs;        // [207][3V]
    const float* J_template;      // [24][3]      Jreg . v_template
    const float* J_shapedirs;     // [24*3][L]    Jreg . shapedirs
    const uint8_t* skin_idx;      // [V][K] joint ids of the K largest-support weights
    const float* skin_w;          // [V][K]
    const int* jl_off;            // [25]  per-joint vertex lists (transpose of the skin weights)
    const int* jl_vert;
    const float* jl_w;
    const int* ex_off;            // [NE+1] extra-joint regressor rows, non-zeros only
    const int* ex_vert;
    const float* ex_w;
    const int* vj_off;            // [V+1] per-vertex joint contributions (picked: w=1; extras: w)
    const int* vj_joint;          //       index into the 54-joint set
    const float* vj_w;
    const int* extra_vertex_ids;  // [NX]
    const int* joint_map;         // [NO]
    const uint16_t* tc_model;     // bf16 x 3 model operand of lbs_tc.cu, or NULL (K > LBS_TC_MAXK or L > 17)
    const uint16_t* tcb_model;    // bf16 x 3 model operand of lbs_tc_bwd.cu (posedirs, K-major over the coordinates)
};

// per-batch intermediates kept for the backward pass (all [B, ...])
struct LbsBuffers {
    float *R, *Jrest, *G, *A, *pf;          // [24*9] [24*3] [24*12] [24*12] [207]
    float *v_posed;                         // [V*3]
    float *g_comb, *g_vposed;               // [V*3] backward scratch
    float *g_pf, *g_beta_vert, *gA;         // [207] [L] [24*12]
    uint16_t* featop;                       // feature operand of lbs_tc.cu: [ceil(B/128)][14][3][2][128][8] bf16
    uint16_t* gradop;                       // gradient operand of lbs_tc_bwd.cu: [ceil(B/128)][slabs * 11][3][2][128][8] bf16
};

// torch.optim.Adam on SMPLify-DC's stage-2 parameters fused into the last kernel of the LBS backward
// (lbs_bwd_chain_kernel); body_pose == NULL = off.  The step counters hold the step being taken: the first kernel
// of the iteration (lbs_pose_kernel) advances them.
struct LbsAdam {
    float *body_pose = nullptr, *global_orient = nullptr;     // [B,69], [B,3] parameters, updated in place
    float *m_pose = nullptr, *v_pose = nullptr, *m_orient = nullptr, *v_orient = nullptr;
    const int *step_pose = nullptr, *step_orient = nullptr;
    const float* g_extra_pose = nullptr;                       // optional [B,69] added to the body_pose gradient
    float *g_out_pose = nullptr, *g_out_orient = nullptr;      // optional: the gradients Adam consumed
    double lr = 0.0, beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
};

// optional second stream for the LBS backward: the per-joint reduction (lbs_bwd_joint_kernel) and the pose-blend
// contraction both consume the vertex pass and are independent of each other
struct LbsSide {
    cudaStream_t stream;
    cudaEvent_t fork, join;
};

size_t lbs_workspace_floats(int V, int L, int B);     // whole workspace of a batch
void lbs_carve(float* base, int B, int V, int L, LbsBuffers& w);
void lbs_tc_pack_model(int V, int L, const float* shapedirs, const float* posedirs, std::vector<uint16_t>& blob);
int launch_lbs_skin_tc(const SmplDev& m, const uint16_t* featop, const float* A, int B, float* verts, float* v_posed,
                       cudaStream_t st);
int lbs_tcb_slabs(int V);
void lbs_tcb_pack_model(int V, const float* posedirs, std::vector<uint16_t>& blob);
// partial: [slabs][B][256] floats of scratch; g_pf [B][207] out
int launch_lbs_tc_bwd(const SmplDev& m, const uint16_t* gradop, int B, float* partial, float* g_pf, cudaStream_t st);

// orient != NULL: split axis-angle pose (pose = body_pose [B,69], orient = global_orient [B,3]);
// step_a / step_b: optional Adam step counters the first kernel advances (see LbsAdam)
int launch_lbs_forward(const SmplDev& m, const float* betas, const float* pose, int pose_is_rotmat, int B,
                       const LbsBuffers& w, float* verts, float* joints, cudaStream_t st, const float* orient = nullptr,
                       int* step_a = nullptr, int* step_b = nullptr);
// the 49 output joints alone (launch_lbs_forward with joints == NULL skips them)
int launch_lbs_joints(const SmplDev& m, const float* verts, const LbsBuffers& w, int B, float* joints, cudaStream_t st);
int launch_lbs_backward(const SmplDev& m, const float* pose, int pose_is_rotmat, int B, const LbsBuffers& w,
                        const float* gV, const float* gJ49, float* g_pose, float* g_betas, cudaStream_t st,
                        const LbsAdam* adam = nullptr, const LbsSide* side = nullptr);

}  // namespace tuch

// opaque handle of include/tuch_b200.h
struct tuch_smpl {
    int device = 0;
    tuch::SmplDev dev{};
    std::vector<void*> owned;       // every device allocation behind `dev`
    int parents[24];
    int depth[24];
    std::vector<int> faces;         // [F][3] host copy (SMPL.faces)
};
