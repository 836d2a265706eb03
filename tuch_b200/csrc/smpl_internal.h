// Internal: device-side view of a SMPL body model and the per-batch LBS workspace.
#pragma once
#include <vector>

#include "kernels.h"

namespace tuch {

constexpr int SMPL_MAX_BETAS = 32;
constexpr int SMPL_MAX_JOINTS54 = 64;   // 24 posed + picked vertices + regressed extras

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
};

// per-batch intermediates kept for the backward pass (all [B, ...])
struct LbsBuffers {
    float *R, *Jrest, *G, *A, *pf;          // [24*9] [24*3] [24*12] [24*12] [207]
    float *v_posed;                         // [V*3]
    float *g_comb, *g_vposed;               // [V*3] backward scratch
    float *g_pf, *g_beta_vert, *gA;         // [207] [L] [24*12]
};

size_t lbs_buffer_floats(int V, int L);     // floats per body
void lbs_carve(float* base, int B, int V, int L, LbsBuffers& w);

int launch_lbs_forward(const SmplDev& m, const float* betas, const float* pose, int pose_is_rotmat, int B,
                       const LbsBuffers& w, float* verts, float* joints, cudaStream_t st);
int launch_lbs_backward(const SmplDev& m, const float* pose, int pose_is_rotmat, int B, const LbsBuffers& w,
                        const float* gV, const float* gJ49, float* g_pose, float* g_betas, cudaStream_t st);

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
