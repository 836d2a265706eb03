// C ABI for the SMPL body model (include/tuch_b200.h, section a11).
#include "api_internal.h"
#include "smpl_internal.h"

#include <algorithm>
#include <cmath>

namespace tuch {

size_t lbs_workspace_floats(int V, int L, int B) {
    (void)L;
    // feature operand (whole 128-body tiles, first: cp.async.bulk needs 16-byte alignment) |
    // per body: R 216 | Jrest 72 | G 288 | A 288 | pf 207(->208) | g_pf 208 | gA 288 | g_beta_vert L(->32) | 3 x V*3
    const size_t per_body = 216 + 72 + 288 + 288 + 208 + 208 + 288 + SMPL_MAX_BETAS + 3 * (size_t)V * 3;
    const size_t per_tile = LBS_TC_FEAT_BYTES_PER_TILE / 4 + (size_t)lbs_tcb_slabs(V) * LBS_TCB_KSTEPS * (LBS_TCB_STEP_BYTES / 4);
    return (size_t)cdiv(B, LBS_TC_NB) * per_tile + per_body * (size_t)B;
}

void lbs_carve(float* base, int B, int V, int L, LbsBuffers& w) {
    (void)L;
    const size_t tiles = (size_t)cdiv(B, LBS_TC_NB);
    w.featop = reinterpret_cast<uint16_t*>(base);
    w.gradop = reinterpret_cast<uint16_t*>(base + tiles * (LBS_TC_FEAT_BYTES_PER_TILE / 4));
    float* p = base + tiles * (LBS_TC_FEAT_BYTES_PER_TILE / 4 + (size_t)lbs_tcb_slabs(V) * LBS_TCB_KSTEPS * (LBS_TCB_STEP_BYTES / 4));
    auto take = [&](size_t per_body) { float* r = p; p += per_body * (size_t)B; return r; };
    w.R = take(216); w.Jrest = take(72); w.G = take(288); w.A = take(288);
    w.pf = take(208); w.g_pf = take(208); w.gA = take(288); w.g_beta_vert = take(SMPL_MAX_BETAS);
    w.v_posed = take((size_t)V * 3); w.g_comb = take((size_t)V * 3); w.g_vposed = take((size_t)V * 3);
}

}  // namespace tuch

using namespace tuch;

namespace {
template <typename T>
int to_device(tuch_smpl* s, const std::vector<T>& h, const T** out) {
    void* d = nullptr;
    TUCH_CUDA(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    if (!h.empty()) TUCH_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    s->owned.push_back(d);
    *out = (const T*)d;
    return 0;
}
}  // namespace

TUCH_EXPORT int tuch_smpl_create(int V, int L, const float* v_template, const float* shapedirs,
                                 const float* posedirs, const float* J_regressor, const float* lbs_weights,
                                 const int32_t* parents, int n_extra_verts, const int32_t* extra_vertex_ids,
                                 int n_extra_reg, const float* J_regressor_extra, int n_out,
                                 const int32_t* joint_map, tuch_smpl** out) {
    TUCH_REQUIRE(out != nullptr, "tuch_smpl_create: out is null");
    *out = nullptr;
    TUCH_REQUIRE(V > 0 && L > 0 && L <= SMPL_MAX_BETAS, "tuch_smpl_create: need V > 0 and 0 < num_betas <= %d", SMPL_MAX_BETAS);
    TUCH_REQUIRE(v_template && shapedirs && posedirs && J_regressor && lbs_weights && parents,
                 "tuch_smpl_create: null model array");
    TUCH_REQUIRE(n_extra_verts >= 0 && n_extra_reg >= 0 && 24 + n_extra_verts + n_extra_reg <= SMPL_MAX_JOINTS54,
                 "tuch_smpl_create: too many extra joints");
    TUCH_REQUIRE(n_out > 0 && joint_map, "tuch_smpl_create: joint_map is required");
    TUCH_REQUIRE(parents[0] < 0, "tuch_smpl_create: parents[0] must be -1");
    for (int k = 1; k < 24; ++k)
        TUCH_REQUIRE(parents[k] >= 0 && parents[k] < k, "tuch_smpl_create: parents[%d]=%d is not an earlier joint", k, parents[k]);
    for (int i = 0; i < n_extra_verts; ++i)
        TUCH_REQUIRE(extra_vertex_ids[i] >= 0 && extra_vertex_ids[i] < V, "tuch_smpl_create: extra vertex id out of range");
    const int n54 = 24 + n_extra_verts + n_extra_reg;
    for (int i = 0; i < n_out; ++i)
        TUCH_REQUIRE(joint_map[i] >= 0 && joint_map[i] < n54, "tuch_smpl_create: joint_map[%d]=%d out of range [0,%d)", i, joint_map[i], n54);

    tuch_smpl* s = new tuch_smpl();
    TUCH_CUDA(cudaGetDevice(&s->device));
    SmplDev& d = s->dev;
    d.V = V; d.L = L; d.NX = n_extra_verts; d.NE = n_extra_reg; d.NO = n_out;
    for (int k = 0; k < 24; ++k) {
        s->parents[k] = parents[k];
        s->depth[k] = k == 0 ? 0 : s->depth[parents[k]] + 1;
        d.parents[k] = (int8_t)s->parents[k];
        d.depth[k] = (int8_t)s->depth[k];
    }
    const size_t V3 = (size_t)V * 3;
    // transposed shape basis [L][3V]
    std::vector<float> ST((size_t)L * V3);
    for (size_t c = 0; c < V3; ++c)
        for (int l = 0; l < L; ++l) ST[(size_t)l * V3 + c] = shapedirs[c * L + l];
    // joint regressor folded into the template and the shape basis (double accumulation)
    std::vector<float> JT(72), JS((size_t)72 * L);
    for (int k = 0; k < 24; ++k)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int v = 0; v < V; ++v) acc += (double)J_regressor[(size_t)k * V + v] * v_template[(size_t)v * 3 + c];
            JT[k * 3 + c] = (float)acc;
            for (int l = 0; l < L; ++l) {
                double a2 = 0.0;
                for (int v = 0; v < V; ++v)
                    a2 += (double)J_regressor[(size_t)k * V + v] * shapedirs[((size_t)v * 3 + c) * L + l];
                JS[(size_t)(k * 3 + c) * L + l] = (float)a2;
            }
        }
    // skinning weights: keep the non-zeros, K = max per vertex
    int K = 1;
    for (int v = 0; v < V; ++v) {
        int n = 0;
        for (int k = 0; k < 24; ++k) n += lbs_weights[(size_t)v * 24 + k] != 0.f;
        K = std::max(K, n);
    }
    std::vector<uint8_t> sidx((size_t)V * K, 0);
    std::vector<float> sw((size_t)V * K, 0.f);
    std::vector<std::vector<std::pair<int, float>>> per_joint(24);
    for (int v = 0; v < V; ++v) {
        int n = 0;
        for (int k = 0; k < 24; ++k) {
            const float w = lbs_weights[(size_t)v * 24 + k];
            if (w != 0.f) {
                sidx[(size_t)v * K + n] = (uint8_t)k;
                sw[(size_t)v * K + n] = w;
                ++n;
                per_joint[k].push_back({v, w});
            }
        }
    }
    std::vector<int> jl_off(25, 0), jl_vert;
    std::vector<float> jl_w;
    for (int k = 0; k < 24; ++k) {
        for (auto& e : per_joint[k]) { jl_vert.push_back(e.first); jl_w.push_back(e.second); }
        jl_off[k + 1] = (int)jl_vert.size();
    }
    // extra regressors (rows, non-zeros) and the per-vertex transpose incl. picked joints
    std::vector<int> ex_off(n_extra_reg + 1, 0), ex_vert;
    std::vector<float> ex_w;
    std::vector<std::vector<std::pair<int, float>>> per_vertex(V);
    for (int i = 0; i < n_extra_verts; ++i) per_vertex[extra_vertex_ids[i]].push_back({24 + i, 1.f});
    for (int e = 0; e < n_extra_reg; ++e) {
        for (int v = 0; v < V; ++v) {
            const float w = J_regressor_extra[(size_t)e * V + v];
            if (w != 0.f) {
                ex_vert.push_back(v); ex_w.push_back(w);
                per_vertex[v].push_back({24 + n_extra_verts + e, w});
            }
        }
        ex_off[e + 1] = (int)ex_vert.size();
    }
    std::vector<int> vj_off(V + 1, 0), vj_joint;
    std::vector<float> vj_w;
    for (int v = 0; v < V; ++v) {
        for (auto& e : per_vertex[v]) { vj_joint.push_back(e.first); vj_w.push_back(e.second); }
        vj_off[v + 1] = (int)vj_joint.size();
    }
    d.K = K;
    int rc = 0;
    d.tc_model = nullptr;
    d.tcb_model = nullptr;
    if (K <= LBS_TC_MAXK && 207 + L <= LBS_TC_K) {
        std::vector<uint16_t> blob;
        lbs_tc_pack_model(V, L, shapedirs, posedirs, blob);
        rc = to_device(s, blob, &d.tc_model);
    }
    if (!rc) {
        std::vector<uint16_t> blob;
        lbs_tcb_pack_model(V, posedirs, blob);
        rc = to_device(s, blob, &d.tcb_model);
    }
    rc = rc ? rc : to_device(s, std::vector<float>(v_template, v_template + V3), &d.v_template);
    rc = rc ? rc : to_device(s, ST, &d.shapedirsT);
    rc = rc ? rc : to_device(s, std::vector<float>(posedirs, posedirs + 207 * V3), &d.posedirs);
    rc = rc ? rc : to_device(s, JT, &d.J_template);
    rc = rc ? rc : to_device(s, JS, &d.J_shapedirs);
    rc = rc ? rc : to_device(s, sidx, &d.skin_idx);
    rc = rc ? rc : to_device(s, sw, &d.skin_w);
    rc = rc ? rc : to_device(s, jl_off, &d.jl_off);
    rc = rc ? rc : to_device(s, jl_vert, &d.jl_vert);
    rc = rc ? rc : to_device(s, jl_w, &d.jl_w);
    rc = rc ? rc : to_device(s, ex_off, &d.ex_off);
    rc = rc ? rc : to_device(s, ex_vert, &d.ex_vert);
    rc = rc ? rc : to_device(s, ex_w, &d.ex_w);
    rc = rc ? rc : to_device(s, vj_off, &d.vj_off);
    rc = rc ? rc : to_device(s, vj_joint, &d.vj_joint);
    rc = rc ? rc : to_device(s, vj_w, &d.vj_w);
    rc = rc ? rc : to_device(s, std::vector<int>(extra_vertex_ids, extra_vertex_ids + n_extra_verts), &d.extra_vertex_ids);
    rc = rc ? rc : to_device(s, std::vector<int>(joint_map, joint_map + n_out), &d.joint_map);
    if (rc) { tuch_smpl_destroy(s); return rc; }
    *out = s;
    return 0;
}

TUCH_EXPORT void tuch_smpl_destroy(tuch_smpl* s) {
    if (!s) return;
    for (void* p : s->owned) if (p) cudaFree(p);
    delete s;
}

TUCH_EXPORT int tuch_smpl_num_verts(const tuch_smpl* s) { return s ? s->dev.V : -1; }
TUCH_EXPORT int tuch_smpl_num_joints(const tuch_smpl* s) { return s ? s->dev.NO : -1; }

TUCH_EXPORT size_t tuch_smpl_workspace_floats(const tuch_smpl* s, int B) {
    if (!s || B <= 0) return 0;
    return lbs_workspace_floats(s->dev.V, s->dev.L, B);
}

TUCH_EXPORT int tuch_smpl_forward(const tuch_smpl* s, const float* betas, const float* pose, int pose_is_rotmat,
                                  int B, float* workspace, float* vertices, float* joints, void* stream) {
    TUCH_REQUIRE(s != nullptr, "tuch_smpl_forward: null model");
    TUCH_REQUIRE(B >= 0, "tuch_smpl_forward: negative batch");
    if (B == 0) return 0;
    TUCH_REQUIRE(betas && pose && workspace && vertices, "tuch_smpl_forward: null pointer");
    TUCH_REQUIRE(((uintptr_t)workspace & 15) == 0, "tuch_smpl_forward: the workspace must be 16-byte aligned");
    LbsBuffers w;
    lbs_carve(workspace, B, s->dev.V, s->dev.L, w);
    return launch_lbs_forward(s->dev, betas, pose, pose_is_rotmat, B, w, vertices, joints, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_smpl_backward(const tuch_smpl* s, const float* pose, int pose_is_rotmat, int B,
                                   float* workspace, const float* g_vertices, const float* g_joints,
                                   float* g_pose, float* g_betas, void* stream) {
    TUCH_REQUIRE(s != nullptr, "tuch_smpl_backward: null model");
    TUCH_REQUIRE(B >= 0, "tuch_smpl_backward: negative batch");
    if (B == 0) return 0;
    TUCH_REQUIRE(pose && workspace, "tuch_smpl_backward: null pointer");
    LbsBuffers w;
    lbs_carve(workspace, B, s->dev.V, s->dev.L, w);
    return launch_lbs_backward(s->dev, pose, pose_is_rotmat, B, w, g_vertices, g_joints, g_pose, g_betas,
                               (cudaStream_t)stream);
}
