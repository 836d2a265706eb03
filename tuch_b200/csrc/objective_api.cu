// C ABI of the SMPLify-DC objective terms (include/tuch_b200.h, sections a6-a9, a13).
#include "objective_internal.h"

using namespace tuch;

TUCH_EXPORT int tuch_reprojection_loss(const float* joints, const float* cam_t, const float* center,
                                       const float* joints_2d, const float* conf, int B, int J, float focal_length,
                                       float sigma, const float* cam_t_est, float depth_loss_weight,
                                       const float* g_loss, float* loss, float* depth_loss, float* g_joints,
                                       float* g_cam_t, void* stream) {
    TUCH_REQUIRE(B >= 0 && J >= 0, "tuch_reprojection_loss: negative size");
    if (B == 0) return 0;
    TUCH_REQUIRE(joints && cam_t && center && joints_2d && conf, "tuch_reprojection_loss: null input");
    return launch_reprojection(joints, cam_t, center, joints_2d, conf, B, J, focal_length, sigma, cam_t_est,
                               depth_loss_weight, g_loss, loss, depth_loss, g_joints, g_cam_t, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_prior_create(int M, int D, const float* means_host, const float* precisions_host,
                                  const float* nll_weights_host, tuch_prior** out) {
    TUCH_REQUIRE(out != nullptr, "tuch_prior_create: out is null");
    *out = nullptr;
    TUCH_REQUIRE(M > 0 && M <= 16 && D > 0 && D <= 96, "tuch_prior_create: need 0 < M <= 16 and 0 < D <= 96 (got %d, %d)", M, D);
    TUCH_REQUIRE(means_host && precisions_host && nll_weights_host, "tuch_prior_create: null array");
    for (int m = 0; m < M; ++m)
        TUCH_REQUIRE(nll_weights_host[m] > 0.f, "tuch_prior_create: nll_weights[%d] must be positive", m);
    tuch_prior* p = new tuch_prior();
    p->M = M; p->D = D;
    TUCH_CUDA(cudaGetDevice(&p->device));
    auto up = [](const float* h, size_t n, float** d) -> int {
        TUCH_CUDA(cudaMalloc((void**)d, n * sizeof(float)));
        TUCH_CUDA(cudaMemcpy(*d, h, n * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    };
    int rc = up(means_host, (size_t)M * D, &p->d_means);
    if (!rc) rc = up(precisions_host, (size_t)M * D * D, &p->d_precisions);
    if (!rc) rc = up(nll_weights_host, (size_t)M, &p->d_nll_weights);
    if (rc) { tuch_prior_destroy(p); return rc; }
    *out = p;
    return 0;
}

TUCH_EXPORT void tuch_prior_destroy(tuch_prior* p) {
    if (!p) return;
    if (p->d_means) cudaFree(p->d_means);
    if (p->d_precisions) cudaFree(p->d_precisions);
    if (p->d_nll_weights) cudaFree(p->d_nll_weights);
    delete p;
}

TUCH_EXPORT int tuch_pose_terms(const tuch_prior* prior, const float* pose, const float* betas, int B, int D, int L,
                                float pose_prior_weight, float angle_prior_weight, float shape_prior_weight,
                                float* value, float* prior_value, int32_t* component, float* g_pose,
                                float* g_betas, void* stream) {
    TUCH_REQUIRE(B >= 0 && D > 0 && L >= 0, "tuch_pose_terms: bad size");
    if (B == 0) return 0;
    TUCH_REQUIRE(pose != nullptr, "tuch_pose_terms: pose is null");
    TUCH_REQUIRE(prior == nullptr || prior->D == D, "tuch_pose_terms: pose has %d entries, the prior %d", D, prior ? prior->D : 0);
    TUCH_REQUIRE(prior != nullptr || pose_prior_weight == 0.f, "tuch_pose_terms: pose_prior_weight != 0 needs a prior");
    TUCH_REQUIRE(betas != nullptr || (shape_prior_weight == 0.f && g_betas == nullptr), "tuch_pose_terms: betas is null");
    return launch_pose_terms(prior ? prior->d_means : nullptr, prior ? prior->d_precisions : nullptr,
                             prior ? prior->d_nll_weights : nullptr, prior ? prior->M : 0, D, pose, betas, L, B,
                             pose_prior_weight * pose_prior_weight, angle_prior_weight * angle_prior_weight,
                             shape_prior_weight * shape_prior_weight, value, prior_value, component, g_pose,
                             g_betas, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_contact_loss(const float* points, const int32_t* argmin, const uint8_t* exterior,
                                  const uint8_t* body_active, const int32_t* counts, int B, int N, float euclthres,
                                  int pull_mode, int reduce_mode, float weight, const float* g_loss, float* loss,
                                  float* parts, float* g_points, void* stream) {
    TUCH_REQUIRE(B >= 0 && N >= 0, "tuch_contact_loss: negative size");
    if (B == 0) return 0;
    TUCH_REQUIRE(N == 0 || (points && argmin && exterior), "tuch_contact_loss: null input");
    TUCH_REQUIRE(pull_mode == TUCH_PULL_THRESHOLD || pull_mode == TUCH_PULL_ALL, "tuch_contact_loss: bad pull_mode %d", pull_mode);
    TUCH_REQUIRE(reduce_mode == TUCH_REDUCE_SUM || reduce_mode == TUCH_REDUCE_MEAN, "tuch_contact_loss: bad reduce_mode %d", reduce_mode);
    return launch_contact_loss(points, argmin, exterior, body_active, counts, B, N, euclthres, pull_mode,
                               reduce_mode, weight, g_loss, loss, parts, g_points, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_region_sum(const float* verts, int B, int V, int n_pairs, const float* min_sq,
                                const int32_t* arg_i, const int32_t* arg_j, const uint8_t* body_active, float weight,
                                const float* g_loss, float* r2r, float* g_verts, void* stream) {
    TUCH_REQUIRE(B >= 0 && V > 0 && n_pairs >= 0, "tuch_region_sum: bad size");
    if (B == 0) return 0;
    TUCH_REQUIRE(n_pairs == 0 || (verts && min_sq && arg_i && arg_j), "tuch_region_sum: null input");
    return launch_region_sum(verts, B, V, n_pairs, min_sq, arg_i, arg_j, body_active, weight, g_loss, r2r, g_verts,
                             (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                               int32_t* step_dev, double lr, double beta1, double beta2, double eps, void* stream) {
    TUCH_REQUIRE(n >= 0, "tuch_adam_step: negative size");
    TUCH_REQUIRE(step_dev != nullptr, "tuch_adam_step: step counter is null");
    if (n > 0) {
        TUCH_REQUIRE(param && grad && exp_avg && exp_avg_sq, "tuch_adam_step: null pointer");
        if (int rc = launch_adam(param, grad, exp_avg, exp_avg_sq, n, step_dev, 1, lr, beta1, beta2, eps,
                                 (cudaStream_t)stream)) return rc;
    }
    return launch_step_advance(step_dev, 1, (cudaStream_t)stream);
}
