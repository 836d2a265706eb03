// Internal: launchers of the SMPLify-DC objective kernels (objective_kernels.cu).
#pragma once
#include "api_internal.h"

struct tuch_prior {
    int device = 0;
    int M = 0, D = 0;
    float *d_means = nullptr, *d_precisions = nullptr, *d_nll_weights = nullptr;
};

namespace tuch {

enum { PULL_THRESHOLD = 0, PULL_ALL = 1 };
enum { REDUCE_SUM = 0, REDUCE_MEAN = 1 };

int launch_reprojection(const float* joints, const float* cam_t, const float* center, const float* joints_2d,
                        const float* conf, int B, int J, float focal, float sigma, const float* cam_t_est,
                        float depth_weight, const float* g_loss, float* loss, float* extra, float* g_joints,
                        float* g_cam_t, cudaStream_t st);
int launch_pose_terms(const float* means, const float* precisions, const float* nll_weights, int M, int D,
                      const float* pose, const float* betas, int L, int B, float wp, float wa, float ws,
                      float* value, float* prior_value, int* which, float* g_pose, float* g_betas,
                      cudaStream_t st);
int launch_contact_loss(const float* points, const int* argmin, const uint8_t* exterior,
                        const uint8_t* body_active, const int* counts, int B, int N, float euclthres,
                        int pull_mode, int reduce_mode, float weight, const float* g_loss, float* loss,
                        float* parts, float* g_points, cudaStream_t st);
int launch_region_sum(const float* verts, int B, int V, int n_pairs, const float* min_sq, const int* arg_i,
                      const int* arg_j, const uint8_t* body_active, float weight, const float* g_loss,
                      float* r2r, float* g_verts, cudaStream_t st);
int launch_adam(float* param, const float* grad, float* m, float* v, long long n, const int* step_dev,
                int step_add, double lr, double beta1, double beta2, double eps, cudaStream_t st);
int launch_step_advance(int* step_dev, int add, cudaStream_t st);
int launch_combine(const float* rep, int J, const float* terms, const float* contact, float w_contact,
                   const float* r2r, float w_r2r, const float* extra, int B, float* per_body, float* total,
                   cudaStream_t st);

}  // namespace tuch
