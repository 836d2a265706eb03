// Pose / camera bookkeeping around the SMPLify-DC call of the train step:
//   estimate_translation        tuch/utils/geometry.py:114-205   (per-sample numpy least squares, B D2H syncs)
//   rotation_matrix_to_angle_axis   torchgeometry 0.1.2, call site tuch/train/train_module.py:208-211
//   FitsDict.rotate_pose / flip_pose   tuch/train/fits_dict.py:89-119   (per-sample cv2.Rodrigues on the host)
// All of them are a few dozen floats per body; the point of doing them on the device is that the train
// step no longer synchronises with the host 2 B + B times between the contact kernels.
#include "api_internal.h"

namespace tuch {

// ------------------------------------------------------------------------------------------
// estimate_translation: weighted least squares of the pinhole equations in the camera translation.
// The reference promotes the fp32 inputs to fp64 (numpy arithmetic with Python floats) and rounds the
// 3-vector back to fp32; the weights are sqrt(conf) taken in fp32 (geometry.py:133).
// One thread per body.
// ------------------------------------------------------------------------------------------
__global__ void estimate_translation_kernel(const float* __restrict__ S, const float* __restrict__ joints_2d,
                                            const uint8_t* __restrict__ has_anno, int B, int J, int n_op,
                                            double focal, double img_size, float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // ground-truth joints n_op.. when annotated, the OpenPose joints 0..n_op otherwise (geometry.py:192-199)
    const int j0 = has_anno[b] ? n_op : 0, j1 = has_anno[b] ? J : n_op;
    const float* s = S + (size_t)b * J * 3;
    const float* k = joints_2d + (size_t)b * J * 3;
    float conf_sum = 0.f;
    for (int j = j0; j < j1; ++j) conf_sum += k[3 * j + 2];
    float* o = out + (size_t)b * 3;
    if (!(conf_sum > 0.f)) { o[0] = o[1] = o[2] = 0.f; return; }          // geometry.py:201
    const double c0 = img_size / 2.0;
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, rhs[3] = {0, 0, 0};
    for (int j = j0; j < j1; ++j) {
        const double w = (double)sqrtf(k[3 * j + 2]);
        const double X = s[3 * j], Y = s[3 * j + 1], Z = s[3 * j + 2];
        const double u = k[3 * j], v = k[3 * j + 1];
        // rows of W Q and W c for the x and the y equation of this joint (geometry.py:136-143)
        const double qx[3] = {w * focal, 0.0, w * (c0 - u)};
        const double qy[3] = {0.0, w * focal, w * (c0 - v)};
        const double cx = w * ((u - c0) * Z - focal * X), cy = w * ((v - c0) * Z - focal * Y);
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) A[r][c] += qx[r] * qx[c] + qy[r] * qy[c];
            rhs[r] += qx[r] * cx + qy[r] * cy;
        }
    }
    // Gaussian elimination with partial pivoting (what LAPACK gesv does for np.linalg.solve); a singular
    // system (the reference raises LinAlgError) yields non-finite output
    int p[3] = {0, 1, 2};
    for (int c = 0; c < 3; ++c) {
        int best = c;
        for (int r = c + 1; r < 3; ++r) if (fabs(A[p[r]][c]) > fabs(A[p[best]][c])) best = r;
        const int t = p[c]; p[c] = p[best]; p[best] = t;
        const double piv = A[p[c]][c];
        for (int r = c + 1; r < 3; ++r) {
            const double f = A[p[r]][c] / piv;
            for (int cc = c; cc < 3; ++cc) A[p[r]][cc] -= f * A[p[c]][cc];
            rhs[p[r]] -= f * rhs[p[c]];
        }
    }
    double x[3];
    for (int r = 2; r >= 0; --r) {
        double acc = rhs[p[r]];
        for (int c = r + 1; c < 3; ++c) acc -= A[p[r]][c] * x[c];
        x[r] = acc / A[p[r]][r];
    }
    o[0] = (float)x[0]; o[1] = (float)x[1]; o[2] = (float)x[2];
}

// ------------------------------------------------------------------------------------------
// torchgeometry 0.1.2 conversions (third-party, not in the reference tree; restated from the published
// package): rotation_matrix_to_quaternion (eps = 1e-6, branches on the TRANSPOSED matrix) followed by
// quaternion_to_angle_axis, and angle_axis_to_rotation_matrix (Rodrigues, first-order Taylor below
// theta^2 = 1e-6).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void rotmat_to_angle_axis(const float* R, int row_stride, float* aa) {
    // rmat_t = transpose(rotation_matrix): t(i, j) = R[j][i]
#define T_(i, j) R[(j) * row_stride + (i)]
    const float eps = 1e-6f;
    const bool d2 = T_(2, 2) < eps;
    const bool d0_d1 = T_(0, 0) > T_(1, 1);
    const bool d0_nd1 = T_(0, 0) < -T_(1, 1);
    float q[4], t;
    if (d2 && d0_d1) {
        t = 1.f + T_(0, 0) - T_(1, 1) - T_(2, 2);
        q[0] = T_(1, 2) - T_(2, 1); q[1] = t; q[2] = T_(0, 1) + T_(1, 0); q[3] = T_(2, 0) + T_(0, 2);
    } else if (d2) {
        t = 1.f - T_(0, 0) + T_(1, 1) - T_(2, 2);
        q[0] = T_(2, 0) - T_(0, 2); q[1] = T_(0, 1) + T_(1, 0); q[2] = t; q[3] = T_(1, 2) + T_(2, 1);
    } else if (d0_nd1) {
        t = 1.f - T_(0, 0) - T_(1, 1) + T_(2, 2);
        q[0] = T_(0, 1) - T_(1, 0); q[1] = T_(2, 0) + T_(0, 2); q[2] = T_(1, 2) + T_(2, 1); q[3] = t;
    } else {
        t = 1.f + T_(0, 0) + T_(1, 1) + T_(2, 2);
        q[0] = t; q[1] = T_(1, 2) - T_(2, 1); q[2] = T_(2, 0) - T_(0, 2); q[3] = T_(0, 1) - T_(1, 0);
    }
#undef T_
    const float sc = 0.5f / sqrtf(t);
    for (int i = 0; i < 4; ++i) q[i] *= sc;
    const float s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const float sn = sqrtf(s2), cs = q[0];
    const float two_theta = 2.f * (cs < 0.f ? atan2f(-sn, -cs) : atan2f(sn, cs));
    const float k = s2 > 0.f ? two_theta / sn : 2.f;
    aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
}

__device__ __forceinline__ void angle_axis_to_rotmat(const float* aa, float R[3][3]) {
    const float theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > 1e-6f) {
        const float theta = sqrtf(theta2);
        const float wx = aa[0] / (theta + 1e-6f), wy = aa[1] / (theta + 1e-6f), wz = aa[2] / (theta + 1e-6f);
        const float c = cosf(theta), s = sinf(theta), k = 1.f - c;
        R[0][0] = c + wx * wx * k;       R[0][1] = wx * wy * k - wz * s;  R[0][2] = wy * s + wx * wz * k;
        R[1][0] = wz * s + wx * wy * k;  R[1][1] = c + wy * wy * k;       R[1][2] = -wx * s + wy * wz * k;
        R[2][0] = -wy * s + wx * wz * k; R[2][1] = wx * s + wy * wz * k;  R[2][2] = c + wz * wz * k;
    } else {
        R[0][0] = 1.f;    R[0][1] = -aa[2]; R[0][2] = aa[1];
        R[1][0] = aa[2];  R[1][1] = 1.f;    R[1][2] = -aa[0];
        R[2][0] = -aa[1]; R[2][1] = aa[0];  R[2][2] = 1.f;
    }
}

// rotmat[N][3][cols] (cols = 3, or 4 for the homogeneous [R | t] of train_module.py:208-210) -> [N][3]
__global__ void rotmat_to_angle_axis_kernel(const float* __restrict__ rotmat, int N, int cols, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float aa[3];
    rotmat_to_angle_axis(rotmat + (size_t)i * 3 * cols, cols, aa);
    out[3 * i] = aa[0]; out[3 * i + 1] = aa[1]; out[3 * i + 2] = aa[2];
}

// rotation vector of a rotation matrix the way cv2.Rodrigues computes it (fits_dict.py:113-117), without
// OpenCV's preliminary SVD re-orthonormalisation (the input is a product of two rotations)
__device__ __forceinline__ void rodrigues_vector(const double R[3][3], double r[3]) {
    double rx = R[2][1] - R[1][2], ry = R[0][2] - R[2][0], rz = R[1][0] - R[0][1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0][0] + R[1][1] + R[2][2] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0.0) { r[0] = r[1] = r[2] = 0.0; return; }
        double t = (R[0][0] + 1.0) * 0.5;
        rx = sqrt(t > 0.0 ? t : 0.0);
        t = (R[1][1] + 1.0) * 0.5;
        ry = sqrt(t > 0.0 ? t : 0.0) * (R[0][1] < 0.0 ? -1.0 : 1.0);
        t = (R[2][2] + 1.0) * 0.5;
        rz = sqrt(t > 0.0 ? t : 0.0) * (R[0][2] < 0.0 ? -1.0 : 1.0);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((R[1][2] > 0.0) != (ry * rz > 0.0))) rz = -rz;
        theta /= sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
        return;
    }
    const double vth = theta / (2.0 * s);
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

// FitsDict.rotate_pose (rot in degrees, in-plane rotation of the global orientation) and flip_pose
// (SMPL_POSE_FLIP_PERM + sign flips), in the order `flip_first` selects:
//   flip_first == 0: out = flip(rotate(pose, rot))      (__getitem__, fits_dict.py:73)
//   flip_first != 0: out = rotate(flip(pose), rot)      (__setitem__ calls it with -rot, fits_dict.py:83)
__global__ void fits_pose_kernel(const float* __restrict__ pose, const float* __restrict__ rot_deg,
                                 const uint8_t* __restrict__ is_flipped, const int* __restrict__ flip_perm, int B,
                                 int D, int flip_first, float* __restrict__ out) {
    const int b = blockIdx.x;
    if (b >= B) return;
    extern __shared__ float s_pose[];                        // [2][D]
    float* cur = s_pose;
    float* tmp = s_pose + D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) cur[i] = pose[(size_t)b * D + i];
    __syncthreads();
    const bool flip = is_flipped != nullptr && is_flipped[b] != 0;
    for (int stage = 0; stage < 2; ++stage) {
        const bool do_flip = (stage == 0) == (flip_first != 0);
        if (do_flip) {
            if (flip) {
                for (int i = threadIdx.x; i < D; i += blockDim.x) {
                    const float v = cur[flip_perm[i]];
                    tmp[i] = (i % 3 == 0) ? v : -v;
                }
                __syncthreads();
                for (int i = threadIdx.x; i < D; i += blockDim.x) cur[i] = tmp[i];
                __syncthreads();
            }
        } else {
            if (threadIdx.x == 0 && rot_deg != nullptr) {
                const float a = -3.14159265358979323846f * rot_deg[b] / 180.f;       // fp32 like torch.cos(-np.pi * rot / 180.)
                const float cs = cosf(a), sn = sinf(a);
                float G[3][3];
                angle_axis_to_rotmat(cur, G);
                double M[3][3];
                for (int c = 0; c < 3; ++c) {                 // R @ G in fp32 (torch.matmul), read back as fp32
                    M[0][c] = (double)(cs * G[0][c] - sn * G[1][c]);
                    M[1][c] = (double)(sn * G[0][c] + cs * G[1][c]);
                    M[2][c] = (double)G[2][c];
                }
                double r[3];
                rodrigues_vector(M, r);
                cur[0] = (float)r[0]; cur[1] = (float)r[1]; cur[2] = (float)r[2];
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) out[(size_t)b * D + i] = cur[i];
}

}  // namespace tuch

using namespace tuch;

TUCH_EXPORT int tuch_estimate_translation(const float* S, const float* joints_2d, const uint8_t* has_2d_kp_anno,
                                          int B, int J, int n_openpose, float focal_length, float img_size,
                                          float* out, void* stream) {
    TUCH_REQUIRE(B >= 0 && J > 0 && n_openpose >= 0 && n_openpose <= J, "tuch_estimate_translation: bad size");
    if (B == 0) return 0;
    TUCH_REQUIRE(S && joints_2d && has_2d_kp_anno && out, "tuch_estimate_translation: null pointer");
    estimate_translation_kernel<<<cdiv(B, 64), 64, 0, (cudaStream_t)stream>>>(S, joints_2d, has_2d_kp_anno, B, J,
                                                                              n_openpose, (double)focal_length,
                                                                              (double)img_size, out);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

TUCH_EXPORT int tuch_rotmat_to_angle_axis(const float* rotmat, int N, int cols, float* out, void* stream) {
    TUCH_REQUIRE(N >= 0 && (cols == 3 || cols == 4), "tuch_rotmat_to_angle_axis: need N >= 0 and 3 or 4 columns");
    if (N == 0) return 0;
    TUCH_REQUIRE(rotmat && out, "tuch_rotmat_to_angle_axis: null pointer");
    rotmat_to_angle_axis_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(rotmat, N, cols, out);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

TUCH_EXPORT int tuch_fits_pose_transform(const float* pose, const float* rot_deg, const uint8_t* is_flipped,
                                         const int32_t* flip_perm, int B, int D, int flip_first, float* out,
                                         void* stream) {
    TUCH_REQUIRE(B >= 0 && D >= 3 && D <= 1024, "tuch_fits_pose_transform: bad size");
    if (B == 0) return 0;
    TUCH_REQUIRE(pose && out, "tuch_fits_pose_transform: null pointer");
    TUCH_REQUIRE(is_flipped == nullptr || flip_perm != nullptr, "tuch_fits_pose_transform: flipping needs the permutation");
    fits_pose_kernel<<<B, 96, sizeof(float) * 2 * (size_t)D, (cudaStream_t)stream>>>(pose, rot_deg, is_flipped, flip_perm,
                                                                                      B, D, flip_first, out);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}
