// SMPLify-DC objective terms, forward value and analytic gradient in one pass (sm_100a):
//   reprojection + GMoF            tuch/smplify/losses.py:25-32,56-61,138-143,178-183 and
//                                  tuch/utils/geometry.py:83-111 (identity rotation)
//   max-mixture pose prior         tuch/smplify/prior.py:117-132
//   angle / shape / depth priors   losses.py:146-149,155-162,189-192
//   push / pull contact terms      losses.py:96-105, tuch/train/loss.py:299-315, tuch/eft/loss.py:158-166
//   region-to-region gradient      losses.py:108-117 (autograd through batch_pairwise_dist)
//   Adam                           torch.optim.Adam as used at tuch/smplify/smplifydc.py:117,150,197
// The reference obtains every gradient from autograd over ~80 ATen launches per body; here each
// term writes its value and its gradient directly.
#include "objective_internal.h"

namespace tuch {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (blockDim.x multiple of 32, <= 1024); result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += s_red[w];
    return t;
}

// ------------------------------------------------------------------------------------------
// reprojection: loss[b,j] = conf^2 * sum_xy gmof(f * (p/p_z)_xy + c_xy - target_xy), p = joint + t
// one CTA per body.  g_loss (optional) is the upstream gradient of loss[b,j].
// depth term (camera_fitting_loss, losses.py:146): extra[b] = dw^2 (t_z - t_est_z)^2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
reprojection_kernel(const float* __restrict__ joints, const float* __restrict__ cam_t,
                    const float* __restrict__ center, const float* __restrict__ joints_2d,
                    const float* __restrict__ conf, int J, float focal, float sigma,
                    const float* __restrict__ cam_t_est, float depth_weight,
                    const float* __restrict__ g_loss, float* __restrict__ loss, float* __restrict__ extra,
                    float* __restrict__ g_joints, float* __restrict__ g_cam_t) {
    __shared__ float s_red[2];
    const int b = blockIdx.x;
    const float tx = cam_t[3 * b], ty = cam_t[3 * b + 1], tz = cam_t[3 * b + 2];
    const float cx = center[2 * b], cy = center[2 * b + 1];
    const float s2 = sigma * sigma;
    float gtx = 0.f, gty = 0.f, gtz = 0.f;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const size_t o = (size_t)b * J + j;
        const float px = joints[3 * o] + tx, py = joints[3 * o + 1] + ty, pz = joints[3 * o + 2] + tz;
        const float u = px / pz, v = py / pz;
        const float rx = fmaf(focal, u, cx) - joints_2d[2 * o];
        const float ry = fmaf(focal, v, cy) - joints_2d[2 * o + 1];
        const float rx2 = rx * rx, ry2 = ry * ry;
        const float c = conf[o], c2 = c * c;
        if (loss != nullptr) loss[o] = c2 * ((s2 * rx2) / (s2 + rx2) + (s2 * ry2) / (s2 + ry2));
        if (g_joints != nullptr || g_cam_t != nullptr) {
            const float up = (g_loss != nullptr ? g_loss[o] : 1.f) * c2;
            const float dx = s2 + rx2, dy = s2 + ry2;
            const float gx = up * 2.f * s2 * s2 * rx / (dx * dx);       // d gmof / d r
            const float gy = up * 2.f * s2 * s2 * ry / (dy * dy);
            const float fz = focal / pz;
            const float jx = gx * fz, jy = gy * fz;
            const float jz = -(jx * u + jy * v);
            if (g_joints != nullptr) { g_joints[3 * o] = jx; g_joints[3 * o + 1] = jy; g_joints[3 * o + 2] = jz; }
            gtx += jx; gty += jy; gtz += jz;
        }
    }
    if (g_cam_t != nullptr || extra != nullptr) {
        gtx = block_sum(gtx, s_red);
        gty = block_sum(gty, s_red);
        gtz = block_sum(gtz, s_red);
        if (threadIdx.x == 0) {
            float e = 0.f;
            if (cam_t_est != nullptr) {
                const float dz = tz - cam_t_est[3 * b + 2];
                const float w2 = depth_weight * depth_weight;
                e = w2 * (dz * dz);
                gtz = fmaf(2.f * w2, dz, gtz);
            }
            if (extra != nullptr) extra[b] = e;
            if (g_cam_t != nullptr) { g_cam_t[3 * b] = gtx; g_cam_t[3 * b + 1] = gty; g_cam_t[3 * b + 2] = gtz; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// pose terms of one body (one CTA, 256 threads):
//   value[b] = wp * min_m [0.5 (th - mu_m)^T P_m (th - mu_m) - log nllw_m]        prior.py:117-132
//            + wa * sum_k exp(s_k th[idx_k])^2                                     losses.py:155-162
//            + ws * sum beta^2                                                     losses.py:149,192
// wp / wa / ws are the already-squared weights.  Gradients: g_pose[b,D], g_betas[b,L] (optional).
// ------------------------------------------------------------------------------------------
constexpr int PT_THREADS = 256;
constexpr int PT_MAXD = 96;
constexpr int PT_MAXM = 16;

__global__ void __launch_bounds__(PT_THREADS)
pose_terms_kernel(const float* __restrict__ means, const float* __restrict__ precisions,
                  const float* __restrict__ nll_weights, int M, int D,
                  const float* __restrict__ pose, const float* __restrict__ betas, int L,
                  float wp, float wa, float ws, float* __restrict__ value, float* __restrict__ prior_value,
                  int* __restrict__ which, float* __restrict__ g_pose, float* __restrict__ g_betas) {
    __shared__ float s_d[PT_MAXD];
    __shared__ float s_quad[PT_MAXM];
    __shared__ float s_red[PT_THREADS / 32];
    __shared__ int s_best;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = PT_THREADS / 32;
    const float* th = pose + (size_t)b * D;
    float prior = 0.f;
    if (wp != 0.f && M > 0) {
        if (threadIdx.x < M) s_quad[threadIdx.x] = 0.f;
        for (int m = 0; m < M; ++m) {
            __syncthreads();
            if (threadIdx.x < D) s_d[threadIdx.x] = th[threadIdx.x] - means[(size_t)m * D + threadIdx.x];
            __syncthreads();
            const float* P = precisions + (size_t)m * D * D;
            float part = 0.f;
            for (int i = warp; i < D; i += nw) {
                float t = 0.f;
                for (int j = lane; j < D; j += 32) t = fmaf(P[(size_t)i * D + j], s_d[j], t);
                t = warp_sum(t);
                part = fmaf(t, s_d[i], part);
            }
            if (lane == 0) s_red[warp] = part;
            __syncthreads();
            if (threadIdx.x == 0) {
                float q = 0.f;
                for (int w = 0; w < nw; ++w) q += s_red[w];
                s_quad[m] = q;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float best = INFINITY;
            int bi = 0;
            for (int m = 0; m < M; ++m) {
                const float ll = 0.5f * s_quad[m] - logf(nll_weights[m]);
                if (ll < best) { best = ll; bi = m; }
            }
            s_best = bi;
            s_quad[0] = best;
        }
        __syncthreads();
        prior = s_quad[0];
        if (g_pose != nullptr) {
            const int m = s_best;
            __syncthreads();
            if (threadIdx.x < D) s_d[threadIdx.x] = th[threadIdx.x] - means[(size_t)m * D + threadIdx.x];
            __syncthreads();
            const float* P = precisions + (size_t)m * D * D;
            // g = 0.5 (P + P^T) d, as autograd yields for the einsum form
            for (int i = warp; i < D; i += nw) {
                float t = 0.f;
                for (int j = lane; j < D; j += 32)
                    t = fmaf(P[(size_t)i * D + j] + P[(size_t)j * D + i], s_d[j], t);
                t = warp_sum(t);
                if (lane == 0) g_pose[(size_t)b * D + i] = wp * 0.5f * t;
            }
        }
    } else if (g_pose != nullptr) {
        for (int i = threadIdx.x; i < D; i += PT_THREADS) g_pose[(size_t)b * D + i] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = wp * prior;
        if (wa != 0.f) {
            // body-pose entries 55-3, 58-3, 12-3, 15-3 with signs (+,-,-,-)   (losses.py:161-162)
            const int idx[4] = {52, 55, 9, 12};
            const float sg[4] = {1.f, -1.f, -1.f, -1.f};
            for (int k = 0; k < 4; ++k) {
                const float e = expf(sg[k] * th[idx[k]]);
                v = fmaf(wa, e * e, v);
                if (g_pose != nullptr) g_pose[(size_t)b * D + idx[k]] += wa * 2.f * e * e * sg[k];
            }
        }
        if (betas != nullptr) {
            float s = 0.f;
            for (int l = 0; l < L; ++l) {
                const float be = betas[(size_t)b * L + l];
                s = fmaf(be, be, s);
                if (g_betas != nullptr) g_betas[(size_t)b * L + l] = ws * 2.f * be;
            }
            v = fmaf(ws, s, v);
        }
        if (value != nullptr) value[b] = v;
        if (prior_value != nullptr) prior_value[b] = prior;
        if (which != nullptr) which[b] = (wp != 0.f && M > 0) ? s_best : -1;
    }
}

// ------------------------------------------------------------------------------------------
// push / pull contact terms of one body (one CTA):
//   d_i = |p_i - p_argmin(i)|
//   push  a_in  tanh(d / s_in)^2   for interior points                  (1.0, 0.04)
//   pull  a_out tanh(d / s_out)^2  for exterior points                  (0.005, 0.005)
//         mode PULL_THRESHOLD: only where d < euclthres                 losses.py:99-104
//         mode PULL_ALL:       every exterior point                     loss.py:306-308
//   reduce REDUCE_SUM:  loss = sum push + sum pull                      losses.py:105, loss.py:315
//          REDUCE_MEAN: loss = mean push + mean pull (empty -> 0)        eft/loss.py:158-166
// gradient (scaled by weight * g_loss[b]) is scattered into g_points (two rows per term: the point and its
// nearest partner) through 64-bit fixed-point accumulators: exact sums, reproducible run to run.  counts[b] (optional) = valid points of body b.
// ------------------------------------------------------------------------------------------
constexpr int CL_THREADS = 512;

__global__ void __launch_bounds__(CL_THREADS)
contact_loss_kernel(const float* __restrict__ points, const int* __restrict__ argmin,
                    const uint8_t* __restrict__ exterior, const uint8_t* __restrict__ body_active,
                    const int* __restrict__ counts, int N, float euclthres, int pull_mode, int reduce_mode,
                    float weight, const float* __restrict__ g_loss, float* __restrict__ loss,
                    float* __restrict__ parts, float* __restrict__ g_points, long long* __restrict__ g_fix) {
    __shared__ float s_red[CL_THREADS / 32];
    const int b = blockIdx.x;
    if (body_active != nullptr && !body_active[b]) {
        if (threadIdx.x == 0) {
            if (loss != nullptr) loss[b] = 0.f;
            if (parts != nullptr) { parts[4 * b] = parts[4 * b + 1] = parts[4 * b + 2] = parts[4 * b + 3] = 0.f; }
        }
        return;
    }
    const int n = counts != nullptr ? min(counts[b], N) : N;
    const float* p = points + (size_t)b * N * 3;
    const int* am = argmin + (size_t)b * N;
    const uint8_t* ext = exterior + (size_t)b * N;

    float push = 0.f, pull = 0.f, n_push = 0.f, n_pull = 0.f;
    for (int i = threadIdx.x; i < n; i += CL_THREADS) {
        const int j0 = am[i], j = j0 < 0 ? i : j0;       // -1: no allowed vertex within the query's limit = infinitely far
        const float dx = p[3 * i] - p[3 * j], dy = p[3 * i + 1] - p[3 * j + 1], dz = p[3 * i + 2] - p[3 * j + 2];
        const float d = j0 < 0 ? INFINITY : sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        if (!ext[i]) {
            const float t = tanhf(d / 0.04f);
            push += t * t; n_push += 1.f;
        } else if (pull_mode == PULL_ALL || d < euclthres) {
            const float t = tanhf(d / 0.005f);
            pull += 0.005f * (t * t); n_pull += 1.f;
        }
    }
    push = block_sum(push, s_red);
    pull = block_sum(pull, s_red);
    n_push = block_sum(n_push, s_red);
    n_pull = block_sum(n_pull, s_red);
    float w_push = 1.f, w_pull = 1.f;
    if (reduce_mode == REDUCE_MEAN) {
        w_push = n_push > 0.f ? 1.f / n_push : 0.f;
        w_pull = n_pull > 0.f ? 1.f / n_pull : 0.f;
    }
    if (threadIdx.x == 0) {
        if (loss != nullptr) loss[b] = push * w_push + pull * w_pull;
        if (parts != nullptr) { parts[4 * b] = push; parts[4 * b + 1] = pull; parts[4 * b + 2] = n_push; parts[4 * b + 3] = n_pull; }
    }
    if (g_points == nullptr) return;
    const float up = weight * (g_loss != nullptr ? g_loss[b] : 1.f);
    if (up == 0.f) return;
    float* g = g_points + (size_t)b * N * 3;
    // the scatter goes through 64-bit fixed-point accumulators (exact, order-independent); this CTA owns the
    // body, so it zeroes them, scatters, and folds them into g_points itself
    long long* gf = g_fix + (size_t)b * N * 3;
    for (int k = threadIdx.x; k < 3 * N; k += CL_THREADS) gf[k] = 0;      // partners may lie beyond counts[b]
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += CL_THREADS) {
        const int j0 = am[i], j = j0 < 0 ? i : j0;       // -1: no allowed vertex within the query's limit = infinitely far
        const float dx = p[3 * i] - p[3 * j], dy = p[3 * i + 1] - p[3 * j + 1], dz = p[3 * i + 2] - p[3 * j + 2];
        const float d = j0 < 0 ? INFINITY : sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        float a, s, w;
        if (!ext[i]) { a = 1.f; s = 0.04f; w = w_push; }
        else if (pull_mode == PULL_ALL || d < euclthres) { a = 0.005f; s = 0.005f; w = w_pull; }
        else continue;
        if (!(d > 0.f)) continue;                       // torch.norm backward is 0 at 0
        const float t = tanhf(d / s);
        // d/dd [a tanh(d/s)^2] = 2 a t (1 - t^2) / s ;  d d / d p_i = (p_i - p_j) / d
        const float c = up * w * 2.f * a * t * (1.f - t * t) / (s * d);
        if (c == 0.f) continue;
        fix_add(&gf[3 * i], &g[3 * i], c * dx); fix_add(&gf[3 * i + 1], &g[3 * i + 1], c * dy);
        fix_add(&gf[3 * i + 2], &g[3 * i + 2], c * dz);
        fix_add(&gf[3 * j], &g[3 * j], -c * dx); fix_add(&gf[3 * j + 1], &g[3 * j + 1], -c * dy);
        fix_add(&gf[3 * j + 2], &g[3 * j + 2], -c * dz);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * N; k += CL_THREADS) {
        const long long a = gf[k];
        if (a != 0) g[k] += fix_value(a);
    }
}

// REDUCE_SUM form spread over gridDim.x CTAs per body (the per-point weights do not depend on the body's
// totals): chunk x of body y adds its terms into partials[y][x][4] and scatters its gradient rows into the
// fixed-point accumulators (zeroed by the launcher); contact_loss_fold_kernel then sums the partials in chunk
// order and folds the accumulators into g_points.  Same values as contact_loss_kernel up to the order of the
// fp32 loss sum; the gradient is bit-identical (exact accumulation).
__global__ void __launch_bounds__(CL_THREADS)
contact_loss_chunk_kernel(const float* __restrict__ points, const int* __restrict__ argmin,
                          const uint8_t* __restrict__ exterior, const uint8_t* __restrict__ body_active,
                          const int* __restrict__ counts, int N, float euclthres, int pull_mode, float weight,
                          const float* __restrict__ g_loss, float* __restrict__ partials,
                          float* __restrict__ g_points, long long* __restrict__ g_fix) {
    __shared__ float s_red[CL_THREADS / 32];
    const int b = blockIdx.y;
    float* part = partials + ((size_t)b * gridDim.x + blockIdx.x) * 4;
    if (body_active != nullptr && !body_active[b]) {
        if (threadIdx.x < 4) part[threadIdx.x] = 0.f;
        return;
    }
    const int n = counts != nullptr ? min(counts[b], N) : N;
    const float* p = points + (size_t)b * N * 3;
    const float up = g_points != nullptr ? weight * (g_loss != nullptr ? g_loss[b] : 1.f) : 0.f;
    float* g = g_points != nullptr ? g_points + (size_t)b * N * 3 : nullptr;
    long long* gf = g_fix != nullptr ? g_fix + (size_t)b * N * 3 : nullptr;

    float push = 0.f, pull = 0.f, n_push = 0.f, n_pull = 0.f;
    for (int i = blockIdx.x * CL_THREADS + threadIdx.x; i < n; i += gridDim.x * CL_THREADS) {
        const int j0 = argmin[(size_t)b * N + i], j = j0 < 0 ? i : j0;   // -1: nothing within the query's limit
        const float dx = p[3 * i] - p[3 * j], dy = p[3 * i + 1] - p[3 * j + 1], dz = p[3 * i + 2] - p[3 * j + 2];
        const float d = j0 < 0 ? INFINITY : sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        float a, s, t;
        if (!exterior[(size_t)b * N + i]) {
            a = 1.f; s = 0.04f; t = tanhf(d / 0.04f);
            push += t * t; n_push += 1.f;
        } else if (pull_mode == PULL_ALL || d < euclthres) {
            a = 0.005f; s = 0.005f; t = tanhf(d / 0.005f);
            pull += 0.005f * (t * t); n_pull += 1.f;
        } else continue;
        if (up == 0.f || !(d > 0.f)) continue;          // torch.norm backward is 0 at 0
        const float c = up * 2.f * a * t * (1.f - t * t) / (s * d);
        if (c == 0.f) continue;
        fix_add(&gf[3 * i], &g[3 * i], c * dx); fix_add(&gf[3 * i + 1], &g[3 * i + 1], c * dy);
        fix_add(&gf[3 * i + 2], &g[3 * i + 2], c * dz);
        fix_add(&gf[3 * j], &g[3 * j], -c * dx); fix_add(&gf[3 * j + 1], &g[3 * j + 1], -c * dy);
        fix_add(&gf[3 * j + 2], &g[3 * j + 2], -c * dz);
    }
    push = block_sum(push, s_red);
    pull = block_sum(pull, s_red);
    n_push = block_sum(n_push, s_red);
    n_pull = block_sum(n_pull, s_red);
    if (threadIdx.x == 0) { part[0] = push; part[1] = pull; part[2] = n_push; part[3] = n_pull; }
}

__global__ void __launch_bounds__(CL_THREADS)
contact_loss_fold_kernel(const float* __restrict__ partials, int chunks, int N, float* __restrict__ loss,
                         float* __restrict__ parts, float* __restrict__ g_points,
                         const long long* __restrict__ g_fix) {
    const int b = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        float v = 0.f;
        for (int c = 0; c < chunks; ++c) v += partials[((size_t)b * chunks + c) * 4 + threadIdx.x];
        if (parts != nullptr) parts[4 * b + threadIdx.x] = v;
        const float other = __shfl_xor_sync(0xfu, v, 1);
        if (threadIdx.x == 0 && loss != nullptr) loss[b] = v + other;
    }
    if (g_points == nullptr) return;
    float* g = g_points + (size_t)b * N * 3;
    const long long* gf = g_fix + (size_t)b * N * 3;
    for (int k = blockIdx.x * CL_THREADS + threadIdx.x; k < 3 * N; k += gridDim.x * CL_THREADS) {
        const long long a = gf[k];
        if (a != 0) g[k] += fix_value(a);
    }
}

// ------------------------------------------------------------------------------------------
// region-to-region: r2r[b] = sum over active pairs of min_sq[b,p] (losses.py:116-117) and the
// gradient of the attaining entry P[i,j] = |x_i|^2 + |x_j|^2 - 2 x_i.x_j (contact.py:42):
// dP/dx_i = 2 x_i - 2 x_j, dP/dx_j = 2 x_j - 2 x_i.  A fully masked pair has min = +inf and, as
// in the reference (the inf was written by index_put_), no gradient.
// one CTA per body.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
region_sum_kernel(const float* __restrict__ verts, int V, int n_pairs, const float* __restrict__ min_sq,
                  const int* __restrict__ arg_i, const int* __restrict__ arg_j,
                  const uint8_t* __restrict__ body_active, float weight, const float* __restrict__ g_loss,
                  float* __restrict__ r2r, float* __restrict__ g_verts) {
    __shared__ float s_red[4];
    const int b = blockIdx.x;
    const bool on = body_active == nullptr || body_active[b];
    float acc = 0.f;
    const float up = weight * (g_loss != nullptr ? g_loss[b] : 1.f);
    const float* vb = verts + (size_t)b * V * 3;
    if (on) {
        for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
            const size_t o = (size_t)b * n_pairs + p;
            if (arg_i[o] < 0) continue;                 // pair not annotated for this body
            acc += min_sq[o];
        }
        // gradient of the attaining entries: a handful of pairs per body, added by one thread in class
        // order (two pairs may share a vertex; a fixed order keeps the sum reproducible)
        if (threadIdx.x == 0 && g_verts != nullptr && up != 0.f) {
            float* g = g_verts + (size_t)b * V * 3;
            for (int p = 0; p < n_pairs; ++p) {
                const size_t o = (size_t)b * n_pairs + p;
                const int i = arg_i[o], j = arg_j[o];
                if (i < 0 || i == j || isinf(min_sq[o])) continue;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float d = 2.f * vb[3 * i + k] - 2.f * vb[3 * j + k];
                    g[3 * i + k] += up * d;
                    g[3 * j + k] -= up * d;
                }
            }
        }
    }
    // fixed-order sum over pairs (the reference accumulates `mindists +=` in class order)
    acc = block_sum(acc, s_red);
    if (threadIdx.x == 0 && r2r != nullptr) r2r[b] = on ? acc : 0.f;
}

// ------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, no weight decay / amsgrad).  `step` lives on the device so that a
// captured CUDA graph can be replayed: the kernel with advance != 0 increments it first.
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, long long n, const int* __restrict__ step_dev, int step_add,
                            double lr, double beta1, double beta2, double eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = *step_dev + step_add;
    const float g = grad[i];
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2);
    const float mi = fmaf(w1, g - m[i], m[i]);                    // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(w2, g * g, b2 * v[i]);                  // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi; v[i] = vi;
    // the bias corrections are Python doubles in torch.optim.Adam
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    const float denom = sqrtf(vi) / (float)sqrt(bc2) + (float)eps;
    param[i] = fmaf(-(float)(lr / bc1), mi / denom, param[i]);
}

__global__ void step_advance_kernel(int* step_dev, int add) { *step_dev += add; }

// total[0] += sum_b (sum_j rep[b,j] + w_terms * terms[b] + w_contact * contact[b] + w_r2r * r2r[b] + extra[b])
__global__ void __launch_bounds__(256)
combine_kernel(const float* __restrict__ rep, int J, const float* __restrict__ terms, const float* __restrict__ contact,
               float w_contact, const float* __restrict__ r2r, float w_r2r, const float* __restrict__ extra, int B,
               float* __restrict__ per_body, float* __restrict__ total) {
    __shared__ float s_red[8];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float v = 0.f;
        if (rep != nullptr) for (int j = 0; j < J; ++j) v += rep[(size_t)b * J + j];
        if (contact != nullptr) v = fmaf(w_contact, contact[b], v);
        if (terms != nullptr) v += terms[b];
        if (r2r != nullptr && w_r2r != 0.f) v = fmaf(w_r2r, r2r[b], v);
        if (extra != nullptr) v += extra[b];
        if (per_body != nullptr) per_body[b] = v;
        acc += v;
    }
    acc = block_sum(acc, s_red);
    if (threadIdx.x == 0 && total != nullptr) *total = acc;
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
int launch_reprojection(const float* joints, const float* cam_t, const float* center, const float* joints_2d,
                        const float* conf, int B, int J, float focal, float sigma, const float* cam_t_est,
                        float depth_weight, const float* g_loss, float* loss, float* extra, float* g_joints,
                        float* g_cam_t, cudaStream_t st) {
    KernelTimer timer("objective_kernels", st);
    if (B == 0) return 0;
    reprojection_kernel<<<B, 64, 0, st>>>(joints, cam_t, center, joints_2d, conf, J, focal, sigma, cam_t_est,
                                          depth_weight, g_loss, loss, extra, g_joints, g_cam_t);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_pose_terms(const float* means, const float* precisions, const float* nll_weights, int M, int D,
                      const float* pose, const float* betas, int L, int B, float wp, float wa, float ws,
                      float* value, float* prior_value, int* which, float* g_pose, float* g_betas,
                      cudaStream_t st) {
    KernelTimer timer("objective_kernels", st);
    if (B == 0) return 0;
    TUCH_REQUIRE(D <= PT_MAXD && M <= PT_MAXM, "pose prior: D=%d (max %d) or M=%d (max %d) too large", D, PT_MAXD, M, PT_MAXM);
    TUCH_REQUIRE(wa == 0.f || D >= 56, "angle prior needs a 69-dimensional body pose (got D=%d)", D);
    pose_terms_kernel<<<B, PT_THREADS, 0, st>>>(means, precisions, nll_weights, M, D, pose, betas, L, wp, wa, ws,
                                                value, prior_value, which, g_pose, g_betas);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_contact_loss(const float* points, const int* argmin, const uint8_t* exterior,
                        const uint8_t* body_active, const int* counts, int B, int N, float euclthres,
                        int pull_mode, int reduce_mode, float weight, const float* g_loss, float* loss,
                        float* parts, float* g_points, cudaStream_t st) {
    KernelTimer timer("objective_kernels", st);
    if (B == 0) return 0;
    const size_t fix_bytes = g_points != nullptr ? sizeof(long long) * 3 * (size_t)B * N : 0;
    const int chunks = reduce_mode == REDUCE_SUM ? std::min((N + CL_THREADS - 1) / CL_THREADS, 16) : 1;
    const size_t part_bytes = chunks > 1 ? sizeof(float) * 4 * (size_t)B * chunks : 0;
    char* scratch = nullptr;
    if (fix_bytes + part_bytes > 0) {
        void* p = nullptr;
        if (int rc = arena_get(st, fix_bytes + part_bytes, &p, 3)) return rc;
        scratch = (char*)p;
    }
    long long* g_fix = g_points != nullptr ? (long long*)scratch : nullptr;   // fixed-point scatter accumulators
    if (chunks > 1) {
        float* partials = (float*)(scratch + fix_bytes);
        if (g_fix != nullptr) TUCH_CUDA(cudaMemsetAsync(g_fix, 0, fix_bytes, st));
        contact_loss_chunk_kernel<<<dim3(chunks, B), CL_THREADS, 0, st>>>(points, argmin, exterior, body_active, counts, N,
                                                                          euclthres, pull_mode, weight, g_loss, partials,
                                                                          g_points, g_fix);
        TUCH_LAUNCH_CHECK(); count_launch();
        contact_loss_fold_kernel<<<dim3(chunks, B), CL_THREADS, 0, st>>>(partials, chunks, N, loss, parts, g_points, g_fix);
        TUCH_LAUNCH_CHECK(); count_launch();
        return 0;
    }
    contact_loss_kernel<<<B, CL_THREADS, 0, st>>>(points, argmin, exterior, body_active, counts, N, euclthres,
                                                  pull_mode, reduce_mode, weight, g_loss, loss, parts, g_points, g_fix);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_region_sum(const float* verts, int B, int V, int n_pairs, const float* min_sq, const int* arg_i,
                      const int* arg_j, const uint8_t* body_active, float weight, const float* g_loss,
                      float* r2r, float* g_verts, cudaStream_t st) {
    KernelTimer timer("objective_kernels", st);
    if (B == 0) return 0;
    region_sum_kernel<<<B, 128, 0, st>>>(verts, V, n_pairs, min_sq, arg_i, arg_j, body_active, weight, g_loss,
                                         r2r, g_verts);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_adam(float* param, const float* grad, float* m, float* v, long long n, const int* step_dev,
                int step_add, double lr, double beta1, double beta2, double eps, cudaStream_t st) {
    KernelTimer timer("adam_kernels", st);
    if (n == 0) return 0;
    adam_kernel<<<cdiv(n, 256), 256, 0, st>>>(param, grad, m, v, n, step_dev, step_add, lr, beta1, beta2, eps);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_step_advance(int* step_dev, int add, cudaStream_t st) {
    step_advance_kernel<<<1, 1, 0, st>>>(step_dev, add);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_combine(const float* rep, int J, const float* terms, const float* contact, float w_contact,
                   const float* r2r, float w_r2r, const float* extra, int B, float* per_body, float* total,
                   cudaStream_t st) {
    combine_kernel<<<1, 256, 0, st>>>(rep, J, terms, contact, w_contact, r2r, w_r2r, extra, B, per_body, total);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

}  // namespace tuch
