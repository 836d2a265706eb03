// Self-contact geometry kernels (sm_100a): triangle/vertex repacking, generalized winding
// numbers, geodesically-masked nearest vertex, dense pairwise distances, solid angles and
// region-pair minima.  Replaces the reference's materialising tensor algebra in
// tuch/utils/contact.py:23-147 and the masked argmin of tuch/smplify/losses.py:92-93.
//
// Layout in HBM
//   tri12 [B][Fp]   3 x float4 per triangle (corner a | b | c, w unused), Fp = F rounded up to
//                   WN_TILE_F; padding triangles are all-zero and contribute exactly 0.
//   vert4 [B][Vp]   float4 (x, y, z, |v|^2), Vp = V rounded up to 32.
//   maskT [W][Vq]   bit-packed geodesic mask, W = ceil(V/32): bit k of maskT[w][c] is
//                   geomask[32w + k][c]  (candidate row r = 32w+k, query column c), so a warp of
//                   consecutive queries reads consecutive words.
#include "kernels.h"

namespace tuch {

// ------------------------------------------------------------------------------------------
// repacking
// ------------------------------------------------------------------------------------------
__global__ void pack_mesh_kernel(const float* __restrict__ verts, const int* __restrict__ faces,
                                 int V, int F, int Fp, int Vp,
                                 float4* __restrict__ tri12, float4* __restrict__ vert4) {
    const int b = blockIdx.y;
    const float* vb = verts + (size_t)b * V * 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (tri12 != nullptr && i < Fp) {
        float4 A = make_float4(0, 0, 0, 0), Bv = A, C = A;
        if (i < F) {
            const int i0 = faces[3 * i], i1 = faces[3 * i + 1], i2 = faces[3 * i + 2];
            A = make_float4(vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2], 0.f);
            Bv = make_float4(vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2], 0.f);
            C = make_float4(vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2], 0.f);
        }
        float4* t = tri12 + ((size_t)b * Fp + i) * 3;
        t[0] = A; t[1] = Bv; t[2] = C;
    }
    if (vert4 != nullptr && i < Vp) {
        float4 o = make_float4(0, 0, 0, 0);
        if (i < V) {
            const float x = vb[3 * i], y = vb[3 * i + 1], z = vb[3 * i + 2];
            // |v|^2 accumulated like a K=3 GEMM diagonal (contact.py:27,37)
            o = make_float4(x, y, z, fmaf(z, z, fmaf(y, y, x * x)));
        }
        vert4[(size_t)b * Vp + i] = o;
    }
}

__global__ void pack_triangles_kernel(const float* __restrict__ tris, int F, int Fp,
                                      float4* __restrict__ tri12) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Fp) return;
    float4 A = make_float4(0, 0, 0, 0), Bv = A, C = A;
    if (i < F) {
        const float* t = tris + ((size_t)b * F + i) * 9;
        A = make_float4(t[0], t[1], t[2], 0.f);
        Bv = make_float4(t[3], t[4], t[5], 0.f);
        C = make_float4(t[6], t[7], t[8], 0.f);
    }
    float4* o = tri12 + ((size_t)b * Fp + i) * 3;
    o[0] = A; o[1] = Bv; o[2] = C;
}

// ------------------------------------------------------------------------------------------
// generalized winding numbers
// grid (query tiles, F splits, bodies); block WN_THREADS; each thread owns WN_QPT queries and
// streams the body's triangles through a WN_STAGES-deep TMA (cp.async.bulk) ring in shared memory.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WN_THREADS)
winding_kernel(const float4* __restrict__ tri12, const float* __restrict__ points,
               float* __restrict__ partial, int Q, int Fp, int tiles_per_split, long long tri_stride,
               long long point_stride, long long partial_stride, const uint8_t* __restrict__ body_active,
               const int* __restrict__ q_counts) {
    __shared__ __align__(128) float4 s_tri[WN_STAGES][WN_TILE_F * 3];
    __shared__ __align__(8) uint64_t s_bar[WN_STAGES];

    const int b = blockIdx.z;
    if (body_active != nullptr && !body_active[b]) return;     // uniform per CTA
    const int q_stride = Q;                                    // layout of `partial` is [S][Q]
    if (q_counts != nullptr) {                                 // data-dependent query count (HD selection)
        Q = min(Q, q_counts[b]);
        if ((int)blockIdx.x * (WN_THREADS * WN_QPT) >= Q) return;
    }
    const int split = blockIdx.y;
    const int n_tiles_total = Fp / WN_TILE_F;
    const int tile0 = split * tiles_per_split;
    const int n_tiles = min(tiles_per_split, n_tiles_total - tile0);
    const float4* src = tri12 + (size_t)b * tri_stride + (size_t)tile0 * WN_TILE_F * 3;
    constexpr uint32_t TILE_BYTES = WN_TILE_F * 3 * sizeof(float4);

    if (threadIdx.x == 0) {
        for (int s = 0; s < WN_STAGES; ++s) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < WN_STAGES && s < n_tiles; ++s) {
            mbar_expect_tx(&s_bar[s], TILE_BYTES);
            tma_load_1d(s_tri[s], src + (size_t)s * WN_TILE_F * 3, TILE_BYTES, &s_bar[s]);
        }
    }

    float px[WN_QPT], py[WN_QPT], pz[WN_QPT], acc[WN_QPT];
    const float* pb = points + (size_t)b * point_stride;
    const int q0 = blockIdx.x * (WN_THREADS * WN_QPT) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < WN_QPT; ++k) {
        const int q = min(q0 + k * WN_THREADS, Q - 1);
        px[k] = pb[3 * q]; py[k] = pb[3 * q + 1]; pz[k] = pb[3 * q + 2];
        acc[k] = 0.f;
    }

    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % WN_STAGES;
        mbar_wait(&s_bar[s], (t / WN_STAGES) & 1);
        const float4* tile = s_tri[s];
#pragma unroll 2
        for (int f = 0; f < WN_TILE_F; ++f) {
            const float4 A = tile[3 * f], Bv = tile[3 * f + 1], C = tile[3 * f + 2];
#pragma unroll
            for (int k = 0; k < WN_QPT; ++k) acc[k] += half_solid_angle(px[k], py[k], pz[k], A, Bv, C);
        }
        __syncthreads();   // every thread is done with stage s
        if (threadIdx.x == 0 && t + WN_STAGES < n_tiles) {
            mbar_expect_tx(&s_bar[s], TILE_BYTES);
            tma_load_1d(s_tri[s], src + (size_t)(t + WN_STAGES) * WN_TILE_F * 3, TILE_BYTES, &s_bar[s]);
        }
    }

    float* out = partial + (size_t)b * partial_stride + (size_t)split * q_stride;
#pragma unroll
    for (int k = 0; k < WN_QPT; ++k) {
        const int q = q0 + k * WN_THREADS;
        if (q < Q) out[q] = acc[k];
    }
}

// sums the F-split partials in a fixed order and applies 2 / (4 pi)   (contact.py:109,146-147)
__global__ void winding_finalize_kernel(const float* __restrict__ partial, int Q, int S,
                                        long long partial_stride, long long out_stride,
                                        float* __restrict__ winding, const uint8_t* __restrict__ body_active,
                                        const int* __restrict__ q_counts) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    if ((body_active != nullptr && !body_active[b]) || (q_counts != nullptr && q >= q_counts[b])) {
        winding[(size_t)b * out_stride + q] = 0.f;
        return;
    }
    const float* p = partial + (size_t)b * partial_stride + q;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += p[(size_t)s * Q];
    winding[(size_t)b * out_stride + q] = acc * 0.159154943091895336f;   // 1 / (2 pi)
}

// ------------------------------------------------------------------------------------------
// geodesically-masked nearest vertex: for every query column c the first row r minimising
// P[r,c] = (|v_r|^2 + |v_c|^2) - 2 v_r.v_c among rows with geomask[r,c] set
// (losses.py:76,92-93; loss.py:256,269-270).  Fully masked column -> (0, +inf).
// grid (query tiles, bodies); block NN_THREADS; candidates stream through a TMA ring.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NN_THREADS)
nearest_kernel(const float4* __restrict__ vert4, const uint32_t* __restrict__ maskT,
               int V, int Vp, int Vq, int* __restrict__ argmin_out, float* __restrict__ min_out) {
    __shared__ __align__(128) float4 s_v[NN_STAGES][NN_TILE_V];
    __shared__ __align__(8) uint64_t s_bar[NN_STAGES];

    const int b = blockIdx.y;
    const float4* src = vert4 + (size_t)b * Vp;
    const int n_tiles = (Vp + NN_TILE_V - 1) / NN_TILE_V;
    auto tile_bytes = [&](int t) -> uint32_t {
        return (uint32_t)(min(NN_TILE_V, Vp - t * NN_TILE_V) * sizeof(float4));
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NN_STAGES; ++s) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < NN_STAGES && s < n_tiles; ++s) {
            mbar_expect_tx(&s_bar[s], tile_bytes(s));
            tma_load_1d(s_v[s], src + (size_t)s * NN_TILE_V, tile_bytes(s), &s_bar[s]);
        }
    }

    const int c = blockIdx.x * NN_THREADS + threadIdx.x;
    const int cc = min(c, V - 1);
    const float4 q = src[cc];
    float best = INFINITY;
    int bi = 0;

    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % NN_STAGES;
        mbar_wait(&s_bar[s], (t / NN_STAGES) & 1);
        const float4* tile = s_v[s];
        const int r0 = t * NN_TILE_V;
        const int n_words = min(NN_TILE_V, Vp - r0) / 32;
        uint32_t m_next = maskT[(size_t)(r0 / 32) * Vq + cc];
        for (int w = 0; w < n_words; ++w) {
            const uint32_t m = m_next;
            if (w + 1 < n_words) m_next = maskT[(size_t)(r0 / 32 + w + 1) * Vq + cc];
            if (m == 0u) continue;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const float4 v = tile[w * 32 + k];
                const float zz = fmaf(v.z, q.z, fmaf(v.y, q.y, v.x * q.x));
                float p = fmaf(-2.f, zz, v.w + q.w);
                p = ((m >> k) & 1u) ? p : INFINITY;
                if (p < best) { best = p; bi = r0 + w * 32 + k; }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && t + NN_STAGES < n_tiles) {
            mbar_expect_tx(&s_bar[s], tile_bytes(t + NN_STAGES));
            tma_load_1d(s_v[s], src + (size_t)(t + NN_STAGES) * NN_TILE_V, tile_bytes(t + NN_STAGES), &s_bar[s]);
        }
    }
    if (c < V) {
        argmin_out[(size_t)b * V + c] = bi;
        min_out[(size_t)b * V + c] = best;
    }
}

// bool [V][V] (row-major, 1 byte per entry) -> maskT [W][Vq]
__global__ void pack_mask_kernel(const uint8_t* __restrict__ mask, int V, int Vq, int W,
                                 uint32_t* __restrict__ maskT) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (c >= Vq) return;
    uint32_t bits = 0;
    if (c < V) {
        for (int k = 0; k < 32; ++k) {
            const int r = w * 32 + k;
            if (r < V && mask[(size_t)r * V + c]) bits |= (1u << k);
        }
    }
    maskT[(size_t)w * Vq + c] = bits;
}

// float geodesic distances [V][V] + threshold -> maskT (geomask = geodist > thres,
// smplifydc.py:65, loss.py:71)
__global__ void pack_mask_from_dist_kernel(const float* __restrict__ dist, float thres, int V, int Vq,
                                           int W, uint32_t* __restrict__ maskT) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (c >= Vq) return;
    uint32_t bits = 0;
    if (c < V) {
        for (int k = 0; k < 32; ++k) {
            const int r = w * 32 + k;
            if (r < V && dist[(size_t)r * V + c] > thres) bits |= (1u << k);
        }
    }
    maskT[(size_t)w * Vq + c] = bits;
}

// ------------------------------------------------------------------------------------------
// API-parity kernels that materialise their full output (small problem sizes only)
// ------------------------------------------------------------------------------------------
// contact.py:23-47
__global__ void pairwise_dist_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                     int nx, int ny, int squared, float* __restrict__ P) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const float* xi = x + ((size_t)b * nx + i) * 3;
    const float* yj = y + ((size_t)b * ny + j) * 3;
    const float rx = fmaf(xi[2], xi[2], fmaf(xi[1], xi[1], xi[0] * xi[0]));
    const float ry = fmaf(yj[2], yj[2], fmaf(yj[1], yj[1], yj[0] * yj[0]));
    const float zz = fmaf(xi[2], yj[2], fmaf(xi[1], yj[1], xi[0] * yj[0]));
    float p = fmaf(-2.f, zz, rx + ry);
    if (!squared) p = sqrtf(p);
    P[((size_t)b * nx + i) * ny + j] = p;
}

// backward of pairwise_dist w.r.t. x (rows) -- call twice with swapped roles for y.
// gx[i] = sum_j w_ij (2 x_i - 2 y_j), w_ij = gP_ij (squared) or gP_ij / (2 P_ij) (sqrt form).
// transposed != 0 reads gP/P as [ny,nx] (i.e. the y-gradient pass).  One warp per row.
__global__ void pairwise_dist_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                         const float* __restrict__ P, const float* __restrict__ gP,
                                         int nx, int ny, int squared, int transposed, float* __restrict__ gx) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    if (i >= nx) return;
    const float* xi = x + ((size_t)b * nx + i) * 3;
    const float x0 = xi[0], x1 = xi[1], x2 = xi[2];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int j = lane; j < ny; j += 32) {
        const size_t e = transposed ? ((size_t)b * ny + j) * nx + i : ((size_t)b * nx + i) * ny + j;
        float w = gP[e];
        if (!squared) w = w / (2.f * P[e]);
        const float* yj = y + ((size_t)b * ny + j) * 3;
        a0 = fmaf(w, 2.f * x0 - 2.f * yj[0], a0);
        a1 = fmaf(w, 2.f * x1 - 2.f * yj[1], a1);
        a2 = fmaf(w, 2.f * x2 - 2.f * yj[2], a2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        float* g = gx + ((size_t)b * nx + i) * 3;
        g[0] = a0; g[1] = a1; g[2] = a2;
    }
}

// contact.py:49-109 (IEEE sqrt/atan2 here: this entry point exists for API parity, not speed)
__global__ void solid_angles_kernel(const float* __restrict__ points, const float* __restrict__ tris,
                                    int Q, int F, float* __restrict__ out) {
    const int b = blockIdx.z;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y * blockDim.y + threadIdx.y;
    if (q >= Q || f >= F) return;
    const float* p = points + ((size_t)b * Q + q) * 3;
    const float* t = tris + ((size_t)b * F + f) * 9;
    const float ax = t[0] - p[0], ay = t[1] - p[1], az = t[2] - p[2];
    const float bx = t[3] - p[0], by = t[4] - p[1], bz = t[5] - p[2];
    const float cx = t[6] - p[0], cy = t[7] - p[1], cz = t[8] - p[2];
    const float la = sqrtf(ax * ax + ay * ay + az * az);
    const float lb = sqrtf(bx * bx + by * by + bz * bz);
    const float lc = sqrtf(cx * cx + cy * cy + cz * cz);
    const float num = ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
    const float dab = ax * bx + ay * by + az * bz;
    const float dac = ax * cx + ay * cy + az * cz;
    const float dbc = bx * cx + by * cy + bz * cz;
    float den = la * lb * lc;
    den = fmaf(dab, lc, den);
    den = fmaf(dac, lb, den);
    den = fmaf(dbc, la, den);
    out[((size_t)b * Q + q) * F + f] = 2.f * atan2f(num, den);
}

// ------------------------------------------------------------------------------------------
// region-pair minimum: min over idsA x idsB of the (optionally geodesically masked) squared
// expansion-form distance, first flat index on ties (losses.py:113-116, train_module.py:83-90).
// grid (pair slots, bodies); one block per (body, pair).
// ------------------------------------------------------------------------------------------
// The geodesic mask restricted to one annotated pair, bit-packed per row of region A:
// pmask[pair_word_off[p] + a * ceil(nb / 32) + w] bit k = geomask[ia[a]][ib[32 w + k]].  Static per topology
// (regions + mask), so the region minimum reads nb / 32 words per row instead of nb scattered mask words.
__global__ void pair_mask_kernel(const uint32_t* __restrict__ maskT, int Vq, const int* __restrict__ region_ids,
                                 const int* __restrict__ region_off, const int* __restrict__ pair_a,
                                 const int* __restrict__ pair_b, const long long* __restrict__ pair_word_off,
                                 uint32_t* __restrict__ pmask) {
    const int p = blockIdx.y;
    const int ra = pair_a[p], rb = pair_b[p];
    const int* ia = region_ids + region_off[ra];
    const int* ib = region_ids + region_off[rb];
    const int na = region_off[ra + 1] - region_off[ra], nb = region_off[rb + 1] - region_off[rb];
    const int nw = (nb + 31) / 32;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)na * nw) return;
    const int a = (int)(e / nw), w = (int)(e % nw);
    const int i = ia[a];
    uint32_t bits = 0;
    for (int k = 0; k < 32 && 32 * w + k < nb; ++k) {
        const int j = ib[32 * w + k];
        bits |= ((maskT[(size_t)(i >> 5) * Vq + j] >> (i & 31)) & 1u) << k;
    }
    pmask[pair_word_off[p] + e] = bits;
}

__global__ void __launch_bounds__(RM_THREADS)
region_min_kernel(const float4* __restrict__ vert4, int Vp, const uint32_t* __restrict__ maskT, int Vq,
                  const int* __restrict__ region_ids, const int* __restrict__ region_off,
                  const int* __restrict__ pair_a, const int* __restrict__ pair_b,
                  const uint8_t* __restrict__ active, int n_pairs, const uint32_t* __restrict__ pmask,
                  const long long* __restrict__ pair_word_off,
                  float* __restrict__ min_out, int* __restrict__ arg_i, int* __restrict__ arg_j) {
    const int b = blockIdx.y, pidx = blockIdx.x;
    const size_t o = (size_t)b * n_pairs + pidx;
    if (active != nullptr && !active[o]) {
        if (threadIdx.x == 0) { min_out[o] = 0.f; arg_i[o] = -1; arg_j[o] = -1; }
        return;
    }
    const int ra = pair_a[pidx], rb = pair_b[pidx];
    const int* ia = region_ids + region_off[ra];
    const int* ib = region_ids + region_off[rb];
    const int na = region_off[ra + 1] - region_off[ra], nb = region_off[rb + 1] - region_off[rb];
    const float4* v = vert4 + (size_t)b * Vp;
    float best = INFINITY;
    unsigned long long best_flat = 0ull;     // ties and the all-masked case resolve to flat index 0
    bool have = false;
    // region B is staged through shared memory in chunks; every thread owns rows a = tid, tid + T, ...
    // and walks the chunk in ascending c, so its running minimum is the first one in flat order
    __shared__ float4 s_vb[RM_THREADS];
    __shared__ int s_jb[RM_THREADS];
    for (int c0 = 0; c0 < nb; c0 += RM_THREADS) {
        const int nc = min(RM_THREADS, nb - c0);
        __syncthreads();
        if ((int)threadIdx.x < nc) {
            const int j = ib[c0 + threadIdx.x];
            s_jb[threadIdx.x] = j;
            s_vb[threadIdx.x] = v[j];
        }
        __syncthreads();
        for (int a = threadIdx.x; a < na; a += RM_THREADS) {
            const int i = ia[a];
            const float4 x = v[i];
            const uint32_t* mrow = maskT != nullptr ? maskT + (size_t)(i >> 5) * Vq : nullptr;
            const int sh = i & 31;
            // packed row of the pair mask when the topology has it (c0 is a multiple of 32)
            const uint32_t* prow = (maskT != nullptr && pmask != nullptr)
                                       ? pmask + pair_word_off[pidx] + (long long)a * ((nb + 31) / 32) + c0 / 32 : nullptr;
            float rbest = INFINITY;
            int rc = -1;
            uint32_t bits = 0;
            for (int c = 0; c < nc; ++c) {
                const float4 y = s_vb[c];
                const float zz = fmaf(x.z, y.z, fmaf(x.y, y.y, x.x * y.x));
                float p = fmaf(-2.f, zz, x.w + y.w);
                if (prow != nullptr) {
                    if ((c & 31) == 0) bits = prow[c >> 5];
                    if (!((bits >> (c & 31)) & 1u)) p = INFINITY;                             // geomask[i][j]
                } else if (mrow != nullptr && !((mrow[s_jb[c]] >> sh) & 1u)) p = INFINITY;
                if (rc < 0 || p < rbest) { rbest = p; rc = c; }
            }
            const unsigned long long flat = (unsigned long long)a * nb + (unsigned long long)(c0 + rc);
            if (!have || rbest < best || (rbest == best && flat < best_flat)) { best = rbest; best_flat = flat; have = true; }
        }
    }
    // block argmin with lowest-flat-index tie-break
    __shared__ float s_val[RM_THREADS];
    __shared__ unsigned long long s_idx[RM_THREADS];
    s_val[threadIdx.x] = have ? best : INFINITY;
    s_idx[threadIdx.x] = have ? best_flat : ~0ull;
    __syncthreads();
    for (int st = RM_THREADS / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            const float ov = s_val[threadIdx.x + st];
            const unsigned long long oi = s_idx[threadIdx.x + st];
            const float mv = s_val[threadIdx.x];
            const unsigned long long mi = s_idx[threadIdx.x];
            if (ov < mv || (ov == mv && oi < mi)) { s_val[threadIdx.x] = ov; s_idx[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long e = s_idx[0];
        const int a = (int)(e / nb), c = (int)(e - (unsigned long long)a * nb);
        min_out[o] = s_val[0];
        arg_i[o] = ia[a];
        arg_j[o] = ib[c];
    }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
int winding_splits(int B, int Q, int Fp, int sm_count) {
    const int qtiles = cdiv(Q, WN_THREADS * WN_QPT);
    const int n_tiles = Fp / WN_TILE_F;
    const long long want = (long long)sm_count * 8;            // >= 8 CTAs per SM in flight
    int S = (int)((want + (long long)qtiles * B - 1) / ((long long)qtiles * B));
    S = max(1, min(S, n_tiles));
    const int per = cdiv(n_tiles, S);
    return cdiv(n_tiles, per);                                  // no empty split
}

int launch_pack_mesh(const float* verts, const int* faces, int B, int V, int F, int Fp, int Vp,
                     float4* tri12, float4* vert4, cudaStream_t st) {
    const int n = max(tri12 ? Fp : 0, vert4 ? Vp : 0);
    dim3 grid(cdiv(n, 256), B);
    pack_mesh_kernel<<<grid, 256, 0, st>>>(verts, faces, V, F, Fp, Vp, tri12, vert4);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_pack_triangles(const float* tris, int B, int F, int Fp, float4* tri12, cudaStream_t st) {
    dim3 grid(cdiv(Fp, 256), B);
    pack_triangles_kernel<<<grid, 256, 0, st>>>(tris, F, Fp, tri12);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_winding(const WindingJob& j, cudaStream_t st) {
    if (j.B == 0 || j.Q == 0) return 0;
    const int n_tiles = j.Fp / WN_TILE_F;
    const int per = cdiv(n_tiles, j.S);
    dim3 grid(cdiv(j.Q, WN_THREADS * WN_QPT), j.S, j.B);
    {
        KernelTimer timer(j.body_active == nullptr && j.Q >= 1024 ? "winding_kernel" : "winding_kernel_segments", st);
        winding_kernel<<<grid, WN_THREADS, 0, st>>>(j.tri12, j.points, j.partial, j.Q, j.Fp, per, j.tri_stride,
                                                    j.point_stride, (long long)j.S * j.Q, j.body_active, j.q_counts);
    }
    TUCH_LAUNCH_CHECK(); count_launch();
    return launch_winding_finalize(j.partial, j.B, j.Q, j.S, j.out_stride, j.winding, j.body_active, j.q_counts, st);
}

int launch_winding_finalize(const float* partial, int B, int Q, int S, long long out_stride, float* winding,
                            const uint8_t* body_active, const int* q_counts, cudaStream_t st) {
    dim3 g2(cdiv(Q, 256), B);
    winding_finalize_kernel<<<g2, 256, 0, st>>>(partial, Q, S, (long long)S * Q, out_stride, winding, body_active,
                                                q_counts);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_nearest(const float4* vert4, const uint32_t* maskT, int B, int V, int Vp, int Vq,
                   int* argmin, float* minval, cudaStream_t st) {
    dim3 grid(cdiv(V, NN_THREADS), B);
    KernelTimer timer("nearest_kernel", st);
    nearest_kernel<<<grid, NN_THREADS, 0, st>>>(vert4, maskT, V, Vp, Vq, argmin, minval);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_pack_mask(const uint8_t* mask, const float* dist, float thres, int V, int Vq, int W,
                     uint32_t* maskT, cudaStream_t st) {
    dim3 grid(cdiv(Vq, 128), W);
    if (mask != nullptr) pack_mask_kernel<<<grid, 128, 0, st>>>(mask, V, Vq, W, maskT);
    else pack_mask_from_dist_kernel<<<grid, 128, 0, st>>>(dist, thres, V, Vq, W, maskT);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_pairwise_dist(const float* x, const float* y, int bs, int nx, int ny, int squared, float* P,
                         cudaStream_t st) {
    dim3 block(32, 8), grid(cdiv(ny, 32), cdiv(nx, 8), bs);
    pairwise_dist_kernel<<<grid, block, 0, st>>>(x, y, nx, ny, squared, P);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_pairwise_dist_bwd(const float* x, const float* y, const float* P, const float* gP, int bs, int nx,
                             int ny, int squared, float* gx, float* gy, cudaStream_t st) {
    if (gx != nullptr && nx > 0) {
        dim3 grid(cdiv(nx, 8), bs);
        pairwise_dist_bwd_kernel<<<grid, 256, 0, st>>>(x, y, P, gP, nx, ny, squared, 0, gx);
        TUCH_LAUNCH_CHECK(); count_launch();
    }
    if (gy != nullptr && ny > 0) {
        dim3 grid(cdiv(ny, 8), bs);
        pairwise_dist_bwd_kernel<<<grid, 256, 0, st>>>(y, x, P, gP, ny, nx, squared, 1, gy);
        TUCH_LAUNCH_CHECK(); count_launch();
    }
    return 0;
}

int launch_solid_angles(const float* points, const float* tris, int bs, int Q, int F, float* out,
                        cudaStream_t st) {
    dim3 block(32, 8), grid(cdiv(F, 32), cdiv(Q, 8), bs);
    solid_angles_kernel<<<grid, block, 0, st>>>(points, tris, Q, F, out);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_pair_mask(const uint32_t* maskT, int Vq, const int* region_ids, const int* region_off, const int* pair_a,
                     const int* pair_b, const long long* pair_word_off, int n_pairs, long long max_words_per_pair,
                     uint32_t* pmask, cudaStream_t st) {
    if (n_pairs == 0 || max_words_per_pair == 0) return 0;
    dim3 grid(cdiv(max_words_per_pair, 128), n_pairs);
    pair_mask_kernel<<<grid, 128, 0, st>>>(maskT, Vq, region_ids, region_off, pair_a, pair_b, pair_word_off, pmask);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_region_min(const float4* vert4, int Vp, const uint32_t* maskT, int Vq, const int* region_ids,
                      const int* region_off, const int* pair_a, const int* pair_b, const uint8_t* active,
                      int n_pairs, int B, const uint32_t* pmask, const long long* pair_word_off, float* min_out,
                      int* arg_i, int* arg_j, cudaStream_t st) {
    if (n_pairs == 0 || B == 0) return 0;
    KernelTimer timer("region_min_kernel", st);
    dim3 grid(n_pairs, B);
    region_min_kernel<<<grid, RM_THREADS, 0, st>>>(vert4, Vp, maskT, Vq, region_ids, region_off, pair_a,
                                                   pair_b, active, n_pairs, pmask, pair_word_off, min_out, arg_i, arg_j);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

}  // namespace tuch
