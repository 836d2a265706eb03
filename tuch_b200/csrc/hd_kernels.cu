// HD-point contact path of the regressor loss (tuch/train/loss.py:274-315), batched and without host
// synchronisation.  The reference, per body: builds a [3F x n] and a [N_hd x n_f] equality matrix to
// find the faces / HD points touching in-contact vertices (:278-281), multiplies a dense
// [n_sel, 6890] regressor slice (:285), materialises an n_sel^2 distance matrix (:288-291) and runs
// the solid-angle tensor algebra on [n_sel x F] (:297).  Here:
//   hd_select_kernel   vertex -> face -> HD-point selection + order-preserving compaction per body
//   hd_gather_kernel   selected HD points from the (3-sparse) regressor rows, their 1 mm normal offset
//                      copies and the geodesic proxy vertex of each point
//   hd_nearest_kernel  masked nearest selected HD point (mask looked up through the proxy vertices)
//   winding_kernel     (contact_kernels.cu) with per-body query counts
//   hd_scatter_kernel  backward: gradient of the selected HD points back onto the mesh vertices
#include "hd_internal.h"

namespace tuch {

constexpr int HS_THREADS = 1024;

// one CTA per body.  sel_v = (min_sq < euclthres^2) | !exterior (loss.py:278); a face is selected when
// it touches a selected vertex (:279-280); an HD point when its source face is selected (:281).
// idx[b][0..count) lists the selected HD points in increasing order.
__global__ void __launch_bounds__(HS_THREADS)
hd_select_kernel(const float* __restrict__ min_sq, const uint8_t* __restrict__ exterior,
                 const uint8_t* __restrict__ body_active, const int* __restrict__ faces, int V, int N,
                 const int* __restrict__ hd_face, float thres_sq, int* __restrict__ idx, int* __restrict__ counts) {
    __shared__ int s_warp[HS_THREADS / 32];
    __shared__ int s_base;
    const int b = blockIdx.x;
    if (body_active != nullptr && !body_active[b]) {
        if (threadIdx.x == 0) counts[b] = 0;
        return;
    }
    const float* mn = min_sq + (size_t)b * V;
    const uint8_t* ex = exterior + (size_t)b * V;
    int* out = idx + (size_t)b * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int k0 = 0; k0 < N; k0 += HS_THREADS) {
        const int k = k0 + threadIdx.x;
        bool sel = false;
        if (k < N) {
            const int f = hd_face[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int v = faces[3 * f + c];
                sel = sel || (mn[v] < thres_sq) || !ex[v];
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < HS_THREADS / 32; ++w) {
            const int c = s_warp[w];
            if (w < warp) before += c;
            total += c;
        }
        const int base = s_base;
        if (sel) out[base + before + __popc(m & ((1u << lane) - 1u))] = k;
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[b] = s_base;
}

// grid (ceil(N / 256), B): selected point i of body b
//   hd    = sum_k w_k v[col_k]                  (rows of the HD regressor, CSR; loss.py:285)
//   off   = hd + 0.001 n_f, n_f = unit normal of the source face (loss.py:30-41, 295-296)
//   proxy = first vertex of the source face (loss.py:89)
__global__ void hd_gather_kernel(const float* __restrict__ verts, int V, int N, const int* __restrict__ idx,
                                 const int* __restrict__ counts, const int* __restrict__ row_off,
                                 const int* __restrict__ cols, const float* __restrict__ vals,
                                 const int* __restrict__ hd_face, const int* __restrict__ faces,
                                 float4* __restrict__ hd4, float* __restrict__ hd, float* __restrict__ off,
                                 int* __restrict__ proxy) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[b]) return;
    const float* vb = verts + (size_t)b * V * 3;
    const int k = idx[(size_t)b * N + i];
    float x = 0.f, y = 0.f, z = 0.f;
    for (int e = row_off[k]; e < row_off[k + 1]; ++e) {
        const float w = vals[e];
        const float* p = vb + 3 * cols[e];
        x = fmaf(w, p[0], x); y = fmaf(w, p[1], y); z = fmaf(w, p[2], z);
    }
    const int f = hd_face[k];
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float e0x = vb[3 * i1] - vb[3 * i0], e0y = vb[3 * i1 + 1] - vb[3 * i0 + 1], e0z = vb[3 * i1 + 2] - vb[3 * i0 + 2];
    const float e1x = vb[3 * i2] - vb[3 * i0], e1y = vb[3 * i2 + 1] - vb[3 * i0 + 1], e1z = vb[3 * i2 + 2] - vb[3 * i0 + 2];
    const float nx = e0y * e1z - e0z * e1y, ny = e0z * e1x - e0x * e1z, nz = e0x * e1y - e0y * e1x;
    const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
    const size_t o = (size_t)b * N + i;
    hd[3 * o] = x; hd[3 * o + 1] = y; hd[3 * o + 2] = z;
    hd4[o] = make_float4(x, y, z, fmaf(z, z, fmaf(y, y, x * x)));
    off[3 * o] = x + 0.001f * (nx / nn);
    off[3 * o + 1] = y + 0.001f * (ny / nn);
    off[3 * o + 2] = z + 0.001f * (nz / nn);
    proxy[o] = i0;
}

// masked nearest selected HD point: for column j the first row i minimising the expansion-form squared
// distance among rows with geomask[proxy_i][proxy_j] (loss.py:288-291).  grid (ceil(N / 128), B);
// rows stream through shared memory in tiles of 128.
constexpr int HN_THREADS = 128;

__global__ void __launch_bounds__(HN_THREADS)
hd_nearest_kernel(const float4* __restrict__ hd4, const int* __restrict__ proxy, const int* __restrict__ counts,
                  int N, const uint32_t* __restrict__ maskT, int Vq, int* __restrict__ argmin_out) {
    __shared__ float4 s_p[HN_THREADS];
    __shared__ int s_row[HN_THREADS];
    const int b = blockIdx.y;
    const int n = counts[b];
    if ((int)blockIdx.x * HN_THREADS >= n) return;
    const float4* pb = hd4 + (size_t)b * N;
    const int* gb = proxy + (size_t)b * N;
    const int j = blockIdx.x * HN_THREADS + threadIdx.x;
    const bool live = j < n;
    const float4 q = pb[live ? j : 0];
    const int gj = gb[live ? j : 0];
    float best = INFINITY;
    int bi = 0;
    for (int i0 = 0; i0 < n; i0 += HN_THREADS) {
        __syncthreads();
        if (i0 + (int)threadIdx.x < n) {
            s_p[threadIdx.x] = pb[i0 + threadIdx.x];
            s_row[threadIdx.x] = gb[i0 + threadIdx.x];
        }
        __syncthreads();
        const int m = min(HN_THREADS, n - i0);
        for (int t = 0; t < m; ++t) {
            const int gi = s_row[t];
            const bool ok = (maskT[(size_t)(gi >> 5) * Vq + gj] >> (gi & 31)) & 1u;     // geomask[gi][gj]
            const float4 v = s_p[t];
            const float zz = fmaf(v.z, q.z, fmaf(v.y, q.y, v.x * q.x));
            const float p = ok ? fmaf(-2.f, zz, v.w + q.w) : INFINITY;
            if (p < best) { best = p; bi = i0 + t; }
        }
    }
    if (live) argmin_out[(size_t)b * N + j] = bi;
}

// backward of the regressor rows: g_verts[b][col_k] += w_k g_hd[b][i], through 64-bit fixed-point
// accumulators g_fix[B][V][3] (zeroed by the caller) so that the sum does not depend on the arrival order;
// hd_fold_kernel adds them into g_verts
__global__ void hd_scatter_kernel(const float* __restrict__ g_hd, int V, int N, const int* __restrict__ idx,
                                  const int* __restrict__ counts, const int* __restrict__ row_off,
                                  const int* __restrict__ cols, const float* __restrict__ vals,
                                  float* __restrict__ g_verts, long long* __restrict__ g_fix) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[b]) return;
    const size_t o = (size_t)b * N + i;
    const float gx = g_hd[3 * o], gy = g_hd[3 * o + 1], gz = g_hd[3 * o + 2];
    if (gx == 0.f && gy == 0.f && gz == 0.f) return;
    const int k = idx[o];
    float* g = g_verts + (size_t)b * V * 3;
    long long* gf = g_fix + (size_t)b * V * 3;
    for (int e = row_off[k]; e < row_off[k + 1]; ++e) {
        const float w = vals[e];
        const int c = cols[e];
        fix_add(&gf[3 * c], &g[3 * c], w * gx); fix_add(&gf[3 * c + 1], &g[3 * c + 1], w * gy);
        fix_add(&gf[3 * c + 2], &g[3 * c + 2], w * gz);
    }
}

__global__ void hd_fold_kernel(const long long* __restrict__ g_fix, long long n, float* __restrict__ g_verts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = g_fix[i];
    if (a != 0) g_verts[i] += fix_value(a);
}

int launch_hd_select(const float* min_sq, const uint8_t* exterior, const uint8_t* body_active, const int* faces,
                     int B, int V, int N, const int* hd_face, float thres_sq, int* idx, int* counts, cudaStream_t st) {
    if (B == 0) return 0;
    KernelTimer timer("hd_select_kernel", st);
    hd_select_kernel<<<B, HS_THREADS, 0, st>>>(min_sq, exterior, body_active, faces, V, N, hd_face, thres_sq, idx, counts);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_hd_gather(const float* verts, int B, int V, int N, const int* idx, const int* counts, const int* row_off,
                     const int* cols, const float* vals, const int* hd_face, const int* faces, float4* hd4,
                     float* hd, float* off, int* proxy, cudaStream_t st) {
    if (B == 0 || N == 0) return 0;
    KernelTimer timer("hd_gather_kernel", st);
    dim3 grid(cdiv(N, 256), B);
    hd_gather_kernel<<<grid, 256, 0, st>>>(verts, V, N, idx, counts, row_off, cols, vals, hd_face, faces, hd4, hd, off, proxy);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_hd_nearest(const float4* hd4, const int* proxy, const int* counts, int B, int N, const uint32_t* maskT,
                      int Vq, int* argmin, cudaStream_t st) {
    if (B == 0 || N == 0) return 0;
    KernelTimer timer("hd_nearest_kernel", st);
    dim3 grid(cdiv(N, HN_THREADS), B);
    hd_nearest_kernel<<<grid, HN_THREADS, 0, st>>>(hd4, proxy, counts, N, maskT, Vq, argmin);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_hd_scatter(const float* g_hd, int B, int V, int N, const int* idx, const int* counts, const int* row_off,
                      const int* cols, const float* vals, float* g_verts, cudaStream_t st) {
    if (B == 0 || N == 0) return 0;
    KernelTimer timer("hd_scatter_kernels", st);
    void* p = nullptr;
    const size_t n = 3 * (size_t)B * V;
    if (int rc = arena_get(st, sizeof(long long) * n, &p, 3)) return rc;
    long long* g_fix = (long long*)p;
    TUCH_CUDA(cudaMemsetAsync(g_fix, 0, sizeof(long long) * n, st));
    dim3 grid(cdiv(N, 256), B);
    hd_scatter_kernel<<<grid, 256, 0, st>>>(g_hd, V, N, idx, counts, row_off, cols, vals, g_verts, g_fix);
    TUCH_LAUNCH_CHECK(); count_launch();
    hd_fold_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>(g_fix, (long long)n, g_verts);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

}  // namespace tuch
