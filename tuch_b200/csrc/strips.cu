// Triangle-strip stream of a fixed mesh topology and the winding-number kernel that consumes it.
//
// winding_kernel (contact_kernels.cu) evaluates every (query, triangle) pair from three explicit
// corners: 9 subtractions, 3 norms, a cross product and 4 dot products per pair.  A mesh's triangles
// share corners, so the topology (constant across bodies and iterations: smplifydc.py:58-61) is cut
// once, on the host, into triangle strips; each stream element then adds ONE vertex and closes one
// triangle with the two previous vertices, whose query-relative vectors, lengths, mutual dot product
// and cross product are carried in registers:
//     per element: c - q, |c - q|, a.c, b.c, c.(a x b), denominator, b x c        (26 FMA-pipe ops
//     instead of 41 before the atan2).
// Stream layout (per body): strip4[Lp] float4 = (x, y, z, flag bits); flag bit 0 = this element closes
// a triangle, bit 31 = that triangle's corner order is an odd permutation of the mesh face (its signed
// solid angle is negated).  Every tile of WS_TILE elements is self-contained (it starts with two primer
// vertices), so F-splits and the TMA stage ring work at tile granularity without carried state.
#include <algorithm>
#include <unordered_map>
#include <vector>

#include "api_internal.h"
#include "strips.h"

namespace tuch {

// ------------------------------------------------------------------------------------------
// host: greedy stripification
// ------------------------------------------------------------------------------------------
namespace {

struct Mesh {
    const int* f;
    int F;
    std::unordered_map<uint64_t, std::vector<int>> edge_faces;
    static uint64_t key(int u, int v) {
        const uint64_t a = (uint64_t)std::min(u, v), b = (uint64_t)std::max(u, v);
        return (a << 32) | b;
    }
    Mesh(const int* faces, int nf) : f(faces), F(nf) {
        edge_faces.reserve((size_t)nf * 2);
        for (int t = 0; t < nf; ++t)
            for (int e = 0; e < 3; ++e) edge_faces[key(f[3 * t + e], f[3 * t + (e + 1) % 3])].push_back(t);
    }
    // unused face other than `not_face` across the undirected edge {u, v}; -1 if none
    int across(int u, int v, int not_face, const std::vector<char>& used, const std::vector<int>& trial, int stamp) const {
        auto it = edge_faces.find(key(u, v));
        if (it == edge_faces.end()) return -1;
        for (int g : it->second)
            if (g != not_face && !used[g] && trial[g] != stamp) return g;
        return -1;
    }
    int third(int g, int u, int v) const {
        for (int e = 0; e < 3; ++e) {
            const int w = f[3 * g + e];
            if (w != u && w != v) return w;
        }
        return -1;          // degenerate face (repeated vertex)
    }
};

struct Strip { std::vector<int> verts, faces; };   // faces[k] closes with verts[k+2]

// extends `s` forward from its last edge
void extend(const Mesh& m, Strip& s, const std::vector<char>& used, std::vector<int>& trial, int stamp) {
    for (;;) {
        const int n = (int)s.verts.size();
        const int u = s.verts[n - 2], v = s.verts[n - 1];
        const int g = m.across(u, v, s.faces.back(), used, trial, stamp);
        if (g < 0) return;
        const int w = m.third(g, u, v);
        if (w < 0) return;
        trial[g] = stamp;
        s.verts.push_back(w);
        s.faces.push_back(g);
    }
}

}  // namespace

int build_strip_stream(const int* faces, int F, int tile, std::vector<int>& vid, std::vector<uint32_t>& flag,
                       std::vector<int>& fid, int* n_strips_out) {
    vid.clear(); flag.clear(); fid.clear();
    Mesh m(faces, F);
    std::vector<char> used(F, 0);
    std::vector<int> trial(F, 0);
    int stamp = 0, n_strips = 0;
    auto emit = [&](int v, uint32_t fl, int face = -1) { vid.push_back(v); flag.push_back(fl); fid.push_back(face); };
    auto sign_of = [&](int g, int a, int b, int c) -> uint32_t {
        // + when (a, b, c) is a cyclic rotation of face g, - (bit 31) otherwise
        for (int r = 0; r < 3; ++r)
            if (faces[3 * g + r] == a && faces[3 * g + (r + 1) % 3] == b && faces[3 * g + (r + 2) % 3] == c) return 0u;
        return 0x80000000u;
    };
    for (int t0 = 0; t0 < F; ++t0) {
        if (used[t0]) continue;
        Strip best;
        for (int r = 0; r < 3; ++r) {                     // three ways to enter the first face
            Strip s;
            s.verts = {faces[3 * t0 + r], faces[3 * t0 + (r + 1) % 3], faces[3 * t0 + (r + 2) % 3]};
            s.faces = {t0};
            ++stamp;
            trial[t0] = stamp;
            extend(m, s, used, trial, stamp);
            // then grow from the other end: reverse and keep extending
            std::reverse(s.verts.begin(), s.verts.end());
            std::reverse(s.faces.begin(), s.faces.end());
            extend(m, s, used, trial, stamp);
            if (s.faces.size() > best.faces.size()) best = std::move(s);
        }
        for (int g : best.faces) used[g] = 1;
        ++n_strips;
        for (size_t k = 0; k < best.verts.size(); ++k) {
            // every tile is self-contained: a closing element needs its two predecessors inside the
            // same tile, so they are re-emitted as primers when the tile has just started
            if (k >= 2 && vid.size() % (size_t)tile < 2) {
                emit(best.verts[k - 2], 0u);
                emit(best.verts[k - 1], 0u);
            }
            uint32_t fl = 0u;
            if (k >= 2) fl = 1u | sign_of(best.faces[k - 2], best.verts[k - 2], best.verts[k - 1], best.verts[k]);
            emit(best.verts[k], fl, k >= 2 ? best.faces[k - 2] : -1);
        }
    }
    while (vid.size() % (size_t)tile != 0 || vid.empty()) emit(-1, 0u);
    if (n_strips_out) *n_strips_out = n_strips;
    // self-check: every face closed exactly once
    size_t closed = 0;
    for (uint32_t fl : flag) closed += fl & 1u;
    if ((int)closed != F) { set_error("strip builder closed %zu of %d faces", closed, F); return 1; }
    return 0;
}

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
// One CTA packs one tile (WS_TILE elements) of one body:
//   el[2i]   = (x, y, z, m)     m = 1 when the element closes a face, else 0
//   el[2i+1] = (Nx, Ny, Nz, 0)  N = (B - A) x (C - A) of that face in its ORIGINAL corner order.  The
//              numerator of the solid-angle formula, (A-q).((B-q) x (C-q)), is affine in q and equals
//              N.(X - q) for any corner X of the face, so the kernel needs no cross product per pair.
//   info[tile] = (centre of the tile's bounding box, r^2) with r = half diagonal + 1.5 x longest edge:
//              a query farther than r from the centre sees every triangle of the tile under a solid
//              angle < 0.19 sr (tan(omega/2) < 0.1), which licenses the short odd atan series.
__global__ void __launch_bounds__(WS_TILE)
pack_strips_kernel(const float* __restrict__ verts, int V, const int* __restrict__ faces,
                   const int* __restrict__ vid, const int* __restrict__ fid, int Lp,
                   float4* __restrict__ strip8, float4* __restrict__ info) {
    __shared__ float s_red[7][WS_TILE / 32];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int i = tile * WS_TILE + threadIdx.x;
    const float* vb = verts + (size_t)b * V * 3;
    const int v = vid[i], f = fid[i];
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), n = make_float4(0.f, 0.f, 0.f, 0.f);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, e2 = 0.f;
    if (v >= 0) {
        p.x = vb[3 * v]; p.y = vb[3 * v + 1]; p.z = vb[3 * v + 2];
        lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
    }
    if (f >= 0) {
        const float* A = vb + 3 * faces[3 * f];
        const float* Bc = vb + 3 * faces[3 * f + 1];
        const float* Cc = vb + 3 * faces[3 * f + 2];
        const float ux = Bc[0] - A[0], uy = Bc[1] - A[1], uz = Bc[2] - A[2];
        const float wx = Cc[0] - A[0], wy = Cc[1] - A[1], wz = Cc[2] - A[2];
        const float tx = Cc[0] - Bc[0], ty = Cc[1] - Bc[1], tz = Cc[2] - Bc[2];
        n.x = uy * wz - uz * wy; n.y = uz * wx - ux * wz; n.z = ux * wy - uy * wx;
        p.w = 1.f;
        e2 = fmaxf(fmaxf(ux * ux + uy * uy + uz * uz, wx * wx + wy * wy + wz * wz), tx * tx + ty * ty + tz * tz);
    }
    float4* o = strip8 + ((size_t)b * Lp + i) * 2;
    o[0] = p; o[1] = n;
    // block reduction of the bounding box and the longest edge
    float r[7] = {lo[0], lo[1], lo[2], -hi[0], -hi[1], -hi[2], -e2};            // all as minima
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) r[k] = fminf(r[k], __shfl_xor_sync(0xffffffffu, r[k], off));
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 7; ++k) s_red[k][threadIdx.x >> 5] = r[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        float m[7];
        for (int k = 0; k < 7; ++k) {
            m[k] = s_red[k][0];
            for (int w = 1; w < WS_TILE / 32; ++w) m[k] = fminf(m[k], s_red[k][w]);
        }
        float4 out = make_float4(0.f, 0.f, 0.f, -1.f);                          // empty tile: never near
        if (m[0] <= -m[3]) {
            const float dx = -m[3] - m[0], dy = -m[4] - m[1], dz = -m[5] - m[2];
            const float rad = 0.5f * sqrtf(dx * dx + dy * dy + dz * dz) * 1.0001f + 1.5f * sqrtf(-m[6]) + 1e-6f;
            out = make_float4(0.5f * (m[0] - m[3]), 0.5f * (m[1] - m[4]), 0.5f * (m[2] - m[5]), rad * rad);
        }
        info[(size_t)b * (Lp / WS_TILE) + tile] = out;
    }
}

// atan(num / den) for 0 <= |num| <= 0.125 den: x - x^3/3 + x^5/5 - x^7/7 (truncation < 1e-9)
__device__ __forceinline__ float atan_small(float num, float den) {
    const float x = num * rcp_approx(den);
    const float s = x * x;
    float q = fmaf(s, -0.142857142857f, 0.2f);
    q = fmaf(s, q, -0.333333333333f);
    return fmaf(x * s, q, x);
}

// One tile of the stream for WS_QPT queries.  FAST: every lane's queries are outside the tile's near
// radius, so all angles are small and no query coincides with a corner.
template <bool FAST>
__device__ __forceinline__ void strip_tile(const float4* __restrict__ tile, const float (&px)[WS_QPT],
                                           const float (&py)[WS_QPT], const float (&pz)[WS_QPT],
                                           float (&acc)[WS_QPT]) {
    // rolling state of the two previous stream vertices, per query
    float ax[WS_QPT], ay[WS_QPT], az[WS_QPT], bx[WS_QPT], by[WS_QPT], bz[WS_QPT];
    float la[WS_QPT], lb[WS_QPT], dab[WS_QPT], pab[WS_QPT];
#pragma unroll
    for (int k = 0; k < WS_QPT; ++k) {
        ax[k] = ay[k] = az[k] = bx[k] = by[k] = bz[k] = 0.f;
        la[k] = lb[k] = dab[k] = pab[k] = 0.f;
    }
#pragma unroll 4
    for (int e = 0; e < WS_TILE; ++e) {
        const float4 c4 = tile[2 * e];
        float cx[WS_QPT], cy[WS_QPT], cz[WS_QPT], lc[WS_QPT], dbc[WS_QPT];
#pragma unroll
        for (int k = 0; k < WS_QPT; ++k) {
            cx[k] = c4.x - px[k]; cy[k] = c4.y - py[k]; cz[k] = c4.z - pz[k];
            lc[k] = sqrt_approx(fmaf(cz[k], cz[k], fmaf(cy[k], cy[k], cx[k] * cx[k])));
            dbc[k] = fmaf(bz[k], cz[k], fmaf(by[k], cy[k], bx[k] * cx[k]));
        }
        if (c4.w != 0.f) {                                       // uniform: the element closes a face
            const float4 n4 = tile[2 * e + 1];
#pragma unroll
            for (int k = 0; k < WS_QPT; ++k) {
                const float dac = fmaf(az[k], cz[k], fmaf(ay[k], cy[k], ax[k] * cx[k]));
                const float num = fmaf(n4.z, cz[k], fmaf(n4.y, cy[k], n4.x * cx[k]));      // N . (C - q)
                float den = pab[k] * lc[k];                      // exactly +0 when q is a corner
                den = fmaf(dab[k], lc[k], den);
                den = fmaf(dac, lb[k], den);
                den = fmaf(dbc[k], la[k], den);
                if (FAST) {
                    acc[k] += atan_small(num, den);
                } else {
                    // N.(C-q) is only mathematically 0 when q is another corner of the face: the
                    // reference's a.(b x c) is exactly 0 there, and so is den
                    acc[k] += atan2_poly(den == 0.f ? 0.f : num, den);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < WS_QPT; ++k) {                       // shift: a <- b, b <- c
            ax[k] = bx[k]; ay[k] = by[k]; az[k] = bz[k];
            bx[k] = cx[k]; by[k] = cy[k]; bz[k] = cz[k];
            la[k] = lb[k]; lb[k] = lc[k];
            dab[k] = dbc[k];
            pab[k] = la[k] * lb[k];
        }
    }
}

// grid (query tiles, splits, bodies); block WS_THREADS; WS_QPT queries per thread
__global__ void __launch_bounds__(WS_THREADS)
winding_strip_kernel(const float4* __restrict__ strip8, const float4* __restrict__ info,
                     const float* __restrict__ points, float* __restrict__ partial, int Q, int Lp,
                     int tiles_per_split, long long point_stride, long long partial_stride,
                     const uint8_t* __restrict__ body_active, const int* __restrict__ q_counts) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    float4* s_el = reinterpret_cast<float4*>(s_raw);                              // [WS_STAGES][2 * WS_TILE]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_raw + (size_t)WS_STAGES * 2 * WS_TILE * sizeof(float4));

    const int b = blockIdx.z;
    if (body_active != nullptr && !body_active[b]) return;
    const int q_stride = Q;
    if (q_counts != nullptr) {
        Q = min(Q, q_counts[b]);
        if ((int)blockIdx.x * (WS_THREADS * WS_QPT) >= Q) return;
    }
    const int split = blockIdx.y;
    const int n_tiles_total = Lp / WS_TILE;
    const int tile0 = split * tiles_per_split;
    const int n_tiles = min(tiles_per_split, n_tiles_total - tile0);
    const float4* src = strip8 + ((size_t)b * Lp + (size_t)tile0 * WS_TILE) * 2;
    const float4* tinfo = info + (size_t)b * n_tiles_total + tile0;
    constexpr uint32_t TILE_BYTES = 2 * WS_TILE * sizeof(float4);

    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_STAGES; ++s) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_STAGES && s < n_tiles; ++s) {
            mbar_expect_tx(&s_bar[s], TILE_BYTES);
            tma_load_1d(s_el + (size_t)s * 2 * WS_TILE, src + (size_t)s * 2 * WS_TILE, TILE_BYTES, &s_bar[s]);
        }
    }

    float px[WS_QPT], py[WS_QPT], pz[WS_QPT], acc[WS_QPT];
    const float* pb = points + (size_t)b * point_stride;
    const int q0 = blockIdx.x * (WS_THREADS * WS_QPT) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < WS_QPT; ++k) {
        const int q = min(q0 + k * WS_THREADS, Q - 1);
        px[k] = pb[3 * q]; py[k] = pb[3 * q + 1]; pz[k] = pb[3 * q + 2];
        acc[k] = 0.f;
    }

    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % WS_STAGES;
        const float4 ti = __ldg(tinfo + t);
        bool is_near = false;
#pragma unroll
        for (int k = 0; k < WS_QPT; ++k) {
            const float dx = px[k] - ti.x, dy = py[k] - ti.y, dz = pz[k] - ti.z;
            is_near = is_near || (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= ti.w);
        }
        const bool fast = !__any_sync(0xffffffffu, is_near);
        mbar_wait(&s_bar[s], (t / WS_STAGES) & 1);
        const float4* tile = s_el + (size_t)s * 2 * WS_TILE;
        if (fast) strip_tile<true>(tile, px, py, pz, acc);
        else strip_tile<false>(tile, px, py, pz, acc);
        __syncthreads();
        if (threadIdx.x == 0 && t + WS_STAGES < n_tiles) {
            mbar_expect_tx(&s_bar[s], TILE_BYTES);
            tma_load_1d(s_el + (size_t)s * 2 * WS_TILE, src + (size_t)(t + WS_STAGES) * 2 * WS_TILE, TILE_BYTES,
                        &s_bar[s]);
        }
    }

    float* out = partial + (size_t)b * partial_stride + (size_t)split * q_stride;
#pragma unroll
    for (int k = 0; k < WS_QPT; ++k) {
        const int q = q0 + k * WS_THREADS;
        if (q < Q) out[q] = acc[k];
    }
}

constexpr size_t WS_SMEM = (size_t)WS_STAGES * 2 * WS_TILE * sizeof(float4) + WS_STAGES * sizeof(uint64_t);

int strip_splits(int B, int Q, int Lp, int sm_count) {
    const int qtiles = cdiv(Q, WS_THREADS * WS_QPT);
    const int n_tiles = Lp / WS_TILE;
    const long long want = (long long)sm_count * 8;
    int S = (int)((want + (long long)qtiles * B - 1) / ((long long)qtiles * B));
    S = std::max(1, std::min(S, n_tiles));
    const int per = cdiv(n_tiles, S);
    return cdiv(n_tiles, per);
}

int launch_pack_strips(const float* verts, int B, int V, const int* faces, const int* vid, const int* fid, int Lp,
                       float4* strip8, float4* info, cudaStream_t st) {
    if (B == 0 || Lp == 0) return 0;
    dim3 grid(Lp / WS_TILE, B);
    pack_strips_kernel<<<grid, WS_TILE, 0, st>>>(verts, V, faces, vid, fid, Lp, strip8, info);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_winding_strips(const StripJob& j, cudaStream_t st) {
    if (j.B == 0 || j.Q == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        TUCH_CUDA(cudaFuncSetAttribute(winding_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM));
        attr_set = true;
    }
    const int n_tiles = j.Lp / WS_TILE;
    const int per = cdiv(n_tiles, j.S);
    dim3 grid(cdiv(j.Q, WS_THREADS * WS_QPT), j.S, j.B);
    {
        KernelTimer timer("winding_kernel", st);
        winding_strip_kernel<<<grid, WS_THREADS, WS_SMEM, st>>>(j.strip8, j.info, j.points, j.partial, j.Q, j.Lp, per,
                                                                j.point_stride, (long long)j.S * j.Q, j.body_active,
                                                                j.q_counts);
    }
    TUCH_LAUNCH_CHECK(); count_launch();
    return launch_winding_finalize(j.partial, j.B, j.Q, j.S, j.out_stride, j.winding, j.body_active, j.q_counts, st);
}

}  // namespace tuch
