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
                       int* n_strips_out) {
    vid.clear(); flag.clear();
    Mesh m(faces, F);
    std::vector<char> used(F, 0);
    std::vector<int> trial(F, 0);
    int stamp = 0, n_strips = 0;
    auto emit = [&](int v, uint32_t fl) { vid.push_back(v); flag.push_back(fl); };
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
            emit(best.verts[k], fl);
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
__global__ void pack_strips_kernel(const float* __restrict__ verts, int V, const int* __restrict__ vid,
                                   const uint32_t* __restrict__ flag, int Lp, float4* __restrict__ strip4) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Lp) return;
    const int v = vid[i];
    float4 o = make_float4(0.f, 0.f, 0.f, __uint_as_float(flag[i]));
    if (v >= 0) {
        const float* p = verts + ((size_t)b * V + v) * 3;
        o.x = p[0]; o.y = p[1]; o.z = p[2];
    }
    strip4[(size_t)b * Lp + i] = o;
}

// grid (query tiles, splits, bodies); block WS_THREADS; WS_QPT queries per thread
__global__ void __launch_bounds__(WS_THREADS)
winding_strip_kernel(const float4* __restrict__ strip4, const float* __restrict__ points,
                     float* __restrict__ partial, int Q, int Lp, int tiles_per_split, long long point_stride,
                     long long partial_stride, const uint8_t* __restrict__ body_active,
                     const int* __restrict__ q_counts) {
    __shared__ __align__(128) float4 s_el[WS_STAGES][WS_TILE];
    __shared__ __align__(8) uint64_t s_bar[WS_STAGES];

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
    const float4* src = strip4 + (size_t)b * Lp + (size_t)tile0 * WS_TILE;
    constexpr uint32_t TILE_BYTES = WS_TILE * sizeof(float4);

    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_STAGES; ++s) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_STAGES && s < n_tiles; ++s) {
            mbar_expect_tx(&s_bar[s], TILE_BYTES);
            tma_load_1d(s_el[s], src + (size_t)s * WS_TILE, TILE_BYTES, &s_bar[s]);
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
        mbar_wait(&s_bar[s], (t / WS_STAGES) & 1);
        const float4* tile = s_el[s];
        // rolling state of the two previous stream vertices, per query
        float ax[WS_QPT], ay[WS_QPT], az[WS_QPT], bx[WS_QPT], by[WS_QPT], bz[WS_QPT];
        float la[WS_QPT], lb[WS_QPT], dab[WS_QPT], pab[WS_QPT], xx[WS_QPT], xy[WS_QPT], xz[WS_QPT];
#pragma unroll
        for (int k = 0; k < WS_QPT; ++k) {
            ax[k] = ay[k] = az[k] = bx[k] = by[k] = bz[k] = 0.f;
            la[k] = lb[k] = dab[k] = pab[k] = xx[k] = xy[k] = xz[k] = 0.f;
        }
#pragma unroll 6
        for (int e = 0; e < WS_TILE; ++e) {
            const float4 c4 = tile[e];
            const uint32_t bits = __float_as_uint(c4.w);
            const bool close = bits & 1u;
            const uint32_t sgn = bits & 0x80000000u;
#pragma unroll
            for (int k = 0; k < WS_QPT; ++k) {
                const float cx = c4.x - px[k], cy = c4.y - py[k], cz = c4.z - pz[k];
                const float lc = sqrt_approx(fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
                const float dac = fmaf(az[k], cz, fmaf(ay[k], cy, ax[k] * cx));
                const float dbc = fmaf(bz[k], cz, fmaf(by[k], cy, bx[k] * cx));
                float num = fmaf(xz[k], cz, fmaf(xy[k], cy, xx[k] * cx));          // c . (a x b)
                num = __uint_as_float(__float_as_uint(num) ^ sgn);
                float den = pab[k] * lc;                                           // +0 on a corner hit
                den = fmaf(dab[k], lc, den);
                den = fmaf(dac, lb[k], den);
                den = fmaf(dbc, la[k], den);
                const float ang = atan2_poly(num, den);
                acc[k] += close ? ang : 0.f;
                // shift: a <- b, b <- c
                xx[k] = fmaf(by[k], cz, -bz[k] * cy);
                xy[k] = fmaf(bz[k], cx, -bx[k] * cz);
                xz[k] = fmaf(bx[k], cy, -by[k] * cx);
                ax[k] = bx[k]; ay[k] = by[k]; az[k] = bz[k];
                bx[k] = cx; by[k] = cy; bz[k] = cz;
                la[k] = lb[k]; lb[k] = lc;
                dab[k] = dbc;
                pab[k] = la[k] * lb[k];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && t + WS_STAGES < n_tiles) {
            mbar_expect_tx(&s_bar[s], TILE_BYTES);
            tma_load_1d(s_el[s], src + (size_t)(t + WS_STAGES) * WS_TILE, TILE_BYTES, &s_bar[s]);
        }
    }

    float* out = partial + (size_t)b * partial_stride + (size_t)split * q_stride;
#pragma unroll
    for (int k = 0; k < WS_QPT; ++k) {
        const int q = q0 + k * WS_THREADS;
        if (q < Q) out[q] = acc[k];
    }
}

int strip_splits(int B, int Q, int Lp, int sm_count) {
    const int qtiles = cdiv(Q, WS_THREADS * WS_QPT);
    const int n_tiles = Lp / WS_TILE;
    const long long want = (long long)sm_count * 8;
    int S = (int)((want + (long long)qtiles * B - 1) / ((long long)qtiles * B));
    S = std::max(1, std::min(S, n_tiles));
    const int per = cdiv(n_tiles, S);
    return cdiv(n_tiles, per);
}

int launch_pack_strips(const float* verts, int B, int V, const int* vid, const uint32_t* flag, int Lp,
                       float4* strip4, cudaStream_t st) {
    if (B == 0 || Lp == 0) return 0;
    dim3 grid(cdiv(Lp, 256), B);
    pack_strips_kernel<<<grid, 256, 0, st>>>(verts, V, vid, flag, Lp, strip4);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_winding_strips(const StripJob& j, cudaStream_t st) {
    if (j.B == 0 || j.Q == 0) return 0;
    const int n_tiles = j.Lp / WS_TILE;
    const int per = cdiv(n_tiles, j.S);
    dim3 grid(cdiv(j.Q, WS_THREADS * WS_QPT), j.S, j.B);
    {
        KernelTimer timer("winding_kernel", st);
        winding_strip_kernel<<<grid, WS_THREADS, 0, st>>>(j.strip4, j.points, j.partial, j.Q, j.Lp, per, j.point_stride,
                                                          (long long)j.S * j.Q, j.body_active, j.q_counts);
    }
    TUCH_LAUNCH_CHECK(); count_launch();
    return launch_winding_finalize(j.partial, j.B, j.Q, j.S, j.out_stride, j.winding, j.body_active, j.q_counts, st);
}

}  // namespace tuch
