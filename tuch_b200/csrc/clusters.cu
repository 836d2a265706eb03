// Far-field winding numbers over a face-cluster hierarchy ("fast winding numbers", Barill et al. 2018,
// restated for a fixed-topology batch of posed bodies).
//
// The generalized winding number of tuch/utils/contact.py:112-147 is a sum of F signed solid angles per
// query.  The solid angle of a flat triangle t seen from q is the exact surface integral
//     Omega_t(q) = int_t  n . (x - q) / |x - q|^3  dA ,
// so a CLUSTER of triangles far from q can be replaced by the Taylor expansion of that integrand about
// the cluster centre p, whose coefficients are area-weighted polynomial moments of the cluster
// (exact for flat triangles: centroid for the linear term, edge-midpoint quadrature for the quadratic):
//     Omega_C(q) ~ (M0.r + tr M1) / R^3 - (3 r'M1 r + 1.5 u.r) / R^5 + 7.5 T(r,r,r) / R^7,   r = p - q, R = |r|.
// The topology (constant across bodies and iterations, smplifydc.py:58-61) is cut once, on the host,
// into leaves of <= 32 faces grouped into super-clusters of <= 8 leaves; per body and iteration one warp
// per node recomputes centre, radius and moments from the posed vertices.  The winding kernel then
// walks supers -> leaves per warp of 32 neighbouring queries: a node farther than WC_BETA radii from
// every query of the warp costs ~50 instructions per query; a leaf that is near for some queries is
// evaluated exactly for those queries with one LANE PER FACE (same arithmetic as winding_kernel).
// Callers only consume `winding <= 0.99` (losses.py:82, loss.py:262): every query whose approximate
// value lies within WC_MARGIN of the threshold is re-evaluated exactly over all faces, so the flags are
// those of the exact kernel.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <type_traits>
#include <unordered_map>

#include "api_internal.h"
#include "clusters.h"

namespace tuch {

// ------------------------------------------------------------------------------------------
// host: hierarchy
// ------------------------------------------------------------------------------------------
namespace {

// Recursive bisection of a set of items (faces or vertices) that live on the mesh: connected components
// are never mixed, a connected set is cut at a balanced position along the principal axis of its item
// positions, and small pieces cut off by the plane are handed to the other half.  Leaves hold <= `leaf`
// items (ascending ids); optionally consecutive leaves are grouped into mids of <= `mid_leaves` leaves and
// consecutive mids into tops of <= `top_leaves` leaves.
struct TreeBuilder {
    int N = 0, leaf = 32, mid_leaves = 0, top_leaves = 0;
    std::vector<float> cen;                  // [N][3] item positions
    std::vector<int> adj_off, adj;           // item adjacency, CSR
    std::vector<int> stamp;
    int cur_stamp = 0;
    std::vector<int>* leaf_items = nullptr;  // [n_leaves][leaf], -1 = padding
    std::vector<int>* mid_off = nullptr;     // [n_mids + 1] leaf ranges
    std::vector<int>* top_off = nullptr;     // [n_tops + 1] mid ranges
    int n_leaves = 0, n_mids = 0, n_tops = 0;

    void set_adjacency(std::vector<std::vector<int>>& nb) {
        adj_off.assign(N + 1, 0);
        for (int t = 0; t < N; ++t) {
            std::sort(nb[t].begin(), nb[t].end());
            nb[t].erase(std::unique(nb[t].begin(), nb[t].end()), nb[t].end());
            adj_off[t + 1] = adj_off[t] + (int)nb[t].size();
        }
        adj.resize(adj_off[N]);
        for (int t = 0; t < N; ++t) std::copy(nb[t].begin(), nb[t].end(), adj.begin() + adj_off[t]);
        stamp.assign(N, 0);
    }

    // connected components of `fs` (ordered by their smallest id, ids ascending inside)
    std::vector<std::vector<int>> components(const std::vector<int>& fs) {
        const int in = ++cur_stamp;
        for (int f : fs) stamp[f] = in;
        std::vector<std::vector<int>> comps;
        std::vector<int> stack;
        for (int f0 : fs) {
            if (stamp[f0] != in) continue;
            stamp[f0] = 0;
            comps.emplace_back();
            stack.assign(1, f0);
            while (!stack.empty()) {
                const int f = stack.back();
                stack.pop_back();
                comps.back().push_back(f);
                for (int k = adj_off[f]; k < adj_off[f + 1]; ++k)
                    if (stamp[adj[k]] == in) { stamp[adj[k]] = 0; stack.push_back(adj[k]); }
            }
            std::sort(comps.back().begin(), comps.back().end());
        }
        return comps;
    }

    void open_groups(bool& in_mid, bool& in_top, int want_leaves) {
        if (top_off != nullptr && !in_top && want_leaves <= top_leaves) { top_off->push_back(n_mids); ++n_tops; in_top = true; }
        if (mid_off != nullptr && !in_mid && want_leaves <= mid_leaves) { mid_off->push_back(n_leaves); ++n_mids; in_mid = true; }
    }

    void emit_leaf(const std::vector<int>& fs, bool in_mid, bool in_top) {
        open_groups(in_mid, in_top, 1);                          // a leaf outside any group gets its own
        for (int i = 0; i < leaf; ++i) leaf_items->push_back(i < (int)fs.size() ? fs[i] : -1);
        ++n_leaves;
    }

    void split(std::vector<int> fs, bool in_mid, bool in_top) {           // fs ascending
        const int n = (int)fs.size();
        const int want_leaves = (n + leaf - 1) / leaf;
        open_groups(in_mid, in_top, want_leaves);
        std::vector<std::vector<int>> comps = components(fs);
        if (comps.size() > 1) {
            // pack small components together (a leaf may hold several of them) as long as they fit
            std::vector<int> bag;
            for (auto& c : comps) {
                if ((int)c.size() > leaf) { split(std::move(c), in_mid, in_top); continue; }
                if ((int)(bag.size() + c.size()) > leaf) { std::sort(bag.begin(), bag.end()); emit_leaf(bag, in_mid, in_top); bag.clear(); }
                bag.insert(bag.end(), c.begin(), c.end());
            }
            if (!bag.empty()) { std::sort(bag.begin(), bag.end()); emit_leaf(bag, in_mid, in_top); }
            return;
        }
        if (n <= leaf) { emit_leaf(fs, in_mid, in_top); return; }
        // principal axis of the item positions (power iteration on the 3x3 covariance, fp64)
        double mean[3] = {0, 0, 0};
        for (int f : fs) for (int a = 0; a < 3; ++a) mean[a] += cen[3 * f + a];
        for (int a = 0; a < 3; ++a) mean[a] /= n;
        double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int f : fs) {
            double d[3];
            for (int a = 0; a < 3; ++a) d[a] = cen[3 * f + a] - mean[a];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a][b] += d[a] * d[b];
        }
        double v[3] = {0.57, 0.58, 0.59};
        for (int it = 0; it < 64; ++it) {
            double w[3] = {0, 0, 0};
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) w[a] += C[a][b] * v[b];
            const double nrm = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            if (nrm < 1e-300) break;
            for (int a = 0; a < 3; ++a) v[a] = w[a] / nrm;
        }
        std::vector<std::pair<double, int>> key(n);
        for (int i = 0; i < n; ++i) {
            const int f = fs[i];
            key[i] = {(cen[3 * f] - mean[0]) * v[0] + (cen[3 * f + 1] - mean[1]) * v[1] + (cen[3 * f + 2] - mean[2]) * v[2], f};
        }
        std::sort(key.begin(), key.end());
        const int left_leaves = (want_leaves + 1) / 2;
        int n_left = (int)std::llround((double)n * left_leaves / want_leaves);
        n_left = std::max(1, std::min(n - 1, n_left));
        std::vector<int> l(n_left), r(n - n_left);
        for (int i = 0; i < n; ++i) (i < n_left ? l[i] : r[i - n_left]) = key[i].second;
        std::sort(l.begin(), l.end());
        std::sort(r.begin(), r.end());
        // a planar cut through a triangulated patch leaves ragged fragments (items whose neighbours all
        // fell on the other side): hand small cut-off pieces to the other half instead of letting them
        // become leaves of their own
        auto give_fragments = [&](std::vector<int>& from, std::vector<int>& to) {
            std::vector<std::vector<int>> cs = components(from);
            if (cs.size() < 2) return;
            size_t big = 0;
            for (size_t c = 1; c < cs.size(); ++c) if (cs[c].size() > cs[big].size()) big = c;
            const size_t small = std::max<size_t>(3, from.size() / 8);
            std::vector<int> keep;
            for (size_t c = 0; c < cs.size(); ++c) {
                std::vector<int>& dst = (c != big && cs[c].size() <= small) ? to : keep;
                dst.insert(dst.end(), cs[c].begin(), cs[c].end());
            }
            from.swap(keep);
            std::sort(from.begin(), from.end());
            std::sort(to.begin(), to.end());
        };
        give_fragments(l, r);
        give_fragments(r, l);
        split(std::move(l), in_mid, in_top);
        split(std::move(r), in_mid, in_top);
    }

    void run() {
        std::vector<int> all(N);
        std::iota(all.begin(), all.end(), 0);
        split(std::move(all), false, false);
        if (mid_off) mid_off->push_back(n_leaves);
        if (top_off) top_off->push_back(n_mids);
    }

    // ---- optional second stage: rounder leaves (scripts/proto/leaf_refine.py) -------------------------
    // The bisection above cuts along one principal axis per level, which leaves elongated leaves (bounding
    // radius 1.55x that of the equal-area disc on the SMPL-sized body).  Here every pair of leaves that are
    // neighbours in the item graph is pooled and re-split along the best of a fixed set of directions, sizes
    // kept <= leaf, whenever that lowers R_a^3 + R_b^3 (the near set of a leaf grows with R^3).  Leaves keep
    // their index, so the group structure above them is untouched.  `pts` holds ppi points per item that the
    // radius has to cover (the corners of a face; the vertex itself).
    std::vector<float> pts;
    int ppi = 0;

    double radius3(const int* items, int n) const {
        double c[3] = {0, 0, 0};
        for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) c[a] += cen[3 * items[i] + a];
        for (int a = 0; a < 3; ++a) c[a] /= n;
        double r2 = 0;
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < ppi; ++k) {
                const float* q = &pts[((size_t)items[i] * ppi + k) * 3];
                const double dx = q[0] - c[0], dy = q[1] - c[1], dz = q[2] - c[2];
                r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
            }
        return r2 * std::sqrt(r2);
    }

    void refine(int sweeps) {
        if (ppi <= 0 || n_leaves < 2) return;
        std::vector<int>& L = *leaf_items;
        // 13 directions: the axes, the face diagonals and the body diagonals of the cube
        static const double dirs[13][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, -1},
                                           {0, 1, 1}, {0, 1, -1}, {1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {-1, 1, 1}};
        std::vector<int> owner(N, -1);
        // measured on B200: exchanging items across group boundaries makes the leaves rounder still (-28 % near
        // pairs) but inflates the group spheres, more groups open per warp and the winding kernel gets SLOWER
        // (2.77 vs 2.66 ms) -- so the leaves of different groups are left alone
        std::vector<int> group_of_leaf;
        if (mid_off != nullptr && (int)mid_off->size() == n_mids + 1) {
            group_of_leaf.resize(n_leaves);
            for (int m = 0; m < n_mids; ++m)
                for (int l = (*mid_off)[m]; l < (*mid_off)[m + 1]; ++l) group_of_leaf[l] = m;
        }
        for (int sweep = 0; sweep < sweeps; ++sweep) {
            for (int l = 0; l < n_leaves; ++l)
                for (int i = 0; i < leaf; ++i) if (L[(size_t)l * leaf + i] >= 0) owner[L[(size_t)l * leaf + i]] = l;
            std::vector<std::pair<int, int>> pairs;
            for (int t = 0; t < N; ++t)
                for (int k = adj_off[t]; k < adj_off[t + 1]; ++k)
                    if (owner[t] < owner[adj[k]]) pairs.emplace_back(owner[t], owner[adj[k]]);
            std::sort(pairs.begin(), pairs.end());
            pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
            if (!group_of_leaf.empty())                          // exchanges stay inside one group: its item set,
                pairs.erase(std::remove_if(pairs.begin(), pairs.end(),   // hence its sphere and moments, is unchanged
                                           [&](const std::pair<int, int>& q) { return group_of_leaf[q.first] != group_of_leaf[q.second]; }),
                            pairs.end());
            int improved = 0;
            std::vector<int> pool, best;
            std::vector<std::pair<double, int>> key;
            for (const auto& pr : pairs) {
                int* la = &L[(size_t)pr.first * leaf];
                int* lb = &L[(size_t)pr.second * leaf];
                pool.clear();
                int na = 0, nb = 0;
                for (int i = 0; i < leaf; ++i) if (la[i] >= 0) { pool.push_back(la[i]); ++na; }
                for (int i = 0; i < leaf; ++i) if (lb[i] >= 0) { pool.push_back(lb[i]); ++nb; }
                const int n = na + nb;
                double cost = radius3(pool.data(), na) + radius3(pool.data() + na, nb);
                int best_k = -1;
                key.resize(n);
                for (const auto& d : dirs) {
                    for (int i = 0; i < n; ++i) {
                        const int f = pool[i];
                        key[i] = {cen[3 * f] * d[0] + cen[3 * f + 1] * d[1] + cen[3 * f + 2] * d[2], f};
                    }
                    std::sort(key.begin(), key.end());
                    std::vector<int> order(n);
                    for (int i = 0; i < n; ++i) order[i] = key[i].second;
                    for (int k = std::max(n - leaf, 1); k <= std::min(leaf, n - 1); ++k) {
                        const double c = radius3(order.data(), k) + radius3(order.data() + k, n - k);
                        if (c < cost * 0.999) { cost = c; best = order; best_k = k; }
                    }
                }
                if (best_k < 0) continue;
                ++improved;
                std::sort(best.begin(), best.begin() + best_k);
                std::sort(best.begin() + best_k, best.end());
                for (int i = 0; i < leaf; ++i) {
                    la[i] = i < best_k ? best[i] : -1;
                    lb[i] = i < n - best_k ? best[best_k + i] : -1;
                }
                for (int i = 0; i < best_k; ++i) owner[best[i]] = pr.first;
                for (int i = best_k; i < n; ++i) owner[best[i]] = pr.second;
            }
            if (improved == 0) break;
        }
    }
};

// sweeps of TreeBuilder::refine over the face leaves and the vertex tiles: neighbouring leaves INSIDE a group are
// pooled and re-split along the best of 13 directions, which makes them rounder (smaller bounding radii, fewer
// near leaves per query) without touching the groups.  Measured on the B200 at 256 bodies (scripts/time_contact.py):
// 0 / 2 / 4 sweeps -> winding 2.19 / 2.07 / 2.06 ms, nearest 1.48 / 1.43 / 1.43 ms, same far-field error, flags and
// nearest vertices identical.  Two sweeps ship (tree build 0.15 -> 0.4 s, once per topology); TUCH_TREE_REFINE
// overrides the count for experiments.
static int tree_refine_sweeps() {
    static const int n = getenv("TUCH_TREE_REFINE") ? atoi(getenv("TUCH_TREE_REFINE")) : 2;
    return std::max(0, std::min(n, 16));
}

uint64_t edge_key(int u, int v) {
    const uint64_t a = (uint64_t)std::min(u, v), b = (uint64_t)std::max(u, v);
    return (a << 32) | b;
}

}  // namespace

int build_cluster_tree(const int* faces, int F, int V, const float* verts, ClusterTree& out) {
    out = ClusterTree();
    TUCH_REQUIRE(F > 0 && V > 0 && faces && verts, "build_cluster_tree: empty mesh");
    std::unordered_map<uint64_t, std::vector<int>> ef;
    ef.reserve((size_t)F * 2);
    for (int t = 0; t < F; ++t)
        for (int e = 0; e < 3; ++e) ef[edge_key(faces[3 * t + e], faces[3 * t + (e + 1) % 3])].push_back(t);
    {   // faces: neighbours share an edge
        TreeBuilder tb;
        tb.N = F; tb.leaf = WC_LEAF; tb.mid_leaves = WC_MID_LEAVES; tb.top_leaves = WC_TOP_LEAVES;
        // development knobs for the fan-out sweeps (scripts/proto/winding_work_model.py); the constants ship
        if (const char* e = getenv("TUCH_WC_MID")) tb.mid_leaves = std::max(2, atoi(e));
        if (const char* e = getenv("TUCH_WC_TOP")) tb.top_leaves = std::max(tb.mid_leaves, atoi(e));
        tb.leaf_items = &out.leaf_face; tb.mid_off = &out.mid_off; tb.top_off = &out.top_off;
        tb.cen.resize((size_t)F * 3);
        for (int t = 0; t < F; ++t)
            for (int a = 0; a < 3; ++a)
                tb.cen[3 * t + a] = (verts[3 * faces[3 * t] + a] + verts[3 * faces[3 * t + 1] + a] + verts[3 * faces[3 * t + 2] + a]) / 3.f;
        std::vector<std::vector<int>> nb(F);
        for (auto& kv : ef)
            for (int x : kv.second) for (int y : kv.second) if (x != y) nb[x].push_back(y);
        tb.set_adjacency(nb);
        tb.run();
        if (tree_refine_sweeps() > 0) {
            tb.ppi = 3;
            tb.pts.resize((size_t)F * 9);
            for (int t = 0; t < F; ++t)
                for (int e = 0; e < 3; ++e)
                    for (int a = 0; a < 3; ++a) tb.pts[((size_t)t * 3 + e) * 3 + a] = verts[3 * faces[3 * t + e] + a];
            tb.refine(tree_refine_sweeps());
        }
        out.K = tb.n_leaves; out.NM = tb.n_mids; out.NT = tb.n_tops;
        for (int t = 0; t < out.NT; ++t)
            out.max_top_leaves = std::max(out.max_top_leaves, out.mid_off[out.top_off[t + 1]] - out.mid_off[out.top_off[t]]);
    }
    {   // vertices: neighbours share an edge; vertices without faces end up in tiles of their own
        TreeBuilder tb;
        tb.N = V; tb.leaf = 32; tb.mid_leaves = 8;
        tb.leaf_items = &out.vtile; tb.mid_off = &out.vgroup_off;
        tb.cen.assign(verts, verts + (size_t)V * 3);
        std::vector<std::vector<int>> nb(V);
        for (auto& kv : ef) {
            const int u = (int)(kv.first >> 32), v = (int)(kv.first & 0xffffffffu);
            if (u != v) { nb[u].push_back(v); nb[v].push_back(u); }
        }
        tb.set_adjacency(nb);
        tb.run();
        if (tree_refine_sweeps() > 0) {
            tb.ppi = 1;
            tb.pts = tb.cen;
            tb.refine(tree_refine_sweeps());
        }
        out.T = tb.n_leaves; out.NG = tb.n_mids;
    }
    // self-checks: partitions of the faces and of the vertices
    std::vector<char> seen(F, 0);
    size_t n = 0;
    for (int f : out.leaf_face)
        if (f >= 0) { TUCH_REQUIRE(!seen[f], "cluster tree lists face %d twice", f); seen[f] = 1; ++n; }
    TUCH_REQUIRE((int)n == F, "cluster tree covers %zu of %d faces", n, F);
    seen.assign(V, 0);
    n = 0;
    for (int v : out.vtile)
        if (v >= 0) { TUCH_REQUIRE(!seen[v], "vertex tiling lists vertex %d twice", v); seen[v] = 1; ++n; }
    TUCH_REQUIRE((int)n == V, "vertex tiling covers %zu of %d vertices", n, V);
    return 0;
}

// ------------------------------------------------------------------------------------------
// device: per-body node records
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// the same over aligned groups of W lanes (W = 16: the two half-warps reduce independently)
template <int W>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int W>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sums of N per-lane values over aligned groups of W lanes as a butterfly reduce-scatter: in the round with lane
// offset O a lane keeps the half of its values its bit O picks and hands the other half to lane ^ O.  Every total
// is built by the tree group_sum<W> builds (bit-identical), with N (1 - 1 / W) shuffles instead of N log2 W.
// Afterwards, with `sub` the lane's index in its group: N >= W: v[0 .. N / W) hold the totals of the values
// [sub N / W, (sub + 1) N / W);  N < W: v[0] holds the total of value sub / (W / N).
template <int O, int N>
struct ReduceScatter {
    static __device__ __forceinline__ void run(float* v, int lane) {
        if constexpr (O > 0) {
            if constexpr (N > 1) {
                const bool up = (lane & O) != 0;
#pragma unroll
                for (int i = 0; i < N / 2; ++i) {
                    const float keep = up ? v[i + N / 2] : v[i];
                    const float send = up ? v[i] : v[i + N / 2];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
                }
                ReduceScatter<O / 2, N / 2>::run(v, lane);
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], O);
                ReduceScatter<O / 2, 1>::run(v, lane);
            }
        }
    }
};
// the N < W case followed by a broadcast: every lane of the group ends up with all N totals
template <int W, int N>
__device__ __forceinline__ void group_sums(float (&v)[N], int lane) {
    static_assert(N < W, "group_sums: fewer values than lanes");
    float t[N];
#pragma unroll
    for (int k = 0; k < N; ++k) t[k] = v[k];
    ReduceScatter<W / 2, N>::run(t, lane);
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = __shfl_sync(0xffffffffu, t[0], k * (W / N), W);
}

// One warp per node; nodes are laid out [tops | mids | leaves].  The two half-warps walk alternate leaves
// of the node, one lane per face slot.  Record layout (all moments pre-scaled so that the kernel's sum is
// Omega / 2, the quantity the finalize step expects):
//   f0 = (p, (beta R)^2)                    f1 = 0.5 (M0, tr M1)
//   f2 = -1.5 (Qxx, Qyy, Qzz, 2Qxy)         f3 = (-1.5 * 2Qxz, -1.5 * 2Qyz, -0.75 ux, -0.75 uy)
//   f4 = (-0.75 uz, 3.75 Txxx, 3.75 Tyyy, 3.75 Tzzz)
//   f5 = 3.75 (Txxy, Txxz, Tyyx, Tyyz)      f6 = 3.75 (Tzzx, Tzzy, Txyz, 0)
// Q = sym(M1), u_k = 2 (a.S)_k + a_k tr S, T = symmetrised sum_t a_t (x) S_t.
__global__ void __launch_bounds__(128)
cluster_pack_kernel(const float* __restrict__ verts, int V, const int* __restrict__ faces,
                    const int* __restrict__ leaf_face, const int* __restrict__ mid_off,
                    const int* __restrict__ top_off, int K, int NM, int NT, float beta_leaf, float beta_group,
                    float4* __restrict__ ctri, float4* __restrict__ nodes, const uint8_t* __restrict__ body_active) {
    const int b = blockIdx.y;
    if (body_active != nullptr && !body_active[b]) return;
    const int node = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (node >= NT + NM + K) return;
    const bool is_leaf = node >= NT + NM;
    int l0, l1;
    if (is_leaf) { l0 = node - NT - NM; l1 = l0 + 1; }
    else if (node >= NT) { l0 = mid_off[node - NT]; l1 = mid_off[node - NT + 1]; }
    else { l0 = mid_off[top_off[node]]; l1 = mid_off[top_off[node + 1]]; }
    const float* vb = verts + (size_t)b * V * 3;
    const int half = lane >> 4, slot = lane & (WC_LEAF - 1);

    // pass 1: area-weighted centre (falls back to the plain centroid mean for zero-area nodes)
    float wsum = 0.f, cx = 0.f, cy = 0.f, cz = 0.f, ux = 0.f, uy = 0.f, uz = 0.f, cnt = 0.f;
    for (int l = l0 + half; l < l1; l += 2) {
        const int f = leaf_face[(size_t)l * WC_LEAF + slot];
        float4 A = make_float4(0.f, 0.f, 0.f, 0.f), Bv = A, C = A;
        if (f >= 0) {
            const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
            A = make_float4(vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2], 0.f);
            Bv = make_float4(vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2], 0.f);
            C = make_float4(vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2], 0.f);
            const float e1x = Bv.x - A.x, e1y = Bv.y - A.y, e1z = Bv.z - A.z;
            const float e2x = C.x - A.x, e2y = C.y - A.y, e2z = C.z - A.z;
            const float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
            A.w = nx; Bv.w = ny; C.w = nz;                       // face normal for half_solid_angle_n
            const float area = 0.5f * sqrtf(nx * nx + ny * ny + nz * nz);
            const float gx = (A.x + Bv.x + C.x) * (1.f / 3.f), gy = (A.y + Bv.y + C.y) * (1.f / 3.f),
                        gz = (A.z + Bv.z + C.z) * (1.f / 3.f);
            wsum += area; cx += area * gx; cy += area * gy; cz += area * gz;
            ux += gx; uy += gy; uz += gz; cnt += 1.f;
        }
        if (is_leaf) {
            float4* o = ctri + (((size_t)b * K + l) * WC_LEAF + slot) * 3;
            o[0] = A; o[1] = Bv; o[2] = C;
        }
    }
    wsum = warp_sum(wsum); cnt = warp_sum(cnt);
    float px, py, pz;
    if (wsum > 1e-30f) {
        const float inv = 1.f / wsum;
        px = warp_sum(cx) * inv; py = warp_sum(cy) * inv; pz = warp_sum(cz) * inv;
    } else {
        const float inv = 1.f / fmaxf(cnt, 1.f);
        px = warp_sum(ux) * inv; py = warp_sum(uy) * inv; pz = warp_sum(uz) * inv;
    }

    // pass 2: radius and moments about p
    float r2 = 0.f;
    float m0x = 0.f, m0y = 0.f, m0z = 0.f, tr = 0.f;
    float qxx = 0.f, qyy = 0.f, qzz = 0.f, qxy = 0.f, qxz = 0.f, qyz = 0.f;
    float uvx = 0.f, uvy = 0.f, uvz = 0.f;
    float txxx = 0.f, tyyy = 0.f, tzzz = 0.f, txxy = 0.f, txxz = 0.f, tyyx = 0.f, tyyz = 0.f, tzzx = 0.f, tzzy = 0.f,
          txyz = 0.f;
    for (int l = l0 + half; l < l1; l += 2) {
        const int f = leaf_face[(size_t)l * WC_LEAF + slot];
        if (f < 0) continue;
        const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        const float ax = vb[3 * i0] - px, ay = vb[3 * i0 + 1] - py, az = vb[3 * i0 + 2] - pz;
        const float bx = vb[3 * i1] - px, by = vb[3 * i1 + 1] - py, bz = vb[3 * i1 + 2] - pz;
        const float gx = vb[3 * i2] - px, gy = vb[3 * i2 + 1] - py, gz = vb[3 * i2 + 2] - pz;
        r2 = fmaxf(r2, fmaxf(ax * ax + ay * ay + az * az, fmaxf(bx * bx + by * by + bz * bz, gx * gx + gy * gy + gz * gz)));
        const float e1x = bx - ax, e1y = by - ay, e1z = bz - az, e2x = gx - ax, e2y = gy - ay, e2z = gz - az;
        const float nx = 0.5f * (e1y * e2z - e1z * e2y), ny = 0.5f * (e1z * e2x - e1x * e2z),
                    nz = 0.5f * (e1x * e2y - e1y * e2x);                        // area vector
        const float hx = (ax + bx + gx) * (1.f / 3.f), hy = (ay + by + gy) * (1.f / 3.f), hz = (az + bz + gz) * (1.f / 3.f);
        m0x += nx; m0y += ny; m0z += nz;
        tr += nx * hx + ny * hy + nz * hz;
        qxx += nx * hx; qyy += ny * hy; qzz += nz * hz;
        qxy += nx * hy + ny * hx; qxz += nx * hz + nz * hx; qyz += ny * hz + nz * hy;   // = 2 Q_ij
        // S = (1/3) sum over the three edge midpoints of m m^T (exact for quadratics on a triangle)
        const float m1x = 0.5f * (ax + bx), m1y = 0.5f * (ay + by), m1z = 0.5f * (az + bz);
        const float m2x = 0.5f * (bx + gx), m2y = 0.5f * (by + gy), m2z = 0.5f * (bz + gz);
        const float m3x = 0.5f * (gx + ax), m3y = 0.5f * (gy + ay), m3z = 0.5f * (gz + az);
        const float k3 = 1.f / 3.f;
        const float sxx = k3 * (m1x * m1x + m2x * m2x + m3x * m3x), syy = k3 * (m1y * m1y + m2y * m2y + m3y * m3y),
                    szz = k3 * (m1z * m1z + m2z * m2z + m3z * m3z);
        const float sxy = k3 * (m1x * m1y + m2x * m2y + m3x * m3y), sxz = k3 * (m1x * m1z + m2x * m2z + m3x * m3z),
                    syz = k3 * (m1y * m1z + m2y * m2z + m3y * m3z);
        const float trs = sxx + syy + szz;
        uvx += 2.f * (nx * sxx + ny * sxy + nz * sxz) + nx * trs;
        uvy += 2.f * (nx * sxy + ny * syy + nz * syz) + ny * trs;
        uvz += 2.f * (nx * sxz + ny * syz + nz * szz) + nz * trs;
        // (n.r)(r'S r) expanded into its 10 cubic coefficients
        txxx += nx * sxx; tyyy += ny * syy; tzzz += nz * szz;
        txxy += 2.f * nx * sxy + ny * sxx; txxz += 2.f * nx * sxz + nz * sxx;
        tyyx += 2.f * ny * sxy + nx * syy; tyyz += 2.f * ny * syz + nz * syy;
        tzzx += 2.f * nz * sxz + nx * szz; tzzy += 2.f * nz * syz + ny * szz;
        txyz += 2.f * (nx * syz + ny * sxz + nz * sxy);
    }
    r2 = warp_max(r2);
    m0x = warp_sum(m0x); m0y = warp_sum(m0y); m0z = warp_sum(m0z); tr = warp_sum(tr);
    qxx = warp_sum(qxx); qyy = warp_sum(qyy); qzz = warp_sum(qzz);
    qxy = warp_sum(qxy); qxz = warp_sum(qxz); qyz = warp_sum(qyz);
    uvx = warp_sum(uvx); uvy = warp_sum(uvy); uvz = warp_sum(uvz);
    txxx = warp_sum(txxx); tyyy = warp_sum(tyyy); tzzz = warp_sum(tzzz);
    txxy = warp_sum(txxy); txxz = warp_sum(txxz); tyyx = warp_sum(tyyx); tyyz = warp_sum(tyyz);
    tzzx = warp_sum(tzzx); tzzy = warp_sum(tzzy); txyz = warp_sum(txyz);
    if (lane == 0) {
        float4* o = nodes + ((size_t)b * (NT + NM + K) + node) * WC_NODE_F4;
        const float beta = is_leaf ? beta_leaf : beta_group;
        o[0] = make_float4(px, py, pz, r2 * (beta * beta * 1.0002f));
        o[1] = make_float4(0.5f * m0x, 0.5f * m0y, 0.5f * m0z, 0.5f * tr);
        o[2] = make_float4(-1.5f * qxx, -1.5f * qyy, -1.5f * qzz, -1.5f * qxy);
        o[3] = make_float4(-1.5f * qxz, -1.5f * qyz, -0.75f * uvx, -0.75f * uvy);
        o[4] = make_float4(-0.75f * uvz, 3.75f * txxx, 3.75f * tyyy, 3.75f * tzzz);
        o[5] = make_float4(3.75f * txxy, 3.75f * txxz, 3.75f * tyyx, 3.75f * tyyz);
        o[6] = make_float4(3.75f * tzzx, 3.75f * tzzy, 3.75f * txyz, 0.f);
    }
}

// Same records, one CTA per TOP group: the corners of the group's face slots are gathered from the posed
// vertices once into shared memory (9 floats per slot, stride 9 = conflict-free) and every node of the
// group -- the top itself, its mids, its leaves -- is then reduced from there by one warp, instead of
// gathering every face six times (three levels x two passes).
struct NodeMoments {
    float r2 = 0.f, m0x = 0.f, m0y = 0.f, m0z = 0.f, tr = 0.f;
    float qxx = 0.f, qyy = 0.f, qzz = 0.f, qxy = 0.f, qxz = 0.f, qyz = 0.f, uvx = 0.f, uvy = 0.f, uvz = 0.f;
    float txxx = 0.f, tyyy = 0.f, tzzz = 0.f, txxy = 0.f, txxz = 0.f, tyyx = 0.f, tyyz = 0.f, tzzx = 0.f, tzzy = 0.f, txyz = 0.f;
    __device__ __forceinline__ void add(float ax, float ay, float az, float bx, float by, float bz, float gx, float gy, float gz) {
        r2 = fmaxf(r2, fmaxf(ax * ax + ay * ay + az * az, fmaxf(bx * bx + by * by + bz * bz, gx * gx + gy * gy + gz * gz)));
        const float e1x = bx - ax, e1y = by - ay, e1z = bz - az, e2x = gx - ax, e2y = gy - ay, e2z = gz - az;
        const float nx = 0.5f * (e1y * e2z - e1z * e2y), ny = 0.5f * (e1z * e2x - e1x * e2z), nz = 0.5f * (e1x * e2y - e1y * e2x);
        const float hx = (ax + bx + gx) * (1.f / 3.f), hy = (ay + by + gy) * (1.f / 3.f), hz = (az + bz + gz) * (1.f / 3.f);
        m0x += nx; m0y += ny; m0z += nz;
        tr += nx * hx + ny * hy + nz * hz;
        qxx += nx * hx; qyy += ny * hy; qzz += nz * hz;
        qxy += nx * hy + ny * hx; qxz += nx * hz + nz * hx; qyz += ny * hz + nz * hy;
        const float m1x = 0.5f * (ax + bx), m1y = 0.5f * (ay + by), m1z = 0.5f * (az + bz);
        const float m2x = 0.5f * (bx + gx), m2y = 0.5f * (by + gy), m2z = 0.5f * (bz + gz);
        const float m3x = 0.5f * (gx + ax), m3y = 0.5f * (gy + ay), m3z = 0.5f * (gz + az);
        const float k3 = 1.f / 3.f;
        const float sxx = k3 * (m1x * m1x + m2x * m2x + m3x * m3x), syy = k3 * (m1y * m1y + m2y * m2y + m3y * m3y),
                    szz = k3 * (m1z * m1z + m2z * m2z + m3z * m3z);
        const float sxy = k3 * (m1x * m1y + m2x * m2y + m3x * m3y), sxz = k3 * (m1x * m1z + m2x * m2z + m3x * m3z),
                    syz = k3 * (m1y * m1z + m2y * m2z + m3y * m3z);
        const float trs = sxx + syy + szz;
        uvx += 2.f * (nx * sxx + ny * sxy + nz * sxz) + nx * trs;
        uvy += 2.f * (nx * sxy + ny * syy + nz * syz) + ny * trs;
        uvz += 2.f * (nx * sxz + ny * syz + nz * szz) + nz * trs;
        txxx += nx * sxx; tyyy += ny * syy; tzzz += nz * szz;
        txxy += 2.f * nx * sxy + ny * sxx; txxz += 2.f * nx * sxz + nz * sxx;
        tyyx += 2.f * ny * sxy + nx * syy; tyyz += 2.f * ny * syz + nz * syy;
        tzzx += 2.f * nz * sxz + nx * szz; tzzy += 2.f * nz * syz + ny * szz;
        txyz += 2.f * (nx * syz + ny * sxz + nz * sxy);
    }
    static constexpr int N_SUMS = 23;                            // every field but r2
    __device__ __forceinline__ void store_sums(float* p) const {
        const float v[N_SUMS] = {m0x, m0y, m0z, tr, qxx, qyy, qzz, qxy, qxz, qyz, uvx, uvy, uvz,
                                 txxx, tyyy, tzzz, txxy, txxz, tyyx, tyyz, tzzx, tzzy, txyz};
#pragma unroll
        for (int k = 0; k < N_SUMS; ++k) p[k] = v[k];
    }
    __device__ __forceinline__ void add_sums(const float* p) {
        m0x += p[0]; m0y += p[1]; m0z += p[2]; tr += p[3]; qxx += p[4]; qyy += p[5]; qzz += p[6]; qxy += p[7];
        qxz += p[8]; qyz += p[9]; uvx += p[10]; uvy += p[11]; uvz += p[12]; txxx += p[13]; tyyy += p[14];
        tzzz += p[15]; txxy += p[16]; txxz += p[17]; tyyx += p[18]; tyyz += p[19]; tzzx += p[20]; tzzy += p[21];
        txyz += p[22];
    }
    template <int W>
    __device__ __forceinline__ void reduce() {
        r2 = group_max<W>(r2);
        m0x = group_sum<W>(m0x); m0y = group_sum<W>(m0y); m0z = group_sum<W>(m0z); tr = group_sum<W>(tr);
        qxx = group_sum<W>(qxx); qyy = group_sum<W>(qyy); qzz = group_sum<W>(qzz);
        qxy = group_sum<W>(qxy); qxz = group_sum<W>(qxz); qyz = group_sum<W>(qyz);
        uvx = group_sum<W>(uvx); uvy = group_sum<W>(uvy); uvz = group_sum<W>(uvz);
        txxx = group_sum<W>(txxx); tyyy = group_sum<W>(tyyy); tzzz = group_sum<W>(tzzz);
        txxy = group_sum<W>(txxy); txxz = group_sum<W>(txxz); tyyx = group_sum<W>(tyyx); tyyz = group_sum<W>(tyyz);
        tzzx = group_sum<W>(tzzx); tzzy = group_sum<W>(tzzy); txyz = group_sum<W>(txyz);
    }
    __device__ __forceinline__ void write(float4* o, float px, float py, float pz, float beta) const {
        o[0] = make_float4(px, py, pz, r2 * (beta * beta * 1.0002f));
        o[1] = make_float4(0.5f * m0x, 0.5f * m0y, 0.5f * m0z, 0.5f * tr);
        o[2] = make_float4(-1.5f * qxx, -1.5f * qyy, -1.5f * qzz, -1.5f * qxy);
        o[3] = make_float4(-1.5f * qxz, -1.5f * qyz, -0.75f * uvx, -0.75f * uvy);
        o[4] = make_float4(-0.75f * uvz, 3.75f * txxx, 3.75f * tyyy, 3.75f * tzzz);
        o[5] = make_float4(3.75f * txxy, 3.75f * txxz, 3.75f * tyyx, 3.75f * tyyz);
        o[6] = make_float4(3.75f * tzzx, 3.75f * tzzy, 3.75f * txyz, 0.f);
    }
};

// Per-node record kept in shared memory while a top group is packed: centre, area, plain centroid sum and count
// (the centre of a zero-area node), then the 23 raw sums about that centre.
constexpr int PK_REC = 32, PK_SUMS = 8;

// Moments of a child node, taken about its own centre, re-expressed about the parent's centre p and added to `mo`.
// With d = child centre - p every relative position grows by d, a face centroid h -> h + d and the face's second
// moment S -> S + h d' + d h' + d d' (its three edge midpoints average to h), hence
//   M0' = M0,  tr' = tr + M0.d,  Q'_ab = Q_ab + M0_a d_b,
//   u'  = u + 2 (Q + Q') d + 2 d (tr + M0.d) + |d|^2 M0,
//   T'(r) = T(r) + 2 (d.r) r'Q r + (d.r)^2 (M0.r)
// -- exact identities, so a group node costs one such shift per child instead of a second pass over its faces.
__device__ __forceinline__ void add_shifted(NodeMoments& mo, const float* __restrict__ rec, float px, float py, float pz) {
    const float dx = rec[0] - px, dy = rec[1] - py, dz = rec[2] - pz;
    const float* s = rec + PK_SUMS;
    const float m0x = s[0], m0y = s[1], m0z = s[2], tr = s[3], qxx = s[4], qyy = s[5], qzz = s[6], qxy = s[7], qxz = s[8],
                qyz = s[9];
    const float dm = dx * m0x + dy * m0y + dz * m0z, d2 = dx * dx + dy * dy + dz * dz, trd = tr + dm;
    mo.m0x += m0x; mo.m0y += m0y; mo.m0z += m0z;
    mo.tr += trd;
    mo.qxx += qxx + m0x * dx; mo.qyy += qyy + m0y * dy; mo.qzz += qzz + m0z * dz;
    mo.qxy += qxy + m0x * dy + m0y * dx; mo.qxz += qxz + m0x * dz + m0z * dx; mo.qyz += qyz + m0y * dz + m0z * dy;
    mo.uvx += s[10] + 2.f * (2.f * qxx * dx + qxy * dy + qxz * dz) + 2.f * dx * trd + d2 * m0x;
    mo.uvy += s[11] + 2.f * (qxy * dx + 2.f * qyy * dy + qyz * dz) + 2.f * dy * trd + d2 * m0y;
    mo.uvz += s[12] + 2.f * (qxz * dx + qyz * dy + 2.f * qzz * dz) + 2.f * dz * trd + d2 * m0z;
    mo.txxx += s[13] + 2.f * dx * qxx + dx * dx * m0x;
    mo.tyyy += s[14] + 2.f * dy * qyy + dy * dy * m0y;
    mo.tzzz += s[15] + 2.f * dz * qzz + dz * dz * m0z;
    mo.txxy += s[16] + 2.f * (dx * qxy + dy * qxx) + dx * dx * m0y + 2.f * dx * dy * m0x;
    mo.txxz += s[17] + 2.f * (dx * qxz + dz * qxx) + dx * dx * m0z + 2.f * dx * dz * m0x;
    mo.tyyx += s[18] + 2.f * (dy * qxy + dx * qyy) + dy * dy * m0x + 2.f * dx * dy * m0y;
    mo.tyyz += s[19] + 2.f * (dy * qyz + dz * qyy) + dy * dy * m0z + 2.f * dy * dz * m0y;
    mo.tzzx += s[20] + 2.f * (dz * qxz + dx * qzz) + dz * dz * m0x + 2.f * dx * dz * m0z;
    mo.tzzy += s[21] + 2.f * (dz * qyz + dy * qzz) + dz * dz * m0y + 2.f * dy * dz * m0z;
    mo.txyz += s[22] + 2.f * (dx * qyz + dy * qxz + dz * qxy) + 2.f * (dx * dy * m0z + dx * dz * m0y + dy * dz * m0x);
}

__global__ void __launch_bounds__(256)
cluster_pack_top_kernel(const float* __restrict__ verts, int V, const int* __restrict__ faces,
                        const int* __restrict__ leaf_face, const int* __restrict__ mid_off,
                        const int* __restrict__ top_off, int K, int NM, int NT, float beta_leaf, float beta_group,
                        float4* __restrict__ ctri, float4* __restrict__ nodes, const uint8_t* __restrict__ body_active) {
    extern __shared__ float s_c[];                               // [n_slots][9] corners, [n_slots] validity, records
    const int b = blockIdx.y, t = blockIdx.x;
    if (body_active != nullptr && !body_active[b]) return;
    const int m0 = top_off[t], m1 = top_off[t + 1];
    const int l0 = mid_off[m0], l1 = mid_off[m1];
    const int nm = m1 - m0, nl = l1 - l0;
    const int n_slots = nl * WC_LEAF;
    float* s_ok = s_c + (size_t)n_slots * 9;
    float* s_leaf = s_ok + n_slots;                              // [nl][PK_REC]
    float* s_mid = s_leaf + (size_t)nl * PK_REC;                 // [nm][PK_REC]
    const float* vb = verts + (size_t)b * V * 3;
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
        const int f = leaf_face[(size_t)l0 * WC_LEAF + i];
        float c[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (f >= 0) {
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const int v = faces[3 * f + e];
                c[3 * e] = vb[3 * v]; c[3 * e + 1] = vb[3 * v + 1]; c[3 * e + 2] = vb[3 * v + 2];
            }
            const float e1x = c[3] - c[0], e1y = c[4] - c[1], e1z = c[5] - c[2];
            const float e2x = c[6] - c[0], e2y = c[7] - c[1], e2z = c[8] - c[2];
            nx = e1y * e2z - e1z * e2y; ny = e1z * e2x - e1x * e2z; nz = e1x * e2y - e1y * e2x;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) s_c[(size_t)i * 9 + k] = c[k];
        s_ok[i] = f >= 0 ? 1.f : 0.f;
        float4* o = ctri + (((size_t)b * K + l0) * WC_LEAF + i) * 3;
        o[0] = make_float4(c[0], c[1], c[2], nx);                // face normal for half_solid_angle_n
        o[1] = make_float4(c[3], c[4], c[5], ny);
        o[2] = make_float4(c[6], c[7], c[8], nz);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* nb = nodes + (size_t)b * (NT + NM + K) * WC_NODE_F4;

    // scale of raw sum k in the node record (NodeMoments::write), which holds it at float 4 + k
    auto rec_scale = [](int k) { return k < 4 ? 0.5f : k < 10 ? -1.5f : k < 13 ? -0.75f : 3.75f; };
    // ---- leaves from their faces, one per half-warp (lane = face slot); the raw sums stay in shared memory
    for (int pair = warp; 2 * pair < nl; pair += 8) {
        const int j = 2 * pair + (lane >> 4), sub = lane & 15;
        const bool live = j < nl;
        const int i = j * WC_LEAF + sub;
        const bool ok = live && s_ok[i] != 0.f;
        const float* c = s_c + (size_t)(ok ? i : 0) * 9;
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // area, area * centroid, centroid, count
        if (ok) {
            const float e1x = c[3] - c[0], e1y = c[4] - c[1], e1z = c[5] - c[2];
            const float e2x = c[6] - c[0], e2y = c[7] - c[1], e2z = c[8] - c[2];
            const float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
            const float area = 0.5f * sqrtf(nx * nx + ny * ny + nz * nz);
            const float gx = (c[0] + c[3] + c[6]) * (1.f / 3.f), gy = (c[1] + c[4] + c[7]) * (1.f / 3.f),
                        gz = (c[2] + c[5] + c[8]) * (1.f / 3.f);
            cs[0] = area; cs[1] = area * gx; cs[2] = area * gy; cs[3] = area * gz;
            cs[4] = gx; cs[5] = gy; cs[6] = gz; cs[7] = 1.f;
        }
        group_sums<16>(cs, lane);
        // area-weighted centre (plain centroid mean for zero-area nodes)
        const bool weighted = cs[0] > 1e-30f;
        const float inv = weighted ? 1.f / cs[0] : 1.f / fmaxf(cs[7], 1.f);
        const float px = (weighted ? cs[1] : cs[4]) * inv, py = (weighted ? cs[2] : cs[5]) * inv,
                    pz = (weighted ? cs[3] : cs[6]) * inv;
        NodeMoments mo;
        if (ok) mo.add(c[0] - px, c[1] - py, c[2] - pz, c[3] - px, c[4] - py, c[5] - pz, c[6] - px, c[7] - py, c[8] - pz);
        const float r2 = group_max<16>(mo.r2);
        float v[32];
        mo.store_sums(v);
#pragma unroll
        for (int k = NodeMoments::N_SUMS; k < 32; ++k) v[k] = 0.f;
        ReduceScatter<8, 32>::run(v, lane);                      // lane `sub` holds the totals 2 sub, 2 sub + 1
        if (live && sub < 12) {
            float* r = s_leaf + (size_t)j * PK_REC;
            float* o = (float*)(nb + (size_t)(NT + NM + l0 + j) * WC_NODE_F4);
            r[PK_SUMS + 2 * sub] = v[0]; r[PK_SUMS + 2 * sub + 1] = v[1];
            *(float2*)(o + 4 + 2 * sub) = make_float2(rec_scale(2 * sub) * v[0], rec_scale(2 * sub + 1) * v[1]);
            if (sub == 0) {
                r[0] = px; r[1] = py; r[2] = pz; r[3] = cs[0]; r[4] = cs[4]; r[5] = cs[5]; r[6] = cs[6]; r[7] = cs[7];
                *(float4*)o = make_float4(px, py, pz, r2 * (beta_leaf * beta_leaf * 1.0002f));
            }
        }
    }
    __syncthreads();

    // ---- a group node from its children: centre from the children's areas and centres, radius from one pass
    //      over its corners, moments by shifting the children's (add_shifted).  One warp per mid.
    struct Centre { float px, py, pz, w, ux, uy, uz, cnt; };
    auto centre_of = [&](const float* child, int nc) {
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = lane; c < nc; c += 32) {
            const float* r = child + (size_t)c * PK_REC;
            cs[0] += r[3]; cs[1] += r[3] * r[0]; cs[2] += r[3] * r[1]; cs[3] += r[3] * r[2];
            cs[4] += r[4]; cs[5] += r[5]; cs[6] += r[6]; cs[7] += r[7];
        }
        group_sums<32>(cs, lane);
        const bool weighted = cs[0] > 1e-30f;
        const float inv = weighted ? 1.f / cs[0] : 1.f / fmaxf(cs[7], 1.f);
        return Centre{(weighted ? cs[1] : cs[4]) * inv, (weighted ? cs[2] : cs[5]) * inv, (weighted ? cs[3] : cs[6]) * inv,
                      cs[0], cs[4], cs[5], cs[6], cs[7]};
    };
    auto radius2 = [&](const Centre& p, int s0, int s1, int first, int stride) {
        float r2 = 0.f;
        for (int i = s0 + first; i < s1; i += stride) {
            if (s_ok[i] == 0.f) continue;
            const float* c = s_c + (size_t)i * 9;
            const float ax = c[0] - p.px, ay = c[1] - p.py, az = c[2] - p.pz, bx = c[3] - p.px, by = c[4] - p.py,
                        bz = c[5] - p.pz, gx = c[6] - p.px, gy = c[7] - p.py, gz = c[8] - p.pz;
            r2 = fmaxf(r2, fmaxf(ax * ax + ay * ay + az * az, fmaxf(bx * bx + by * by + bz * bz, gx * gx + gy * gy + gz * gz)));
        }
        return group_max<32>(r2);
    };
    // the children's moments about p, summed over the warp: lane k < 23 returns raw sum k
    auto shifted_sum = [&](const float* child, int nc, const Centre& p) {
        NodeMoments mo;
        for (int c = lane; c < nc; c += 32) add_shifted(mo, child + (size_t)c * PK_REC, p.px, p.py, p.pz);
        float v[32];
        mo.store_sums(v);
#pragma unroll
        for (int k = NodeMoments::N_SUMS; k < 32; ++k) v[k] = 0.f;
        ReduceScatter<16, 32>::run(v, lane);
        return v[0];
    };
    for (int n = warp; n < nm; n += 8) {
        const int c0 = mid_off[m0 + n] - l0, c1 = mid_off[m0 + n + 1] - l0;
        const float* child = s_leaf + (size_t)c0 * PK_REC;
        const Centre p = centre_of(child, c1 - c0);
        const float r2 = radius2(p, c0 * WC_LEAF, c1 * WC_LEAF, lane, 32);
        const float total = shifted_sum(child, c1 - c0, p);
        float* r = s_mid + (size_t)n * PK_REC;
        float* o = (float*)(nb + (size_t)(NT + m0 + n) * WC_NODE_F4);
        if (lane < 24) { r[PK_SUMS + lane] = total; o[4 + lane] = rec_scale(lane) * total; }
        if (lane == 0) {
            r[0] = p.px; r[1] = p.py; r[2] = p.pz; r[3] = p.w; r[4] = p.ux; r[5] = p.uy; r[6] = p.uz; r[7] = p.cnt;
            *(float4*)o = make_float4(p.px, p.py, p.pz, r2 * (beta_group * beta_group * 1.0002f));
        }
    }
    __syncthreads();
    // ---- the top node from its mids: every warp takes a share of the radius pass, warp 0 shifts the moments
    {
        __shared__ float s_r2[8];
        const Centre p = centre_of(s_mid, nm);                   // same sums in the same order in every warp
        const float r2w = radius2(p, 0, n_slots, threadIdx.x, blockDim.x);
        if (lane == 0) s_r2[warp] = r2w;
        __syncthreads();
        if (warp == 0) {
            const float total = shifted_sum(s_mid, nm, p);
            float* o = (float*)(nb + (size_t)t * WC_NODE_F4);
            if (lane < 24) o[4 + lane] = rec_scale(lane) * total;
            if (lane == 0) {
                float r2 = 0.f;
                for (int w = 0; w < 8; ++w) r2 = fmaxf(r2, s_r2[w]);
                *(float4*)o = make_float4(p.px, p.py, p.pz, r2 * (beta_group * beta_group * 1.0002f));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// device: hierarchical winding kernel
// ------------------------------------------------------------------------------------------
// half the far-field solid angle of one node; (x, y, z) = node centre - query, d2 = |.|^2 > 0
__device__ __forceinline__ float node_far_field(const float4* __restrict__ rec, float x, float y, float z, float d2) {
    const float4 f1 = __ldg(rec + 1), f2 = __ldg(rec + 2), f3 = __ldg(rec + 3), f4 = __ldg(rec + 4),
                 f5 = __ldg(rec + 5), f6 = __ldg(rec + 6);
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    float inv;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(d2));
    const float i2 = inv * inv, i3 = i2 * inv, i5 = i3 * i2, i7 = i5 * i2;
    const float t0 = fmaf(f1.x, x, fmaf(f1.y, y, fmaf(f1.z, z, f1.w)));
    float t1 = fmaf(f2.x, xx, fmaf(f2.y, yy, f2.z * zz));
    t1 = fmaf(f2.w, xy, fmaf(f3.x, xz, fmaf(f3.y, yz, t1)));
    t1 = fmaf(f3.z, x, fmaf(f3.w, y, fmaf(f4.x, z, t1)));
    const float cxp = fmaf(f4.y, xx, fmaf(f5.x, xy, fmaf(f5.y, xz, fmaf(f5.z, yy, fmaf(f6.x, zz, f6.z * yz)))));
    const float cyp = fmaf(f4.z, yy, fmaf(f5.w, yz, f6.y * zz));
    const float t2 = fmaf(x, cxp, fmaf(y, cyp, z * (f4.w * zz)));
    return fmaf(t0, i3, fmaf(t1, i5, t2 * i7));
}

// grid (groups of WC_WARPS vertex tiles, top splits, bodies); one query per lane, one vertex tile per warp.
// Tops and mids open for the whole warp as soon as one query is near; leaves are near or far per query.
// In the near pass the two half-warps serve two near queries at a time, one lane per face of the leaf.
__global__ void __launch_bounds__(WC_WARPS * 32)
winding_cluster_kernel(const float* __restrict__ points, const int* __restrict__ vtile,
                       const float4* __restrict__ ctri, const float4* __restrict__ nodes,
                       const int* __restrict__ mid_off, const int* __restrict__ top_off,
                       float* __restrict__ partial, int Q, int T, int K, int NM, int NT, int tops_per_split, int S,
                       const int* __restrict__ q_counts, const uint8_t* __restrict__ body_active, float open_leaf,
                       float open_group) {
    // open_leaf / open_group scale the squared opening radii baked into the node records (1 = as packed)
    __shared__ float s_acc[WC_WARPS][32 * (WC_LEAF + 1)];
    const int b = blockIdx.z, split = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (body_active != nullptr && !body_active[b]) return;
    const int tile = blockIdx.x * WC_WARPS + warp;               // one query tile per warp
    if (tile >= T) return;
    int qv, qi;
    if (vtile != nullptr) {                                      // mesh vertices in tile order
        qv = vtile[tile * 32 + lane];                            // -1 = padding lane (a tile holds >= 1 vertex)
        qi = qv >= 0 ? qv : vtile[tile * 32];
    } else {                                                     // 32 consecutive points
        const int n = q_counts != nullptr ? min(Q, q_counts[b]) : Q;
        if (tile * 32 >= n) return;
        qi = tile * 32 + lane;
        qv = qi < n ? qi : -1;
        qi = min(qi, n - 1);
    }
    const float* vb = points + (size_t)b * Q * 3;
    const float px = vb[3 * qi], py = vb[3 * qi + 1], pz = vb[3 * qi + 2];
    float* acc = s_acc[warp];
    for (int k = lane; k < 32 * (WC_LEAF + 1); k += 32) acc[k] = 0.f;
    __syncwarp();

    const float4* nb = nodes + (size_t)b * (NT + NM + K) * WC_NODE_F4;
    const float4* tb = ctri + (size_t)b * K * WC_LEAF * 3;
    const int half = lane >> 4, slot = lane & (WC_LEAF - 1);
    // a query's work sits in the one or two top groups it is near, and neighbouring groups are neighbours in index:
    // the splits take the tops INTERLEAVED (split, split + S, ...) so that the opened groups of a tile are shared
    // out among its splits instead of landing in one of them
    (void)tops_per_split;
    float far = 0.f;
    for (int t = split; t < NT; t += S) {
        const float4* trec = nb + (size_t)t * WC_NODE_F4;
        const float4 tc = __ldg(trec);
        const float tx = tc.x - px, ty = tc.y - py, tz = tc.z - pz;
        const float td2 = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
        if (!__any_sync(0xffffffffu, td2 < tc.w * open_group)) {
            far += node_far_field(trec, tx, ty, tz, td2);
            continue;
        }
        const int m0 = __ldg(top_off + t), m1 = __ldg(top_off + t + 1);
        for (int m = m0; m < m1; ++m) {
            const float4* mrec = nb + (size_t)(NT + m) * WC_NODE_F4;
            const float4 mc = __ldg(mrec);
            const float mx = mc.x - px, my = mc.y - py, mz = mc.z - pz;
            const float md2 = fmaf(mz, mz, fmaf(my, my, mx * mx));
            if (!__any_sync(0xffffffffu, md2 < mc.w * open_group)) {
                far += node_far_field(mrec, mx, my, mz, md2);
                continue;
            }
            const int l0 = __ldg(mid_off + m), l1 = __ldg(mid_off + m + 1);
            for (int l = l0; l < l1; ++l) {
                const float4* lrec = nb + (size_t)(NT + NM + l) * WC_NODE_F4;
                const float4 lc = __ldg(lrec);
                const float lx = lc.x - px, ly = lc.y - py, lz = lc.z - pz;
                const float ld2 = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
                const bool near = ld2 < lc.w * open_leaf;
                unsigned nm = __ballot_sync(0xffffffffu, near);
                if (nm != 0xffffffffu) {
                    const float v = node_far_field(lrec, lx, ly, lz, ld2);
                    far += near ? 0.f : v;
                }
                if (nm != 0u) {                                  // lanes become faces of this leaf
                    const float4* f = tb + ((size_t)l * WC_LEAF + slot) * 3;
                    const float4 A = __ldg(f), Bv = __ldg(f + 1), C = __ldg(f + 2);
                    while (nm != 0u) {                           // two near queries per step, one per half-warp
                        const int q0 = __ffs(nm) - 1;
                        nm &= nm - 1u;
                        const int q1 = nm != 0u ? __ffs(nm) - 1 : -1;
                        nm &= nm - 1u;                           // 0 & anything = 0 when nothing is left
                        const int q = half == 0 ? q0 : q1;
                        const int src = q >= 0 ? q : q0;
                        const float qx = __shfl_sync(0xffffffffu, px, src), qy = __shfl_sync(0xffffffffu, py, src),
                                    qz = __shfl_sync(0xffffffffu, pz, src);
                        if (q >= 0) acc[q * (WC_LEAF + 1) + slot] += half_solid_angle_n(qx, qy, qz, A, Bv, C);
                        // a query served by one half-warp in this step may be served by the other half in a later
                        // one: order the read-modify-writes of the tile across the lanes (racecheck-clean)
                        __syncwarp();
                    }
                }
            }
        }
    }
    __syncwarp();
    float near_sum = 0.f;
#pragma unroll
    for (int j = 0; j < WC_LEAF; ++j) near_sum += acc[lane * (WC_LEAF + 1) + j];
    if (qv >= 0) partial[((size_t)b * S + split) * Q + qi] = far + near_sum;
}

// sums the split partials in a fixed order, scales by 1 / (2 pi) and lists the queries whose value is
// within WC_MARGIN of the 0.99 threshold of losses.py:82 for exact re-evaluation
// early_ext (optional, [B][V]): the exterior flags as far as they are known before the exact re-evaluation -- final
// for every query outside the band, 0 ("interior") for the listed ones, which is the safe side for the consumer that
// starts on them early (the mixed nearest-vertex query searches interior vertices without a limit)
__global__ void cluster_finalize_kernel(const float* __restrict__ partial, int V, int S, float* __restrict__ winding,
                                        int* __restrict__ refine_list, const int* __restrict__ q_counts,
                                        const uint8_t* __restrict__ body_active, float margin,
                                        uint8_t* __restrict__ early_ext) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= V) return;
    if ((body_active != nullptr && !body_active[b]) || (q_counts != nullptr && q >= q_counts[b])) {
        winding[(size_t)b * V + q] = 0.f;
        if (early_ext != nullptr) early_ext[(size_t)b * V + q] = 1;
        return;
    }
    const float* p = partial + (size_t)b * S * V + q;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += p[(size_t)s * V];
    const float w = acc * 0.159154943091895336f;
    winding[(size_t)b * V + q] = w;
    const bool listed = fabsf(w - 0.99f) < margin;
    if (listed) {
        const int k = atomicAdd(refine_list, 1);
        refine_list[1 + k] = b * V + q;
    }
    if (early_ext != nullptr) early_ext[(size_t)b * V + q] = (!listed && w <= 0.99f) ? 1 : 0;
}

// exact winding number (all face slots, one lane per slot) of every listed query; one CTA per entry, its
// warps stride over the packed face slots and their partial sums are combined in a fixed order
__global__ void __launch_bounds__(256)
cluster_refine_kernel(const float* __restrict__ verts, const float4* __restrict__ ctri, int V, int K,
                      const int* __restrict__ refine_list, float* __restrict__ winding, int* __restrict__ stats) {
    __shared__ float s_part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = refine_list[0];
    if (stats != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *stats = count;   // tuch_topology_query_stats
    const int n_slots = K * WC_LEAF;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const int id = refine_list[1 + e];
        const int b = id / V, q = id - b * V;
        const float* p = verts + ((size_t)b * V + q) * 3;
        const float px = p[0], py = p[1], pz = p[2];
        const float4* t = ctri + (size_t)b * n_slots * 3;
        float acc = 0.f;
        // ~60 trips per thread over L2-resident triangles: the loop is load-latency bound, so eight trips' loads are
        // issued together (the adds into `acc` stay in index order: the value does not depend on the unrolling)
#pragma unroll 8
        for (int i = threadIdx.x; i < n_slots; i += 256) {
            const float4 A = __ldg(t + 3 * i), Bv = __ldg(t + 3 * i + 1), C = __ldg(t + 3 * i + 2);
            acc += half_solid_angle_n(px, py, pz, A, Bv, C);
        }
        acc = warp_sum(acc);
        if (lane == 0) s_part[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_part[w];
            winding[(size_t)b * V + q] = tot * 0.159154943091895336f;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int cluster_splits(int B, int T, int NT, int sm_count) {
    const int qtiles = cdiv(T, WC_WARPS);
    // CTAs wanted: several per SM slot, so that a small batch is cut into pieces short enough to balance over the
    // SMs (at 32 bodies one split gave 928 CTAs of ~0.5 ms each on 1184 slots: a single ragged wave).  The splits
    // take the top groups interleaved, so any count up to NT works; 4 bounds the partial-sum traffic.
    const long long want = (long long)sm_count * 16;
    const int S = (int)((want + (long long)qtiles * B - 1) / ((long long)qtiles * B));
    return std::max(1, std::min(S, std::min(NT, 4)));
}

void cluster_pack_betas(const ClusterJob& j, float* beta_leaf, float* beta_group) {
    // development knobs (scripts/sweep_beta.sh sweeps them); the shipped values are the constants
    static const float env_leaf = getenv("TUCH_WC_BETA") ? (float)atof(getenv("TUCH_WC_BETA")) : 0.f;
    static const float env_group = getenv("TUCH_WC_BETA_SUPER") ? (float)atof(getenv("TUCH_WC_BETA_SUPER")) : 0.f;
    *beta_leaf = env_leaf > 0.f ? env_leaf : j.beta_leaf;
    *beta_group = env_group > 0.f ? env_group : j.beta_group;
}

int launch_cluster_pack(const ClusterJob& j, cudaStream_t st) {
    if (j.B == 0) return 0;
    KernelTimer timer("cluster_pack_kernel", st);
    dim3 grid(cdiv(j.NT + j.NM + j.K, 4), j.B);
    float beta_leaf, beta_group;
    cluster_pack_betas(j, &beta_leaf, &beta_group);
    const size_t smem = (size_t)j.max_top_leaves * (WC_LEAF * 10 + 2 * PK_REC) * sizeof(float);
    if (j.max_top_leaves > 0 && smem <= 200 * 1024 && !j.direct_pack) {
        static size_t attr_smem = 0;
        if (smem > attr_smem) {
            TUCH_CUDA(cudaFuncSetAttribute(cluster_pack_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem = smem;
        }
        dim3 g2(j.NT, j.B);
        cluster_pack_top_kernel<<<g2, 256, smem, st>>>(j.verts, j.V, j.faces, j.leaf_face, j.mid_off, j.top_off, j.K, j.NM,
                                                       j.NT, beta_leaf, beta_group, j.ctri, j.nodes, j.body_active);
    } else {                                                     // a top group too large for shared memory
        cluster_pack_kernel<<<grid, 128, 0, st>>>(j.verts, j.V, j.faces, j.leaf_face, j.mid_off, j.top_off, j.K, j.NM, j.NT,
                                                  beta_leaf, beta_group, j.ctri, j.nodes, j.body_active);
    }
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

// the far-field / near-field traversal alone, for the bodies [b0, b0 + nb) of the job
int launch_cluster_traverse(const ClusterJob& j, int b0, int nb, cudaStream_t st) {
    const float* points = j.points != nullptr ? j.points : j.verts;
    const int Q = j.points != nullptr ? j.Q : j.V;
    const int T = j.vtile != nullptr ? j.T : cdiv(Q, 32);
    const float rl = j.packed_beta_leaf > 0.f ? j.beta_leaf / j.packed_beta_leaf : 1.f;
    const float rg = j.packed_beta_group > 0.f ? j.beta_group / j.packed_beta_group : 1.f;
    const float open_leaf = rl * rl, open_group = rg * rg;
    const int per = cdiv(j.NT, j.S);
    dim3 grid(cdiv(T, WC_WARPS), j.S, nb);
    KernelTimer timer(j.vtile != nullptr ? "winding_kernel" : "winding_kernel_points", st);
    winding_cluster_kernel<<<grid, WC_WARPS * 32, 0, st>>>(
        points + (size_t)b0 * Q * 3, j.vtile, j.ctri + (size_t)b0 * j.K * WC_LEAF * 3,
        j.nodes + (size_t)b0 * (j.NT + j.NM + j.K) * WC_NODE_F4, j.mid_off, j.top_off, j.partial + (size_t)b0 * j.S * Q, Q, T,
        j.K, j.NM, j.NT, per, j.S, j.q_counts != nullptr ? j.q_counts + b0 : nullptr,
        j.body_active != nullptr ? j.body_active + b0 : nullptr, open_leaf, open_group);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

// split sum + list of the queries near the threshold ...
int launch_cluster_finalize(const ClusterJob& j, uint8_t* early_ext, cudaStream_t st) {
    const int Q = j.points != nullptr ? j.Q : j.V;
    TUCH_CUDA(cudaMemsetAsync(j.refine_list, 0, sizeof(int), st));
    dim3 grid(cdiv(Q, 256), j.B);
    static const float env_margin = getenv("TUCH_WC_MARGIN") ? (float)atof(getenv("TUCH_WC_MARGIN")) : 0.f;   // dev knob
    cluster_finalize_kernel<<<grid, 256, 0, st>>>(j.partial, Q, j.S, j.winding, j.refine_list, j.q_counts, j.body_active,
                                                  env_margin > 0.f && j.points == nullptr ? env_margin : j.margin, early_ext);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

// ... then their exact re-evaluation
int launch_cluster_refine(const ClusterJob& j, cudaStream_t st) {
    const float* points = j.points != nullptr ? j.points : j.verts;
    const int Q = j.points != nullptr ? j.Q : j.V;
    KernelTimer timer("winding_refine_kernel", st);
    cluster_refine_kernel<<<sm_count() * 4, 256, 0, st>>>(points, j.ctri, Q, j.K, j.refine_list, j.winding, j.stats);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_cluster_finish(const ClusterJob& j, cudaStream_t st) {
    if (int rc = launch_cluster_finalize(j, nullptr, st)) return rc;
    return launch_cluster_refine(j, st);
}

// queries against the packed hierarchy of launch_cluster_pack
int launch_cluster_query(const ClusterJob& j, cudaStream_t st) {
    const int Q = j.points != nullptr ? j.Q : j.V;
    if (j.B == 0 || Q == 0) return 0;
    if (int rc = launch_cluster_traverse(j, 0, j.B, st)) return rc;
    return launch_cluster_finish(j, st);
}

// Tried and not kept (round 2): pack + traverse in chunks of bodies whose packed leaf triangles fit the L2, so that
// the traversal reads what the pack just wrote from L2 instead of from HBM (at 256 bodies they are 191 MB, 9x the
// kernel's algorithmic bytes).  Measured at 256 bodies, ms per iteration: one chunk 3.67, 128 bodies 3.79, 96: 3.91,
// 64: 4.02, 32: 4.50 -- every chunk pays its own ragged last wave, and HBM is not what binds the traversal.
int launch_winding_clusters(const ClusterJob& j, cudaStream_t st) {
    if (j.B == 0 || j.V == 0) return 0;
    if (int rc = launch_cluster_pack(j, st)) return rc;
    return launch_cluster_query(j, st);
}

}  // namespace tuch
