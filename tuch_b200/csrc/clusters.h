// Internal: face-cluster hierarchy and the far-field ("fast winding number") kernels (clusters.cu).
#pragma once
#include <vector>

#include "kernels.h"

namespace tuch {

constexpr int WC_LEAF = 16;          // faces per leaf cluster = one half-warp lane per face in the near pass
constexpr int WC_MID_LEAVES = 8;     // leaves per mid-level group (128 faces) ...
constexpr int WC_TOP_LEAVES = 64;    // ... and per top-level group (1024 faces) of the far field
constexpr int WC_NODE_F4 = 7;        // float4 per node record (centre + radius, scaled moments)
constexpr int WC_WARPS = 8;          // warps per CTA of the winding kernel; one query per lane
// Opening radii: a leaf is "far" for a query beyond WC_BETA x its radius, a mid / top group (larger, so larger
// absolute error) beyond WC_BETA_GROUP x.  Measured on the B200 at 256 bodies (scripts/sweep_beta.sh, lattice body,
// folded-arm poses; time = pack + winding + exact re-evaluation, error = max |w_fast - w_all_faces|):
//   (2.0, 2.5) 3.03 ms 3.9e-3 | (2.0, 2.0) 2.67 ms 7.7e-3 | (1.8, 2.0) 2.40 ms 9.2e-3 | (1.6, 2.0) 2.19 ms 1.05e-2
//   (1.4, 2.0) 2.00 ms 1.76e-2;   exterior flags identical to the all-faces sum in every row.
// Again with the refined leaves and the whole iteration as the clock (ms per stage-2 iteration at 256 bodies, leaf
// radius / margin / max error): 1.6 / 0.06: 2.76, 1.07e-2 | 1.5 / 0.08: 2.70, 1.47e-2 | 1.4 / 0.10: 2.64, 1.93e-2 |
// 1.3 / 0.13: 2.58, 2.69e-2 | 1.2 / 0.16: 2.57, 4.18e-2.
// The flags only need |error| < WC_MARGIN (everything nearer the threshold is re-evaluated over all faces), so the
// radii are chosen for time with the margin at >= 5x the measured error; tests/test_clusters_cpu.py bounds the
// leaf-level error WITHOUT cancellation by half the margin.
constexpr float WC_BETA = 1.4f;
constexpr float WC_BETA_GROUP = 2.0f;
// off-surface point queries (the 1 mm offset HD points of loss.py:295-297): interior points sit at 1.0, only 0.01
// above the 0.99 threshold, so the far field opens later (3 / 3.5 radii) and every value within 0.02 of the
// threshold -- 4x the largest far-field error measured for the looser on-surface radii -- is re-evaluated exactly
constexpr float WC_BETA_POINTS = 3.0f, WC_BETA_GROUP_POINTS = 3.5f, WC_MARGIN_POINTS = 0.02f;
constexpr float WC_MARGIN = 0.10f;   // |w - 0.99| below this is re-evaluated exactly (5.2 x the worst far-field
                                     // error measured at these opening parameters: 1.93e-2)

// Host-side hierarchy of one mesh topology, built from the faces and ONE set of vertex positions
// (the template, or the first body seen).
struct ClusterTree {
    std::vector<int> leaf_face;      // [K][WC_LEAF] face id or -1 (padding)
    std::vector<int> mid_off;        // [NM + 1] leaf ranges of the mid-level groups
    std::vector<int> top_off;        // [NT + 1] mid ranges of the top-level groups
    std::vector<int> vtile;          // [T][32] vertex tiles: vertex id or -1 (padding); the 32 vertices of
                                     // a tile are neighbours on the mesh (queries of one warp, candidate
                                     // rows of one mask word)
    std::vector<int> vgroup_off;     // [NG + 1] tile ranges of the vertex-tile groups (<= 8 tiles each)
    int K = 0, NM = 0, NT = 0, T = 0, NG = 0, max_top_leaves = 0;
};
int build_cluster_tree(const int* faces, int F, int V, const float* verts, ClusterTree& out);

struct ClusterJob {
    const float* verts;              // [B][V][3] mesh vertices
    const int* faces;                // [F][3]
    const int* leaf_face;            // [K][WC_LEAF]
    const int* mid_off;              // [NM + 1]
    const int* top_off;              // [NT + 1]
    const int* vtile;                // [T][32] query order when the queries are the mesh vertices, else NULL
    float4* ctri;                    // [B][K][WC_LEAF][3] scratch: corners a | b | c per face slot, the face's
                                     // normal (b - a) x (c - a) in the three w components
    float4* nodes;                   // [B][NT + NM + K][WC_NODE_F4] scratch: tops, then mids, then leaves
    float* partial;                  // [B][S][Q] scratch
    float* winding;                  // [B][Q] out
    int* refine_list;                // [1 + B * Q] scratch: count, then b * Q + q entries
    int B, V, K, NM, NT, S, T;
    // queries: the mesh vertices themselves (points == verts, Q == V, vtile != NULL), or arbitrary points
    // [B][Q][3] taken 32 consecutive ones per warp, of which only the first q_counts[b] are valid
    const float* points = nullptr;
    int Q = 0;
    const int* q_counts = nullptr;
    const uint8_t* body_active = nullptr;   // optional [B]: 0 = skip body (winding 0)
    // opening radii and re-evaluation band; the defaults suit queries ON the surface (values near 0.5 /
    // 1.5).  Off-surface queries see near-integer winding numbers, i.e. interior points sit at 1.0, only
    // 0.01 above the threshold: they need a tighter far field and a narrower band (WC_*_POINTS).
    float beta_leaf = WC_BETA, beta_group = WC_BETA_GROUP, margin = WC_MARGIN;
    // launch_cluster_query on node records another job packed (the opening radius is baked into them): the radii
    // THAT job packed with (cluster_pack_betas); 0 = the records are this job's own
    float packed_beta_leaf = 0.f, packed_beta_group = 0.f;
    int max_top_leaves = 0;          // largest top group in leaves (shared-memory staging of the pack kernel)
    int* stats = nullptr;            // optional device int: receives the length of the exact re-evaluation list
    bool direct_pack = false;        // every node straight from its faces (cluster_pack_kernel): the test reference
};
// nearest_tiles.cu: masked nearest vertex over cluster-ordered 32-vertex tiles with bounding-sphere pruning
int launch_permute_mask(const uint32_t* maskT, int Vq, const int* vtile, int T, uint32_t* maskP, cudaStream_t st);
// tinfo: [B][T + NG][2] float4 -- tile spheres, then group spheres
int launch_group_mask(const uint32_t* maskP, const int* vgroup_off, int T, int NG, uint32_t* maskG, cudaStream_t st);
// tile_any [T][ceil(T / 32)]: bit (t & 31) of tile_any[qt][t >> 5] = some query of tile qt has an unmasked row in tile t
int launch_tile_any_mask(const uint32_t* maskP, int T, uint32_t* tile_any, cudaStream_t st);
int launch_nearest_tiles(const float* verts, const uint32_t* maskP, const uint32_t* maskG, const uint32_t* tile_any, const int* vtile,
                         const int* vgroup_off, int B, int V, int T, int NG, float4* vert4p, float4* tinfo,
                         int* argmin, float* minval, cudaStream_t st);

int launch_nearest_tiles_pack(const float* verts, const int* vtile, const int* vgroup_off, int B, int V, int T, int NG,
                              float4* vert4p, float4* tinfo, cudaStream_t st);
int launch_nearest_tiles_query(const uint32_t* maskP, const uint32_t* maskG, const uint32_t* tile_any, const int* vtile, const int* vgroup_off,
                               int b0, int nb, int V, int T, int NG, const float4* vert4p, const float4* tinfo,
                               float limit, const uint8_t* exterior, int* todo_list, int* argmin, float* minval, cudaStream_t st);

// the radius-limited query needs 2 T words of shared memory per warp of its single-query kernel (48 KB per CTA)
inline bool nearest_limited_supported(int T) { return (size_t)T * 2 * 4 * sizeof(float) <= 48 * 1024; }

int cluster_splits(int B, int T, int NT, int sm_count);
int launch_cluster_pack(const ClusterJob& job, cudaStream_t st);      // node records + packed leaf triangles
void cluster_pack_betas(const ClusterJob& job, float* beta_leaf, float* beta_group);   // the radii it bakes in
int launch_cluster_traverse(const ClusterJob& job, int b0, int nb, cudaStream_t st);   // the winding kernel alone
int launch_cluster_finalize(const ClusterJob& job, uint8_t* early_ext, cudaStream_t st);   // split sum + refine list
int launch_cluster_refine(const ClusterJob& job, cudaStream_t st);    // exact re-evaluation of the listed queries
int launch_cluster_finish(const ClusterJob& job, cudaStream_t st);    // finalize + exact refine
int launch_cluster_query(const ClusterJob& job, cudaStream_t st);     // winding kernel + finalize + exact refine
int launch_winding_clusters(const ClusterJob& job, cudaStream_t st);  // both

}  // namespace tuch
