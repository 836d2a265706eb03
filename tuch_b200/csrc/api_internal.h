// Internal: opaque handle layouts and the scratch bump allocator behind include/tuch_b200.h.
#pragma once
#include <atomic>
#include <vector>

#include "../../include/tuch_b200.h"
#include "kernels.h"

struct tuch_topology {
    int device = 0;
    int V = 0, F = 0, Fp = 0, Vp = 0, Vq = 0, W = 0;
    int* d_faces = nullptr;            // [F][3]
    // triangle-strip stream of the faces (strips.cu): vertex id / flag per element, Lp elements
    int Lp = 0, n_strips = 0;
    int *d_strip_vid = nullptr, *d_strip_fid = nullptr;
    // face-cluster hierarchy for the far-field winding kernel (clusters.cu); built from the template
    // (tuch_topology_set_template) or lazily from the first body a contact query sees
    int K = 0, NM = 0, NT = 0, T = 0, NG = 0, max_top_leaves = 0;
    int *d_leaf_face = nullptr, *d_mid_off = nullptr, *d_top_off = nullptr, *d_vtile = nullptr, *d_vgroup_off = nullptr;
    std::atomic<bool> has_clusters{false};   // published last (release) by install_clusters; the lazy build of the
                                             // first query may run while other threads read it (acquire)
    uint32_t* d_maskG = nullptr;       // per-group summary of d_maskP (any unmasked row in the tile group)
    uint32_t* d_tile_any = nullptr;    // per (query tile, candidate tile) summary: any unmasked pair at all
    uint32_t* d_maskP = nullptr;       // geodesic mask in cluster order (nearest_tiles.cu); valid when both
    bool has_maskP = false;            // the mask and the hierarchy exist
    int winding_mode = 1;              // TUCH_WINDING_FAST
    int* d_stats = nullptr;            // [2] length of the last exact re-evaluation list: vertex query, point query
    uint32_t* d_maskT = nullptr;       // [W][Vq] bit-packed geodesic mask
    bool has_mask = false;
    // DSC regions (CSR) and annotated pairs
    int n_regions = 0, n_pairs = 0;
    int *d_region_off = nullptr, *d_region_ids = nullptr, *d_pair_a = nullptr, *d_pair_b = nullptr;
    // per-pair bit-packed sub-masks of the geodesic mask (valid when regions and mask both exist)
    uint32_t* d_pair_mask = nullptr;
    long long* d_pair_word_off = nullptr;
    std::vector<long long> h_pair_word_off;   // [n_pairs + 1]
    long long max_pair_words = 0;
    bool has_pair_mask = false;
    // closed body segments
    int n_segments = 0, n_bands = 0, n_sv = 0, n_slots = 0;
    std::vector<int> h_vidx_off;       // [S+1] member-vertex ranges
    std::vector<int> h_slot_off;       // [S+1] packed (padded) triangle ranges
    int *d_seg_vidx = nullptr, *d_seg_faces = nullptr, *d_slot_face = nullptr, *d_slot_band0 = nullptr;
    int *d_loop_off = nullptr, *d_loop_ids = nullptr;
    int *d_member_seg = nullptr, *d_seg_face_off = nullptr, *d_seg_band0 = nullptr;   // whitelist pass
    // HD-point regressor (CSR rows) and the source face of every HD point (loss.py:81-89)
    int n_hd = 0;
    int *d_hd_row_off = nullptr, *d_hd_cols = nullptr, *d_hd_face = nullptr;
    float* d_hd_vals = nullptr;
};

namespace tuch {

// slot 0 is the arena of the leaf entry points; slot 1 belongs to entry points that call them
int arena_get(cudaStream_t st, size_t bytes, void** out, int slot = 0);

// Two-phase bump allocator over the per-(device, stream) arena: plan() every buffer, commit()
// once (may grow the arena), then get<T>().
class Scratch {
public:
    size_t plan(size_t bytes) {
        const size_t off = total_;
        total_ += align_up(bytes, 256);
        return off;
    }
    int commit(cudaStream_t st) { return commit_slot(st, 0); }
    int commit_slot(cudaStream_t st, int slot);
    template <typename T> T* get(size_t handle) const { return (T*)(base_ + handle); }
private:
    size_t total_ = 0;
    char* base_ = nullptr;
};

int launch_segment_apex(const float* verts, int B, int V, const int* loop_off, const int* loop_ids,
                        int n_bands, float* apex, const uint8_t* body_active, cudaStream_t st);
int launch_segment_pack(const float* verts, int B, int V, const float* apex, int n_bands, const int* seg_faces,
                        const int* slot_face, const int* slot_band0, int n_slots, const int* seg_vidx,
                        int n_sv, float4* tri12, float* points, const uint8_t* body_active, cudaStream_t st);
int launch_segment_whitelist(const float* verts, int B, int V, const float* apex, int n_bands, const int* seg_faces,
                             const int* seg_face_off, const int* seg_band0, const int* seg_vidx,
                             const int* member_seg, int n_sv, uint8_t* exterior, const uint8_t* body_active,
                             int* list, cudaStream_t st);
int launch_exterior_init(const float* winding, int B, int V, uint8_t* exterior, uint8_t* any_interior,
                         cudaStream_t st);
int launch_segment_apply(const float* seg_winding, const int* seg_vidx, int n_sv, int B, int V,
                         uint8_t* exterior, uint8_t* seg_ext_out, const uint8_t* body_active, cudaStream_t st);

// the face hierarchy of this batch as packed by the hierarchical winding path (leaf triangles + node records);
// lives in the stream's shared scratch arena: valid until the next call that commits that arena
struct PackedClusters {
    const float4* ctri = nullptr;
    const float4* nodes = nullptr;
    float beta_leaf = 0.f, beta_group = 0.f;      // opening radii baked into the node records
};

// vert4_out: optional caller-owned [B][Vp] float4 buffer that receives the packed vertices
// packed_out: optional, receives the packed hierarchy when the hierarchical winding path ran (else stays null)
// Optional extra streams of one contact query (the fused iteration hands in library-owned streams of different
// priorities).  nn: the masked-nearest-vertex half, independent of the inside test; the caller orders it after the
// vertices and joins it before it reads argmin / min_sq.  trav: the hierarchical winding kernel alone -- the pack
// before it and the finalize / exact re-evaluation / segment pass after it stay on `st`; the two events order them.
struct QueryStreams {
    cudaStream_t nn = nullptr, trav = nullptr;
    bool split_trav = false;
    cudaEvent_t before_trav = nullptr, after_trav = nullptr;
    // nn_limit >= 0 (needs the exterior output): the nearest vertex is searched AFTER the inside test, without a
    // limit for the interior vertices and among the candidates within nn_limit metres for the exterior ones.
    // argmin / min_sq are then exact for every interior vertex and for every vertex with an allowed vertex within
    // the limit, and (-1, +inf) elsewhere: all that losses.py:96-103 consumes.
    float nn_limit = -1.f;
    cudaEvent_t after_ext = nullptr;          // orders that search (on nn) after the flags (on st) when nn != st
};
int contact_query_impl(const tuch_topology* t, const float* verts, int B, int use_segments, int32_t* argmin,
                       float* min_sq, float* winding, uint8_t* exterior, float4* vert4_out, cudaStream_t st,
                       PackedClusters* packed_out = nullptr, const QueryStreams* qs = nullptr);

}  // namespace tuch
