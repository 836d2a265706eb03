// Internal launcher declarations shared by the .cu translation units.
#pragma once
#include "common.cuh"

namespace tuch {

// winding-number kernel tiling
constexpr int WN_THREADS = 128;   // threads per CTA
constexpr int WN_QPT = 2;         // queries per thread
constexpr int WN_TILE_F = 128;    // triangles per shared-memory stage (6 KB)
constexpr int WN_STAGES = 4;
// masked nearest-vertex kernel tiling
constexpr int NN_THREADS = 256;
constexpr int NN_TILE_V = 1024;   // candidate vertices per stage (16 KB)
constexpr int NN_STAGES = 2;
constexpr int RM_THREADS = 256;

inline int padded_faces(int F) { return (int)align_up((size_t)(F > 0 ? F : 1), WN_TILE_F); }
inline int padded_verts(int V) { return (int)align_up((size_t)(V > 0 ? V : 1), 32); }

int winding_splits(int B, int Q, int Fp, int sm_count);
int launch_pack_mesh(const float* verts, const int* faces, int B, int V, int F, int Fp, int Vp,
                     float4* tri12, float4* vert4, cudaStream_t st);
int launch_pack_triangles(const float* tris, int B, int F, int Fp, float4* tri12, cudaStream_t st);
// one batched winding-number problem: B bodies, each Q queries against Fp (padded) triangles
struct WindingJob {
    const float4* tri12; long long tri_stride;     // float4 units per body
    const float* points; long long point_stride;   // floats per body
    float* partial;                                // [B][S][Q] scratch
    float* winding; long long out_stride;          // floats per body
    const uint8_t* body_active;                    // optional [B]: 0 = skip body (output 0)
    int B, Q, Fp, S;
    const int* q_counts = nullptr;                 // optional [B]: only the first q_counts[b] queries are valid
};
int launch_winding(const WindingJob& job, cudaStream_t st);
int launch_winding_finalize(const float* partial, int B, int Q, int S, long long out_stride, float* winding,
                            const uint8_t* body_active, const int* q_counts, cudaStream_t st);
int launch_nearest(const float4* vert4, const uint32_t* maskT, int B, int V, int Vp, int Vq,
                   int* argmin, float* minval, cudaStream_t st);
int launch_pack_mask(const uint8_t* mask, const float* dist, float thres, int V, int Vq, int W,
                     uint32_t* maskT, cudaStream_t st);
int launch_pairwise_dist(const float* x, const float* y, int bs, int nx, int ny, int squared, float* P,
                         cudaStream_t st);
int launch_pairwise_dist_bwd(const float* x, const float* y, const float* P, const float* gP, int bs, int nx,
                             int ny, int squared, float* gx, float* gy, cudaStream_t st);
int launch_solid_angles(const float* points, const float* tris, int bs, int Q, int F, float* out,
                        cudaStream_t st);
int launch_pair_mask(const uint32_t* maskT, int Vq, const int* region_ids, const int* region_off, const int* pair_a,
                     const int* pair_b, const long long* pair_word_off, int n_pairs, long long max_words_per_pair,
                     uint32_t* pmask, cudaStream_t st);
// pmask / pair_word_off: optional per-pair bit-packed geodesic sub-masks (launch_pair_mask), NULL = look the
// mask up entry by entry
int launch_region_min(const float4* vert4, int Vp, const uint32_t* maskT, int Vq, const int* region_ids,
                      const int* region_off, const int* pair_a, const int* pair_b, const uint8_t* active,
                      int n_pairs, int B, const uint32_t* pmask, const long long* pair_word_off, float* min_out,
                      int* arg_i, int* arg_j, cudaStream_t st);

int sm_count();
void count_launch();

// Optional per-kernel device timing (tuch_kernel_timing_*): when enabled, a KernelTimer brackets a
// launch with two CUDA events on the launch stream; the pairs are resolved when the totals are read.
// Nothing is recorded while the stream is being captured into a CUDA graph.
class KernelTimer {
public:
    KernelTimer(const char* name, cudaStream_t st);
    ~KernelTimer();
private:
    int slot_ = -1;
    cudaStream_t st_;
};

}  // namespace tuch
