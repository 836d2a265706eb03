// Internal: triangle-strip stream (strips.cu).
#pragma once
#include <vector>

#include "kernels.h"

namespace tuch {

constexpr int WS_THREADS = 128;   // threads per CTA
constexpr int WS_QPT = 2;         // queries per thread
constexpr int WS_TILE = 256;      // stream elements per shared-memory stage (4 KB)
constexpr int WS_STAGES = 4;

// host: cuts the faces into strips and lays them out as a tile-self-contained vertex stream
// (vid = vertex id or -1 for padding, flag bit 0 = closes a triangle, bit 31 = negate)
// fid = face closed by the element, or -1
int build_strip_stream(const int* faces, int F, int tile, std::vector<int>& vid, std::vector<uint32_t>& flag,
                       std::vector<int>& fid, int* n_strips_out);

struct StripJob {
    const float4* strip8;                          // [B][Lp][2]: (x, y, z, closes) | (face normal, 0)
    const float4* info;                            // [B][Lp / WS_TILE]: tile centre, near radius^2
    const float* points; long long point_stride;   // floats per body
    float* partial;                                // [B][S][Q] scratch
    float* winding; long long out_stride;
    const uint8_t* body_active;
    int B, Q, Lp, S;
    const int* q_counts = nullptr;
};
int strip_splits(int B, int Q, int Lp, int sm_count);
int launch_pack_strips(const float* verts, int B, int V, const int* faces, const int* vid, const int* fid, int Lp,
                       float4* strip8, float4* info, cudaStream_t st);
int launch_winding_strips(const StripJob& job, cudaStream_t st);

}  // namespace tuch
