// Internal: launchers of the HD-point contact path (hd_kernels.cu).
#pragma once
#include "api_internal.h"

namespace tuch {

int launch_hd_select(const float* min_sq, const uint8_t* exterior, const uint8_t* body_active, const int* faces,
                     int B, int V, int N, const int* hd_face, float thres_sq, int* idx, int* counts, cudaStream_t st);
int launch_hd_gather(const float* verts, int B, int V, int N, const int* idx, const int* counts, const int* row_off,
                     const int* cols, const float* vals, const int* hd_face, const int* faces, float4* hd4,
                     float* hd, float* off, int* proxy, cudaStream_t st);
int launch_hd_nearest(const float4* hd4, const int* proxy, const int* counts, int B, int N, const uint32_t* maskT,
                      int Vq, int* argmin, cudaStream_t st);
int launch_hd_scatter(const float* g_hd, int B, int V, int N, const int* idx, const int* counts, const int* row_off,
                      const int* cols, const float* vals, float* g_verts, cudaStream_t st);

}  // namespace tuch
