// Allowed-self-intersection filter: closed body segments (tuch/utils/segmentation.py:29-124) and
// the whitelist write-back of tuch/smplify/losses.py:85-89 / tuch/train/loss.py:264-266, all on
// the device (the reference copies one flag vector per segment per body to the host).
#include "api_internal.h"

namespace tuch {

// apex[b][j] = mean of the j-th band loop's vertices as listed (segmentation.py:73-75)
__global__ void segment_apex_kernel(const float* __restrict__ verts, int V, const int* __restrict__ loop_off,
                                    const int* __restrict__ loop_ids, int n_bands,
                                    float* __restrict__ apex, const uint8_t* __restrict__ body_active) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_bands) return;
    if (body_active != nullptr && !body_active[b]) return;
    const float* vb = verts + (size_t)b * V * 3;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    const int l0 = loop_off[j], l1 = loop_off[j + 1];
    for (int k = l0; k < l1; ++k) {
        const int v = loop_ids[k];
        sx += vb[3 * v]; sy += vb[3 * v + 1]; sz += vb[3 * v + 2];
    }
    const float n = (float)(l1 - l0);
    float* o = apex + ((size_t)b * n_bands + j) * 3;
    o[0] = sx / n; o[1] = sy / n; o[2] = sz / n;
}

// Packs every segment's closed triangle list into the tri12 layout (one padded range per segment)
// and gathers the segment member vertices as query points.
//   slot_face[i]  index into seg_faces for packed triangle slot i, or -1 (padding)
//   slot_band0[i] first band index of the slot's segment (apex lookup for vertex ids >= V)
__global__ void segment_pack_kernel(const float* __restrict__ verts, int V, const float* __restrict__ apex,
                                    int n_bands, const int* __restrict__ seg_faces,
                                    const int* __restrict__ slot_face, const int* __restrict__ slot_band0,
                                    int n_slots, const int* __restrict__ seg_vidx, int n_sv,
                                    float4* __restrict__ tri12, float* __restrict__ points,
                                    const uint8_t* __restrict__ body_active) {
    const int b = blockIdx.y;
    if (body_active != nullptr && !body_active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* vb = verts + (size_t)b * V * 3;
    if (i < n_slots) {
        float4 c[3] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
        const int f = slot_face[i];
        if (f >= 0) {
            const float* ab = apex + ((size_t)b * n_bands + slot_band0[i]) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int v = seg_faces[3 * f + k];
                const float* p = (v < V) ? (vb + 3 * v) : (ab + 3 * (v - V));
                c[k] = make_float4(p[0], p[1], p[2], 0.f);
            }
        }
        float4* t = tri12 + ((size_t)b * n_slots + i) * 3;
        t[0] = c[0]; t[1] = c[1]; t[2] = c[2];
    }
    if (i < n_sv) {
        const int v = seg_vidx[i];
        float* o = points + ((size_t)b * n_sv + i) * 3;
        o[0] = vb[3 * v]; o[1] = vb[3 * v + 1]; o[2] = vb[3 * v + 2];
    }
}

// exterior = winding <= 0.99 (losses.py:82) and per-body "has interior vertex" flag (losses.py:85)
__global__ void exterior_init_kernel(const float* __restrict__ winding, int V, uint8_t* __restrict__ exterior,
                                     uint8_t* __restrict__ any_interior) {
    const int b = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool interior = false;
    if (v < V) {
        const bool ext = winding[(size_t)b * V + v] <= 0.99f;
        exterior[(size_t)b * V + v] = ext ? 1 : 0;
        interior = !ext;
    }
    if (any_interior != nullptr && __syncthreads_or(interior) && threadIdx.x == 0) any_interior[b] = 1;
}

// exterior[seg_vidx[k]] = 1 where the vertex is NOT exterior to its own closed segment
// (losses.py:88-89); optionally also emits has_self_isect's flags (segmentation.py:97).
__global__ void segment_apply_kernel(const float* __restrict__ seg_winding, const int* __restrict__ seg_vidx,
                                     int n_sv, int V, uint8_t* __restrict__ exterior,
                                     uint8_t* __restrict__ seg_ext_out, const uint8_t* __restrict__ body_active) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sv) return;
    if (body_active != nullptr && !body_active[b]) return;
    const bool seg_ext = seg_winding[(size_t)b * n_sv + k] <= 0.99f;
    if (seg_ext_out != nullptr) seg_ext_out[(size_t)b * n_sv + k] = seg_ext ? 1 : 0;
    if (exterior != nullptr && !seg_ext) exterior[(size_t)b * V + seg_vidx[k]] = 1;
}

// Whitelist pass of the fused contact query: only INTERIOR member vertices can change (the write-back sets
// exterior = 1).  They are a few per cent of the (body, member slot) pairs, so they are compacted into a list first
// (one warp per pair that exits unless its vertex is interior kept the SMs full of warps waiting for that one flag:
// 173 us at 256 bodies for 52 us worth of instructions) and a fixed grid of warps then walks the list, each entry
// summing the solid angles of its segment's closed face list with one lane per face.  The order of the list does
// not matter: every entry only ever sets its own vertex's flag.
__global__ void __launch_bounds__(256)
segment_compact_kernel(const uint8_t* __restrict__ exterior, int V, const int* __restrict__ seg_vidx, int n_sv,
                       const uint8_t* __restrict__ body_active, int* __restrict__ list) {
    const int b = blockIdx.y;
    if (body_active != nullptr && !body_active[b]) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool interior = k < n_sv && exterior[(size_t)b * V + seg_vidx[k]] == 0;
    const unsigned vote = __ballot_sync(0xffffffffu, interior);
    if (vote == 0u) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(list, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (interior) list[1 + base + __popc(vote & ((1u << lane) - 1u))] = b * n_sv + k;
}

__global__ void __launch_bounds__(256)
segment_whitelist_kernel(const float* __restrict__ verts, int V, const float* __restrict__ apex, int n_bands,
                         const int* __restrict__ seg_faces, const int* __restrict__ seg_face_off,
                         const int* __restrict__ seg_band0, const int* __restrict__ seg_vidx,
                         const int* __restrict__ member_seg, int n_sv, uint8_t* __restrict__ exterior,
                         const int* __restrict__ list) {
    const int lane = threadIdx.x & 31;
    const int n = list[0];
    for (int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += gridDim.x * 8) {
        const int bk = list[1 + e];
        const int b = bk / n_sv, k = bk - b * n_sv;
        const int v = seg_vidx[k];
        const int s = member_seg[k];
        const float* vb = verts + (size_t)b * V * 3;
        const float* ab = apex + ((size_t)b * n_bands + seg_band0[s]) * 3;
        const float px = vb[3 * v], py = vb[3 * v + 1], pz = vb[3 * v + 2];
        float acc = 0.f;
        // Two dependent loads per face (corner index -> corner) over ~100 trips: the loop is pure load latency, so
        // four trips are software-pipelined by hand -- all twelve indices first, then all thirty-six coordinates,
        // then the arithmetic.  The adds into `acc` stay in face order.
        const int f_end = seg_face_off[s + 1];
        for (int f0 = seg_face_off[s] + lane; f0 < f_end; f0 += 4 * 32) {
            int idx[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int f = f0 + 32 * u;
#pragma unroll
                for (int c = 0; c < 3; ++c) idx[u][c] = f < f_end ? __ldg(seg_faces + 3 * f + c) : 0;
            }
            float c[4][3][3];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int i = idx[u][q];
                    const float* p = (i < V) ? (vb + 3 * i) : (ab + 3 * (i - V));
                    c[u][q][0] = p[0]; c[u][q][1] = p[1]; c[u][q][2] = p[2];
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (f0 + 32 * u < f_end)
                    acc += half_solid_angle(px, py, pz, make_float4(c[u][0][0], c[u][0][1], c[u][0][2], 0.f),
                                            make_float4(c[u][1][0], c[u][1][1], c[u][1][2], 0.f),
                                            make_float4(c[u][2][0], c[u][2][1], c[u][2][2], 0.f));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && !(acc * 0.159154943091895336f <= 0.99f)) exterior[(size_t)b * V + v] = 1;   // inside its own segment
    }
}

// list: [1 + B * n_sv] ints of scratch
int launch_segment_whitelist(const float* verts, int B, int V, const float* apex, int n_bands, const int* seg_faces,
                             const int* seg_face_off, const int* seg_band0, const int* seg_vidx,
                             const int* member_seg, int n_sv, uint8_t* exterior, const uint8_t* body_active,
                             int* list, cudaStream_t st) {
    if (n_sv == 0 || B == 0) return 0;
    TUCH_REQUIRE((long long)B * n_sv < (1LL << 31), "segment whitelist: %d bodies x %d member vertices overflow the list index", B, n_sv);
    KernelTimer timer("winding_kernel_segments", st);
    TUCH_CUDA(cudaMemsetAsync(list, 0, sizeof(int), st));
    dim3 grid(cdiv(n_sv, 256), B);
    segment_compact_kernel<<<grid, 256, 0, st>>>(exterior, V, seg_vidx, n_sv, body_active, list);
    TUCH_LAUNCH_CHECK(); count_launch();
    segment_whitelist_kernel<<<sm_count() * 4, 256, 0, st>>>(verts, V, apex, n_bands, seg_faces, seg_face_off, seg_band0,
                                                             seg_vidx, member_seg, n_sv, exterior, list);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_segment_apex(const float* verts, int B, int V, const int* loop_off, const int* loop_ids,
                        int n_bands, float* apex, const uint8_t* body_active, cudaStream_t st) {
    if (n_bands == 0 || B == 0) return 0;
    dim3 grid(cdiv(n_bands, 64), B);
    segment_apex_kernel<<<grid, 64, 0, st>>>(verts, V, loop_off, loop_ids, n_bands, apex, body_active);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_segment_pack(const float* verts, int B, int V, const float* apex, int n_bands, const int* seg_faces,
                        const int* slot_face, const int* slot_band0, int n_slots, const int* seg_vidx,
                        int n_sv, float4* tri12, float* points, const uint8_t* body_active, cudaStream_t st) {
    const int n = max(n_slots, n_sv);
    if (n == 0 || B == 0) return 0;
    dim3 grid(cdiv(n, 256), B);
    segment_pack_kernel<<<grid, 256, 0, st>>>(verts, V, apex, n_bands, seg_faces, slot_face, slot_band0, n_slots,
                                              seg_vidx, n_sv, tri12, points, body_active);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_exterior_init(const float* winding, int B, int V, uint8_t* exterior, uint8_t* any_interior,
                         cudaStream_t st) {
    if (B == 0 || V == 0) return 0;
    dim3 grid(cdiv(V, 256), B);
    exterior_init_kernel<<<grid, 256, 0, st>>>(winding, V, exterior, any_interior);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_segment_apply(const float* seg_winding, const int* seg_vidx, int n_sv, int B, int V,
                         uint8_t* exterior, uint8_t* seg_ext_out, const uint8_t* body_active, cudaStream_t st) {
    if (n_sv == 0 || B == 0) return 0;
    dim3 grid(cdiv(n_sv, 256), B);
    segment_apply_kernel<<<grid, 256, 0, st>>>(seg_winding, seg_vidx, n_sv, V, exterior, seg_ext_out, body_active);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

}  // namespace tuch
