/* tuch_b200 -- C ABI of the B200-native TUCH self-contact hot path.
 *
 * The reference (muelea/tuch) is pure Python/PyTorch and has no FFI; its de-facto boundary is
 * a set of Python call signatures.  Every entry point below names the reference
 * function (file:line under /root/reference) whose arithmetic it replaces.  The host-side
 * mirror of those signatures lives in tuch_b200/ (Python, ctypes); INTEGRATION.md shows the
 * binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; tuch_last_error() describes the
 *     most recent failure of the calling thread.  There is no CPU fallback anywhere.
 *   - pointers are DEVICE pointers unless the parameter name ends in _host or the function name
 *     ends in _host; all arrays are dense, row-major, fp32 / int32 / uint8.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is enqueued
 *     asynchronously, nothing synchronises unless documented.
 *   - scratch memory comes from a grow-only per-(device, stream) arena inside the library: the
 *     first call at a new problem size may cudaMalloc (and therefore must not happen inside a
 *     CUDA-graph capture); later calls at the same or smaller size never allocate.
 */
#ifndef TUCH_B200_H
#define TUCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TUCH_B200_ABI_VERSION 1

/* ------------------------------------------------------------------ status */
const char* tuch_last_error(void);
int tuch_abi_version(void);
/* sm_count / compute capability of the current device; fails if it is not sm_100. */
int tuch_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched in the calling process (bench.py: gpu_launches) */
long long tuch_launch_count(void);
/* releases every scratch arena of the current device (synchronises the device) */
int tuch_release_scratch(void);

/* ------------------------------------------------------------------ a1  contact.py:23-47
 * batch_pairwise_dist(x, y, squared): P[b,i,j] = (|x_i|^2 + |y_j|^2) - 2 x_i.y_j, sqrt if !squared.
 * x[bs,nx,3], y[bs,ny,3] -> P[bs,nx,ny].  Materialises the full matrix (API parity). */
int tuch_pairwise_dist(const float* x, const float* y, int bs, int nx, int ny, int squared,
                       float* P, void* stream);

/* autograd backward of a1 (the reference differentiates P through torch.bmm, losses.py:76,182):
 * gx[b,i,:] = sum_j w_ij (2 x_i - 2 y_j), gy[b,j,:] = sum_i w_ij (2 y_j - 2 x_i), with
 * w = gP (squared) or gP / (2 P) (sqrt form; P is the forward output).  gx or gy may be NULL. */
int tuch_pairwise_dist_backward(const float* x, const float* y, const float* P, const float* gP,
                                int bs, int nx, int ny, int squared, float* gx, float* gy, void* stream);

/* ------------------------------------------------------------------ a2  contact.py:49-109
 * solid_angles(points, triangles): points[bs,Q,3], triangles[bs,F,3,3] -> out[bs,Q,F]. */
int tuch_solid_angles(const float* points, const float* triangles, int bs, int Q, int F,
                      float* out, void* stream);

/* ------------------------------------------------------------------ a3  contact.py:112-147
 * winding_numbers(points, triangles): -> out[bs,Q]; the [Q,F] solid-angle matrix is never formed. */
int tuch_winding_numbers(const float* points, const float* triangles, int bs, int Q, int F,
                         float* out, void* stream);

/* ------------------------------------------------------------------ mesh topology (constants)
 * Device-resident constants of one mesh topology: faces (smplifydc.py:58-61), the geodesic mask
 * geomask = geodist > geothres (smplifydc.py:65, loss.py:71) bit-packed, DSC region pairs
 * (losses.py:110-113, train_module.py:65-67) and closed body segments (segmentation.py:29-66). */
typedef struct tuch_topology tuch_topology;

int tuch_topology_create(int V, int F, const int32_t* faces_host, tuch_topology** out);
void tuch_topology_destroy(tuch_topology* topo);
int tuch_topology_num_verts(const tuch_topology* topo);
int tuch_topology_num_faces(const tuch_topology* topo);

/* geomask[r][c] = geodist[r][c] > geothres, from a DEVICE float [V,V] matrix ... */
int tuch_topology_set_geodist(tuch_topology* topo, const float* geodist, float geothres, void* stream);
/* ... or directly from a DEVICE bool/uint8 [V,V] mask (non-zero = geodesically far) */
int tuch_topology_set_geomask(tuch_topology* topo, const uint8_t* geomask, void* stream);

/* cdict = {'csig': region -> vertex ids, 'classes': [(regionA, regionB)]}: CSR region lists and
 * pair table (HOST arrays, copied).  region_offsets[n_regions+1], pair_a/pair_b[n_pairs]. */
int tuch_topology_set_regions(tuch_topology* topo, int n_regions, const int32_t* region_offsets_host,
                              const int32_t* region_ids_host, int n_pairs, const int32_t* pair_a_host,
                              const int32_t* pair_b_host);

/* BatchBodySegment (segmentation.py:102-124): per segment its member vertices
 * (segment_vidx, :42), its closed face list (faces inside the segment + cap fans, :50-66; vertex
 * index V + k denotes the centroid of the segment's k-th band loop) and the band loops (:45-46).
 * All HOST CSR arrays, copied. */
int tuch_topology_set_segments(tuch_topology* topo, int n_segments,
                               const int32_t* vidx_offsets_host, const int32_t* vidx_host,
                               const int32_t* face_offsets_host, const int32_t* faces_host,
                               const int32_t* band_offsets_host, const int32_t* loop_offsets_host,
                               const int32_t* loop_ids_host);

/* ------------------------------------------------------------------ fused self-contact query
 * Replaces, for every body of the batch, losses.py:76-93 / loss.py:256-270:
 *   winding[b,v]  = winding_numbers(verts[b], verts[b][faces])           (contact.py:112)
 *   exterior[b,v] = winding <= 0.99, then set to 1 where v is inside its own closed segment
 *                   (losses.py:82-89; segment pass only if use_segments != 0; when
 *                   segments_only_if_interior != 0 a body without interior vertices skips it,
 *                   which is unobservable in the flags but mirrors losses.py:85)
 *   argmin[b,c]   = first r minimising P[r,c] over geomask[r,c]          (losses.py:92-93)
 *   min_sq[b,c]   = that minimum (expansion-form squared distance, +inf if fully masked)
 * verts[B,V,3]; outputs argmin int32 [B,V], min_sq/winding fp32 [B,V], exterior uint8 [B,V].
 * Any output pointer may be NULL. */
int tuch_contact_query(const tuch_topology* topo, const float* verts, int B, int use_segments,
                       int32_t* argmin, float* min_sq, float* winding, uint8_t* exterior, void* stream);

/* has_self_isect for every segment (segmentation.py:81-99,117-124): out[b][k] = 1 where the k-th
 * entry of the concatenated segment vertex lists is EXTERIOR to its closed segment.
 * out uint8 [B, total_segment_verts]. */
int tuch_segment_exterior(const tuch_topology* topo, const float* verts, int B, uint8_t* out,
                          float* winding_out, void* stream);
int tuch_topology_total_segment_verts(const tuch_topology* topo);

/* region-pair minima (losses.py:108-117 with masked != 0; train_module.py:69-91 with masked == 0):
 * min_sq[b,p] = min over csig[a] x csig[b] of P (masked entries = +inf), plus the attaining
 * vertex pair.  active uint8 [B,n_pairs] or NULL (= all pairs). */
int tuch_region_min(const tuch_topology* topo, const float* verts, int B, int masked,
                    const uint8_t* active, float* min_sq, int32_t* arg_i, int32_t* arg_j, void* stream);

/* ------------------------------------------------------------------ a11  SMPL body model
 * tuch/models/smpl.py:34-56 over smplx==0.1.13 `SMPL.forward` / `lbs.lbs` (third-party, not in
 * the reference tree): shape blend, joint regression, Rodrigues, pose blend, 24-joint kinematic
 * chain, skinning, 21 vertex-picked joints, n_extra_reg regressed joints, remap to n_out joints.
 * All model arrays are HOST pointers, copied at creation:
 *   v_template[V,3] shapedirs[V,3,L] posedirs[207,3V] J_regressor[24,V] lbs_weights[V,24]
 *   parents[24] extra_vertex_ids[n_extra_verts] J_regressor_extra[n_extra_reg,V] joint_map[n_out]
 * (joint_map indexes the concatenation [24 posed | picked vertices | regressed extras]). */
typedef struct tuch_smpl tuch_smpl;

int tuch_smpl_create(int V, int num_betas, const float* v_template_host, const float* shapedirs_host,
                     const float* posedirs_host, const float* J_regressor_host,
                     const float* lbs_weights_host, const int32_t* parents_host, int n_extra_verts,
                     const int32_t* extra_vertex_ids_host, int n_extra_reg,
                     const float* J_regressor_extra_host, int n_out, const int32_t* joint_map_host,
                     tuch_smpl** out);
void tuch_smpl_destroy(tuch_smpl* smpl);
int tuch_smpl_num_verts(const tuch_smpl* smpl);
int tuch_smpl_num_joints(const tuch_smpl* smpl);
/* floats of caller-owned device workspace a forward/backward pair at batch B needs; forward fills
 * it with the intermediates backward reads (rotations, chain transforms, posed template). */
size_t tuch_smpl_workspace_floats(const tuch_smpl* smpl, int B);

/* betas[B,L]; pose = axis-angle [B,72] (pose_is_rotmat == 0; smplx batch_rodrigues form) or
 * rotation matrices [B,24,3,3] (pose_is_rotmat != 0, the `pose2rot=False` call of train_module.py:202).
 * -> vertices[B,V,3], joints[B,n_out,3] (joints may be NULL). */
int tuch_smpl_forward(const tuch_smpl* smpl, const float* betas, const float* pose, int pose_is_rotmat,
                      int B, float* workspace, float* vertices, float* joints, void* stream);
/* vector-Jacobian product of tuch_smpl_forward: g_vertices[B,V,3] and/or g_joints[B,n_out,3]
 * (either may be NULL) -> g_pose ([B,72] or [B,24,3,3]) and g_betas[B,L] (either may be NULL). */
int tuch_smpl_backward(const tuch_smpl* smpl, const float* pose, int pose_is_rotmat, int B,
                       float* workspace, const float* g_vertices, const float* g_joints,
                       float* g_pose, float* g_betas, void* stream);

/* ------------------------------------------------------------------ host-buffer conveniences
 * Same as the calls above with HOST buffers; copies in/out on an internal stream and
 * synchronises.  Used by non-torch callers and by the end-to-end benchmark leg. */
int tuch_winding_numbers_host(const float* points_host, const float* triangles_host, int bs, int Q,
                              int F, float* out_host);
int tuch_contact_query_host(const tuch_topology* topo, const float* verts_host, int B, int use_segments,
                            int32_t* argmin_host, float* min_sq_host, float* winding_host,
                            uint8_t* exterior_host);

#ifdef __cplusplus
}
#endif
#endif /* TUCH_B200_H */
