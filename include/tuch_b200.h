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
/* Optional per-kernel device timing for benchmarks: while enabled, the dominant kernels
 * ("winding_kernel", "winding_kernel_segments", "nearest_kernel") are bracketed by CUDA events on
 * their launch stream.  tuch_kernel_timing_read synchronises on the recorded events and returns the
 * accumulated device time and launch count since the last reset. */
int tuch_kernel_timing_enable(int on);
/* comma-separated names of everything timed since the library was loaded (kernels or launch groups) */
int tuch_kernel_timing_names(char* buf, int capacity);
int tuch_kernel_timing_reset(void);
int tuch_kernel_timing_read(const char* name, double* total_ms, long long* launches);
/* Scratch arenas: one grow-only block per (device, stream, purpose) inside the library, so that calls do not
 * allocate after the first one at a size and can be captured into CUDA graphs.  A block that a capture has seen
 * is never freed by growth (it is retired and kept for the graphs that point at it).  tuch_release_scratch()
 * frees every block of the current device (synchronises the device) and INVALIDATES every graph captured over
 * library calls before it: it increments tuch_scratch_generation(), which holders of captured graphs compare
 * with the value they saw at capture time to know that they must capture again. */
int tuch_release_scratch(void);
long long tuch_scratch_generation(void);

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
/* The faces are cut once, on the host, into triangle strips laid out as a vertex stream in
 * self-contained tiles of 256 elements (each element closes one triangle with its two predecessors;
 * flag bit 0 = closes a face, bit 31 = corner order is an odd permutation of the face): the winding
 * kernel then pays one new corner per triangle instead of three.  stats: stream length (elements per
 * body, >= F) and number of strips.  tuch_strip_stream_host runs the same builder without a device
 * (vid_out / flag_out may be NULL to query the length only). */
int tuch_topology_strip_stats(const tuch_topology* topo, int* stream_len, int* n_strips);
int tuch_strip_stream_host(const int32_t* faces_host, int F, int32_t* vid_out, uint32_t* flag_out, int capacity,
                           int* stream_len, int* n_strips);
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

/* Winding-number evaluation inside tuch_contact_query.  Callers of the reference only consume the
 * flag `winding_numbers(...) <= 0.99` (losses.py:82, loss.py:262), so the default mode evaluates the
 * sum hierarchically: the faces are clustered once per topology (leaves of <= 16 faces, mid groups of
 * <= 8 leaves, top groups of <= 64 leaves); per body, nodes farther than 1.4 (leaves) / 2 (groups) radii
 * from a query contribute through a
 * second-order multipole expansion of the solid-angle integrand (max abs error measured 1.05e-2 on the winding
 * number), nearer leaves are summed exactly, and every query whose value falls within 0.10 of the 0.99
 * threshold is re-evaluated exactly over all faces -- the exterior flags are those of the exact sum.
 * TUCH_WINDING_EXACT sums all F solid angles for every query (values within 2e-5 of the reference).
 * The hierarchy is built from the template given to tuch_topology_set_template (HOST [V,3]) or, when
 * none was given, from the first body a query sees (one blocking device->host copy, outside graph
 * capture only; inside a capture the exact kernels run instead).  Give the template whenever the body model is
 * at hand (SMPL.v_template): the lazily built hierarchy depends on which body arrives first, and with it the
 * far-field rounding of the returned `winding` values (never the flags: every value within the band around the
 * threshold is re-evaluated over all faces).  In TUCH_WINDING_FAST mode `winding` is the hierarchical value:
 * within 2.5e-2 of the all-faces sum away from the threshold band, exact (2e-5) inside it. */
#define TUCH_WINDING_EXACT 0
#define TUCH_WINDING_FAST 1
int tuch_topology_set_template(tuch_topology* topo, const float* verts_host);
int tuch_topology_set_winding_mode(tuch_topology* topo, int mode);
int tuch_topology_cluster_stats(const tuch_topology* topo, int* n_leaves, int* n_mids, int* n_tops, int* n_tiles,
                                int* leaf_faces);
/* how many queries the LAST hierarchical call on this topology re-evaluated exactly (values within the band
 * around the 0.99 threshold of losses.py:82): mesh-vertex queries (tuch_contact_query and everything built on
 * it) and HD-point queries (tuch_regressor_contact_loss).  Synchronises `stream`. */
int tuch_topology_query_stats(const tuch_topology* topo, int* refine_vertices, int* refine_points, void* stream);
/* Introspection for the tests: the per-body node records (tops, mids, leaves: 28 floats each = centre, squared
 * opening radius, scaled moments) the hierarchical winding path packs for `verts` [B,V,3] (device) into nodes_out
 * [B, n_tops + n_mids + n_leaves, 28] (device).  direct != 0: every node straight from its own faces (the slow
 * reference); 0: the shipped kernel (group nodes from their children's shifted moments). */
int tuch_topology_pack_nodes(const tuch_topology* topo, const float* verts, int B, int direct, float* nodes_out,
                             void* stream);
/* the hierarchy builder without a device: leaf_face_out[n_leaves][16] (face id or -1),
 * mid_off_out[n_mids + 1] (leaf ranges), top_off_out[n_tops + 1] (mid ranges), vtile_out[n_tiles][32]
 * (vertex tiles: the 32 neighbouring vertices one warp queries / one word of the cluster-ordered geodesic
 * mask covers; -1 = padding); output pointers may be NULL to query the counts only. */
int tuch_cluster_tree_host(const int32_t* faces_host, int F, int V, const float* verts_host,
                           int32_t* leaf_face_out, int leaf_capacity, int32_t* mid_off_out, int mid_capacity,
                           int32_t* top_off_out, int top_capacity, int32_t* vtile_out, int tile_capacity,
                           int* n_leaves, int* n_mids, int* n_tops, int* n_tiles);

/* ------------------------------------------------------------------ fused self-contact query
 * Replaces, for every body of the batch, losses.py:76-93 / loss.py:256-270:
 *   winding[b,v]  = winding_numbers(verts[b], verts[b][faces])           (contact.py:112; see the
 *                   winding mode above: approximate away from the threshold in TUCH_WINDING_FAST)
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

/* The same query with the nearest vertex restricted to what SMPLify-DC's contact term consumes (losses.py:96-103:
 * the nearest allowed vertex of every INTERIOR vertex, and of an exterior vertex only when it is closer than
 * euclthres): argmin / min_sq are exactly tuch_contact_query's for every interior vertex (before the segment
 * whitelist) and for every vertex with an allowed vertex within `radius` metres, and (-1, +inf) -- or, on meshes of
 * more than 49,152 vertices, still tuch_contact_query's answer -- for the others
 * ((0, +inf), as above, for a vertex whose mask column is empty).  tuch_contact_loss reads -1 as "infinitely far".
 * About a fifth of the unlimited query's work at radius = 0.02.  exterior and one of argmin / min_sq are required. */
int tuch_contact_query_within(const tuch_topology* topo, const float* verts, int B, int use_segments, float radius,
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

/* ------------------------------------------------------------------ a7/a8  reprojection + GMoF
 * tuch/utils/geometry.py:83-111 (perspective_projection, identity rotation as at losses.py:56-59)
 * + tuch/smplify/losses.py:25-32,60-61: loss[b,j] = conf^2 * sum_xy gmof(f (p/p_z)_xy + c_xy - target_xy),
 * p = joints[b,j] + cam_t[b].  Forward value and analytic gradient in one launch:
 *   g_joints[B,J,3], g_cam_t[B,3] = d(sum_bj g_loss[b,j] loss[b,j])/d(.)   (g_loss NULL = ones).
 * cam_t_est != NULL adds camera_fitting_loss's depth term (losses.py:146):
 *   depth_loss[b] = depth_loss_weight^2 (cam_t_z - cam_t_est_z)^2, its gradient goes into g_cam_t.
 * Any output may be NULL. */
int tuch_reprojection_loss(const float* joints, const float* cam_t, const float* center,
                           const float* joints_2d, const float* conf, int B, int J, float focal_length,
                           float sigma, const float* cam_t_est, float depth_loss_weight, const float* g_loss,
                           float* loss, float* depth_loss, float* g_joints, float* g_cam_t, void* stream);

/* ------------------------------------------------------------------ a9  MaxMixturePrior
 * tuch/smplify/prior.py:36-132 (use_merged=True): HOST arrays means[M,D], precisions[M,D,D],
 * nll_weights[M] (prior.py:80-96), copied. */
typedef struct tuch_prior tuch_prior;
int tuch_prior_create(int M, int D, const float* means_host, const float* precisions_host,
                      const float* nll_weights_host, tuch_prior** out);
void tuch_prior_destroy(tuch_prior* prior);

/* value[b] = ppw^2 * min_m[0.5 (th-mu_m)^T P_m (th-mu_m) - log nll_w_m]      prior.py:117-132
 *          + apw^2 * sum exp(+-th[52,55,9,12])^2                              losses.py:155-162
 *          + spw^2 * |betas[b]|^2                                             losses.py:149,192
 * pose[B,D] (body pose, D = 69), betas[B,L] (may be NULL when spw == 0).  Outputs (each may be NULL):
 * value[B], prior_value[B] (unweighted mixture term), component[B] (arg min), g_pose[B,D], g_betas[B,L]. */
int tuch_pose_terms(const tuch_prior* prior, const float* pose, const float* betas, int B, int D, int L,
                    float pose_prior_weight, float angle_prior_weight, float shape_prior_weight,
                    float* value, float* prior_value, int32_t* component, float* g_pose, float* g_betas,
                    void* stream);

/* ------------------------------------------------------------------ push / pull contact terms
 * d_i = |p_i - p_argmin[i]|; interior points (exterior == 0) pay tanh(d/0.04)^2, exterior points pay
 * 0.005 tanh(d/0.005)^2 -- only where d < euclthres (TUCH_PULL_THRESHOLD, losses.py:96-105) or
 * everywhere (TUCH_PULL_ALL, loss.py:306-312).  TUCH_REDUCE_SUM adds the terms (losses.py:105,
 * loss.py:315), TUCH_REDUCE_MEAN averages push and pull separately (eft/loss.py:158-166).
 * points[B,N,3], argmin int32 [B,N], exterior uint8 [B,N]; body_active uint8 [B] (NULL = all;
 * inactive bodies get loss 0 and no gradient: ignore_idxs / valid_fit), counts int32 [B] (NULL = N;
 * valid points per body for the data-dependent HD selection).  loss[B]; parts[B,4] = (sum push, sum
 * pull, #push, #pull); g_points[B,N,3] is ACCUMULATED into: += weight * g_loss[b] * d loss[b]/d points. */
#define TUCH_PULL_THRESHOLD 0
#define TUCH_PULL_ALL 1
#define TUCH_REDUCE_SUM 0
#define TUCH_REDUCE_MEAN 1
int tuch_contact_loss(const float* points, const int32_t* argmin, const uint8_t* exterior,
                      const uint8_t* body_active, const int32_t* counts, int B, int N, float euclthres,
                      int pull_mode, int reduce_mode, float weight, const float* g_loss, float* loss,
                      float* parts, float* g_points, void* stream);

/* region-to-region term (losses.py:108-117): r2r[b] = sum over the pairs tuch_region_min reported
 * (arg_i >= 0) of min_sq[b,p]; g_verts[B,V,3] += weight * g_loss[b] * d r2r[b]/d verts through the
 * attaining entry of the expansion-form distance matrix (contact.py:42).  +inf minima (fully masked
 * pair) propagate to r2r but carry no gradient, as in the reference. */
int tuch_region_sum(const float* verts, int B, int V, int n_pairs, const float* min_sq, const int32_t* arg_i,
                    const int32_t* arg_j, const uint8_t* body_active, float weight, const float* g_loss,
                    float* r2r, float* g_verts, void* stream);

/* ------------------------------------------------------------------ a13  torch.optim.Adam
 * One step of Adam (no weight decay, no amsgrad) as configured at smplifydc.py:117,150,197.
 * step_dev is a DEVICE int32 holding the number of steps taken so far; the call uses step+1 for the
 * bias corrections and then increments it (so a captured CUDA graph can be replayed). */
int tuch_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                   int32_t* step_dev, double lr, double beta1, double beta2, double eps, void* stream);

/* ------------------------------------------------------------------ a10  SMPLify-DC stage-2 iteration
 * tuch/smplify/smplifydc.py:155-183, ONE call per iteration for the whole batch:
 *     smpl_output = self.smpl(global_orient, body_pose, betas)            (:157-160)  tuch_smpl_forward
 *     loss = contact_fitting_loss(...)                                      (:162-180)  losses.py:34-123
 *     body_optimizer.zero_grad(); loss.backward(); body_optimizer.step()    (:181-183)  torch.optim.Adam
 * ~27 kernel launches on `stream`, no host synchronisation; bit-identical to composing tuch_smpl_forward,
 * tuch_reprojection_loss, tuch_pose_terms, tuch_contact_query, tuch_contact_loss, tuch_region_min / _sum,
 * tuch_smpl_backward and tuch_adam_step, whose kernels it launches -- minus the pose concatenation, the gradient
 * round trips and the separate Adam launches: the vertex / joint gradients feed the LBS backward directly and Adam
 * runs in the epilogue of its last kernel.  All pointers are DEVICE pointers.  Capturable into a CUDA graph after
 * one warm-up call at the same batch size. */
typedef struct tuch_contact_fit_args {
    /* parameters and torch.optim.Adam state (lr, betas, eps below), all updated in place */
    float* body_pose;            /* [B,69] */
    float* global_orient;        /* [B,3]  */
    float* exp_avg_pose;         /* [B,69] */
    float* exp_avg_sq_pose;      /* [B,69] */
    float* exp_avg_orient;       /* [B,3]  */
    float* exp_avg_sq_orient;    /* [B,3]  */
    int32_t* step_pose;          /* device scalars: steps taken so far, incremented by the call */
    int32_t* step_orient;
    /* fixed inputs of the stage (losses.py:34-41) */
    const float* betas;          /* [B,L]   */
    const float* camera_t;       /* [B,3]   */
    const float* camera_center;  /* [B,2]   */
    const float* joints_2d;      /* [B,J,2] */
    const float* joints_conf;    /* [B,J]   */
    const uint8_t* body_active;  /* [B] or NULL: ~ignore_idxs (losses.py:73) */
    const uint8_t* pair_active;  /* [B,n_pairs] or NULL: gt_contact == 1 & has_discrete_contact & ~ignore (:109-112) */
    /* outputs */
    float* smpl_workspace;       /* tuch_smpl_workspace_floats(smpl, B) floats, 16-byte aligned */
    float* vertices;             /* [B,V,3] of THIS iteration's forward (before the parameter update) */
    float* joints;               /* [B,J,3] */
    float* loss;                 /* scalar: the objective summed over the batch (losses.py:123) */
    float* per_body;             /* [B] or NULL */
    uint8_t* exterior;           /* [B,V] or NULL */
    int32_t* argmin;             /* [B,V] or NULL: as tuch_contact_query_within(radius = euclthres) leaves it */
    float* grad_body_pose;       /* [B,69] or NULL: the gradients Adam consumed (both or neither) */
    float* grad_global_orient;   /* [B,3]  or NULL */
    /* configuration */
    float euclthres, focal_length, sigma, pose_prior_weight, contact_loss_weight;
    int use_segments;
    double lr, beta1, beta2, eps;
} tuch_contact_fit_args;
int tuch_contact_fit_step(const tuch_smpl* smpl, const tuch_topology* topo, const tuch_prior* prior, int B,
                          const tuch_contact_fit_args* args, void* stream);

/* ------------------------------------------------------------------ a12  RegressorLoss.contact_loss
 * tuch/train/loss.py:240-317.  The HD-point regressor (loss.py:81-83 loads it as a dense [N_hd, V]
 * matrix) is handed over in CSR form (HOST arrays, copied) together with faces_vert_is_sampled_from
 * (loss.py:84-88). */
int tuch_topology_set_hd(tuch_topology* topo, int n_hd, const int32_t* row_offsets_host,
                         const int32_t* cols_host, const float* vals_host, const int32_t* hd_face_host);
int tuch_topology_num_hd(const tuch_topology* topo);

/* loss[b] for every body with valid[b] != 0 (valid NULL = all; others get 0):
 *   exterior / nearest geodesically-far vertex of the mesh vertices, segment whitelist always on (:251-270)
 *   use_hd != 0: HD points on faces touching a vertex that is in contact (min_sq < euclthres^2) or interior
 *     (:278-281), regressed from the vertices (:285), masked nearest HD point through the proxy vertices
 *     (:288-291), inside test of the 1 mm normal-offset points against the mesh (:295-297),
 *     loss[b] = sum 0.005 tanh(d/0.005)^2 [exterior] + sum tanh(d/0.04)^2 [interior]   (:299-315)
 *   use_hd == 0: the same terms on the mesh vertices (:303).
 * g_verts[B,V,3] (may be NULL) is ACCUMULATED into: += weight * g_loss[b] * d loss[b] / d verts.
 * Optional debug outputs: counts_out[B] selected HD points per body, sel_out[B,N_hd] their indices
 * (first counts entries), hd_argmin_out / hd_exterior_out [B,N_hd] in selection order.
 * The caller forms contact_loss[valid_fit].mean() (:317). */
int tuch_regressor_contact_loss(const tuch_topology* topo, const float* verts, int B, const uint8_t* valid,
                                float euclthres, int use_hd, float weight, const float* g_loss, float* loss,
                                float* g_verts, int32_t* counts_out, int32_t* sel_out, int32_t* hd_argmin_out,
                                uint8_t* hd_exterior_out, void* stream);

/* ------------------------------------------------------------------ f2  estimate_translation
 * tuch/utils/geometry.py:114-205: camera translation that brings the 3-D joints S closest to their 2-D
 * detections (weighted least squares, weights sqrt(conf)); bodies with has_2d_kp_anno use the joints
 * [n_openpose, J), the others [0, n_openpose) (:192-199); a body whose selected confidences sum to 0 gets
 * zeros (:201).  fp32 inputs, fp64 arithmetic, fp32 result -- as the reference's numpy code.  A singular
 * system (the reference raises LinAlgError) yields non-finite output.
 * S[B,J,3], joints_2d[B,J,3] (x, y, conf), has_2d_kp_anno uint8 [B] -> out[B,3]. */
int tuch_estimate_translation(const float* S, const float* joints_2d, const uint8_t* has_2d_kp_anno, int B, int J,
                              int n_openpose, float focal_length, float img_size, float* out, void* stream);

/* ------------------------------------------------------------------ f3  pose bookkeeping
 * rotation_matrix_to_angle_axis of torchgeometry==0.1.2 (third-party; call site train_module.py:208-211):
 * rotmat[N,3,cols] with cols = 3 or 4 (homogeneous [R | t]) -> out[N,3]. */
int tuch_rotmat_to_angle_axis(const float* rotmat, int N, int cols, float* out, void* stream);
/* FitsDict.rotate_pose / flip_pose (tuch/train/fits_dict.py:89-119) for a batch of poses [B,D]:
 * rot_deg[B] in-plane rotation in degrees (NULL = none), is_flipped uint8 [B] (NULL = none),
 * flip_perm[D] = SMPL_POSE_FLIP_PERM.  flip_first == 0: out = flip(rotate(pose, rot)) (__getitem__, :73);
 * flip_first != 0: out = rotate(flip(pose), rot) (__setitem__ passes -rot, :83). */
int tuch_fits_pose_transform(const float* pose, const float* rot_deg, const uint8_t* is_flipped,
                             const int32_t* flip_perm, int B, int D, int flip_first, float* out, void* stream);

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
