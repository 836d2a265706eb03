// C ABI (include/tuch_b200.h): status, scratch arenas, mesh topology and the contact entry points.
#include "api_internal.h"
#include "strips.h"
#include "clusters.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <string>

namespace tuch {

// ---------------------------------------------------------------- status
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
    return 2;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---------------------------------------------------------------- per-kernel timing
struct TimedPair { cudaEvent_t a, b; };
struct TimedKernel { std::string name; std::vector<TimedPair> pending; double total_ms = 0.0; long long count = 0; };
static std::mutex g_timing_mu;
static bool g_timing_on = false;
static std::vector<TimedKernel> g_timed;

KernelTimer::KernelTimer(const char* name, cudaStream_t st) : st_(st) {
    if (!g_timing_on) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    int k = -1;
    for (size_t i = 0; i < g_timed.size(); ++i) if (g_timed[i].name == name) { k = (int)i; break; }
    if (k < 0) { g_timed.push_back(TimedKernel{name}); k = (int)g_timed.size() - 1; }
    TimedPair p{};
    if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
    cudaEventRecord(p.a, st);
    g_timed[k].pending.push_back(p);
    slot_ = k;
}
KernelTimer::~KernelTimer() {
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    cudaEventRecord(g_timed[slot_].pending.back().b, st_);
}

// ---------------------------------------------------------------- scratch arenas
// One grow-only block per (device, slot, stream).  A CUDA graph captured over a call bakes the block's address
// into its kernel nodes, so a block that a capture has seen is never freed while the library lives: when such an
// arena has to grow (a later, larger eager call on the same stream) the old block is RETIRED -- kept allocated
// for the graphs that still point at it -- and a new one serves the calls from then on.  tuch_release_scratch()
// is the one call that frees everything; it bumps tuch_scratch_generation() so that holders of captured graphs
// can tell that they must re-capture (tuch_b200/smplify/smplifydc.py does).
struct Arena { void* ptr = nullptr; size_t cap = 0; bool captured = false; };
static std::mutex g_arena_mu;
static std::map<std::pair<std::pair<int, int>, cudaStream_t>, Arena> g_arenas;
static std::vector<std::pair<int, void*>> g_retired;          // (device, block) kept alive for captured graphs
static std::atomic<long long> g_scratch_generation{0};

int arena_get(cudaStream_t st, size_t bytes, void** out, int slot) {
    int dev = 0;
    TUCH_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_arena_mu);
    Arena& a = g_arenas[{{dev, slot}, st}];
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (a.cap < bytes) {
        TUCH_REQUIRE(cs == cudaStreamCaptureStatusNone,
                     "scratch arena must grow (%zu -> %zu bytes) during CUDA-graph capture; run the call once "
                     "outside the capture first", a.cap, bytes);
        if (a.ptr != nullptr) {
            if (a.captured) {
                g_retired.emplace_back(dev, a.ptr);           // a graph may still read and write it
            } else {
                TUCH_CUDA(cudaStreamSynchronize(st));
                TUCH_CUDA(cudaFree(a.ptr));
            }
            a.ptr = nullptr; a.cap = 0; a.captured = false;
        }
        const size_t want = align_up(bytes + bytes / 4, (size_t)1 << 20);
        TUCH_CUDA(cudaMalloc(&a.ptr, want));
        a.cap = want;
    }
    if (cs != cudaStreamCaptureStatusNone) a.captured = true;
    *out = a.ptr;
    return 0;
}

int Scratch::commit_slot(cudaStream_t st, int slot) {
    void* p = nullptr;
    if (int rc = arena_get(st, total_ > 0 ? total_ : 256, &p, slot)) return rc;
    base_ = (char*)p;
    return 0;
}

}  // namespace tuch

using namespace tuch;

// ================================================================ status
TUCH_EXPORT const char* tuch_last_error(void) { return g_err; }
TUCH_EXPORT int tuch_abi_version(void) { return TUCH_B200_ABI_VERSION; }
TUCH_EXPORT long long tuch_launch_count(void) { return g_launches.load(); }

TUCH_EXPORT int tuch_device_info(int* sms, int* cc_major, int* cc_minor) {
    int dev = 0, mj = 0, mn = 0, n = 0;
    TUCH_CUDA(cudaGetDevice(&dev));
    TUCH_CUDA(cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev));
    TUCH_CUDA(cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, dev));
    TUCH_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (sms) *sms = n;
    if (cc_major) *cc_major = mj;
    if (cc_minor) *cc_minor = mn;
    TUCH_REQUIRE(mj == 10, "tuch_b200 is built for sm_100a only; current device is sm_%d%d", mj, mn);
    return 0;
}

TUCH_EXPORT int tuch_release_scratch(void) {
    int dev = 0;
    TUCH_CUDA(cudaGetDevice(&dev));
    TUCH_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_arena_mu);
    for (auto it = g_arenas.begin(); it != g_arenas.end();) {
        if (it->first.first.first == dev) {
            if (it->second.ptr) cudaFree(it->second.ptr);
            it = g_arenas.erase(it);
        } else {
            ++it;
        }
    }
    for (auto it = g_retired.begin(); it != g_retired.end();) {
        if (it->first == dev) { cudaFree(it->second); it = g_retired.erase(it); } else { ++it; }
    }
    g_scratch_generation.fetch_add(1);          // every graph captured before this call is now invalid
    return 0;
}

TUCH_EXPORT long long tuch_scratch_generation(void) { return g_scratch_generation.load(); }

TUCH_EXPORT int tuch_kernel_timing_enable(int on) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timing_on = on != 0;
    return 0;
}

TUCH_EXPORT int tuch_kernel_timing_reset(void) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (auto& k : g_timed) {
        for (auto& p : k.pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        k.pending.clear(); k.total_ms = 0.0; k.count = 0;
    }
    return 0;
}

TUCH_EXPORT int tuch_kernel_timing_names(char* buf, int capacity) {
    TUCH_REQUIRE(buf != nullptr && capacity > 0, "tuch_kernel_timing_names: need a buffer");
    std::lock_guard<std::mutex> lk(g_timing_mu);
    std::string all;
    for (auto& k : g_timed) all += (all.empty() ? "" : ",") + k.name;
    TUCH_REQUIRE((int)all.size() < capacity, "tuch_kernel_timing_names: buffer of %d bytes is too small (%zu needed)", capacity, all.size() + 1);
    std::memcpy(buf, all.c_str(), all.size() + 1);
    return 0;
}

TUCH_EXPORT int tuch_kernel_timing_read(const char* name, double* total_ms, long long* launches) {
    TUCH_REQUIRE(name != nullptr, "tuch_kernel_timing_read: name is null");
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (total_ms) *total_ms = 0.0;
    if (launches) *launches = 0;
    for (auto& k : g_timed) {
        if (k.name != name) continue;
        for (auto& p : k.pending) {
            TUCH_CUDA(cudaEventSynchronize(p.b));
            float ms = 0.f;
            TUCH_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
            k.total_ms += ms; k.count += 1;
            cudaEventDestroy(p.a); cudaEventDestroy(p.b);
        }
        k.pending.clear();
        if (total_ms) *total_ms = k.total_ms;
        if (launches) *launches = k.count;
    }
    return 0;
}

// ================================================================ a1..a3
TUCH_EXPORT int tuch_pairwise_dist(const float* x, const float* y, int bs, int nx, int ny, int squared,
                                   float* P, void* stream) {
    TUCH_REQUIRE(bs >= 0 && nx >= 0 && ny >= 0, "tuch_pairwise_dist: negative size");
    if (bs == 0 || nx == 0 || ny == 0) return 0;
    TUCH_REQUIRE(x && y && P, "tuch_pairwise_dist: null pointer");
    return launch_pairwise_dist(x, y, bs, nx, ny, squared, P, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_pairwise_dist_backward(const float* x, const float* y, const float* P, const float* gP,
                                            int bs, int nx, int ny, int squared, float* gx, float* gy,
                                            void* stream) {
    TUCH_REQUIRE(bs >= 0 && nx >= 0 && ny >= 0, "tuch_pairwise_dist_backward: negative size");
    if (bs == 0) return 0;
    TUCH_REQUIRE(x && y && gP && (squared || P), "tuch_pairwise_dist_backward: null pointer");
    return launch_pairwise_dist_bwd(x, y, P, gP, bs, nx, ny, squared, gx, gy, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_solid_angles(const float* points, const float* triangles, int bs, int Q, int F,
                                  float* out, void* stream) {
    TUCH_REQUIRE(bs >= 0 && Q >= 0 && F >= 0, "tuch_solid_angles: negative size");
    if (bs == 0 || Q == 0 || F == 0) return 0;
    TUCH_REQUIRE(points && triangles && out, "tuch_solid_angles: null pointer");
    return launch_solid_angles(points, triangles, bs, Q, F, out, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_winding_numbers(const float* points, const float* triangles, int bs, int Q, int F,
                                     float* out, void* stream) {
    TUCH_REQUIRE(bs >= 0 && Q >= 0 && F >= 0, "tuch_winding_numbers: negative size");
    if (bs == 0 || Q == 0) return 0;
    TUCH_REQUIRE(points && out, "tuch_winding_numbers: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (F == 0) {                                     // empty sum (contact.py:146-147)
        TUCH_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)bs * Q, st));
        return 0;
    }
    TUCH_REQUIRE(triangles, "tuch_winding_numbers: null pointer");
    const int Fp = padded_faces(F);
    const int S = winding_splits(bs, Q, Fp, sm_count());
    Scratch sc;
    auto h_tri = sc.plan(sizeof(float4) * 3 * (size_t)bs * Fp);
    auto h_par = sc.plan(sizeof(float) * (size_t)bs * S * Q);
    if (int rc = sc.commit(st)) return rc;
    float4* tri12 = sc.get<float4>(h_tri);
    float* partial = sc.get<float>(h_par);
    if (int rc = launch_pack_triangles(triangles, bs, F, Fp, tri12, st)) return rc;
    WindingJob j{tri12, (long long)Fp * 3, points, (long long)Q * 3, partial, out, (long long)Q, nullptr, bs, Q, Fp, S};
    return launch_winding(j, st);
}

// ================================================================ topology
static void free_dev(void* p) { if (p) cudaFree(p); }

template <typename T>
static int upload(const T* host, size_t n, T** dev) {
    *dev = nullptr;
    if (n == 0) return 0;
    TUCH_CUDA(cudaMalloc((void**)dev, n * sizeof(T)));
    TUCH_CUDA(cudaMemcpy(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

TUCH_EXPORT int tuch_topology_create(int V, int F, const int32_t* faces_host, tuch_topology** out) {
    TUCH_REQUIRE(out != nullptr, "tuch_topology_create: out is null");
    *out = nullptr;
    TUCH_REQUIRE(V > 0 && F >= 0, "tuch_topology_create: need V > 0, F >= 0 (got V=%d F=%d)", V, F);
    TUCH_REQUIRE(F == 0 || faces_host != nullptr, "tuch_topology_create: faces is null");
    for (size_t i = 0; i < (size_t)F * 3; ++i)
        TUCH_REQUIRE(faces_host[i] >= 0 && faces_host[i] < V, "tuch_topology_create: face index %d out of range [0,%d)",
                     faces_host[i], V);
    tuch_topology* t = new tuch_topology();
    TUCH_CUDA(cudaGetDevice(&t->device));
    t->V = V; t->F = F;
    t->Fp = padded_faces(F); t->Vp = padded_verts(V); t->Vq = padded_verts(V); t->W = cdiv(V, 32);
    if (int rc = upload(faces_host, (size_t)F * 3, &t->d_faces)) { delete t; return rc; }
    if (cudaMalloc((void**)&t->d_stats, 2 * sizeof(int)) != cudaSuccess || cudaMemset(t->d_stats, 0, 2 * sizeof(int)) != cudaSuccess) {
        tuch_topology_destroy(t);
        return cuda_fail(cudaGetLastError(), "tuch_topology_create: stats", __FILE__, __LINE__);
    }
    if (F > 0) {
        std::vector<int> vid, fid;
        std::vector<uint32_t> flag;
        int rc = build_strip_stream(faces_host, F, WS_TILE, vid, flag, fid, &t->n_strips);
        if (!rc) rc = upload(vid.data(), vid.size(), &t->d_strip_vid);
        if (!rc) rc = upload(fid.data(), fid.size(), &t->d_strip_fid);
        if (rc) { tuch_topology_destroy(t); return rc; }
        t->Lp = (int)vid.size();
    }
    *out = t;
    return 0;
}

static void free_regions(tuch_topology* t) {
    free_dev(t->d_region_off); free_dev(t->d_region_ids); free_dev(t->d_pair_a); free_dev(t->d_pair_b);
    free_dev(t->d_pair_mask); free_dev(t->d_pair_word_off);
    t->d_region_off = t->d_region_ids = t->d_pair_a = t->d_pair_b = nullptr;
    t->d_pair_mask = nullptr; t->d_pair_word_off = nullptr;
    t->h_pair_word_off.clear(); t->max_pair_words = 0; t->has_pair_mask = false;
    t->n_regions = t->n_pairs = 0;
}

// (re)builds the per-pair sub-masks once the regions and the geodesic mask both exist
static int refresh_pair_mask(tuch_topology* t, cudaStream_t st) {
    t->has_pair_mask = false;
    if (!t->has_mask || t->n_pairs == 0 || t->d_pair_word_off == nullptr) return 0;
    if (int rc = launch_pair_mask(t->d_maskT, t->Vq, t->d_region_ids, t->d_region_off, t->d_pair_a, t->d_pair_b,
                                  t->d_pair_word_off, t->n_pairs, t->max_pair_words, t->d_pair_mask, st)) return rc;
    t->has_pair_mask = true;
    return 0;
}
static void free_segments(tuch_topology* t) {
    free_dev(t->d_seg_vidx); free_dev(t->d_seg_faces); free_dev(t->d_slot_face); free_dev(t->d_slot_band0);
    free_dev(t->d_loop_off); free_dev(t->d_loop_ids);
    free_dev(t->d_member_seg); free_dev(t->d_seg_face_off); free_dev(t->d_seg_band0);
    t->d_member_seg = t->d_seg_face_off = t->d_seg_band0 = nullptr;
    t->d_seg_vidx = t->d_seg_faces = t->d_slot_face = t->d_slot_band0 = t->d_loop_off = t->d_loop_ids = nullptr;
    t->n_segments = t->n_bands = t->n_sv = t->n_slots = 0;
    t->h_vidx_off.clear(); t->h_slot_off.clear();
}

TUCH_EXPORT void tuch_topology_destroy(tuch_topology* t) {
    if (!t) return;
    free_dev(t->d_faces); free_dev(t->d_maskT); free_dev(t->d_strip_vid); free_dev(t->d_strip_fid);
    free_dev(t->d_leaf_face); free_dev(t->d_mid_off); free_dev(t->d_top_off); free_dev(t->d_vtile); free_dev(t->d_vgroup_off);
    free_dev(t->d_maskP); free_dev(t->d_maskG); free_dev(t->d_tile_any); free_dev(t->d_stats);
    free_dev(t->d_hd_row_off); free_dev(t->d_hd_cols); free_dev(t->d_hd_face); free_dev(t->d_hd_vals);
    free_regions(t); free_segments(t);
    delete t;
}
TUCH_EXPORT int tuch_topology_strip_stats(const tuch_topology* t, int* stream_len, int* n_strips) {
    TUCH_REQUIRE(t != nullptr, "tuch_topology_strip_stats: null topology");
    if (stream_len) *stream_len = t->Lp;
    if (n_strips) *n_strips = t->n_strips;
    return 0;
}

TUCH_EXPORT int tuch_strip_stream_host(const int32_t* faces_host, int F, int32_t* vid_out, uint32_t* flag_out,
                                       int capacity, int* stream_len, int* n_strips) {
    TUCH_REQUIRE(faces_host != nullptr && F > 0, "tuch_strip_stream_host: need faces");
    std::vector<int> vid, fid;
    std::vector<uint32_t> flag;
    if (int rc = build_strip_stream(faces_host, F, WS_TILE, vid, flag, fid, n_strips)) return rc;
    if (stream_len) *stream_len = (int)vid.size();
    if (vid_out != nullptr && flag_out != nullptr) {
        TUCH_REQUIRE(capacity >= (int)vid.size(), "tuch_strip_stream_host: capacity %d < stream length %zu", capacity, vid.size());
        std::copy(vid.begin(), vid.end(), vid_out);
        std::copy(flag.begin(), flag.end(), flag_out);
    }
    return 0;
}

// (re)builds the cluster-ordered copy of the geodesic mask once both inputs exist
static int refresh_permuted_mask(tuch_topology* t, cudaStream_t st) {
    t->has_maskP = false;
    if (!t->has_mask || !t->has_clusters) return 0;
    const int T = t->T;
    free_dev(t->d_maskP);
    t->d_maskP = nullptr;
    TUCH_CUDA(cudaMalloc((void**)&t->d_maskP, sizeof(uint32_t) * (size_t)T * T * 32));
    if (int rc = launch_permute_mask(t->d_maskT, t->Vq, t->d_vtile, T, t->d_maskP, st)) return rc;
    free_dev(t->d_maskG);
    t->d_maskG = nullptr;
    TUCH_CUDA(cudaMalloc((void**)&t->d_maskG, sizeof(uint32_t) * (size_t)cdiv(t->NG, 32) * T * 32));
    if (int rc = launch_group_mask(t->d_maskP, t->d_vgroup_off, T, t->NG, t->d_maskG, st)) return rc;
    free_dev(t->d_tile_any);
    t->d_tile_any = nullptr;
    TUCH_CUDA(cudaMalloc((void**)&t->d_tile_any, sizeof(uint32_t) * (size_t)T * cdiv(T, 32)));
    if (int rc = launch_tile_any_mask(t->d_maskP, T, t->d_tile_any, st)) return rc;
    t->has_maskP = true;
    return 0;
}

static int install_clusters(tuch_topology* t, const float* verts_host, cudaStream_t st = nullptr) {
    std::vector<int> faces((size_t)t->F * 3);
    TUCH_CUDA(cudaMemcpy(faces.data(), t->d_faces, faces.size() * sizeof(int), cudaMemcpyDeviceToHost));
    ClusterTree tree;
    if (int rc = build_cluster_tree(faces.data(), t->F, t->V, verts_host, tree)) return rc;
    TUCH_CUDA(cudaDeviceSynchronize());            // queued work may still read the old hierarchy
    free_dev(t->d_leaf_face); free_dev(t->d_mid_off); free_dev(t->d_top_off); free_dev(t->d_vtile); free_dev(t->d_vgroup_off);
    t->d_leaf_face = t->d_mid_off = t->d_top_off = t->d_vtile = t->d_vgroup_off = nullptr;
    t->has_clusters = false;
    if (int rc = upload(tree.leaf_face.data(), tree.leaf_face.size(), &t->d_leaf_face)) return rc;
    if (int rc = upload(tree.mid_off.data(), tree.mid_off.size(), &t->d_mid_off)) return rc;
    if (int rc = upload(tree.top_off.data(), tree.top_off.size(), &t->d_top_off)) return rc;
    if (int rc = upload(tree.vtile.data(), tree.vtile.size(), &t->d_vtile)) return rc;
    if (int rc = upload(tree.vgroup_off.data(), tree.vgroup_off.size(), &t->d_vgroup_off)) return rc;
    t->K = tree.K; t->NM = tree.NM; t->NT = tree.NT; t->T = tree.T; t->NG = tree.NG; t->max_top_leaves = tree.max_top_leaves;
    t->has_clusters = true;
    return refresh_permuted_mask(t, st);
}

TUCH_EXPORT int tuch_topology_set_template(tuch_topology* t, const float* verts_host) {
    TUCH_REQUIRE(t != nullptr && verts_host != nullptr, "tuch_topology_set_template: null pointer");
    TUCH_REQUIRE(t->F > 0, "tuch_topology_set_template: topology has no faces");
    for (size_t i = 0; i < (size_t)t->V * 3; ++i)
        TUCH_REQUIRE(std::isfinite(verts_host[i]), "tuch_topology_set_template: non-finite coordinate at %zu", i);
    return install_clusters(t, verts_host);
}

TUCH_EXPORT int tuch_cluster_tree_host(const int32_t* faces_host, int F, int V, const float* verts_host,
                                       int32_t* leaf_face_out, int leaf_capacity, int32_t* mid_off_out,
                                       int mid_capacity, int32_t* top_off_out, int top_capacity,
                                       int32_t* vtile_out, int tile_capacity, int* n_leaves, int* n_mids,
                                       int* n_tops, int* n_tiles) {
    TUCH_REQUIRE(faces_host != nullptr && verts_host != nullptr && F > 0 && V > 0, "tuch_cluster_tree_host: need a mesh");
    for (size_t i = 0; i < (size_t)F * 3; ++i)
        TUCH_REQUIRE(faces_host[i] >= 0 && faces_host[i] < V, "tuch_cluster_tree_host: face index %d out of range", faces_host[i]);
    ClusterTree tree;
    if (int rc = build_cluster_tree(faces_host, F, V, verts_host, tree)) return rc;
    if (n_leaves) *n_leaves = tree.K;
    if (n_mids) *n_mids = tree.NM;
    if (n_tops) *n_tops = tree.NT;
    if (n_tiles) *n_tiles = tree.T;
    if (leaf_face_out != nullptr) {
        TUCH_REQUIRE(leaf_capacity >= tree.K, "tuch_cluster_tree_host: leaf capacity %d < %d", leaf_capacity, tree.K);
        std::copy(tree.leaf_face.begin(), tree.leaf_face.end(), leaf_face_out);
    }
    if (mid_off_out != nullptr) {
        TUCH_REQUIRE(mid_capacity >= tree.NM, "tuch_cluster_tree_host: mid capacity %d < %d", mid_capacity, tree.NM);
        std::copy(tree.mid_off.begin(), tree.mid_off.end(), mid_off_out);
    }
    if (top_off_out != nullptr) {
        TUCH_REQUIRE(top_capacity >= tree.NT, "tuch_cluster_tree_host: top capacity %d < %d", top_capacity, tree.NT);
        std::copy(tree.top_off.begin(), tree.top_off.end(), top_off_out);
    }
    if (vtile_out != nullptr) {
        TUCH_REQUIRE(tile_capacity >= tree.T, "tuch_cluster_tree_host: tile capacity %d < %d", tile_capacity, tree.T);
        std::copy(tree.vtile.begin(), tree.vtile.end(), vtile_out);
    }
    return 0;
}

TUCH_EXPORT int tuch_topology_set_winding_mode(tuch_topology* t, int mode) {
    TUCH_REQUIRE(t != nullptr, "tuch_topology_set_winding_mode: null topology");
    TUCH_REQUIRE(mode == TUCH_WINDING_EXACT || mode == TUCH_WINDING_FAST, "tuch_topology_set_winding_mode: unknown mode %d", mode);
    t->winding_mode = mode;
    return 0;
}

TUCH_EXPORT int tuch_topology_cluster_stats(const tuch_topology* t, int* n_leaves, int* n_mids, int* n_tops,
                                            int* n_tiles, int* leaf_faces) {
    TUCH_REQUIRE(t != nullptr, "tuch_topology_cluster_stats: null topology");
    if (n_leaves) *n_leaves = t->has_clusters ? t->K : 0;
    if (n_mids) *n_mids = t->has_clusters ? t->NM : 0;
    if (n_tops) *n_tops = t->has_clusters ? t->NT : 0;
    if (n_tiles) *n_tiles = t->has_clusters ? t->T : 0;
    if (leaf_faces) *leaf_faces = WC_LEAF;
    return 0;
}

TUCH_EXPORT int tuch_topology_query_stats(const tuch_topology* t, int* refine_vertices, int* refine_points, void* stream) {
    TUCH_REQUIRE(t != nullptr, "tuch_topology_query_stats: null topology");
    int h[2] = {0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    TUCH_CUDA(cudaMemcpyAsync(h, t->d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    TUCH_CUDA(cudaStreamSynchronize(st));
    if (refine_vertices) *refine_vertices = h[0];
    if (refine_points) *refine_points = h[1];
    return 0;
}

TUCH_EXPORT int tuch_topology_pack_nodes(const tuch_topology* t, const float* verts, int B, int direct, float* nodes_out,
                                         void* stream) {
    TUCH_REQUIRE(t != nullptr && verts != nullptr && nodes_out != nullptr, "tuch_topology_pack_nodes: null pointer");
    TUCH_REQUIRE(t->has_clusters, "tuch_topology_pack_nodes: the topology has no face hierarchy (tuch_topology_set_template)");
    TUCH_REQUIRE(B >= 0 && B <= 65535, "tuch_topology_pack_nodes: batch out of range");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_nodes = (size_t)t->NT + t->NM + t->K;
    Scratch sc;
    const size_t h_tri = sc.plan(sizeof(float4) * 3 * (size_t)B * t->K * WC_LEAF);
    const size_t h_info = sc.plan(sizeof(float4) * WC_NODE_F4 * (size_t)B * n_nodes);
    if (int rc = sc.commit(st)) return rc;
    ClusterJob j{verts, t->d_faces, t->d_leaf_face, t->d_mid_off, t->d_top_off, t->d_vtile, sc.get<float4>(h_tri),
                 sc.get<float4>(h_info), nullptr, nullptr, nullptr, B, t->V, t->K, t->NM, t->NT, 1, t->T};
    j.max_top_leaves = t->max_top_leaves;
    j.direct_pack = direct != 0;
    if (int rc = launch_cluster_pack(j, st)) return rc;
    TUCH_CUDA(cudaMemcpyAsync(nodes_out, j.nodes, sizeof(float4) * WC_NODE_F4 * (size_t)B * n_nodes, cudaMemcpyDeviceToDevice, st));
    return 0;
}

TUCH_EXPORT int tuch_topology_num_verts(const tuch_topology* t) { return t ? t->V : -1; }
TUCH_EXPORT int tuch_topology_num_faces(const tuch_topology* t) { return t ? t->F : -1; }
TUCH_EXPORT int tuch_topology_total_segment_verts(const tuch_topology* t) { return t ? t->n_sv : -1; }

static int ensure_mask(tuch_topology* t) {
    if (t->d_maskT == nullptr)
        TUCH_CUDA(cudaMalloc((void**)&t->d_maskT, sizeof(uint32_t) * (size_t)t->W * t->Vq));
    return 0;
}

TUCH_EXPORT int tuch_topology_set_geodist(tuch_topology* t, const float* geodist, float geothres, void* stream) {
    TUCH_REQUIRE(t && geodist, "tuch_topology_set_geodist: null pointer");
    if (int rc = ensure_mask(t)) return rc;
    if (int rc = launch_pack_mask(nullptr, geodist, geothres, t->V, t->Vq, t->W, t->d_maskT, (cudaStream_t)stream)) return rc;
    t->has_mask = true;
    if (int rc = refresh_pair_mask(t, (cudaStream_t)stream)) return rc;
    return refresh_permuted_mask(t, (cudaStream_t)stream);
}
TUCH_EXPORT int tuch_topology_set_geomask(tuch_topology* t, const uint8_t* geomask, void* stream) {
    TUCH_REQUIRE(t && geomask, "tuch_topology_set_geomask: null pointer");
    if (int rc = ensure_mask(t)) return rc;
    if (int rc = launch_pack_mask(geomask, nullptr, 0.f, t->V, t->Vq, t->W, t->d_maskT, (cudaStream_t)stream)) return rc;
    t->has_mask = true;
    if (int rc = refresh_pair_mask(t, (cudaStream_t)stream)) return rc;
    return refresh_permuted_mask(t, (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_topology_set_regions(tuch_topology* t, int n_regions, const int32_t* off, const int32_t* ids,
                                          int n_pairs, const int32_t* pa, const int32_t* pb) {
    TUCH_REQUIRE(t != nullptr, "tuch_topology_set_regions: null topology");
    TUCH_REQUIRE(n_regions >= 0 && n_pairs >= 0, "tuch_topology_set_regions: negative count");
    TUCH_REQUIRE(n_regions == 0 || (off && ids), "tuch_topology_set_regions: null region arrays");
    TUCH_REQUIRE(n_pairs == 0 || (pa && pb), "tuch_topology_set_regions: null pair arrays");
    for (int r = 0; r < n_regions; ++r)
        TUCH_REQUIRE(off[r + 1] > off[r], "tuch_topology_set_regions: region %d is empty", r);
    const int total = n_regions ? off[n_regions] : 0;
    for (int i = 0; i < total; ++i)
        TUCH_REQUIRE(ids[i] >= 0 && ids[i] < t->V, "tuch_topology_set_regions: vertex id %d out of range", ids[i]);
    for (int p = 0; p < n_pairs; ++p)
        TUCH_REQUIRE(pa[p] >= 0 && pa[p] < n_regions && pb[p] >= 0 && pb[p] < n_regions,
                     "tuch_topology_set_regions: pair %d names an unknown region", p);
    free_regions(t);
    if (int rc = upload(off, (size_t)n_regions + 1, &t->d_region_off)) return rc;
    if (int rc = upload(ids, (size_t)total, &t->d_region_ids)) return rc;
    if (int rc = upload(pa, (size_t)n_pairs, &t->d_pair_a)) return rc;
    if (int rc = upload(pb, (size_t)n_pairs, &t->d_pair_b)) return rc;
    t->n_regions = n_regions; t->n_pairs = n_pairs;
    t->h_pair_word_off.assign(1, 0);
    for (int p = 0; p < n_pairs; ++p) {
        const long long na = off[pa[p] + 1] - off[pa[p]], nb = off[pb[p] + 1] - off[pb[p]];
        const long long words = na * ((nb + 31) / 32);
        t->max_pair_words = std::max(t->max_pair_words, words);
        t->h_pair_word_off.push_back(t->h_pair_word_off.back() + words);
    }
    if (n_pairs > 0) {
        if (int rc = upload(t->h_pair_word_off.data(), t->h_pair_word_off.size(), &t->d_pair_word_off)) return rc;
        TUCH_CUDA(cudaMalloc((void**)&t->d_pair_mask, sizeof(uint32_t) * (size_t)std::max<long long>(1, t->h_pair_word_off.back())));
    }
    return refresh_pair_mask(t, nullptr);
}

TUCH_EXPORT int tuch_topology_set_segments(tuch_topology* t, int n_segments, const int32_t* vidx_off,
                                           const int32_t* vidx, const int32_t* face_off, const int32_t* faces,
                                           const int32_t* band_off, const int32_t* loop_off,
                                           const int32_t* loop_ids) {
    TUCH_REQUIRE(t != nullptr, "tuch_topology_set_segments: null topology");
    TUCH_REQUIRE(n_segments >= 0, "tuch_topology_set_segments: negative count");
    free_segments(t);
    if (n_segments == 0) return 0;
    TUCH_REQUIRE(vidx_off && vidx && face_off && faces && band_off && loop_off && loop_ids,
                 "tuch_topology_set_segments: null array");
    const int n_sv = vidx_off[n_segments], n_sf = face_off[n_segments], n_bands = band_off[n_segments];
    const int n_loop = loop_off[n_bands];
    for (int i = 0; i < n_sv; ++i)
        TUCH_REQUIRE(vidx[i] >= 0 && vidx[i] < t->V, "tuch_topology_set_segments: member vertex %d out of range", vidx[i]);
    for (int i = 0; i < n_loop; ++i)
        TUCH_REQUIRE(loop_ids[i] >= 0 && loop_ids[i] < t->V, "tuch_topology_set_segments: loop vertex %d out of range", loop_ids[i]);
    for (int j = 0; j < n_bands; ++j)
        TUCH_REQUIRE(loop_off[j + 1] > loop_off[j], "tuch_topology_set_segments: band %d has an empty loop", j);
    std::vector<int> slot_face, slot_band0;
    t->h_vidx_off.assign(vidx_off, vidx_off + n_segments + 1);
    t->h_slot_off.assign(1, 0);
    for (int s = 0; s < n_segments; ++s) {
        const int nb = band_off[s + 1] - band_off[s];
        for (int f = face_off[s]; f < face_off[s + 1]; ++f)
            for (int k = 0; k < 3; ++k)
                TUCH_REQUIRE(faces[3 * f + k] >= 0 && faces[3 * f + k] < t->V + nb,
                             "tuch_topology_set_segments: segment %d face index %d out of range", s, faces[3 * f + k]);
        const int nf = face_off[s + 1] - face_off[s];
        const int nfp = padded_faces(nf);
        for (int i = 0; i < nfp; ++i) {
            slot_face.push_back(i < nf ? face_off[s] + i : -1);
            slot_band0.push_back(band_off[s]);
        }
        t->h_slot_off.push_back((int)slot_face.size());
    }
    if (int rc = upload(vidx, (size_t)n_sv, &t->d_seg_vidx)) return rc;
    if (int rc = upload(faces, (size_t)n_sf * 3, &t->d_seg_faces)) return rc;
    if (int rc = upload(slot_face.data(), slot_face.size(), &t->d_slot_face)) return rc;
    if (int rc = upload(slot_band0.data(), slot_band0.size(), &t->d_slot_band0)) return rc;
    if (int rc = upload(loop_off, (size_t)n_bands + 1, &t->d_loop_off)) return rc;
    if (int rc = upload(loop_ids, (size_t)n_loop, &t->d_loop_ids)) return rc;
    std::vector<int> member_seg((size_t)n_sv);
    for (int s = 0; s < n_segments; ++s)
        for (int k = vidx_off[s]; k < vidx_off[s + 1]; ++k) member_seg[k] = s;
    if (int rc = upload(member_seg.data(), member_seg.size(), &t->d_member_seg)) return rc;
    if (int rc = upload(face_off, (size_t)n_segments + 1, &t->d_seg_face_off)) return rc;
    if (int rc = upload(band_off, (size_t)n_segments, &t->d_seg_band0)) return rc;
    t->n_segments = n_segments; t->n_bands = n_bands; t->n_sv = n_sv; t->n_slots = (int)slot_face.size();
    return 0;
}

// ================================================================ fused self-contact query
namespace tuch {

// Segment pass shared by tuch_contact_query and tuch_segment_exterior.
struct SegPlan {
    size_t apex, tri, pts, wind;
    std::vector<size_t> partial;
    std::vector<int> S;
};

static void plan_segments(const tuch_topology* t, int B, Scratch& sc, SegPlan& p) {
    p.apex = sc.plan(sizeof(float) * 3 * (size_t)B * (t->n_bands > 0 ? t->n_bands : 1));
    p.tri = sc.plan(sizeof(float4) * 3 * (size_t)B * t->n_slots);
    p.pts = sc.plan(sizeof(float) * 3 * (size_t)B * t->n_sv);
    p.wind = sc.plan(sizeof(float) * (size_t)B * t->n_sv);
    for (int s = 0; s < t->n_segments; ++s) {
        const int Q = t->h_vidx_off[s + 1] - t->h_vidx_off[s];
        const int Fp = t->h_slot_off[s + 1] - t->h_slot_off[s];
        const int S = winding_splits(B, Q, Fp, sm_count());
        p.S.push_back(S);
        p.partial.push_back(sc.plan(sizeof(float) * (size_t)B * S * Q));
    }
}

static int run_segments(const tuch_topology* t, const float* verts, int B, Scratch& sc, const SegPlan& p,
                        const uint8_t* body_active, uint8_t* exterior, uint8_t* seg_ext_out, float* seg_w_out,
                        cudaStream_t st) {
    float* apex = sc.get<float>(p.apex);
    float4* tri = sc.get<float4>(p.tri);
    float* pts = sc.get<float>(p.pts);
    float* wind = seg_w_out ? seg_w_out : sc.get<float>(p.wind);
    if (int rc = launch_segment_apex(verts, B, t->V, t->d_loop_off, t->d_loop_ids, t->n_bands, apex, body_active, st)) return rc;
    if (int rc = launch_segment_pack(verts, B, t->V, apex, t->n_bands, t->d_seg_faces, t->d_slot_face, t->d_slot_band0,
                                     t->n_slots, t->d_seg_vidx, t->n_sv, tri, pts, body_active, st)) return rc;
    for (int s = 0; s < t->n_segments; ++s) {
        const int Q = t->h_vidx_off[s + 1] - t->h_vidx_off[s];
        const int Fp = t->h_slot_off[s + 1] - t->h_slot_off[s];
        WindingJob j{tri + (size_t)t->h_slot_off[s] * 3, (long long)t->n_slots * 3,
                     pts + (size_t)t->h_vidx_off[s] * 3, (long long)t->n_sv * 3,
                     sc.get<float>(p.partial[s]), wind + t->h_vidx_off[s], (long long)t->n_sv,
                     body_active, B, Q, Fp, p.S[s]};
        if (int rc = launch_winding(j, st)) return rc;
    }
    return launch_segment_apply(wind, t->d_seg_vidx, t->n_sv, B, t->V, exterior, seg_ext_out, body_active, st);
}

int contact_query_impl(const tuch_topology* t, const float* verts, int B, int use_segments, int32_t* argmin,
                       float* min_sq, float* winding, uint8_t* exterior, float4* vert4_out, cudaStream_t st,
                       PackedClusters* packed_out, const QueryStreams* qs) {
    const cudaStream_t st_nn = qs != nullptr ? qs->nn : st;
    TUCH_REQUIRE(t != nullptr, "tuch_contact_query: null topology");
    TUCH_REQUIRE(B >= 0, "tuch_contact_query: negative batch");
    TUCH_REQUIRE(B <= 65535, "tuch_contact_query: at most 65535 bodies per call (the batch is a grid dimension), got %d", B);
    if (B == 0) return 0;
    TUCH_REQUIRE(verts != nullptr, "tuch_contact_query: verts is null");
    const bool want_w = winding != nullptr || exterior != nullptr;
    const bool want_nn = argmin != nullptr || min_sq != nullptr;
    TUCH_REQUIRE(!want_nn || t->has_mask,
                 "tuch_contact_query: nearest-vertex outputs requested but the topology has no geodesic mask "
                 "(call tuch_topology_set_geodist / tuch_topology_set_geomask)");
    TUCH_REQUIRE(!want_w || t->F > 0, "tuch_contact_query: winding requested on a topology without faces");
    const bool segs = exterior != nullptr && use_segments && t->n_segments > 0;
    const int V = t->V, Fp = t->Fp, Vp = t->Vp, Lp = t->Lp;
    if (want_w && t->winding_mode == TUCH_WINDING_FAST && !t->has_clusters) {
        // no template was given: cluster the faces on the first body seen (one blocking copy, once);
        // inside a CUDA-graph capture the exact kernel is used instead
        static std::mutex lazy_mu;                        // two threads may race to the first query
        std::lock_guard<std::mutex> lk(lazy_mu);
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        TUCH_CUDA(cudaStreamIsCapturing(st, &cs));
        if (cs == cudaStreamCaptureStatusNone && !t->has_clusters) {
            std::vector<float> h((size_t)V * 3);
            TUCH_CUDA(cudaMemcpyAsync(h.data(), verts, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
            TUCH_CUDA(cudaStreamSynchronize(st));
            bool finite = true;
            for (float x : h) finite = finite && std::isfinite(x);
            if (finite)
                if (int rc = install_clusters(const_cast<tuch_topology*>(t), h.data(), st)) return rc;   // lazily built cache
        }
    }
    const bool fast = want_w && t->winding_mode == TUCH_WINDING_FAST && t->has_clusters;
    const int S = !want_w ? 1 : fast ? cluster_splits(B, t->T, t->NT, sm_count()) : strip_splits(B, V, Lp, sm_count());

    Scratch sc;
    const size_t h_tri = sc.plan(!want_w ? 0 : fast ? sizeof(float4) * 3 * (size_t)B * t->K * WC_LEAF
                                                    : sizeof(float4) * 2 * (size_t)B * Lp);
    const size_t h_info = sc.plan(!want_w ? 0 : fast ? sizeof(float4) * WC_NODE_F4 * (size_t)B * (t->NT + t->NM + t->K)
                                                     : sizeof(float4) * (size_t)B * (Lp / WS_TILE));
    const size_t h_ref = sc.plan(fast ? sizeof(int) * ((size_t)B * V + 1) : 0);
    const bool nn_tiles = want_nn && t->has_maskP && vert4_out == nullptr;
    const int T = t->T;
    const size_t h_v4 = sc.plan(!want_nn ? 0 : nn_tiles ? sizeof(float4) * (size_t)B * T * 32
                                                        : (vert4_out ? 0 : sizeof(float4) * (size_t)B * Vp));
    const size_t h_tinfo = sc.plan(nn_tiles ? sizeof(float4) * 2 * (size_t)B * (T + t->NG) : 0);
    const size_t h_par = sc.plan(want_w ? sizeof(float) * (size_t)B * S * V : 0);
    const size_t h_w = sc.plan((want_w && !winding) ? sizeof(float) * (size_t)B * V : 0);
    const size_t h_any = sc.plan(segs ? (size_t)B : 0);
    const size_t h_am = sc.plan((want_nn && !argmin) ? sizeof(int) * (size_t)B * V : 0);
    const size_t h_mn = sc.plan((want_nn && !min_sq) ? sizeof(float) * (size_t)B * V : 0);
    const size_t h_apex = sc.plan(segs ? sizeof(float) * 3 * (size_t)B * (t->n_bands > 0 ? t->n_bands : 1) : 0);
    const size_t h_wl = sc.plan(segs ? sizeof(int) * ((size_t)B * t->n_sv + 1) : 0);
    const size_t h_todo = sc.plan(qs != nullptr && qs->nn_limit >= 0.f ? sizeof(int) * ((size_t)B * V + 1) : 0);
    if (int rc = sc.commit(st)) return rc;

    float4* strip4 = want_w ? sc.get<float4>(h_tri) : nullptr;
    float4* vert4 = want_nn ? (vert4_out ? vert4_out : sc.get<float4>(h_v4)) : vert4_out;
    if (vert4 != nullptr && !nn_tiles)
        if (int rc = launch_pack_mesh(verts, t->d_faces, B, V, t->F, Fp, Vp, nullptr, vert4, st)) return rc;

    const bool nn_after_flags = nn_tiles && qs != nullptr && qs->nn_limit >= 0.f && exterior != nullptr && want_w &&
                                nearest_limited_supported(T);      // else: the unlimited query, which answers every vertex
    bool nn_launched = false;
    int* am_nn = want_nn ? (argmin ? argmin : sc.get<int>(h_am)) : nullptr;
    float* mn_nn = want_nn ? (min_sq ? min_sq : sc.get<float>(h_mn)) : nullptr;
    if (nn_after_flags)                                   // tile / group spheres: independent of the inside test
        if (int rc = launch_nearest_tiles_pack(verts, t->d_vtile, t->d_vgroup_off, B, V, T, t->NG, vert4, sc.get<float4>(h_tinfo),
                                               st_nn)) return rc;
    if (want_w) {
        float* w = winding ? winding : sc.get<float>(h_w);
        float4* info = sc.get<float4>(h_info);
        if (fast) {
            ClusterJob j{verts, t->d_faces, t->d_leaf_face, t->d_mid_off, t->d_top_off, t->d_vtile, strip4, info,
                         sc.get<float>(h_par), w, sc.get<int>(h_ref), B, V, t->K, t->NM, t->NT, S, t->T};
            j.max_top_leaves = t->max_top_leaves;
            j.stats = t->d_stats;
            if (qs != nullptr && qs->split_trav) {
                if (int rc = launch_cluster_pack(j, st)) return rc;
                TUCH_CUDA(cudaEventRecord(qs->before_trav, st));
                TUCH_CUDA(cudaStreamWaitEvent(qs->trav, qs->before_trav, 0));
                if (int rc = launch_cluster_traverse(j, 0, j.B, qs->trav)) return rc;
                TUCH_CUDA(cudaEventRecord(qs->after_trav, qs->trav));
                TUCH_CUDA(cudaStreamWaitEvent(st, qs->after_trav, 0));
                const bool nn_early = nn_after_flags && st_nn != st;
                if (int rc = launch_cluster_finalize(j, nn_early ? exterior : nullptr, st)) return rc;
                if (nn_early) {
                    // The nearest vertex starts on the flags as the finalize pass knows them -- final outside the
                    // band, "interior" (= searched without a limit, the safe side) for the few queries the exact
                    // re-evaluation still has to decide -- instead of waiting for that re-evaluation and the flag
                    // pass (~60 us of the chain behind the winding kernel).  It reads the flags while the kernels
                    // below rewrite them (same value, or interior -> exterior): either value is fine.
                    TUCH_CUDA(cudaEventRecord(qs->after_ext, st));
                    TUCH_CUDA(cudaStreamWaitEvent(st_nn, qs->after_ext, 0));
                    if (int rc = launch_nearest_tiles_query(t->d_maskP, t->d_maskG, t->d_tile_any, t->d_vtile, t->d_vgroup_off, 0, B, V, T,
                                                            t->NG, vert4, sc.get<float4>(h_tinfo), qs->nn_limit, exterior,
                                                            sc.get<int>(h_todo), am_nn, mn_nn, st_nn)) return rc;
                    nn_launched = true;
                }
                if (int rc = launch_cluster_refine(j, st)) return rc;
            } else {
                if (int rc = launch_winding_clusters(j, st)) return rc;
            }
            if (packed_out != nullptr) {
                packed_out->ctri = strip4; packed_out->nodes = info;
                cluster_pack_betas(j, &packed_out->beta_leaf, &packed_out->beta_group);
            }
        } else {
            if (int rc = launch_pack_strips(verts, B, V, t->d_faces, t->d_strip_vid, t->d_strip_fid, Lp, strip4, info, st)) return rc;
            StripJob j{strip4, info, verts, (long long)V * 3, sc.get<float>(h_par), w, (long long)V, nullptr, B, V, Lp, S};
            if (int rc = launch_winding_strips(j, st)) return rc;
        }
        if (exterior) {
            uint8_t* any = segs ? sc.get<uint8_t>(h_any) : nullptr;
            if (any) TUCH_CUDA(cudaMemsetAsync(any, 0, (size_t)B, st));
            if (int rc = launch_exterior_init(w, B, V, exterior, any, st)) return rc;
            if (nn_after_flags && !nn_launched) {
                // the nearest vertex once the flags exist: interior vertices without a limit, exterior ones within
                // nn_limit.  It reads the flags while the segment pass below may still turn some of them to
                // "exterior": either value is fine, a vertex read as interior merely gets the unlimited answer
                // where the limited one would have done
                if (st_nn != st) {
                    TUCH_CUDA(cudaEventRecord(qs->after_ext, st));
                    TUCH_CUDA(cudaStreamWaitEvent(st_nn, qs->after_ext, 0));
                }
                if (int rc = launch_nearest_tiles_query(t->d_maskP, t->d_maskG, t->d_tile_any, t->d_vtile, t->d_vgroup_off, 0, B, V, T, t->NG,
                                                        vert4, sc.get<float4>(h_tinfo), qs->nn_limit, exterior, sc.get<int>(h_todo),
                                                        am_nn, mn_nn, st_nn)) return rc;
            }
            if (segs) {
                float* apex = sc.get<float>(h_apex);
                if (int rc = launch_segment_apex(verts, B, V, t->d_loop_off, t->d_loop_ids, t->n_bands, apex, any, st)) return rc;
                if (int rc = launch_segment_whitelist(verts, B, V, apex, t->n_bands, t->d_seg_faces, t->d_seg_face_off,
                                                      t->d_seg_band0, t->d_seg_vidx, t->d_member_seg, t->n_sv, exterior,
                                                      any, sc.get<int>(h_wl), st)) return rc;
            }
        }
    }
    if (want_nn && !nn_after_flags) {
        int* am = am_nn;
        float* mn = mn_nn;
        if (nn_tiles) {
            if (int rc = launch_nearest_tiles(verts, t->d_maskP, t->d_maskG, t->d_tile_any, t->d_vtile, t->d_vgroup_off, B, V, T, t->NG, vert4,
                                              sc.get<float4>(h_tinfo), am, mn, st_nn)) return rc;
        } else {
            if (int rc = launch_nearest(vert4, t->d_maskT, B, V, Vp, t->Vq, am, mn, st)) return rc;
        }
    }
    return 0;
}

}  // namespace tuch

TUCH_EXPORT int tuch_contact_query(const tuch_topology* topo, const float* verts, int B, int use_segments,
                                   int32_t* argmin, float* min_sq, float* winding, uint8_t* exterior,
                                   void* stream) {
    return contact_query_impl(topo, verts, B, use_segments, argmin, min_sq, winding, exterior, nullptr,
                              (cudaStream_t)stream);
}

TUCH_EXPORT int tuch_contact_query_within(const tuch_topology* topo, const float* verts, int B, int use_segments,
                                          float radius, int32_t* argmin, float* min_sq, float* winding, uint8_t* exterior,
                                          void* stream) {
    TUCH_REQUIRE(radius >= 0.f, "tuch_contact_query_within: negative radius");
    TUCH_REQUIRE(exterior != nullptr && (argmin != nullptr || min_sq != nullptr),
                 "tuch_contact_query_within: needs the exterior output and a nearest-vertex output");
    TUCH_REQUIRE(topo != nullptr && topo->has_maskP, "tuch_contact_query_within: the topology needs a geodesic mask and vertex tiles");
    QueryStreams qs;
    qs.nn = (cudaStream_t)stream;
    qs.nn_limit = radius;
    return contact_query_impl(topo, verts, B, use_segments, argmin, min_sq, winding, exterior, nullptr, (cudaStream_t)stream,
                              nullptr, &qs);
}

TUCH_EXPORT int tuch_segment_exterior(const tuch_topology* t, const float* verts, int B, uint8_t* out,
                                      float* winding_out, void* stream) {
    TUCH_REQUIRE(t != nullptr, "tuch_segment_exterior: null topology");
    TUCH_REQUIRE(B >= 0, "tuch_segment_exterior: negative batch");
    if (B == 0 || t->n_segments == 0) return 0;
    TUCH_REQUIRE(verts && out, "tuch_segment_exterior: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc;
    SegPlan sp;
    plan_segments(t, B, sc, sp);
    if (int rc = sc.commit(st)) return rc;
    return run_segments(t, verts, B, sc, sp, nullptr, nullptr, out, winding_out, st);
}

TUCH_EXPORT int tuch_region_min(const tuch_topology* t, const float* verts, int B, int masked,
                                const uint8_t* active, float* min_sq, int32_t* arg_i, int32_t* arg_j,
                                void* stream) {
    TUCH_REQUIRE(t != nullptr, "tuch_region_min: null topology");
    TUCH_REQUIRE(B >= 0, "tuch_region_min: negative batch");
    if (B == 0 || t->n_pairs == 0) return 0;
    TUCH_REQUIRE(verts && min_sq, "tuch_region_min: null pointer");
    TUCH_REQUIRE(!masked || t->has_mask, "tuch_region_min: masked minimum requested but no geodesic mask is set");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc;
    const size_t h_v4 = sc.plan(sizeof(float4) * (size_t)B * t->Vp);
    const size_t h_i = sc.plan(arg_i ? 0 : sizeof(int) * (size_t)B * t->n_pairs);
    const size_t h_j = sc.plan(arg_j ? 0 : sizeof(int) * (size_t)B * t->n_pairs);
    if (int rc = sc.commit(st)) return rc;
    float4* v4 = sc.get<float4>(h_v4);
    if (int rc = launch_pack_mesh(verts, t->d_faces, B, t->V, t->F, t->Fp, t->Vp, nullptr, v4, st)) return rc;
    return launch_region_min(v4, t->Vp, masked ? t->d_maskT : nullptr, t->Vq, t->d_region_ids, t->d_region_off,
                             t->d_pair_a, t->d_pair_b, active, t->n_pairs, B,
                             t->has_pair_mask ? t->d_pair_mask : nullptr, t->d_pair_word_off, min_sq,
                             arg_i ? arg_i : sc.get<int>(h_i), arg_j ? arg_j : sc.get<int>(h_j), st);
}

// ================================================================ host-buffer conveniences
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { TUCH_CUDA(cudaMalloc(&p, n ? n : 1)); return 0; }
    template <typename T> T* as() { return (T*)p; }
};
}  // namespace

TUCH_EXPORT int tuch_winding_numbers_host(const float* points_host, const float* triangles_host, int bs, int Q,
                                          int F, float* out_host) {
    TUCH_REQUIRE(bs >= 0 && Q >= 0 && F >= 0, "tuch_winding_numbers_host: negative size");
    if (bs == 0 || Q == 0) return 0;
    TUCH_REQUIRE(points_host && out_host && (F == 0 || triangles_host), "tuch_winding_numbers_host: null pointer");
    DevBuf p, t, o;
    const size_t np = sizeof(float) * 3 * (size_t)bs * Q, nt = sizeof(float) * 9 * (size_t)bs * F,
                 no = sizeof(float) * (size_t)bs * Q;
    if (int rc = p.alloc(np)) return rc;
    if (int rc = t.alloc(nt)) return rc;
    if (int rc = o.alloc(no)) return rc;
    TUCH_CUDA(cudaMemcpyAsync(p.p, points_host, np, cudaMemcpyHostToDevice, 0));
    if (nt) TUCH_CUDA(cudaMemcpyAsync(t.p, triangles_host, nt, cudaMemcpyHostToDevice, 0));
    if (int rc = tuch_winding_numbers(p.as<float>(), t.as<float>(), bs, Q, F, o.as<float>(), nullptr)) return rc;
    TUCH_CUDA(cudaMemcpyAsync(out_host, o.p, no, cudaMemcpyDeviceToHost, 0));
    TUCH_CUDA(cudaStreamSynchronize(0));
    return 0;
}

TUCH_EXPORT int tuch_contact_query_host(const tuch_topology* topo, const float* verts_host, int B, int use_segments,
                                        int32_t* argmin_host, float* min_sq_host, float* winding_host,
                                        uint8_t* exterior_host) {
    TUCH_REQUIRE(topo != nullptr, "tuch_contact_query_host: null topology");
    TUCH_REQUIRE(B >= 0, "tuch_contact_query_host: negative batch");
    if (B == 0) return 0;
    TUCH_REQUIRE(verts_host != nullptr, "tuch_contact_query_host: verts is null");
    const size_t n = (size_t)B * topo->V;
    DevBuf v, am, mn, w, e;
    if (int rc = v.alloc(n * 12)) return rc;
    if (argmin_host) if (int rc = am.alloc(n * 4)) return rc;
    if (min_sq_host) if (int rc = mn.alloc(n * 4)) return rc;
    if (winding_host) if (int rc = w.alloc(n * 4)) return rc;
    if (exterior_host) if (int rc = e.alloc(n)) return rc;
    TUCH_CUDA(cudaMemcpyAsync(v.p, verts_host, n * 12, cudaMemcpyHostToDevice, 0));
    if (int rc = tuch_contact_query(topo, v.as<float>(), B, use_segments, am.as<int32_t>(), mn.as<float>(),
                                    w.as<float>(), e.as<uint8_t>(), nullptr)) return rc;
    if (argmin_host) TUCH_CUDA(cudaMemcpyAsync(argmin_host, am.p, n * 4, cudaMemcpyDeviceToHost, 0));
    if (min_sq_host) TUCH_CUDA(cudaMemcpyAsync(min_sq_host, mn.p, n * 4, cudaMemcpyDeviceToHost, 0));
    if (winding_host) TUCH_CUDA(cudaMemcpyAsync(winding_host, w.p, n * 4, cudaMemcpyDeviceToHost, 0));
    if (exterior_host) TUCH_CUDA(cudaMemcpyAsync(exterior_host, e.p, n, cudaMemcpyDeviceToHost, 0));
    TUCH_CUDA(cudaStreamSynchronize(0));
    return 0;
}
