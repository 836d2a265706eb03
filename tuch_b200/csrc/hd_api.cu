// C ABI of the regressor contact loss (include/tuch_b200.h, section a12): tuch/train/loss.py:240-317.
#include "hd_internal.h"
#include "clusters.h"
#include "objective_internal.h"
#include "strips.h"

using namespace tuch;

TUCH_EXPORT int tuch_topology_set_hd(tuch_topology* t, int n_hd, const int32_t* row_offsets_host,
                                     const int32_t* cols_host, const float* vals_host, const int32_t* hd_face_host) {
    TUCH_REQUIRE(t != nullptr, "tuch_topology_set_hd: null topology");
    TUCH_REQUIRE(n_hd >= 0, "tuch_topology_set_hd: negative count");
    auto freep = [](void* p) { if (p) cudaFree(p); };
    freep(t->d_hd_row_off); freep(t->d_hd_cols); freep(t->d_hd_vals); freep(t->d_hd_face);
    t->d_hd_row_off = t->d_hd_cols = t->d_hd_face = nullptr; t->d_hd_vals = nullptr; t->n_hd = 0;
    if (n_hd == 0) return 0;
    TUCH_REQUIRE(row_offsets_host && cols_host && vals_host && hd_face_host, "tuch_topology_set_hd: null array");
    TUCH_REQUIRE(row_offsets_host[0] == 0, "tuch_topology_set_hd: row_offsets[0] must be 0");
    for (int k = 0; k < n_hd; ++k) {
        TUCH_REQUIRE(row_offsets_host[k + 1] >= row_offsets_host[k], "tuch_topology_set_hd: row offsets must not decrease");
        TUCH_REQUIRE(hd_face_host[k] >= 0 && hd_face_host[k] < t->F, "tuch_topology_set_hd: source face %d out of range", hd_face_host[k]);
    }
    const int nnz = row_offsets_host[n_hd];
    for (int e = 0; e < nnz; ++e)
        TUCH_REQUIRE(cols_host[e] >= 0 && cols_host[e] < t->V, "tuch_topology_set_hd: column %d out of range", cols_host[e]);
    auto up = [](const void* h, size_t bytes, void** d) -> int {
        TUCH_CUDA(cudaMalloc(d, bytes ? bytes : 4));
        if (bytes) TUCH_CUDA(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
        return 0;
    };
    if (int rc = up(row_offsets_host, sizeof(int) * ((size_t)n_hd + 1), (void**)&t->d_hd_row_off)) return rc;
    if (int rc = up(cols_host, sizeof(int) * (size_t)nnz, (void**)&t->d_hd_cols)) return rc;
    if (int rc = up(vals_host, sizeof(float) * (size_t)nnz, (void**)&t->d_hd_vals)) return rc;
    if (int rc = up(hd_face_host, sizeof(int) * (size_t)n_hd, (void**)&t->d_hd_face)) return rc;
    t->n_hd = n_hd;
    return 0;
}

TUCH_EXPORT int tuch_topology_num_hd(const tuch_topology* t) { return t ? t->n_hd : -1; }

TUCH_EXPORT int tuch_regressor_contact_loss(const tuch_topology* t, const float* verts, int B, const uint8_t* valid,
                                            float euclthres, int use_hd, float weight, const float* g_loss,
                                            float* loss, float* g_verts, int32_t* counts_out, int32_t* sel_out,
                                            int32_t* hd_argmin_out, uint8_t* hd_exterior_out, void* stream) {
    TUCH_REQUIRE(t != nullptr, "tuch_regressor_contact_loss: null topology");
    TUCH_REQUIRE(B >= 0, "tuch_regressor_contact_loss: negative batch");
    if (B == 0) return 0;
    TUCH_REQUIRE(verts && loss, "tuch_regressor_contact_loss: null pointer");
    TUCH_REQUIRE(t->has_mask && t->F > 0, "tuch_regressor_contact_loss: the topology needs faces and a geodesic mask");
    TUCH_REQUIRE(!use_hd || t->n_hd > 0, "tuch_regressor_contact_loss: use_hd requested but no HD regressor is set "
                                         "(tuch_topology_set_hd)");
    cudaStream_t st = (cudaStream_t)stream;
    const int V = t->V, N = use_hd ? t->n_hd : 0;
    const size_t BV = (size_t)B * V, BN = (size_t)B * N;

    // the caller-visible scratch of this entry point lives in its own arena slot: contact_query_impl
    // commits the shared per-stream arena, so everything that must survive it is planned FIRST there
    // through `reserve` (see contact_query_impl)
    Scratch sc;
    const size_t h_am = sc.plan(sizeof(int) * BV), h_mn = sc.plan(sizeof(float) * BV), h_ex = sc.plan(BV);
    const size_t h_idx = sc.plan(sizeof(int) * BN), h_cnt = sc.plan(sizeof(int) * (size_t)B);
    const size_t h_hd4 = sc.plan(sizeof(float4) * BN), h_hd = sc.plan(sizeof(float) * 3 * BN);
    const size_t h_off = sc.plan(sizeof(float) * 3 * BN), h_px = sc.plan(sizeof(int) * BN);
    const size_t h_ham = sc.plan(sizeof(int) * BN), h_hw = sc.plan(sizeof(float) * BN), h_hex = sc.plan(BN);
    const size_t h_ghd = sc.plan(sizeof(float) * 3 * BN);
    const int Lp = t->Lp;
    // inside test of the HD points: hierarchical far field when the topology has its face hierarchy
    // (the default), else the all-faces strip kernel
    const bool fast = use_hd && t->winding_mode == TUCH_WINDING_FAST && t->has_clusters;
    const int S = !use_hd ? 1 : fast ? cluster_splits(B, cdiv(N, 32), t->NT, sm_count()) : strip_splits(B, N, Lp, sm_count());
    // hierarchical: the point query re-uses the hierarchy the vertex query below packs for this batch
    const size_t h_tri = sc.plan(!use_hd || fast ? 0 : sizeof(float4) * 2 * (size_t)B * Lp);
    const size_t h_info = sc.plan(!use_hd || fast ? 0 : sizeof(float4) * (size_t)B * (Lp / WS_TILE));
    const size_t h_par = sc.plan(use_hd ? sizeof(float) * (size_t)B * S * N : 0);
    const size_t h_ref = sc.plan(fast ? sizeof(int) * (BN + 1) : 0);
    if (int rc = sc.commit_slot(st, 1)) return rc;
    int* am = sc.get<int>(h_am);
    float* mn = sc.get<float>(h_mn);
    uint8_t* ex = sc.get<uint8_t>(h_ex);

    // loss.py:251-270 -- exterior flags (winding + segment whitelist) and the masked nearest vertex
    // With HD points the vertices' nearest vertex only decides which of them are "in contact" (min_sq < euclthres^2,
    // hd_select): the search is limited to that radius; without, every exterior vertex is pulled (PULL_ALL below)
    // and the search is unlimited.
    PackedClusters packed;
    QueryStreams qs;
    qs.nn = st;
    if (use_hd && t->has_maskP && euclthres >= 0.f) qs.nn_limit = euclthres;
    if (int rc = contact_query_impl(t, verts, B, 1, am, mn, nullptr, ex, nullptr, st, &packed, &qs)) return rc;

    if (!use_hd) {                                                    // loss.py:303-315
        return launch_contact_loss(verts, am, ex, valid, nullptr, B, V, 0.f, PULL_ALL, REDUCE_SUM, weight, g_loss,
                                   loss, nullptr, g_verts, st);
    }
    int* idx = sel_out ? sel_out : sc.get<int>(h_idx);
    int* cnt = counts_out ? counts_out : sc.get<int>(h_cnt);
    float4* hd4 = sc.get<float4>(h_hd4);
    float* hd = sc.get<float>(h_hd);
    float* off = sc.get<float>(h_off);
    int* proxy = sc.get<int>(h_px);
    int* ham = hd_argmin_out ? hd_argmin_out : sc.get<int>(h_ham);
    float* hw = sc.get<float>(h_hw);
    uint8_t* hex = hd_exterior_out ? hd_exterior_out : sc.get<uint8_t>(h_hex);
    if (int rc = launch_hd_select(mn, ex, valid, t->d_faces, B, V, N, t->d_hd_face, euclthres * euclthres, idx, cnt, st)) return rc;
    if (int rc = launch_hd_gather(verts, B, V, N, idx, cnt, t->d_hd_row_off, t->d_hd_cols, t->d_hd_vals, t->d_hd_face,
                                  t->d_faces, hd4, hd, off, proxy, st)) return rc;
    if (int rc = launch_hd_nearest(hd4, proxy, cnt, B, N, t->d_maskT, t->Vq, ham, st)) return rc;
    // loss.py:297 -- inside test of the offset HD points against the full mesh
    if (fast) {
        // nothing between the vertex query and here commits the shared arena, so its packed hierarchy is intact
        TUCH_REQUIRE(packed.ctri != nullptr && packed.nodes != nullptr,
                     "tuch_regressor_contact_loss: the vertex query did not leave a packed face hierarchy");
        ClusterJob j{verts, t->d_faces, t->d_leaf_face, t->d_mid_off, t->d_top_off, nullptr,
                     const_cast<float4*>(packed.ctri), const_cast<float4*>(packed.nodes),
                     sc.get<float>(h_par), hw, sc.get<int>(h_ref), B, V, t->K, t->NM, t->NT, S, t->T};
        j.points = off; j.Q = N; j.q_counts = cnt; j.body_active = valid; j.max_top_leaves = t->max_top_leaves;
        j.beta_leaf = WC_BETA_POINTS; j.beta_group = WC_BETA_GROUP_POINTS; j.margin = WC_MARGIN_POINTS;
        j.packed_beta_leaf = packed.beta_leaf; j.packed_beta_group = packed.beta_group;
        j.stats = t->d_stats + 1;
        if (int rc = launch_cluster_query(j, st)) return rc;
    } else {
        float4* strip4 = sc.get<float4>(h_tri);
        float4* info = sc.get<float4>(h_info);
        if (int rc = launch_pack_strips(verts, B, V, t->d_faces, t->d_strip_vid, t->d_strip_fid, Lp, strip4, info, st)) return rc;
        StripJob j{strip4, info, off, (long long)N * 3, sc.get<float>(h_par), hw, (long long)N, valid, B, N, Lp, S};
        j.q_counts = cnt;
        if (int rc = launch_winding_strips(j, st)) return rc;
    }
    if (int rc = launch_exterior_init(hw, B, N, hex, nullptr, st)) return rc;
    // loss.py:299-315
    float* g_hd = nullptr;
    if (g_verts != nullptr) {
        g_hd = sc.get<float>(h_ghd);
        TUCH_CUDA(cudaMemsetAsync(g_hd, 0, sizeof(float) * 3 * BN, st));
    }
    if (int rc = launch_contact_loss(hd, ham, hex, valid, cnt, B, N, 0.f, PULL_ALL, REDUCE_SUM, weight, g_loss, loss,
                                     nullptr, g_hd, st)) return rc;
    if (g_verts != nullptr)
        if (int rc = launch_hd_scatter(g_hd, B, V, N, idx, cnt, t->d_hd_row_off, t->d_hd_cols, t->d_hd_vals, g_verts, st)) return rc;
    return 0;
}
