// C ABI: one whole SMPLify-DC stage-2 iteration (tuch/smplify/smplifydc.py:155-183) as ONE call.
//
//   SMPL forward -> contact_fitting_loss (losses.py:34-123) -> backward -> torch.optim.Adam on [body_pose, global_orient]
//
// The reference spends ~80 ATen launches per BODY on it plus autograd's bookkeeping and Adam's per-parameter loop.
// Here the whole batch is ~27 kernel launches on the caller's stream, no host synchronisation and nothing but our
// kernels: the values and analytic gradients of every term come out of the same kernels the per-term entry points
// use (so the iteration is bit-identical to composing those), the vertex / joint gradients go straight into the
// LBS backward, and Adam runs in the epilogue of its last kernel.  Scratch comes from the grow-only arenas, so
// after one warm-up call the step can be captured into a CUDA graph.
#include "api_internal.h"
#include "objective_internal.h"
#include "smpl_internal.h"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

using namespace tuch;

namespace {
// Library-owned streams of one iteration, per (device, caller's stream).  The iteration is two big kernels -- the
// hierarchical winding kernel and the masked nearest vertex, ~80 % of the work -- and ~25 smaller ones before,
// between and after them:
//   hi   the chain (SMPL forward, hierarchy pack, finalize .. segment pass, losses, SMPL backward)
//   s2   the joint-side terms (output joints, reprojection, pose prior, region minima; the per-joint reduction of
//        the SMPL backward)
//   mid  the winding kernel
//   lo   the masked nearest vertex when it is searched without a limit (small batches); the limited search of
//        larger batches runs behind the winding kernel on `mid`
// The caller's stream only forks into `hi` at the start and joins it at the end.  Events and waits are capturable:
// in a CUDA graph the forks / joins become plain dependencies and the kernel nodes keep their stream's priority.
//
// What the CUPTI timeline shows (scripts/diag/timeline.py, 256 bodies): the block scheduler hands out the CTAs of
// equal-priority kernels in launch order, so the two big kernels run one after the other and only overlap at their
// tails, and a small kernel launched behind a big one on another stream waits until that kernel's whole grid has
// been DISPATCHED.  The iteration is bound by the sum of its kernels' work, not by exposed latency: giving the
// chain priority moves its kernels forward but slows the nearest-vertex kernel by as much.  Measured with the
// unlimited nearest-vertex query, ms per iteration, all streams at one priority / hi = s2 = mid above lo: 8 bodies
// 0.388 / 0.390, 32: 0.753 / 0.733, 64: 1.179 / 1.130, 128: 1.955 / 1.991, 256: 3.607 / 3.690; with the limited
// query: 64: 1.066 / 1.033, 128: 1.691 / 1.600, 256: 2.937 / 2.894 (more than two levels change nothing).  The
// streams are prioritised at every batch size.
// Also tried at 256 bodies: the nearest-vertex query in two launches, one ahead of the hierarchy pack and one behind
// the winding kernel, so that the pack is overlapped too: 3.52 against 3.53 ms eager, no gain (the pack kernel's
// CTAs fill the register file, nothing co-resides with them).
struct Side {
    cudaStream_t hi = nullptr, mid = nullptr, lo = nullptr, s2 = nullptr;
    cudaEvent_t begin = nullptr, fork = nullptr, before_trav = nullptr, after_trav = nullptr, join1 = nullptr,
                join2 = nullptr, fork_b = nullptr, join_b = nullptr, done = nullptr, after_ext = nullptr;
};
constexpr int FIT_NN_LIMIT_FROM = 48;
std::mutex g_side_mu;
std::map<std::tuple<int, cudaStream_t, bool>, Side> g_side;

int side_streams(cudaStream_t caller, bool prioritised, Side** out) {
    int dev = 0;
    TUCH_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_side_mu);
    Side& s = g_side[std::make_tuple(dev, caller, prioritised)];
    if (s.hi == nullptr) {
        int least = 0, greatest = 0;
        TUCH_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));     // numerically lower = higher priority
        // TUCH_FIT_PRIORITIES="h,m,l": levels above the default priority for hi (= s2) / mid / lo, at every batch
        // size (A/B measurements)
        int lv[3] = {prioritised ? 1 : 0, prioritised ? 1 : 0, 0};
        if (const char* e = getenv("TUCH_FIT_PRIORITIES")) sscanf(e, "%d,%d,%d", &lv[0], &lv[1], &lv[2]);
        auto level = [&](int l) { return least - l < greatest ? greatest : least - l; };
        const int p_hi = level(lv[0]), p_mid = level(lv[1]), p_lo = level(lv[2]);
        TUCH_CUDA(cudaStreamCreateWithPriority(&s.hi, cudaStreamNonBlocking, p_hi));
        TUCH_CUDA(cudaStreamCreateWithPriority(&s.s2, cudaStreamNonBlocking, p_hi));
        TUCH_CUDA(cudaStreamCreateWithPriority(&s.mid, cudaStreamNonBlocking, p_mid));
        TUCH_CUDA(cudaStreamCreateWithPriority(&s.lo, cudaStreamNonBlocking, p_lo));
        for (cudaEvent_t* e : {&s.begin, &s.fork, &s.before_trav, &s.after_trav, &s.join1, &s.join2, &s.fork_b, &s.join_b,
                               &s.done, &s.after_ext})
            TUCH_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    *out = &s;
    return 0;
}
}  // namespace

TUCH_EXPORT int tuch_contact_fit_step(const tuch_smpl* smpl, const tuch_topology* topo, const tuch_prior* prior, int B,
                                      const tuch_contact_fit_args* a, void* stream) {
    TUCH_REQUIRE(smpl != nullptr && topo != nullptr && a != nullptr, "tuch_contact_fit_step: null handle");
    TUCH_REQUIRE(B >= 0, "tuch_contact_fit_step: negative batch");
    if (B == 0) return 0;
    TUCH_REQUIRE(B <= 65535, "tuch_contact_fit_step: at most 65535 bodies per call, got %d", B);
    const SmplDev& m = smpl->dev;
    const int V = m.V, J = m.NO, P = topo->n_pairs;
    TUCH_REQUIRE(topo->V == V, "tuch_contact_fit_step: the topology has %d vertices, the body model %d", topo->V, V);
    TUCH_REQUIRE(topo->has_mask && topo->F > 0, "tuch_contact_fit_step: the topology needs faces and a geodesic mask");
    TUCH_REQUIRE(a->body_pose && a->global_orient && a->exp_avg_pose && a->exp_avg_sq_pose && a->exp_avg_orient &&
                     a->exp_avg_sq_orient && a->step_pose && a->step_orient,
                 "tuch_contact_fit_step: null parameter / optimiser-state pointer");
    TUCH_REQUIRE(a->betas && a->camera_t && a->camera_center && a->joints_2d && a->joints_conf,
                 "tuch_contact_fit_step: null input pointer");
    TUCH_REQUIRE(a->smpl_workspace && a->vertices && a->joints && a->loss, "tuch_contact_fit_step: null output pointer");
    TUCH_REQUIRE(((uintptr_t)a->smpl_workspace & 15) == 0, "tuch_contact_fit_step: the SMPL workspace must be 16-byte aligned");
    TUCH_REQUIRE(prior == nullptr || prior->D == 69, "tuch_contact_fit_step: the pose prior must be over 69 pose entries");
    TUCH_REQUIRE(a->pair_active == nullptr || P > 0, "tuch_contact_fit_step: pair_active given but the topology has no region pairs");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t BV = (size_t)B * V;

    Scratch sc;
    const size_t h_rep = sc.plan(sizeof(float) * (size_t)B * J), h_gj = sc.plan(sizeof(float) * 3 * (size_t)B * J);
    const size_t h_val = sc.plan(sizeof(float) * B), h_pv = sc.plan(sizeof(float) * B), h_cmp = sc.plan(sizeof(int) * B);
    const size_t h_gp = sc.plan(sizeof(float) * (size_t)B * 69);
    const size_t h_con = sc.plan(sizeof(float) * B), h_r2r = sc.plan(sizeof(float) * B);
    const size_t h_mn = sc.plan(sizeof(float) * (size_t)B * (P > 0 ? P : 1));
    const size_t h_ai = sc.plan(sizeof(int) * (size_t)B * (P > 0 ? P : 1)), h_aj = sc.plan(sizeof(int) * (size_t)B * (P > 0 ? P : 1));
    const size_t h_gv = sc.plan(sizeof(float) * 3 * BV);
    const size_t h_am = sc.plan(a->argmin ? 0 : sizeof(int) * BV), h_ex = sc.plan(a->exterior ? 0 : BV);
    const size_t h_v4 = sc.plan(a->pair_active ? sizeof(float4) * (size_t)B * topo->Vp : 0);
    if (int rc = sc.commit_slot(st, 4)) return rc;
    float* rep = sc.get<float>(h_rep);
    float* g_joints = sc.get<float>(h_gj);
    float* g_prior = sc.get<float>(h_gp);
    float* contact = sc.get<float>(h_con);
    float* r2r = sc.get<float>(h_r2r);
    float* g_verts = sc.get<float>(h_gv);
    int* am = a->argmin ? a->argmin : sc.get<int>(h_am);
    uint8_t* ext = a->exterior ? a->exterior : sc.get<uint8_t>(h_ex);

    // TUCH_FIT_STREAMS=0 keeps the whole iteration on the caller's stream (A/B measurements)
    static const bool one_stream = getenv("TUCH_FIT_STREAMS") != nullptr && atoi(getenv("TUCH_FIT_STREAMS")) == 0;
    Side* side = nullptr;
    if (!one_stream) if (int rc = side_streams((cudaStream_t)stream, true, &side)) return rc;
    const cudaStream_t caller = (cudaStream_t)stream;
    if (side) {                                  // from here on `st` is the high-priority chain
        TUCH_CUDA(cudaEventRecord(side->begin, caller));
        TUCH_CUDA(cudaStreamWaitEvent(side->hi, side->begin, 0));
        st = side->hi;
    }
    // The contact term (launch_contact_loss below, PULL_THRESHOLD) reads the nearest allowed vertex of interior
    // vertices and of exterior vertices closer than euclthres only: the query is limited accordingly (nn_limit),
    // which leaves loss, gradients and parameters bit-identical (tests/test_objective_gpu.py compares with the
    // composition over the unlimited query) at a fifth of the nearest-vertex work.
    // (Below FIT_NN_LIMIT_FROM bodies the unlimited query hides behind the inside test anyway and a search behind
    // the flags would only lengthen the chain: 32 bodies 0.687 against 0.716 ms.)
    static const bool nn_full = getenv("TUCH_FIT_NN_FULL") != nullptr && atoi(getenv("TUCH_FIT_NN_FULL")) != 0;   // A/B
    static const int nn_from = getenv("TUCH_FIT_NN_FROM") != nullptr ? atoi(getenv("TUCH_FIT_NN_FROM")) : FIT_NN_LIMIT_FROM;
    const bool nn_limited = !nn_full && topo->has_maskP && B >= nn_from;
    // the limited search runs behind the winding kernel, on its stream: at the chain's priority, so that it is not
    // locked out by the segment pass, whose CTAs fill the register file (TUCH_FIT_NN_LO=1: the low-priority stream)
    static const bool nn_lo = getenv("TUCH_FIT_NN_LO") != nullptr && atoi(getenv("TUCH_FIT_NN_LO")) != 0;
    const cudaStream_t s_nn = side ? (nn_limited && !nn_lo ? side->mid : side->lo) : st, s2 = side ? side->s2 : st;

    // ---- SMPL forward (split pose; advances the Adam step counters); the output joints follow on side 2
    LbsBuffers w;
    lbs_carve(a->smpl_workspace, B, V, m.L, w);
    if (int rc = launch_lbs_forward(m, a->betas, a->body_pose, 0, B, w, a->vertices, nullptr, st, a->global_orient,
                                    a->step_pose, a->step_orient)) return rc;
    if (side) {
        TUCH_CUDA(cudaEventRecord(side->fork, st));
        TUCH_CUDA(cudaStreamWaitEvent(s_nn, side->fork, 0));
        TUCH_CUDA(cudaStreamWaitEvent(s2, side->fork, 0));
    }
    // ---- side 2: joints, losses.py:56-64 (reprojection, pose prior) and the region minima of :108-117
    if (int rc = launch_lbs_joints(m, a->vertices, w, B, a->joints, s2)) return rc;
    if (int rc = launch_reprojection(a->joints, a->camera_t, a->camera_center, a->joints_2d, a->joints_conf, B, J,
                                     a->focal_length, a->sigma, nullptr, 0.f, nullptr, rep, nullptr, g_joints, nullptr, s2)) return rc;
    const float wp = a->pose_prior_weight * a->pose_prior_weight;
    const bool with_prior = prior != nullptr && wp != 0.f;
    if (with_prior)
        if (int rc = launch_pose_terms(prior->d_means, prior->d_precisions, prior->d_nll_weights, prior->M, 69, a->body_pose,
                                       nullptr, 0, B, wp, 0.f, 0.f, sc.get<float>(h_val), sc.get<float>(h_pv),
                                       sc.get<int>(h_cmp), g_prior, nullptr, s2)) return rc;
    const bool with_r2r = a->pair_active != nullptr && P > 0;
    float* mn = sc.get<float>(h_mn);
    int* ai = sc.get<int>(h_ai);
    int* aj = sc.get<int>(h_aj);
    if (with_r2r) {
        float4* v4 = sc.get<float4>(h_v4);
        TUCH_CUDA(cudaMemsetAsync(mn, 0, sizeof(float) * (size_t)B * P, s2));
        TUCH_CUDA(cudaMemsetAsync(ai, 0xff, sizeof(int) * (size_t)B * P, s2));
        TUCH_CUDA(cudaMemsetAsync(aj, 0xff, sizeof(int) * (size_t)B * P, s2));
        if (int rc = launch_pack_mesh(a->vertices, topo->d_faces, B, V, topo->F, topo->Fp, topo->Vp, nullptr, v4, s2)) return rc;
        if (int rc = launch_region_min(v4, topo->Vp, topo->d_maskT, topo->Vq, topo->d_region_ids, topo->d_region_off,
                                       topo->d_pair_a, topo->d_pair_b, a->pair_active, P, B,
                                       topo->has_pair_mask ? topo->d_pair_mask : nullptr, topo->d_pair_word_off, mn, ai, aj, s2)) return rc;
    }
    if (side) TUCH_CUDA(cudaEventRecord(side->join2, s2));
    // ---- losses.py:73-105: inside test + allowed self-intersections (winding kernel on `mid`, the rest on the
    //      chain), masked nearest vertex (`lo`)
    QueryStreams qs;
    qs.nn = st;
    if (nn_limited) qs.nn_limit = a->euclthres > 0.f ? a->euclthres : 0.f;
    if (side) {
        qs.nn = s_nn; qs.trav = side->mid; qs.split_trav = true;
        qs.before_trav = side->before_trav; qs.after_trav = side->after_trav; qs.after_ext = side->after_ext;
    }
    if (int rc = contact_query_impl(topo, a->vertices, B, a->use_segments, am, nullptr, nullptr, ext, nullptr, st, nullptr,
                                    &qs)) return rc;
    if (side) TUCH_CUDA(cudaEventRecord(side->join1, s_nn));
    TUCH_CUDA(cudaMemsetAsync(g_verts, 0, sizeof(float) * 3 * BV, st));
    if (side) TUCH_CUDA(cudaStreamWaitEvent(st, side->join1, 0));
    if (int rc = launch_contact_loss(a->vertices, am, ext, a->body_active, nullptr, B, V, a->euclthres, PULL_THRESHOLD,
                                     REDUCE_SUM, 10.f, nullptr, contact, nullptr, g_verts, st)) return rc;
    if (side) TUCH_CUDA(cudaStreamWaitEvent(st, side->join2, 0));
    if (with_r2r)
        if (int rc = launch_region_sum(a->vertices, B, V, P, mn, ai, aj, a->body_active, a->contact_loss_weight, nullptr, r2r,
                                       g_verts, st)) return rc;
    // ---- losses.py:120-123: per-body totals and their sum
    if (int rc = launch_combine(rep, J, with_prior ? sc.get<float>(h_val) : nullptr, contact, 10.f, with_r2r ? r2r : nullptr,
                                a->contact_loss_weight, nullptr, B, a->per_body, a->loss, st)) return rc;
    // ---- backward through the SMPL forward; Adam on [body_pose, global_orient] in its last kernel
    LbsAdam ad;
    ad.body_pose = a->body_pose; ad.global_orient = a->global_orient;
    ad.m_pose = a->exp_avg_pose; ad.v_pose = a->exp_avg_sq_pose; ad.m_orient = a->exp_avg_orient; ad.v_orient = a->exp_avg_sq_orient;
    ad.step_pose = a->step_pose; ad.step_orient = a->step_orient;
    ad.g_extra_pose = with_prior ? g_prior : nullptr;
    ad.g_out_pose = a->grad_body_pose; ad.g_out_orient = a->grad_global_orient;
    ad.lr = a->lr; ad.beta1 = a->beta1; ad.beta2 = a->beta2; ad.eps = a->eps;
    TUCH_REQUIRE((a->grad_body_pose == nullptr) == (a->grad_global_orient == nullptr),
                 "tuch_contact_fit_step: give both gradient outputs or neither");
    LbsSide bs{s2, side ? side->fork_b : nullptr, side ? side->join_b : nullptr};
    if (int rc = launch_lbs_backward(m, a->body_pose, 0, B, w, g_verts, g_joints, nullptr, nullptr, st, &ad,
                                     side ? &bs : nullptr)) return rc;
    if (side) {
        TUCH_CUDA(cudaEventRecord(side->done, st));
        TUCH_CUDA(cudaStreamWaitEvent(caller, side->done, 0));
    }
    return 0;
}
