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

#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

using namespace tuch;

namespace {
// Two library-owned side streams per device.  Inside one iteration three chains are independent once the vertices
// exist: the inside test (face hierarchy pack -> winding numbers -> exact re-evaluation -> segment whitelist), the
// masked nearest vertex, and the joint-side terms (output joints, reprojection, pose prior, region minima).  The
// inside test is the critical path, so it runs on a HIGH-PRIORITY stream (`s1`): its CTAs are dispatched first and
// the nearest-vertex kernel on the caller's stream fills whatever the winding kernel leaves idle; the joint-side
// terms run on `s2`.  They fork after the LBS forward and join before the losses that consume them.  At small
// batches, where no single kernel fills the GPU, this is ~25 % of the iteration.  Events and waits are capturable:
// in a CUDA graph the fork / join become plain dependencies and the kernel nodes keep their stream's priority.
// At large batches both big kernels fill the GPU on their own and what overlap buys is co-residency (the
// issue-bound winding kernel and the latency-bound nearest kernel share SMs better than either does alone): there
// the two run at EQUAL priority (inside test on the caller's stream, nearest vertex on `s0`).
struct Side {
    cudaStream_t s0 = nullptr, s1 = nullptr, s2 = nullptr;
    cudaEvent_t fork = nullptr, join1 = nullptr, join2 = nullptr, fork_b = nullptr, join_b = nullptr;
};
constexpr int FIT_PRIORITY_BELOW = 192;      // bodies: below this the inside test gets the high-priority stream
std::mutex g_side_mu;
// one set per (device, caller's stream): fits that run concurrently on different streams must not meet on a shared
// side stream.  (Tried, round 2: a batch of 256 as two concurrent fits of 128 on two streams, so that one half's
// latency-bound head and tail overlap the other half's big kernels: 3.77 ms against 3.65 for the one fit, four fits
// of 64: 3.93 -- the iteration is bound by its total instruction count, scripts/diag/two_halves.py.)
std::map<std::pair<int, cudaStream_t>, Side> g_side;

int side_streams(cudaStream_t caller, Side** out) {
    int dev = 0;
    TUCH_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_side_mu);
    Side& s = g_side[std::make_pair(dev, caller)];
    if (s.s1 == nullptr) {
        int least = 0, greatest = 0;
        TUCH_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        TUCH_CUDA(cudaStreamCreateWithFlags(&s.s0, cudaStreamNonBlocking));
        TUCH_CUDA(cudaStreamCreateWithPriority(&s.s1, cudaStreamNonBlocking, greatest));
        TUCH_CUDA(cudaStreamCreateWithFlags(&s.s2, cudaStreamNonBlocking));
        TUCH_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        TUCH_CUDA(cudaEventCreateWithFlags(&s.join1, cudaEventDisableTiming));
        TUCH_CUDA(cudaEventCreateWithFlags(&s.join2, cudaEventDisableTiming));
        TUCH_CUDA(cudaEventCreateWithFlags(&s.fork_b, cudaEventDisableTiming));
        TUCH_CUDA(cudaEventCreateWithFlags(&s.join_b, cudaEventDisableTiming));
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
    if (!one_stream) if (int rc = side_streams(st, &side)) return rc;
    static const int prio_below = getenv("TUCH_FIT_PRIORITY_BELOW") != nullptr ? atoi(getenv("TUCH_FIT_PRIORITY_BELOW")) : FIT_PRIORITY_BELOW;
    const bool prio = side != nullptr && B < prio_below;
    // s_in: stream of the inside test, s_nn: stream of the masked nearest vertex, s2: joint-side terms
    cudaStream_t s_in = !side ? st : prio ? side->s1 : st, s_nn = !side ? st : prio ? st : side->s0;
    cudaStream_t s2 = side ? side->s2 : st;

    // ---- SMPL forward (split pose; advances the Adam step counters); the output joints follow on side 2
    LbsBuffers w;
    lbs_carve(a->smpl_workspace, B, V, m.L, w);
    if (int rc = launch_lbs_forward(m, a->betas, a->body_pose, 0, B, w, a->vertices, nullptr, st, a->global_orient,
                                    a->step_pose, a->step_orient)) return rc;
    if (side) {
        TUCH_CUDA(cudaEventRecord(side->fork, st));
        TUCH_CUDA(cudaStreamWaitEvent(prio ? s_in : s_nn, side->fork, 0));
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
    // ---- losses.py:73-105: inside test + allowed self-intersections (high-priority side 1), masked nearest vertex
    //      (caller's stream)
    if (int rc = contact_query_impl(topo, a->vertices, B, a->use_segments, am, nullptr, nullptr, ext, nullptr, s_in, nullptr, &s_nn)) return rc;
    if (side) TUCH_CUDA(cudaEventRecord(side->join1, prio ? s_in : s_nn));
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
    return launch_lbs_backward(m, a->body_pose, 0, B, w, g_verts, g_joints, nullptr, nullptr, st, &ad, side ? &bs : nullptr);
}
