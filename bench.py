"""SMPLify-DC throughput benchmark (BASELINE.json: "SMPLify-DC iters/sec (batch-256) at 1/2/4/8 B200;
contact-kernel HBM GB/s").

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference ...                     # CPU port of the reference (oracle/), host cores

A "step" is ONE stage-2 SMPLify-DC iteration (tuch/smplify/smplifydc.py:155-183) over a batch of 256
synthetic SMPL-sized bodies per GPU: SMPL forward -> contact_fitting_loss (winding-number inside test,
segment whitelist, geodesically-masked nearest vertex, push/pull + region-to-region terms, reprojection,
GMM prior) -> backward -> Adam.  Bodies are independent, so N GPUs run N shards of 256 bodies with no
data-path collective (weak scaling); the reported value is whole-job iterations/s in batch-256 units,
i.e. (bodies processed by all ranks / 256) / time.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 256                      # bodies per GPU (the batch BASELINE.json's metric is quoted on)
GEOTHRES, EUCLTHRES = 0.3, 0.02  # configs/config.py:90-91 as passed by train.py:72-76
CONTACT_W = 2000.0               # configs/smplify_dc_options.py:37
METRIC = 'smplify_dc_iters_per_sec_batch256'
UNIT = 'iters/s'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def make_assets(batch, seed):
    from tuch_b200 import synthetic as syn
    model = syn.make_lattice_body_model(seed=0)                                   # V = 6890, F = 13776
    geo = syn.make_geodesics(model['v_template'], model['faces'],
                             cache_dir=os.environ.get('TUCH_B200_CACHE', '/tmp/tuch_b200_cache'))
    regions = syn.make_regions(model)
    segs = syn.make_segments(model)
    gmm = syn.make_gmm()
    inp = syn.make_smplify_inputs(model, regions, batch, seed=seed)
    return dict(model=model, geo=geo, regions=regions, segs=segs, gmm=gmm, inp=inp)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled back to back (~20 Hz) while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


# ====================================================================================== our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from tuch_b200 import ops
    from tuch_b200.models.smpl import SMPL
    from tuch_b200.smplify.prior import MaxMixturePrior
    from tuch_b200.smplify.smplifydc import SMPLifyDC
    from tuch_b200.utils.segmentation import BatchBodySegment
    from tuch_b200 import synthetic as syn

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('--gpus %d needs torchrun with %d ranks (see the module docstring)' % (args.gpus, args.gpus))
        raise SystemExit('WORLD_SIZE=%d does not match --gpus %d' % (world, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    a = make_assets(B, seed=1000 + rank)                                       # every rank owns its own shard
    model, inp = a['model'], a['inp']
    V, F = len(model['v_template']), len(model['faces'])

    smpl = SMPL(model_arrays=model, batch_size=B).to(dev)
    prior = MaxMixturePrior(gmm=a['gmm'], num_gaussians=8).to(dev)
    faces = torch.tensor(model['faces'], device=dev)
    segments = BatchBodySegment(list(a['segs'].keys()), faces, segment_data=a['segs'])
    geod = torch.tensor(a['geo'], device=dev)
    ign = [syn.JOINT_IDS[n] for n in syn.IGN_JOINTS]
    smplify = SMPLifyDC(step_size=1e-2, batch_size=B, num_iters=args.steps, focal_length=syn.FOCAL_LENGTH,
                        geodistssmpl=geod, geothres=GEOTHRES, euclthres=EUCLTHRES, device=dev, smpl=smpl,
                        pose_prior=prior, ign_joints=ign)

    # pinned host copies of the per-batch inputs (what a caller hands to SMPLifyDC.__call__)
    host = {k: torch.tensor(np.ascontiguousarray(inp[k])).pin_memory()
            for k in ('init_pose', 'init_betas', 'init_cam_t', 'camera_center', 'keypoints_2d', 'gt_contact',
                      'has_discrete_contact', 'ignore_idxs')}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())

    def upload():
        return {k: t.to(dev, non_blocking=True) for k, t in host.items()}

    def begin(d):
        kp = d['keypoints_2d']
        conf = kp[:, :, 2].clone()
        conf[:, ign] = 0.0
        return smplify.begin_contact_fit(d['init_pose'][:, 3:].clone(), d['init_pose'][:, :3].clone(),
                                         d['init_betas'], d['init_cam_t'], d['camera_center'],
                                         kp[:, :, :2].contiguous(), conf, a['regions'], [d['gt_contact'], None],
                                         d['ignore_idxs'], d['has_discrete_contact'], CONTACT_W, 'sum', segments)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- device-resident leg (`value`)
    fit = begin(upload())
    for _ in range(max(args.warmup, 3)):
        fit.step()
    ops.kernel_timing(enable=True, reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        loss = fit.step()
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ops.launch_count() - launches0
    wind_ms, wind_n = ops.kernel_time('winding_kernel')
    near_ms, near_n = ops.kernel_time('nearest_kernel')
    seg_ms, seg_n = ops.kernel_time('winding_kernel_segments')
    kernel_ms = {k: round(v[0] / args.steps, 4) for k, v in sorted(ops.kernel_times().items()) if v[1]}
    ops.kernel_timing(enable=False)
    final_loss = float(loss.item())
    assert np.isfinite(final_loss), 'objective diverged'
    ms_per_step = ms / args.steps
    value = (world * B / float(BATCH)) / (ms_per_step * 1e-3)

    # ---------------------------------------------------------------- end-to-end leg (`e2e`)
    # every step: pinned-host -> device copy of the batch inputs, one iteration through the public
    # SMPLifyDC API (ContactFit.load + step, the iteration replayed as a CUDA graph), device -> host read of
    # the loss and the updated pose
    out_host = torch.empty(B, 72).pin_memory()
    loss_host = torch.empty(()).pin_memory()
    d2h_bytes = out_host.numel() * 4 + 4

    # the fit object and its captured CUDA graph are built once; every step loads a fresh batch of host
    # inputs into it (ContactFit.load = begin_contact_fit() for a new batch of the same size)
    fit_e = begin(upload()).capture()

    def e2e_step():
        fit_e.load(host['init_pose'], host['init_betas'], host['init_cam_t'], host['camera_center'],
                   host['keypoints_2d'], host['gt_contact'], host['ignore_idxs'], host['has_discrete_contact'])
        l = fit_e.step()
        out_host[:, :3].copy_(fit_e.global_orient.detach(), non_blocking=True)
        out_host[:, 3:].copy_(fit_e.body_pose.detach(), non_blocking=True)
        loss_host.copy_(l.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), 0.0)) / args.steps
    e2e_wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    clocks = sampler.finish() if rank == 0 else None      # sampled over both timed regions (value + e2e legs)
    e2e_ms = max(e2e_ms, e2e_wall_ms)          # host-side work between launches counts end to end
    e2e_value = (world * B / float(BATCH)) / (e2e_ms * 1e-3)

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = peaks()
    # algorithmic bytes of one winding launch: vertices in (B*V*12) + faces (F*12) + winding out (B*V*4)
    alg_bytes = B * V * 12 + F * 12 + B * V * 4
    wind_avg_ms = wind_ms / max(wind_n, 1)
    achieved = alg_bytes / (wind_avg_ms * 1e-3) / 1e9 if wind_n else None
    pairs = float(B) * V * F
    roofline = dict(bound='hbm', kernel='winding_cluster_kernel', achieved=achieved, peak=peak, unit='GB/s',
                    frac=(achieved / peak) if achieved else None, traffic=None, peak_source=peak_src,
                    algorithmic_bytes_per_launch=alg_bytes, avg_launch_ms=wind_avg_ms, launches_timed=wind_n,
                    share_of_step=(wind_ms / ms) if ms > 0 else None,
                    equivalent_pair_evals_per_s=(pairs / (wind_avg_ms * 1e-3)) if wind_n else None,
                    nearest_kernel_avg_ms=near_ms / max(near_n, 1), segment_winding_ms_per_step=seg_ms / args.steps,
                    clusters=fit.topo.cluster_stats(), kernel_ms_per_step=kernel_ms,
                    note='the winding numbers never touch HBM as a [V,F] tensor: the kernel is bound by fp32 issue slots '
                         '(see roofline_compute), not by HBM; its algorithmic bytes are the vertices in, the faces and '
                         'the winding numbers out.  See DESIGN.md section 5 and profiles/')
    # the binding roof: warp-instruction issue slots (148 SMs x 4 schedulers x SM clock).  The issued warp
    # instructions per launch come from the committed ncu capture of the same workload shape
    # (profiles/winding_traffic.json); the launch time is measured live above.
    sm_hz = (clocks or {}).get('sm_mhz') if clocks else None
    traffic_file = os.path.join(ROOT, 'profiles', 'winding_traffic.json')
    tf = None
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            tf = json.load(f)
    roofline_compute = None
    if tf and tf.get('batch') == B and tf.get('kernel') == 'winding_cluster_kernel':
        roofline['traffic'] = tf.get('dram_bytes_per_launch')
        wi = tf.get('warp_instr_per_launch')
        if wi and sm_hz and wind_n:
            peak_issue = 148 * 4 * sm_hz * 1e6
            ach = wi / (wind_avg_ms * 1e-3)
            roofline_compute = dict(bound='fp32-issue', kernel='winding_cluster_kernel', achieved=ach, peak=peak_issue,
                                    unit='warp-instr/s', frac=ach / peak_issue, warp_instr_per_launch=wi, sm_mhz=sm_hz,
                                    ncu_issue_active_pct=tf.get('issue_active_pct'),
                                    note='peak = 148 SM x 4 issue slots x SM clock; instructions per launch from the ncu '
                                         'capture in profiles/ (same batch and body), launch time measured live')

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_iteration(a, n_bodies=args.cpu_bodies, repeats=1)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                    data='synthetic',
                    config=dict(workload='SMPLify-DC stage-2 iteration (SMPL fwd + contact_fitting_loss + bwd + Adam), '
                                         'batch=%d bodies per GPU, V=%d, F=%d, synthetic DSC contact pairs' % (B, V, F),
                                bodies_per_gpu=B, total_bodies=world * B, geothres=GEOTHRES, euclthres=EUCLTHRES,
                                contact_loss_weight=CONTACT_W, segments=len(a['segs']), region_pairs=len(a['regions']['classes']),
                                l2_policy='per-step working set (packed leaf triangles %.0f MB + node records + partials) '
                                          'exceeds the 126 MB L2' % (B * fit.topo.cluster_stats()['leaves'] * fit.topo.cluster_stats()['leaf_faces'] * 48 / 1e6),
                                body_iters_per_s=world * B / (ms_per_step * 1e-3), final_loss=final_loss),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=d2h_bytes,
                             ms_per_step=e2e_ms),
                    gpu_launches=int(launches), clocks=clocks, roofline=roofline)
        if roofline_compute is not None:
            line['roofline_compute'] = roofline_compute
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ====================================================================================== CPU port
def cpu_port_iteration(a, n_bodies, repeats=1):
    """The same stage-2 iteration in the CPU oracle (torch-CPU LBS + autograd, OpenMP C pair loops) on a
    bounded sample of bodies; the reference loops over bodies, so its cost is linear in the batch."""
    import torch
    from oracle import clib, lbs as olbs, losses as ol, segments as oseg
    model, inp = a['model'], a['inp']
    n = min(n_bodies, len(inp['init_pose']))
    tm = olbs.to_torch_model(model)
    geomask = a['geo'] > GEOTHRES
    segs = oseg.build_segments(a['segs'], model['faces'])
    prior = ol.GMMPrior(a['gmm'])
    from tuch_b200 import synthetic as syn
    ign = [syn.JOINT_IDS[k] for k in syn.IGN_JOINTS]
    bp = torch.tensor(inp['init_pose'][:n, 3:]).requires_grad_(True)
    go = torch.tensor(inp['init_pose'][:n, :3]).requires_grad_(True)
    betas = torch.tensor(inp['init_betas'][:n])
    kp = torch.tensor(inp['keypoints_2d'][:n])
    conf = kp[:, :, 2].clone()
    conf[:, ign] = 0
    opt = torch.optim.Adam([bp, go], lr=1e-2)
    clib.build()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        verts, joints, _ = olbs.smpl_forward(tm, betas, bp, go)
        loss = ol.contact_fitting_loss(bp, betas, joints, geomask, EUCLTHRES, torch.tensor(inp['init_cam_t'][:n]),
                                       torch.tensor(inp['camera_center'][:n]), kp[:, :, :2], conf, prior, a['regions'],
                                       inp['gt_contact'][:n], inp['ignore_idxs'][:n], inp['has_discrete_contact'][:n],
                                       verts, model['faces'], contact_loss_weight=CONTACT_W, segments=segs)
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    value = (n / float(BATCH)) / best
    return dict(value=value, unit=UNIT, cores=max(clib.num_threads(), torch.get_num_threads()), kind='port',
                sample='%d of %d bodies, one stage-2 iteration in oracle/ (C/OpenMP pair loops, torch-CPU LBS + autograd), '
                       '%.2f s; scaled linearly to batch %d (the reference loops over bodies)' % (n, BATCH, best, BATCH),
                host_cpus=os.cpu_count(), seconds=best, loss=float(loss.item()))


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    a = make_assets(max(args.cpu_bodies, 1), seed=1000)
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    vals = []
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_port_iteration(a, args.cpu_bodies)
    for _ in range(steps):
        vals.append(cpu_port_iteration(a, args.cpu_bodies))
    cpu = min(vals, key=lambda c: c['seconds'])
    mean_s = float(np.mean([c['seconds'] for c in vals]))
    value = (args.cpu_bodies / float(BATCH)) / mean_s
    cpu = dict(cpu, value=value)
    V, F = len(a['model']['v_template']), len(a['model']['faces'])
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps,
                warmup=max(0, min(args.warmup, 1)), ms_per_step=mean_s * 1e3 * BATCH / args.cpu_bodies,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload='SMPLify-DC stage-2 iteration (SMPL fwd + contact_fitting_loss + bwd + Adam), '
                                     'batch=%d bodies, V=%d, F=%d, synthetic DSC contact pairs; CPU port of the reference '
                                     'timed on %d of the %d bodies per step and scaled linearly' % (BATCH, V, F, args.cpu_bodies, BATCH),
                            bodies_per_gpu=BATCH, geothres=GEOTHRES, euclthres=EUCLTHRES, contact_loss_weight=CONTACT_W),
                cpu_baseline=cpu,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=time.perf_counter() - t0,
                note='the reference is pure Python/PyTorch on un-shipped third-party smplx and data files and cannot '
                     'travel to the GPU box; this arm times oracle/, the CPU restatement pinned to the reference by '
                     'tests/golden (kind=port)')
    print(json.dumps(line))


def main():
    if '--impl' in sys.argv and 'reference' in sys.argv:
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core
        os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH, help='bodies per GPU')
    ap.add_argument('--cpu-bodies', type=int, default=32, help='bodies in the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
