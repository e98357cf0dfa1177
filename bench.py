#!/usr/bin/env python
"""bench.py - SK_GS hot path benchmark (FK + LBS + rasterize forward + backward), one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Metric (BASELINE.json): train steps/s, one step = FK + LBS + assembly + preprocess + binning/sort + composite forward +
full backward to the parameter gradients for ONE 800x800 view of the 100K-Gaussian / 32-joint scene (`c2`); with N > 1
ranks every rank renders its own view of the same scene (view sharding, weak scaling) and the Gaussian + skeleton
gradients are all-reduced with NCCL inside the step.  `value` = views processed by all ranks per second.

Timing: W warm-up steps, then K steps; every step is bracketed by its own CUDA-event pair on the launching stream and a
256 MiB memset between steps evicts L2 (excluded from the step time); ranks are aligned by a barrier + synchronize on
both sides and the slowest rank's time counts.  `e2e` repeats the measurement through the public API with the per-step
inputs (camera, joint rotations, upstream image gradient) coming from pinned HOST memory and the step's scalar result
read back to the host inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def _dist():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return world, rank, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.idx), '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def algorithmic_bytes(P, M, K, C, R, W, H, tiles):
    """SURVEY.md 8(d) per-unit figures x the units one launch processes."""
    nb = 6 if tiles <= 65536 else 7
    return {
        'fk_lbs_fwd_kernel': P * (12 + 4 * K + 40) + P * 12 * K,  # xyz in, K sp_W gathers, d_* out, weights+idx out
        'assemble_fwd_kernel': P * (44 + 40 + 44),
        'preprocess_scan_kernel': P * (44 + 12 * C + 75),
        'duplicate_keys_kernel': P * 8 + R * 12,
        'onesweep_pass_kernel': R * 24,  # per pass: read key+value, write key+value
        'tile_ranges_kernel': R * 8,
        'composite_fwd_kernel': R * 44 + H * W * 28,
        'composite_bwd_kernel': R * (44 + 40) + H * W * (20 + 8),
        'preprocess_bwd_kernel': P * (44 + 12 * C + 75 + 40) + P * (12 + 12 + 16 + 4 + 12 + 12 * C),
        'assemble_bwd_kernel': P * (44 + 44 + 44),
        'lbs_bwd_kernel': P * (40 + 12 + 12 * K) + P * 4 * M,
        'fk_bwd_kernel': M * 44 * 4,
        '_sort_passes': nb,
    }


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, world, rank, local):
    import torch.distributed as dist
    from sk_gs_b200 import _lib
    from sk_gs_b200 import scene as S
    from sk_gs_b200.pipeline import HotPath

    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    cfg = S.CONFIGS[args.workload]
    sc = S.make_scene(cfg, views=max(world, 1))
    # SH coefficients live as one [P,16,3] parameter (what the rasterizer reads and what skgs_adam_step updates with two
    # learning rates); the reference keeps f_dc / f_rest apart and concatenates them on every step
    hp = HotPath(sc, dev, mode='W', merged_sh=not args.autograd)
    view = rank % len(sc.cameras)
    H, W = cfg.H, cfg.W
    gen = torch.Generator().manual_seed(1234 + rank)
    dL_host = (torch.randn(3, H, W, generator=gen) / (3 * H * W)).pin_memory()
    dL_dev = dL_host.to(dev)
    # per-step host inputs of the e2e path: what the joint MLP would emit + the camera
    joint_host = {n: getattr(sc, n).clone().pin_memory() for n in ('sk_r', 'sk_d_rot', 'sk_d_scale')}
    cam = sc.cameras[view]
    cam_host = {'viewmatrix': cam.viewmatrix.pin_memory(), 'projmatrix': cam.projmatrix.pin_memory(),
                'campos': cam.campos.clone().pin_memory()}
    rs = hp.settings[view]
    result_host = torch.zeros(1).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    params = list(hp.params.values())

    from sk_gs_b200.dist import GradArena, SymmGradArena, allreduce_max_
    from sk_gs_b200.fk_lbs import scatter_sp_W_grad
    arena = None
    if world > 1:
        # flat fp32 exchange buffer; the backward kernels write into it directly (no packing copies).  sp_W travels in
        # compact [P, K] form (the KNN pattern is identical on every rank); SH gradients as one [P, 16, 3] block.
        shapes = {'shs': (cfg.P, 16, 3), 'xyz': (cfg.P, 3), 'viewspace_points': (cfg.P, 3), 'scaling': (cfg.P, 3),
                  'rotation': (cfg.P, 4), 'opacity': (cfg.P, 1), 'sp_W': (cfg.P, sc.K), 'joints': (cfg.M, 3),
                  'sk_r': (cfg.M, 4), 'sk_d_rot': (cfg.M, 4), 'sk_d_scale': (cfg.M, 3), 'g_tr': (7,)}
        arena = (GradArena if args.allreduce == 'nccl' else SymmGradArena)(shapes, dev, order=list(shapes))
        exchange_kind = 'NVLS multimem all-reduce kernel (symmetric memory)' if getattr(arena, 'multimem', False) \
            else 'NCCL all-reduce'
        dL_dev.mul_(1.0 / world)   # mean over the views of the step: folded into the upstream gradient
        dL_host.mul_(1.0 / world)

    def exchange(out, grads):
        """The one exchange step of a data-parallel iteration: SUM of all gradients (the MAX of the screen radii was
        already started on a side stream right after the forward, see after_forward)."""
        if world == 1:
            return
        if split_exchange:  # the rasterizer-side blocks were reduced under the LBS / FK backward (mid_backward)
            arena.allreduce_range(arena.block_start('sp_W'), arena.flat_padded.numel(), channel=1)
        else:
            arena.allreduce(chunks=1)  # one call: at 27 MB two NCCL chunks cost more latency than they overlap
        scatter_sp_W_grad(arena.view('sp_W'), out['_sk'][8], cfg.M)  # dense [P, M] gradient for the optimizer

    side = torch.cuda.Stream(dev) if world > 1 else None

    def after_forward(radii):
        """radii are final after the forward: their MAX all-reduce runs on a side stream under the whole backward."""
        if world == 1:
            return None
        main = torch.cuda.current_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            allreduce_max_(radii)
        return lambda: main.wait_stream(side)

    split_exchange = world > 1 and getattr(arena, 'multimem', False)
    side2 = torch.cuda.Stream(dev) if split_exchange else None

    def mid_backward():
        """Called when every rasterizer-side gradient (SH, means, scales, rotations, opacity: 92 % of the bytes) is final:
        their in-switch reduction runs on a side stream while the LBS and FK backward kernels execute."""
        main = torch.cuda.current_stream(dev)
        side2.wait_stream(main)
        with torch.cuda.stream(side2):
            arena.allreduce_range(0, arena.block_start('sp_W'), channel=0)
        return lambda: main.wait_stream(side2)

    uploads = [(hp.params[n].data, t) for n, t in joint_host.items()] + \
        [(rs.viewmatrix, cam_host['viewmatrix']), (rs.projmatrix, cam_host['projmatrix']), (rs.campos, cam_host['campos'])]

    def upload():
        for dst, src in uploads:
            dst.copy_(src, non_blocking=True)
        dL_dev.copy_(dL_host, non_blocking=True)

    def download(img):
        result_host.copy_((img.detach() * dL_dev).sum().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if args.graph and hp.overflowed():
            raise RuntimeError('binning capacity of the captured graph exceeded')
        return float(result_host[0])

    graph_state = {}
    compact = world > 1

    def step(e2e: bool):
        if args.graph:
            key = 'e2e' if e2e else 'dev'
            if key not in graph_state:  # the e2e graph contains the host->device uploads, the device graph does not
                graph_state[key] = hp.capture_step(view, dL_dev, compact_sp_W=compact,
                                                   uploads=uploads if e2e else None, dL_host=dL_host if e2e else None,
                                                   epilogue=exchange if world > 1 else None, arena=arena,
                                                   after_forward=after_forward if world > 1 else None,
                                                   mid_backward=mid_backward if split_exchange else None)
            g, out, grads = graph_state[key]
            g.replay()  # with N > 1 the NCCL all-reduces are nodes of the same graph
            return download(out['images']) if e2e else None
        elif args.autograd:  # the drop-in autograd API (render_gs_offical + fk_lbs + assemble Functions)
            if e2e:
                upload()
            hp.zero_grad()
            out = hp.render(view)
            out['images'].backward(dL_dev)
            grads = dict(hp.grads())
            grads['viewspace_points'] = out['viewspace_points'].grad
            if compact:
                grads['sp_W'] = torch.gather(grads['sp_W'], 1, out['_sk'][8])
                grads['shs'] = torch.cat((grads['f_dc'], grads['f_rest']), 1)
                arena.pack(grads)
                allreduce_max_(out['radii'])
        else:
            if e2e:
                upload()
            out, grads = hp.step_grads(view, dL_dev, compact_sp_W=compact, arena=arena,
                                       after_forward=after_forward if world > 1 else None,
                                       mid_backward=mid_backward if split_exchange else None)
        exchange(out, grads)
        return download(out['images']) if e2e else None

    def timed(e2e: bool, K: int, Wu: int):
        for _ in range(Wu):
            step(e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        launches0 = _lib.launch_count()
        for _ in range(K):
            flush.zero_()  # evict L2 (126 MB) between steps; outside the event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(e2e)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - launches0
        if args.graph:
            launches = K * getattr(hp, 'launches_per_step', 0)  # replays do not pass through the launch counter
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(True, args.steps, max(args.warmup, 3))

    # ---- render FPS (forward only, the reference's test.py --fps protocol: CUDA events around N renders, no_grad)
    fps = None
    if rank == 0:
        with torch.no_grad():
            for _ in range(3):
                hp.render(view)
            torch.cuda.synchronize()
            gfps = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gfps):
                hp.render(view)
            torch.cuda.synchronize()
            nf = max(50, args.steps)
            evs = []
            for _ in range(nf):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                gfps.replay()
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            fps = round(nf / (sum(a.elapsed_time(b) for a, b in evs) * 1e-3), 1)

    # ---- per-kernel device times for the roofline (separate pass, events around every launch)
    kern = {}
    R = 0
    if rank == 0:
        _lib.profile_enable(True)
        nprof = 5
        for _ in range(nprof):
            flush.zero_()
            out = hp.render(view)
            out['images'].backward(dL_dev)
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        from sk_gs_b200 import diff_gaussian_rasterization as DGR
        R = DGR._capacity.get((dev.index, cfg.P, W, H)) or 0
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        ab = algorithmic_bytes(cfg.P, cfg.M, sc.K, 16, R, W, H, tiles)
        peak, peak_src = measured_peak_hbm()
        for name, (n, us) in prof.items():
            per = us / n
            key = 'onesweep_pass_kernel' if name.startswith('onesweep_pass') else name
            gbs = ab.get(key, 0) / (per * 1e-6) / 1e9 if per > 0 else 0.0
            kern[name] = {'launches_per_step': n / nprof, 'us_per_launch': round(per, 3),
                          'us_per_step': round(us / nprof, 3), 'algorithmic_GBps': round(gbs, 1),
                          'frac_of_peak': round(gbs / peak, 4)}
        dom = max(kern, key=lambda k: kern[k]['us_per_step'])
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r1_roofline_traffic.json')
        if os.path.exists(tpath) and args.workload == 'c2':
            with open(tpath) as f:
                tj = json.load(f).get(dom)
            if tj:
                traffic = tj['dram_bytes_read'] + tj['dram_bytes_write']  # bytes per launch, one ncu --set full capture
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': kern[dom]['algorithmic_GBps'], 'peak': peak,
                    'peak_source': peak_src, 'unit': 'GB/s', 'frac': kern[dom]['frac_of_peak'],
                    'traffic': traffic, 'algorithmic_bytes': ab.get(dom),
                    'note': 'algorithmic bytes (SURVEY 8d) / CUDA-event time of the kernel; compositing is FP32-issue '
                            'bound (about 145 flop/B), see DESIGN.md'}
    # ---- widening rows: the complete iteration render -> L1+SSIM loss -> backward -> Adam as one
    # CUDA graph.  Extra information only; the headline metric above is BASELINE.json's (loss and optimizer excluded).
    iteration = None
    if rank == 0 and world == 1 and not args.no_iteration:
        try:
            iteration = full_iteration(args, sc, cfg, dev, view, flush, kern)
        except Exception as e:  # noqa: never let the extra section take the contract line down
            iteration = {'error': f'{type(e).__name__}: {e}'[:300]}
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    # ---- CPU baseline: the oracle port on the host cores, bounded sample (same workload, 1 view)
    cpu = cpu_baseline(args.workload, steps=args.cpu_steps) if (world == 1 and not args.no_cpu) else None
    K = args.steps
    line = {
        'metric': 'train_steps_per_sec (FK+LBS+render fwd+bwd, one 800x800 view per step)',
        'value': round(world * K / (ms_dev * 1e-3), 2), 'unit': 'steps/s', 'n_gpus': world, 'steps': K,
        'warmup': args.warmup, 'ms_per_step': round(ms_dev / K, 4), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{cfg.name}: {cfg.P} Gaussians, {cfg.M} joints, {W}x{H}, 1 view per GPU, SH degree 3, '
                               f'K=5 LBS mode W, fwd+bwd', 'num_rendered': R, 'views_per_step': world,
                   'parallelism': f'view-sharded dp{world}' + (f' + {exchange_kind}' if world > 1 else ''),
                   'l2_flush': '256 MiB memset between steps, outside the per-step CUDA-event pairs',
                   'sh_layout': 'f_dc/f_rest parameters + cat per step' if args.autograd else 'one [P,16,3] parameter',
                   'launch': 'CUDA graph replay (fixed binning capacity, overflow flag checked)' if args.graph
                   else ('eager launches through the autograd API' if args.autograd else 'eager launches')},
        'clocks': clocks,
        'e2e': {'value': round(world * K / (ms_e2e * 1e-3), 2), 'unit': 'steps/s',
                'ms_per_step': round(ms_e2e / K, 4),
                'h2d_bytes_per_step': int(dL_host.numel() * 4 + sum(t.numel() for t in joint_host.values()) * 4 +
                                          sum(t.numel() for t in cam_host.values()) * 4),
                'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'render_fps': {'value': fps, 'unit': 'frames/s', 'note': 'forward only (FK+LBS+assembly+rasterize), 1 GPU, CUDA graph'},
        'roofline': roofline,
        'kernels': kern,
        'full_iteration': iteration,
        'cpu_baseline': cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def full_iteration(args, sc, cfg, dev, view, flush, kern):
    """render -> fused L1 + SSIM loss -> backward -> one-launch Adam on one GPU (sk_gs_b200.train.TrainLoop), replayed as a
    CUDA graph; then an eager profiled pass for the per-kernel times of the loss and optimizer kernels."""
    from sk_gs_b200 import _lib
    from sk_gs_b200 import diff_gaussian_rasterization as DGR
    from sk_gs_b200.pipeline import HotPath
    from sk_gs_b200.train import TrainLoop
    H, W = cfg.H, cfg.W
    fixed = DGR._capacity.fixed
    # joint rotations come from the joint-rotation network (8f-1); head_std 0.02 gives rotations of ~10 degrees, the
    # regime of the synthetic scene (the reference's 1e-6 init would make every joint rotation the identity)
    hp = HotPath(sc, dev, mode='W', requires_grad=False, merged_sh=True, joint_mlp=True, head_std=0.02)
    # Adam moves every network weight by +-lr per step; 1e-5 instead of the reference's 1e-3 keeps the synthetic scene
    # (and with it R, the work per iteration) stationary over the timed replays - same kernels, same bytes
    loop = TrainLoop(hp, lrs={'theta': 1e-5})
    # target = the scene's own rendering + noise: the near-converged regime, parameters (and with them the number of
    # (Gaussian, tile) pairs the fixed-capacity graph must hold) drift slowly
    with torch.no_grad():
        target = hp.render(view)['images'].detach().clone()
    target = (target + 0.05 * torch.randn(3, H, W, generator=torch.Generator().manual_seed(99)).to(dev)).clamp_(0, 1)
    try:
        loop.capture(view, target, headroom=2.0)
        K = max(20, min(args.steps, 200))
        for _ in range(5):
            loop.replay(wait=False)
        torch.cuda.synchronize()
        evs = []
        for _ in range(K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loop.replay(wait=False)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / K
        overflow = hp.overflowed()
        R_last = int(DGR.last_header_words(dev)[0])
        terms = [round(float(x), 6) for x in loop.out['loss_terms'].cpu()]
    finally:
        DGR.set_fixed_capacity(fixed)
    # per-kernel pass (eager, events around every launch)
    _lib.profile_enable(True)
    nprof = 5
    for _ in range(nprof):
        flush.zero_()
        loop.step(view, target)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    peak, _ = measured_peak_hbm()
    n_params = sum(hp.params[n].numel() for n in loop.names)
    ab = {'ssim_stats_kernel': 3 * H * W * (8 + 12), 'ssim_grad_kernel': 3 * H * W * (12 + 8 + 4),
          'adam_kernel': 28 * n_params}
    for name in prof:
        if name in ab or name.startswith('joint_'):
            n, us = prof[name]
            per = us / n
            gbs = ab.get(name, 0) / (per * 1e-6) / 1e9
            kern[name] = {'launches_per_step': n / nprof, 'us_per_launch': round(per, 3),
                          'us_per_step': round(us / nprof, 3), 'algorithmic_GBps': round(gbs, 1),
                          'frac_of_peak': round(gbs / peak, 4)}
    return {'value': round(1e3 / ms, 2), 'unit': 'iterations/s', 'ms_per_iteration': round(ms, 4),
            'what': 'joint-rotation MLP -> FK+LBS+render fwd -> L1+SSIM loss fwd+bwd -> render/LBS/FK/MLP bwd -> Adam over all parameters '
                    f'({n_params} floats), one CUDA graph, 1 GPU',
            'loss_terms_last': terms, 'num_rendered_last': R_last, 'overflow': bool(overflow)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline(workload: str, steps: int = 2, threads: int = 0):
    """Oracle port (torch FK/LBS + C rasterizer) on the host cores: `steps` full fwd+bwd steps of the same workload."""
    from oracle import raster as OR
    from sk_gs_b200 import scene as S
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from skgs_test_util import np32, oracle_deform, oracle_settings
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    OR.set_num_threads(cores)
    cfg = S.CONFIGS[workload]
    sc = S.make_scene(cfg, views=1)
    s = oracle_settings(sc.cameras[0])
    rng = np.random.default_rng(0)
    dC = (rng.standard_normal((3, cfg.H, cfg.W)) / (3 * cfg.H * cfg.W)).astype(np.float32)

    def one():
        net, sk_out, leaves = oracle_deform(sc, requires_grad=True)
        img, g, b = OR.render_forward(s, np32(net['points']), np32(net['opacity']), np32(net['scales']),
                                      np32(net['rotations']), np32(net['sh_features']))
        gr = OR.render_backward(s, g, b, img, dC, np32(net['points']), np32(net['scales']), np32(net['rotations']),
                                np32(net['sh_features']))
        # back through assembly + LBS + FK with torch autograd
        torch.autograd.backward(
            [net['points'], net['scales'], net['rotations'], net['opacity'], net['sh_features']],
            [torch.from_numpy(gr.dL_dmeans3D), torch.from_numpy(gr.dL_dscales), torch.from_numpy(gr.dL_drotations),
             torch.from_numpy(gr.dL_dopacity).reshape(-1, 1), torch.from_numpy(gr.dL_dsh)])
        return b.R

    one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return {'value': round(steps / dt, 4), 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
            'sample': f'{steps} full fwd+bwd steps of {cfg.name} after 1 warm-up, oracle/fk_lbs.py (torch, {cores} threads)'
                      f' + oracle/raster_oracle.c (OpenMP, {OR.num_threads()} threads)'}


def run_reference(args, world, rank, local):
    """Reference arm.  The reference has NO CPU implementation of this path: its rasterizer is a CUDA extension.  When the
    extension compiled here from /root/reference (oracle/_ref) is loadable and a GPU is present, it is what runs
    (kind "reference", device cuda) with the torch-op FK/LBS of oracle/fk_lbs.py on the GPU; otherwise the oracle port on
    the host cores (kind "port").  --ref-device cpu forces the latter."""
    if rank != 0:
        return
    from oracle import ref_ext
    use_gpu = args.ref_device != 'cpu' and torch.cuda.is_available() and ref_ext.available()
    K = args.steps
    if not use_gpu:
        steps = min(K, args.cpu_steps)
        cpu = cpu_baseline(args.workload, steps=steps)
        from sk_gs_b200 import scene as S
        cfg = S.CONFIGS[args.workload]
        line = {'impl': 'reference', 'metric': 'train_steps_per_sec (FK+LBS+render fwd+bwd, one 800x800 view per step)',
                'value': cpu['value'], 'unit': 'steps/s', 'n_gpus': world, 'steps': steps, 'warmup': 1,
                'ms_per_step': round(1000.0 / cpu['value'], 3), 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': cfg.name},
                'cpu_baseline': cpu, 'e2e': {'value': cpu['value'], 'unit': 'steps/s', 'h2d_bytes_per_step': 0,
                                             'd2h_bytes_per_step': 0}}
        print(json.dumps(line), flush=True)
        return
    from oracle import fk_lbs as OF
    from sk_gs_b200 import scene as S
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    cfg = S.CONFIGS[args.workload]
    sc = S.make_scene(cfg, views=1)
    cam = sc.cameras[0]
    names = ['xyz', 'scaling', 'rotation', 'opacity', 'f_dc', 'f_rest', 'sp_W', 'joints', 'sk_r', 'sk_d_rot',
             'sk_d_scale', 'g_tr']
    p = {n: getattr(sc, n).to(dev).clone().requires_grad_(True) for n in names}
    parents = sc.parents.to(dev).long()
    H, W = cfg.H, cfg.W
    gen = torch.Generator().manual_seed(1234)
    dL_host = (torch.randn(3, H, W, generator=gen) / (3 * H * W)).pin_memory()
    dL_dev = dL_host.to(dev)
    joint_host = {n: getattr(sc, n).clone().pin_memory() for n in ('sk_r', 'sk_d_rot', 'sk_d_scale')}
    result_host = torch.zeros(1).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(e2e):
        for t in p.values():
            t.grad = None
        if e2e:
            for n, t in joint_host.items():
                p[n].data.copy_(t, non_blocking=True)
            dL_dev.copy_(dL_host, non_blocking=True)
        out = OF.sk_stage(p['xyz'], p['joints'], p['sk_r'], p['sk_d_rot'], p['sk_d_scale'], p['g_tr'], parents, sc.root,
                          K=sc.K, mode='W', sp_W=p['sp_W'])
        pts, scl, rot, op, sh = OF.assemble(p['xyz'], p['scaling'], p['rotation'], p['opacity'], p['f_dc'], p['f_rest'],
                                            *out[:3])
        r = ref_ext.render(pts, op, scl, rot, sh, cam)
        r['images'].backward(dL_dev)
        if e2e:
            result_host.copy_((r['images'].detach() * dL_dev).sum().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def timed(e2e, K, Wu):
        for _ in range(Wu):
            step(e2e)
        torch.cuda.synchronize()
        evs = []
        for _ in range(K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(e2e)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    ms_dev = timed(False, K, args.warmup)
    ms_e2e = timed(True, K, max(args.warmup, 3))
    line = {'impl': 'reference', 'metric': 'train_steps_per_sec (FK+LBS+render fwd+bwd, one 800x800 view per step)',
            'value': round(K / (ms_dev * 1e-3), 2), 'unit': 'steps/s', 'n_gpus': 1, 'steps': K, 'warmup': args.warmup,
            'ms_per_step': round(ms_dev / K, 4), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{cfg.name}: {cfg.P} Gaussians, {cfg.M} joints, {W}x{H}, 1 view, fwd+bwd',
                       'l2_flush': '256 MiB memset between steps, outside the per-step CUDA-event pairs'},
            'cpu_baseline': {'value': round(K / (ms_dev * 1e-3), 2), 'unit': 'steps/s', 'cores': 0, 'kind': 'reference',
                             'device': 'cuda',
                             'sample': "the reference's own CUDA rasterizer (my_ext/_C/src/nerf/gaussian_*.cu compiled "
                                       'unmodified into oracle/_ref, colmap=True) + torch-op FK/LBS on the GPU; the '
                                       'reference has no CPU implementation of this path'},
            'e2e': {'value': round(K / (ms_e2e * 1e-3), 2), 'unit': 'steps/s', 'ms_per_step': round(ms_e2e / K, 4),
                    'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--ref-device', default='auto', choices=['auto', 'cpu', 'cuda'])
    ap.add_argument('--cpu-steps', type=int, default=40,
                    help='bounded CPU sample: full fwd+bwd oracle steps of the same workload (about 10-20 s of host time)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-iteration', action='store_true', help='skip the loss+Adam full-iteration section')
    ap.add_argument('--allreduce', default='multimem', choices=['multimem', 'nccl'],
                    help='gradient exchange for N > 1: in-switch multimem kernel over symmetric memory, or NCCL')
    ap.add_argument('--autograd', action='store_true', help='with --no-graph: time the drop-in autograd API path')
    ap.add_argument('--no-graph', dest='graph', action='store_false',
                    help='launch every step eagerly instead of replaying a captured CUDA graph')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world, rank, local = _dist()
    if args.impl == 'reference':
        # rank 0 alone runs and prints the reference arm; the other ranks exit 0 without work (no process group needed)
        if rank == 0:
            run_reference(args, world, rank, local)
        return
    if world > 1:
        import torch.distributed as dist
        # measured on the 8xB200 NVSwitch box for the 26.8 MB gradient arena (tools/ar_sweep.sh): Ring 111 us,
        # default (NVLS) 140 us, Tree 158 us
        os.environ.setdefault('NCCL_ALGO', 'Ring')
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    try:
        run_ours(args, world, rank, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
