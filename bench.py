#!/usr/bin/env python
"""bench.py - SK_GS hot path benchmark (FK + LBS + rasterize forward + backward), one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Metric (BASELINE.json): train steps/s, one step = FK + LBS + assembly + preprocess + binning/sort + composite forward +
full backward to the parameter gradients for ONE 800x800 view of the 100K-Gaussian / 32-joint scene (`c2`); with N > 1
ranks every rank renders its own view of the same scene (view sharding, weak scaling) and the Gaussian + skeleton
gradients are all-reduced inside the step.  `value` = views processed by all ranks per second.

Timing: W warm-up steps, then K steps; every step is bracketed by its own CUDA-event pair on the launching stream and a
256 MiB memset between steps evicts L2 (excluded from the step time); ranks are aligned by a barrier + synchronize on
both sides and the slowest rank's time counts.  `e2e` repeats the measurement through the public API with the per-step
inputs (camera, joint rotations, upstream image gradient) coming from pinned HOST memory and the step's scalar result
read back to the host inside the timed region.

Besides the headline the line carries (both arms print the same keys, so ratios can be taken per block):
  `dropin`     the same step through the drop-in operator API exactly as the reference calls it - fk_lbs -> assemble ->
               render_gs_offical under torch autograd, separate f_dc / f_rest parameters, eager launches;
  `raster_only` rasterizer forward + backward alone (the part of the reference arm that is the reference's own code);
  `stages`     CUDA-event times of {FK/LBS fwd, raster fwd, raster bwd, FK/LBS bwd} of one eager step;
  `workloads`  the other named shapes: `ns` (north-star target, 300K Gaussians @ 800x800), `c3` (200K @ 512x512, 8 views
               per step sharded over the ranks: strong scaling), `c4` (300K @ 1024x1024, 4 views per step), `c5` (3M
               Gaussians @ 1080p, 64 poses sharded over the ranks, forward only, frames/s).
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

METRIC = 'train_steps_per_sec (FK+LBS+render fwd+bwd, one 800x800 view per step)'


# what the watchdog needs: the section in progress and, once it exists, the headline line
_STATE = {'section': 'setup', 'line': None, 'done': False}


def _start_watchdog(seconds: float, rank: int):
    """A benchmark must never hang the box (a collective that dead-locks spins on the GPU for ever): after `seconds`
    the process prints what it has - the headline line if it was measured, marked `aborted` - and exits.  Every rank
    runs its own watchdog, so torchrun sees all of them leave."""
    import threading
    import time

    def run():
        time.sleep(seconds)
        if _STATE['done']:
            return
        line = _STATE['line']
        if rank == 0:
            if line is not None:
                line = dict(line)
                line['aborted'] = {'after_seconds': seconds, 'section': _STATE['section']}
                print(json.dumps(line), flush=True)
            else:
                print(json.dumps({'error': f'watchdog: no headline after {seconds} s (in {_STATE["section"]})'}),
                      flush=True)
        sys.stdout.flush()
        os._exit(0 if line is not None else 124)

    threading.Thread(target=run, daemon=True).start()


def shards_evenly(views_total, world: int) -> bool:
    """Can this workload run on `world` ranks?  Weak scaling (views_total None: one view per rank) always; a fixed
    number of views per step only when every rank gets the same number (a rank with fewer views - or none - would have
    to mirror the other ranks' exchanges one for one, including those of their capture warm-up)."""
    return views_total is None or (views_total >= world and views_total % world == 0)


def _dist():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return world, rank, local


def workload_string(cfg, K=5):
    """Identical in both arms: the scene, not the implementation."""
    return (f'{cfg.name}: {cfg.P} Gaussians, {cfg.M} joints, {cfg.W}x{cfg.H}, SH degree 3, K={K} LBS mode W, '
            f'{"fwd+bwd" if cfg.backward else "forward only"}')


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


def algorithmic_bytes(P, M, K, C, R, W, H):
    """SURVEY.md 8(d) per-unit figures x the units one launch processes."""
    return {
        'fk_lbs_fwd_kernel': P * (12 + 4 * K + 40) + P * 12 * K,  # xyz in, K sp_W gathers, d_* out, weights+idx out
        'assemble_fwd_kernel': P * (44 + 40 + 44),
        'preprocess_scan_kernel': P * (44 + 12 * C + 75) + P * 8 + R * 12,  # + scan r/w + key emission
        'duplicate_keys_kernel': P * 8 + R * 12,
        # fused per-Gaussian forward: canonical parameters + K sp_W gathers + SH in; assembled Gaussians, LBS outputs for
        # the backward, geometry records (incl. the 16-byte culling record), scan words and the R keys / values out
        'deform_preprocess_kernel': P * (44 + 4 * K + 12 * C) + P * (44 + 16 + 12 * K) + P * (75 + 16 + 8) + R * 12,
        'fk_table_kernel': M * (44 + 28 + 96),
        'tile_plan_kernel': (W // 16 + 1) * (H // 16 + 1) * 4 + (W // 16) * (H // 16) * 24,
        'tile_scatter_kernel': R * (12 + 8),          # key + value in, (depth, id) word out
        'tile_sort_kernel': R * (8 + 12),             # word in, sorted key + value out (both sort kernels together)
        'tile_sort_large_kernel': 0,
        'composite_fwd_kernel': R * 44 + H * W * 28,
        'composite_bwd_kernel': R * (44 + 40) + H * W * (20 + 8),
        'preprocess_bwd_kernel': P * (44 + 12 * C + 75 + 40) + P * (12 + 12 + 16 + 4 + 12 + 12 * C),
        'assemble_bwd_kernel': P * (44 + 44 + 44),
        'lbs_bwd_kernel': P * (40 + 12 + 12 * K) + P * 4 * M,
        'fk_bwd_kernel': M * 44 * 4,
    }


def cuda_time_ms(fn, K, flush=None):
    """Sum of per-call CUDA-event times over K calls (L2 flushed between calls, outside the event pairs)."""
    evs = []
    for _ in range(K):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs)


# ---------------------------------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------------------------------
class Runner:
    """One workload on one rank: the scene, the rank's views, the captured step graph(s)."""

    def __init__(self, args, name, world, rank, dev, views_total=None):
        import torch.distributed as dist
        from sk_gs_b200 import scene as S
        from sk_gs_b200.dist import GradArena, SymmGradArena, allreduce_max_, shard_views
        from sk_gs_b200.pipeline import HotPath
        self.args, self.world, self.rank, self.dev = args, world, rank, dev
        cfg = self.cfg = S.CONFIGS[name]
        self.strong = views_total is not None          # a fixed number of views per step, sharded over the ranks
        V = views_total if self.strong else world      # weak scaling: one view per rank
        if V % world != 0:
            # a rank without views (or with fewer than the others) would have to mirror the exchanges of the other ranks'
            # capture warm-up one for one - not supported: fail on every rank alike instead of dead-locking a barrier
            raise RuntimeError(f'{name}: {V} views per step cannot be sharded evenly over {world} ranks')
        self.V = V
        sc = self.sc = S.make_scene(cfg, views=V)
        self.my_views = shard_views(V, world, rank) if self.strong else [rank]
        hp = self.hp = HotPath(sc, dev, mode='W', merged_sh=True, requires_grad=False)
        H, W = cfg.H, cfg.W
        nv = max(len(self.my_views), 1)
        gen = torch.Generator().manual_seed(1234 + rank)
        self.dL_host = [(torch.randn(3, H, W, generator=gen) / (3 * H * W * V)).pin_memory() for _ in range(nv)]
        self.dL_dev = [t.to(dev) for t in self.dL_host]
        # per-step host inputs of the e2e path: what the joint MLP would emit + the camera(s)
        self.joint_host = {n: getattr(sc, n).clone().pin_memory() for n in ('sk_r', 'sk_d_rot', 'sk_d_scale')}
        self.uploads = [(hp.params[n].data, t) for n, t in self.joint_host.items()]
        for v in self.my_views:
            cam, rs = sc.cameras[v], hp.settings[v]
            self.uploads += [(rs.viewmatrix, cam.viewmatrix.pin_memory()), (rs.projmatrix, cam.projmatrix.pin_memory()),
                             (rs.campos, cam.campos.clone().pin_memory())]
        self.result_host = torch.zeros(1).pin_memory()
        # flat fp32 gradient arena: the backward kernels write into it directly (no packing copies); sp_W travels in
        # compact [P, K] form (the KNN pattern is identical on every rank); SH gradients as one [P, 16, 3] block
        shapes = {'shs': (cfg.P, 16, 3), 'xyz': (cfg.P, 3), 'viewspace_points': (cfg.P, 3), 'scaling': (cfg.P, 3),
                  'rotation': (cfg.P, 4), 'opacity': (cfg.P, 1), 'sp_W': (cfg.P, sc.K), 'joints': (cfg.M, 3),
                  'sk_r': (cfg.M, 4), 'sk_d_rot': (cfg.M, 4), 'sk_d_scale': (cfg.M, 3), 'g_tr': (7,)}
        self.shapes = shapes
        self.arena = self.scratch = None
        self.exchange_kind = None
        multi_view = len(self.my_views) > 1
        if world > 1:
            self.arena = (GradArena if args.allreduce == 'nccl' else SymmGradArena)(shapes, dev, order=list(shapes))
            if getattr(self.arena, 'multimem', False):
                self.exchange_kind = 'NVLS multimem all-reduce kernel over symmetric memory between two signal-pad ' \
                                     'barrier kernels' 
            else:
                self.exchange_kind = 'NCCL all-reduce'
        elif multi_view:
            self.arena = GradArena(shapes, dev, order=list(shapes))
        if multi_view:
            self.scratch = GradArena(shapes, dev, order=list(shapes))
        # one view on every rank (weak scaling, or V == world): the exchange is split and overlapped as in the headline
        self.split_exchange = world > 1 and getattr(self.arena, 'multimem', False) and V == world
        self.side = torch.cuda.Stream(dev) if world > 1 else None
        self.side2 = torch.cuda.Stream(dev) if self.split_exchange else None
        self.graphs = {}
        self._dist, self._allreduce_max = dist, allreduce_max_
        self._idle_radii = torch.zeros(cfg.P, dtype=torch.int32, device=dev)

    # ---- the one exchange step of a data-parallel iteration
    def exchange(self, out, grads):
        if self.world == 1 or os.environ.get('SKGS_BENCH_NO_EXCHANGE') == '1':  # diagnostic: view imbalance alone
            return
        a = self.arena
        if self.split_exchange:  # the rasterizer-side blocks were reduced under the LBS / FK backward (mid_backward)
            a.allreduce_range(a.block_start('sp_W'), a.flat_padded.numel(), channel=1)
        else:
            a.allreduce(chunks=1)
            self._allreduce_max(out['radii'])

    def after_forward(self, radii):
        """radii are final after the forward: their MAX all-reduce runs on a side stream under the whole backward."""
        main = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self._allreduce_max(radii)
        return lambda: main.wait_stream(self.side)

    def mid_backward(self):
        """Called when every rasterizer-side gradient (SH, means, scales, rotations, opacity: 92 % of the bytes) is
        final: their in-switch reduction runs on a side stream while the LBS and FK backward kernels execute."""
        main = torch.cuda.current_stream(self.dev)
        if os.environ.get('SKGS_BENCH_NO_EXCHANGE') == '1':
            return lambda: None
        self.side2.wait_stream(main)
        with torch.cuda.stream(self.side2):
            # no exit barrier: exchange() runs a full allreduce_range on every rank after the join below
            self.arena.allreduce_range(0, self.arena.block_start('sp_W'), channel=0, exit_barrier=False)
        return lambda: main.wait_stream(self.side2)

    def capture(self, e2e: bool):
        hp, views = self.hp, self.my_views
        if not views:  # strong scaling with more ranks than views: this rank only takes part in the exchange
            return None
        multi = len(views) > 1
        kw = dict(compact_sp_W=True, uploads=self.uploads if e2e else None, arena=self.arena,
                  epilogue=self.exchange if self.world > 1 else None)
        if multi:
            return hp.capture_step(views, self.dL_dev, dL_host=self.dL_host if e2e else None, scratch=self.scratch,
                                   **kw)
        return hp.capture_step(views[0], self.dL_dev[0], dL_host=self.dL_host[0] if e2e else None,
                               after_forward=self.after_forward if self.split_exchange else None,
                               mid_backward=self.mid_backward if self.split_exchange else None, **kw)

    def step(self, e2e: bool):
        key = 'e2e' if e2e else 'dev'
        if key not in self.graphs:
            self.graphs[key] = self.capture(e2e)
            self.launches = getattr(self.hp, 'launches_per_step', 0)
        g = self.graphs[key]
        if g is None:
            self.arena.flat.zero_()
            self.exchange({'radii': self._idle_radii}, None)
            return None
        graph, out, grads = g
        graph.replay()  # with N > 1 the gradient exchange is part of the same graph
        if e2e:
            self.result_host.copy_((out['images'] * self.dL_dev[-1]).sum().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if self.hp.overflowed():
                raise RuntimeError('binning capacity of the captured graph exceeded')
            return float(self.result_host[0])
        return None

    def timed(self, e2e: bool, K: int, Wu: int, flush):
        dist, world = self._dist, self.world
        for _ in range(Wu):
            self.step(e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = cuda_time_ms(lambda: self.step(e2e), K, flush)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if self.hp.overflowed():
            raise RuntimeError('binning capacity of the captured graph exceeded')
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def h2d_bytes(self):
        return int(sum(t.numel() for t in self.dL_host) * 4 + sum(s.numel() * 4 for _, s in self.uploads))

    def exchange_selfcheck(self):
        """One step's gradients reduced by the exchange the benchmark uses (in-switch multimem kernel over symmetric
        memory, split over two streams) against a plain NCCL all-reduce of the same per-rank gradients."""
        if self.world == 1:
            return None
        dist = self._dist
        hp = self.hp
        from sk_gs_b200.dist import GradArena
        local = GradArena(self.shapes, self.dev, order=list(self.shapes))
        radii = self._idle_radii.clone()
        if len(self.my_views) > 1:
            scratch = GradArena(self.shapes, self.dev, order=list(self.shapes))
            out, _ = hp.step_views(self.my_views, self.dL_dev, local, scratch, compact_sp_W=True)
            radii = out['radii'].clone()
        elif self.my_views:
            out, _ = hp.step_grads(self.my_views[0], self.dL_dev[0], compact_sp_W=True, arena=local)
            radii = out['radii'].clone()
        ref = local.flat.clone()
        dist.all_reduce(ref)
        dist.all_reduce(radii, op=dist.ReduceOp.MAX)
        err, reps = 0.0, int(os.environ.get('SKGS_CHECK_REPS', '3'))
        for _ in range(reps):  # an intermittent race would not show in a single replay
            self.step(False)  # the benchmarked graph: same inputs, same parameters
            torch.cuda.synchronize()
            got = self.arena.flat[:ref.numel()].clone()
            err = max(err, float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
            dist.barrier()
        worst, bad = None, {}  # the block of the arena with the largest deviation (diagnostic)
        for name in self.shapes:
            b0 = self.arena.block_start(name)
            n = 1
            for d_ in self.shapes[name]:
                n *= d_
            e_ = float((got[b0:b0 + n] - ref[b0:b0 + n]).abs().max() / ref[b0:b0 + n].abs().max().clamp_min(1e-30))
            if e_ > 1e-5:
                bad[name] = e_
            if worst is None or e_ > worst[1]:
                worst = (name, e_)
        radii_ok = True
        if self.graphs.get('dev') is not None:
            radii_ok = bool(torch.equal(self.graphs['dev'][1]['radii'], radii))
        t = torch.tensor([err, 0.0 if radii_ok else 1.0], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {'max_rel_err_vs_nccl_allreduce': float(t[0]), 'radii_max_equal': bool(t[1] == 0),
                'ok': bool(t[0] <= 1e-5 and t[1] == 0), 'replays_checked': reps, 'worst_block_rank0': list(worst), 'bad_blocks_rank0': bad, 'what': 'arena after the in-graph exchange vs '
                'dist.all_reduce(SUM) of the same per-rank gradients; fp32 sums in a different order'}


def per_kernel_profile(hp, view, dL, flush, cfg, K_lbs):
    """Per-kernel device times (separate eager pass, a CUDA-event pair around every launch)."""
    from sk_gs_b200 import _lib
    from sk_gs_b200 import diff_gaussian_rasterization as DGR
    for _ in range(2):
        hp.step_grads(view, dL)
    _lib.profile_enable(True)
    nprof = 5
    for _ in range(nprof):
        flush.zero_()
        hp.step_grads(view, dL)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    R = int(DGR.last_header_words(hp.device)[0]) & 0xffffffff
    ab = algorithmic_bytes(cfg.P, cfg.M, K_lbs, 16, R, cfg.W, cfg.H)
    peak, peak_src = measured_peak_hbm()
    kern = {}
    for name, (n, us) in prof.items():
        per = us / n
        gbs = ab.get(name, 0) / (per * 1e-6) / 1e9 if per > 0 else 0.0
        kern[name] = {'launches_per_step': n / nprof, 'us_per_launch': round(per, 3),
                      'us_per_step': round(us / nprof, 3), 'algorithmic_GBps': round(gbs, 1),
                      'frac_of_peak': round(gbs / peak, 4)}
    return kern, R, ab, peak, peak_src


def run_ours(args, world, rank, local):
    import torch.distributed as dist
    from sk_gs_b200 import scene as S

    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    name = args.workload
    cfg = S.CONFIGS[name]
    K = args.steps
    sampler = ClockSampler(local)

    # ------------------------------------------------------------------------------------------------ headline
    strong_views = {'c3': 8, 'c4': 4}.get(name)
    run = Runner(args, name, world, rank, dev, views_total=strong_views)
    if rank == 0:
        sampler.start()
    _STATE['section'] = 'headline'
    ms_dev = run.timed(False, K, args.warmup, flush)
    clocks = sampler.stop() if rank == 0 else None
    launches = K * getattr(run, 'launches', 0)
    _STATE['section'] = 'e2e'
    ms_e2e = run.timed(True, K, max(args.warmup, 3), flush)
    _STATE['section'] = 'exchange self-check'
    check = run.exchange_selfcheck()
    per_step_units = 1 if run.strong else run.V  # weak scaling: every rank's view is one step of the metric
    line = {
        'metric': METRIC, 'value': round(per_step_units * K / (ms_dev * 1e-3), 2),
        'unit': 'steps/s', 'n_gpus': world, 'steps': K, 'warmup': args.warmup, 'ms_per_step': round(ms_dev / K, 4),
        'higher_is_better': True, 'scaling': 'strong' if run.strong else 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': workload_string(cfg), 'views_per_step': run.V, 'views_per_rank': len(run.my_views),
                   'parallelism': f'view-sharded dp{world}' + (f' + {run.exchange_kind}' if world > 1 else ''),
                   'l2_flush': '256 MiB memset between steps, outside the per-step CUDA-event pairs',
                   'sh_layout': 'one [P,16,3] parameter (f_dc | f_rest interleaved per Gaussian); `dropin` keeps them apart',
                   'launch': 'CUDA graph replay (fixed binning capacity per graph, overflow flag checked)'},
        'clocks': clocks,
        'e2e': {'value': round(per_step_units * K / (ms_e2e * 1e-3), 2), 'unit': 'steps/s',
                'ms_per_step': round(ms_e2e / K, 4), 'h2d_bytes_per_step': run.h2d_bytes(), 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
    }
    if check is not None:
        line['exchange_check'] = check
    _STATE['line'], _STATE['section'] = line, 'per-kernel profile'

    extra_ok = not args.headline_only
    # ---- per-kernel table + roofline (rank 0, eager pass with events around every launch)
    kern = {}
    if rank == 0 and run.my_views:
        hp, view = run.hp, run.my_views[0]
        kern, R, ab, peak, peak_src = per_kernel_profile(hp, view, run.dL_dev[0], flush, cfg, run.sc.K)
        dom = max(kern, key=lambda k: kern[k]['us_per_step'])
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r2_roofline_traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f).get(name, {}).get(dom)
            if tj:
                traffic = tj['dram_bytes_read'] + tj['dram_bytes_write']
                traffic_src = tj.get('source')
        line['config']['num_rendered'] = R
        line['roofline'] = {'bound': 'hbm', 'kernel': dom, 'achieved': kern[dom]['algorithmic_GBps'], 'peak': peak,
                            'peak_source': peak_src, 'unit': 'GB/s', 'frac': kern[dom]['frac_of_peak'],
                            'traffic': traffic, 'traffic_source': traffic_src, 'algorithmic_bytes': ab.get(dom),
                            'note': 'algorithmic bytes (SURVEY 8d) / CUDA-event time of the kernel, measured in this '
                                    'run; compositing is FP32-issue bound (about 145 flop/B), see DESIGN.md'}
        line['kernels'] = kern
        # stages of one eager step (the reference arm prints the same keys)
        line['stages'] = stage_times(hp, view, run.dL_dev[0], flush)
        line['raster_only'] = {'ms_per_step': round(line['stages']['raster_fwd_ms'] + line['stages']['raster_bwd_ms'], 4),
                               'what': 'rasterizer forward + backward of the same view, eager launches'}

    # ---- forward-only FPS (the reference's test.py --fps protocol: CUDA events around N renders)
    if rank == 0 and extra_ok and run.my_views:
        g, out = run.hp.capture_render(run.my_views[0])
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        nf = max(50, K)
        ms = cuda_time_ms(g.replay, nf, flush)
        line['render_fps'] = {'value': round(nf / (ms * 1e-3), 1), 'unit': 'frames/s',
                              'note': 'forward only (FK+LBS+assembly+rasterize), 1 GPU, CUDA graph'}

    # ---- the drop-in operator API, eager, separate f_dc / f_rest (what a user of the reference switches to)
    if rank == 0 and extra_ok and world == 1:
        try:
            line['dropin'] = dropin_block(args, run.sc, cfg, dev, flush)
        except Exception as e:  # noqa: never let an extra section take the contract line down
            line['dropin'] = {'error': f'{type(e).__name__}: {e}'[:300]}
    del run
    torch.cuda.empty_cache()

    # ---- other named shapes (device-resident graph replay; fewer steps)
    if extra_ok and not args.no_workloads and name == 'c2':
        wl = {}
        for wname, views_total in (('ns', None), ('c3', 8), ('c4', 4)):
            if not shards_evenly(views_total, world):
                wl[wname] = {'skipped': f'{views_total} views per step do not shard evenly over {world} ranks'}
                continue
            _STATE['section'] = f'workload {wname}'
            try:
                wl[wname] = secondary_workload(args, wname, views_total, world, rank, dev, flush)
            except Exception as e:  # noqa
                wl[wname] = {'error': f'{type(e).__name__}: {e}'[:300]}
                if world > 1:
                    raise  # a rank that drops out of a collective would hang the others
            torch.cuda.empty_cache()
        _STATE['section'] = 'workload c5'
        try:
            wl['c5'] = c5_fps(args, world, rank, dev, flush)
        except Exception as e:  # noqa
            wl['c5'] = {'error': f'{type(e).__name__}: {e}'[:300]}
            if world > 1:
                raise
        torch.cuda.empty_cache()
        line['workloads'] = wl

    # ---- widening rows: the complete iteration MLP -> render -> L1+SSIM -> backward -> Adam as one graph
    if rank == 0 and world == 1 and extra_ok and not args.no_iteration:
        try:
            sc = S.make_scene(cfg, views=1)
            line['full_iteration'] = full_iteration(args, sc, cfg, dev, 0, flush, kern)
        except Exception as e:  # noqa
            line['full_iteration'] = {'error': f'{type(e).__name__}: {e}'[:300]}
    if rank == 0 and world == 1 and extra_ok and not args.no_workloads:
        try:
            line['widening'] = widening_rows(dev, 'ours', flush)
        except Exception as e:  # noqa
            line['widening'] = {'error': f'{type(e).__name__}: {e}'[:300]}
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    # ---- CPU baseline: the oracle port on the host cores, bounded sample (same workload, 1 view)
    line['cpu_baseline'] = cpu_baseline(name, steps=args.cpu_steps) if (world == 1 and not args.no_cpu) else None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def stage_times(hp, view, dL, flush, n=10):
    """CUDA-event times of the four stages of one eager step (hand-driven operators)."""
    from sk_gs_b200 import diff_gaussian_rasterization as DGR
    from sk_gs_b200.fk_lbs import (assemble_backward_raw, assemble_forward_raw, fk_lbs_backward_raw, fk_lbs_forward_raw)
    p = hp.params
    acc = {'fk_lbs_fwd_ms': 0.0, 'raster_fwd_ms': 0.0, 'raster_bwd_ms': 0.0, 'fk_lbs_bwd_ms': 0.0}
    sh = p['shs'] if 'shs' in p else torch.cat((p['f_dc'], p['f_rest']), 1)
    for it in range(n + 2):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        with torch.no_grad():
            ev[0].record()
            (d_xyz, d_rot, d_scale, sk_T, weights, indices), c1 = fk_lbs_forward_raw(
                p['xyz'], p['joints'], p['sk_r'], p['sk_d_rot'], p['sk_d_scale'], p['g_tr'], hp.parents, hp.root,
                K=hp.K, mode='W', sp_W=p['sp_W'])
            (points, scales, rotations, opacity), c2 = assemble_forward_raw(p['xyz'], p['scaling'], p['rotation'],
                                                                            p['opacity'], d_xyz, d_rot, d_scale)
            ev[1].record()
            color, depth, alpha, radii, st = DGR.rasterize_forward(hp.settings[view], points, opacity, shs=sh,
                                                                   scales=scales, rotations=rotations, quat_wxyz=False)
            ev[2].record()
            g = DGR.rasterize_backward(st, dL)
            ev[3].record()
            r = assemble_backward_raw(c2, g['means3D'], g['scales'], g['rotations'], g['opacities'],
                                      need=[False, True, True, True, True, True, True])
            fk_lbs_backward_raw(c1, r[4], r[5], r[6])
            ev[4].record()
        torch.cuda.synchronize()
        if it >= 2:
            for k, (a, b) in zip(acc, zip(ev[:-1], ev[1:])):
                acc[k] += a.elapsed_time(b)
    return {k: round(v / n, 4) for k, v in acc.items()}


def dropin_block(args, sc, cfg, dev, flush):
    """The step exactly as a user of the reference would run it after switching: `fk_lbs` -> `assemble` ->
    `render_gs_offical` (sk_gs_b200's drop-ins for sk_stage / the assembly lines / the renderer adapter) under torch
    autograd, f_dc and f_rest as separate leaf parameters concatenated per step (gaussian_splatting.py:155-157),
    eager launches, `.grad` filled by the autograd engine."""
    from sk_gs_b200.pipeline import HotPath
    hp = HotPath(sc, dev, mode='W', merged_sh=False, requires_grad=True)
    H, W = cfg.H, cfg.W
    dL_host = (torch.randn(3, H, W, generator=torch.Generator().manual_seed(1234)) / (3 * H * W)).pin_memory()
    dL = dL_host.to(dev)
    joint_host = {n: getattr(sc, n).clone().pin_memory() for n in ('sk_r', 'sk_d_rot', 'sk_d_scale')}
    result_host = torch.zeros(1).pin_memory()

    def step(e2e):
        if e2e:
            for n, t in joint_host.items():
                hp.params[n].data.copy_(t, non_blocking=True)
            dL.copy_(dL_host, non_blocking=True)
        hp.zero_grad()
        out = hp.render(0)
        out['images'].backward(dL)
        if e2e:
            result_host.copy_((out['images'].detach() * dL).sum().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    K = max(20, min(args.steps, 100))
    res = {}
    for key, e2e in (('value', False), ('e2e', True)):
        for _ in range(5):
            step(e2e)
        torch.cuda.synchronize()
        ms = cuda_time_ms(lambda: step(e2e), K, flush)
        res[key] = round(K / (ms * 1e-3), 2)
        res[('ms_per_step' if not e2e else 'e2e_ms_per_step')] = round(ms / K, 4)
    res.update(unit='steps/s', steps=K,
               what='fk_lbs -> assemble -> render_gs_offical under torch autograd, separate f_dc / f_rest parameters + '
                    'cat per step, eager launches (no CUDA graph), gradients in .grad')
    return res


def secondary_workload(args, name, views_total, world, rank, dev, flush):
    from sk_gs_b200 import scene as S
    cfg = S.CONFIGS[name]
    run = Runner(args, name, world, rank, dev, views_total=views_total)
    K = max(10, min(args.steps, 30))
    ms = run.timed(False, K, 3, flush)
    ms_e2e = run.timed(True, K, 3, flush) if name == 'ns' else None
    steps_per_s = K / (ms * 1e-3)
    res = {'workload': workload_string(cfg), 'value': round(steps_per_s * (1 if run.strong else world), 2),
           'unit': 'steps/s', 'ms_per_step': round(ms / K, 4), 'steps': K,
           'views_per_step': run.V, 'views_per_rank': len(run.my_views),
           'view_steps_per_s': round(steps_per_s * run.V, 2),
           'scaling': 'strong' if run.strong else 'weak', 'n_gpus': world}
    if ms_e2e is not None:
        res['e2e'] = {'value': round(K / (ms_e2e * 1e-3) * world, 2), 'unit': 'steps/s',
                      'h2d_bytes_per_step': run.h2d_bytes(), 'd2h_bytes_per_step': 4}
    if rank == 0 and run.my_views:
        w = run.graphs['dev'][1]['_header_words']
        res['num_rendered'] = int((w[0] if isinstance(w, list) else w)[0])
    chk = run.exchange_selfcheck()
    if chk is not None:
        res['exchange_check'] = chk
    del run
    return res


def c5_fps(args, world, rank, dev, flush):
    """Reposing render stress (gui.py:575-591, test.py:102-123): 3M Gaussians, 64 joints, 1920x1080, 64 novel poses /
    views per sweep sharded over the ranks, forward only (FK + LBS + assembly + rasterize per frame), no collective."""
    import torch.distributed as dist
    from sk_gs_b200 import scene as S
    from sk_gs_b200.dist import shard_views
    from sk_gs_b200.pipeline import HotPath
    cfg = S.CONFIGS['c5']
    poses = 64
    mine = shard_views(poses, world, rank)
    ncam = min(len(mine), 2)  # two distinct cameras per rank, cycled (each graph pins its own arenas: ~1.5 GB)
    sc = S.make_scene(cfg, views=max(ncam, 1))
    hp = HotPath(sc, dev, mode='W', merged_sh=True, requires_grad=False)
    graphs = [hp.capture_render(v, headroom=1.2)[0] for v in range(ncam)]

    def sweep():
        for i in range(len(mine)):
            graphs[i % ncam].replay()

    sweep()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = 2
    ms = cuda_time_ms(sweep, n)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if hp.overflowed():
        raise RuntimeError('binning capacity exceeded')
    return {'workload': workload_string(cfg), 'value': round(n * poses / (ms * 1e-3), 2), 'unit': 'frames/s',
            'ms_per_frame_per_gpu': round(ms / (n * max(len(mine), 1)), 4), 'poses_per_sweep': poses, 'n_gpus': world,
            'scaling': 'strong', 'collective': 'none (forward only)'}


def full_iteration(args, sc, cfg, dev, view, flush, kern):
    """joint MLP -> render -> fused L1 + SSIM loss -> backward -> one-launch Adam on one GPU
    (sk_gs_b200.train.TrainLoop), replayed as a CUDA graph; then an eager profiled pass for the per-kernel times of the
    loss, network and optimizer kernels."""
    from sk_gs_b200 import _lib
    from sk_gs_b200.pipeline import HotPath
    from sk_gs_b200.train import TrainLoop
    H, W = cfg.H, cfg.W
    # joint rotations come from the joint-rotation network (8f-1); head_std 0.02 gives rotations of ~10 degrees, the
    # regime of the synthetic scene (the reference's 1e-6 init would make every joint rotation the identity)
    hp = HotPath(sc, dev, mode='W', requires_grad=False, merged_sh=True, joint_mlp=True, head_std=0.02)
    # Adam moves every network weight by +-lr per step; 1e-5 instead of the reference's 1e-3 keeps the synthetic scene
    # (and with it R, the work per iteration) stationary over the timed replays - same kernels, same bytes
    loop = TrainLoop(hp, lrs={'theta': 1e-5})
    # target = the scene's own rendering + noise: the near-converged regime
    with torch.no_grad():
        target = hp.forward_raw(view)[0]['images'].detach().clone()
    target = (target + 0.05 * torch.randn(3, H, W, generator=torch.Generator().manual_seed(99)).to(dev)).clamp_(0, 1)
    loop.capture(view, target, headroom=2.0)
    K = max(20, min(args.steps, 200))
    for _ in range(5):
        loop.replay(wait=False)
    torch.cuda.synchronize()
    ms = cuda_time_ms(lambda: loop.replay(wait=False), K, flush) / K
    overflow = hp.overflowed()
    R_last = int(loop.out['_header_words'][0])
    terms = [round(float(x), 6) for x in loop.out['loss_terms'].cpu()]
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
            'loss_terms_last': terms, 'num_rendered_last': R_last, 'overflow': bool(overflow),
            'recaptures': loop.recaptures}


# ---------------------------------------------------------------------------------------------------------------------
def widening_rows(dev, impl: str, flush=None):
    """SURVEY 8 rows f-4 (sp-stage LBS at the default 512 superpoints) and f-3 (densification bookkeeping) at the
    headline Gaussian count, eager calls, CUDA-event timed.  impl 'ours': libskgs_b200.so through sk_gs_b200.sp_lbs /
    sk_gs_b200.densify; impl 'reference': the torch op sequences the reference executes for the same functions
    (networks/sk_gs.py:751-828, networks/gaussian_splatting.py:589-660) as restated in oracle/ - lietorch / pytorch3d
    are not installable, so this is kind "port" (its rigid action and KNN are plain torch ops)."""
    P, M, K = 100_000, 512, 5
    g = torch.Generator().manual_seed(7)
    r = lambda *s, scale=1.0, shift=0.0: (torch.randn(*s, generator=g) * scale + shift).to(dev)  # noqa: E731
    points, sp_points, sp_t = r(P, 3, scale=0.5), r(M, 3, scale=0.5), r(M, 3, scale=0.05)
    sp_r = torch.nn.functional.normalize(r(M, 4, scale=0.2) + torch.tensor([0, 0, 0, 1.0], device=dev), dim=-1)
    sp_rot = torch.nn.functional.normalize(r(M, 4, scale=0.2) + torch.tensor([0, 0, 0, 1.0], device=dev), dim=-1)
    sp_scale, sp_W = r(M, 3, scale=0.01), r(P, M)
    cots = [r(P, 3), r(P, 4), r(P, 3)]
    out = {}
    if impl == 'ours':
        from sk_gs_b200.sp_lbs import sp_warp_backward_raw, sp_warp_forward_raw
        keep = {}

        def fwd():
            keep['o'], keep['c'] = sp_warp_forward_raw(points, sp_points, sp_t, sp_r, sp_rot, sp_scale, K=K, mode='W',
                                                       sp_W=sp_W, method='LBS')

        def bwd():
            sp_warp_backward_raw(keep['c'], *cots, compact_sp_W=True)
    else:
        from oracle import fk_lbs as OF
        leaves = [t.clone().requires_grad_() for t in (sp_points, sp_t, sp_r, sp_rot, sp_scale, sp_W)]
        keep = {}

        def fwd():
            keep['o'] = OF.sp_stage(points, leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], K=K, mode='W',
                                    sp_W=leaves[5], method='LBS')

        def bwd():
            o = keep['o']
            torch.autograd.grad(sum((a * c).sum() for a, c in zip(o[:3], cots)), leaves, allow_unused=True)
    reps = 20 if impl == 'ours' else 3
    for _ in range(2):
        fwd(); bwd()
    torch.cuda.synchronize()
    t_f = cuda_time_ms(fwd, reps, flush) / reps
    fwd()
    t_fb = cuda_time_ms(lambda: (fwd(), bwd()), reps, flush) / reps
    out['sp_stage'] = {'what': f'sp-stage LBS forward + backward, P={P} Gaussians, M={M} superpoints, K={K}, mode W, '
                               f'method LBS, sep_rot', 'forward_us': round(t_f * 1e3, 1),
                       'forward_backward_us': round(t_fb * 1e3, 1)}
    # ---- densification: clone + split + prune of the whole per-Gaussian state (6 tensors + 2 Adam moments each)
    params = dict(xyz=r(P, 3, scale=0.5), shs=r(P, 16, 3), scaling=r(P, 3, shift=-3.6), rotation=r(P, 4),
                  opacity=r(P, 1, scale=3.0, shift=-2.0), sp_W=r(P, 32))
    trip = {n: (p, torch.zeros_like(p), torch.zeros_like(p)) for n, p in params.items()}
    accum = (torch.rand(P, generator=g) * 1.2e-3).to(dev)
    denom = torch.randint(0, 4, (P,), generator=g).float().to(dev)
    radii = torch.randint(0, 40, (P,), generator=g).float().to(dev)
    noise = r(2 * P, 3)
    kw = dict(do_densify=True, do_prune=True, grad_threshold=0.0002, densify_extent=0.02, min_opacity=0.005,
              max_screen_size=20.0, prune_extent=0.2)
    if impl == 'ours':
        from sk_gs_b200.densify import DensifyStats, add_densification_stats, densify_and_prune
        st = DensifyStats(P, dev)
        st.grad_accum, st.denom, st.max_radii2D = accum, denom, radii
        res = {}

        def dens():
            res['r'] = densify_and_prune(trip, st, noise=noise, **kw)
        vs, rad = r(P, 3, scale=1e-4), torch.randint(0, 40, (P,), generator=g).int().to(dev)
        st2 = DensifyStats(P, dev)
        t_s = cuda_time_ms(lambda: add_densification_stats(st2, rad, vs), 20, flush) / 20
        n_new = lambda: res['r'].counts['n_new']  # noqa: E731
    else:
        from oracle import densify as OD
        res = {}

        def dens():
            res['r'] = OD.densify_and_prune(trip, accum.clone(), denom.clone(), radii.clone(), noise=noise, **kw)
        vs, rad = r(P, 3, scale=1e-4), torch.randint(0, 40, (P,), generator=g).int().to(dev)
        a2, d2, m2 = torch.zeros(P, device=dev), torch.zeros(P, device=dev), torch.zeros(P, device=dev)
        t_s = cuda_time_ms(lambda: OD.add_densification_stats(a2, d2, m2, rad, vs), 5, flush) / 5
        n_new = lambda: int(res['r'][0]['xyz'][0].shape[0])  # noqa: E731
    dens()
    torch.cuda.synchronize()
    reps = 10 if impl == 'ours' else 3
    t_d = cuda_time_ms(dens, reps, flush) / reps
    out['densify'] = {'what': f'clone + split + prune of P={P} Gaussians (6 parameter tensors incl. a 32-column skinning '
                              f'table, each with both Adam moments) including the host read of the new count',
                      'P_new': n_new(), 'densify_and_prune_us': round(t_d * 1e3, 1),
                      'per_step_statistics_us': round(t_s * 1e3, 1)}
    return out


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


# ---------------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------------
class RefRunner:
    """The reference's own CUDA rasterizer (oracle/_ref, compiled unmodified, colmap=True) + the torch-op FK/LBS of
    oracle/fk_lbs.py on the GPU (lietorch / pytorch3d are not installable), driven through torch autograd like the
    reference does."""

    def __init__(self, name, dev, views):
        from oracle import fk_lbs as OF
        from oracle import ref_ext
        from sk_gs_b200 import scene as S
        self.OF, self.ref_ext, self.dev = OF, ref_ext, dev
        cfg = self.cfg = S.CONFIGS[name]
        sc = self.sc = S.make_scene(cfg, views=views)
        names = ['xyz', 'scaling', 'rotation', 'opacity', 'f_dc', 'f_rest', 'sp_W', 'joints', 'sk_r', 'sk_d_rot',
                 'sk_d_scale', 'g_tr']
        self.p = {n: getattr(sc, n).to(dev).clone().requires_grad_(True) for n in names}
        self.parents = sc.parents.to(dev).long()
        H, W = cfg.H, cfg.W
        self.dL_host = (torch.randn(3, H, W, generator=torch.Generator().manual_seed(1234)) / (3 * H * W)).pin_memory()
        self.dL = self.dL_host.to(dev)
        self.joint_host = {n: getattr(sc, n).clone().pin_memory() for n in ('sk_r', 'sk_d_rot', 'sk_d_scale')}
        self.result_host = torch.zeros(1).pin_memory()

    def deform(self):
        p, OF, sc = self.p, self.OF, self.sc
        out = OF.sk_stage(p['xyz'], p['joints'], p['sk_r'], p['sk_d_rot'], p['sk_d_scale'], p['g_tr'], self.parents,
                          sc.root, K=sc.K, mode='W', sp_W=p['sp_W'])
        return OF.assemble(p['xyz'], p['scaling'], p['rotation'], p['opacity'], p['f_dc'], p['f_rest'], *out[:3])

    def step(self, e2e, view=0, backward=True):
        p = self.p
        for t in p.values():
            t.grad = None
        if e2e:
            for n, t in self.joint_host.items():
                p[n].data.copy_(t, non_blocking=True)
            self.dL.copy_(self.dL_host, non_blocking=True)
        pts, scl, rot, op, sh = self.deform()
        r = self.ref_ext.render(pts, op, scl, rot, sh, self.sc.cameras[view])
        if backward:
            r['images'].backward(self.dL)
        if e2e:
            self.result_host.copy_((r['images'].detach() * self.dL).sum().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def stages(self, flush, n=5):
        """CUDA-event times of {FK/LBS fwd, raster fwd, raster bwd, FK/LBS bwd} (autograd split at the rasterizer)."""
        acc = {'fk_lbs_fwd_ms': 0.0, 'raster_fwd_ms': 0.0, 'raster_bwd_ms': 0.0, 'fk_lbs_bwd_ms': 0.0}
        for it in range(n + 1):
            flush.zero_()
            for t in self.p.values():
                t.grad = None
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            outs = self.deform()
            ev[1].record()
            leaves = [o.detach().requires_grad_(True) for o in outs]
            pts, scl, rot, op, sh = leaves
            r = self.ref_ext.render(pts, op, scl, rot, sh, self.sc.cameras[0])
            ev[2].record()
            r['images'].backward(self.dL)
            ev[3].record()
            torch.autograd.backward(list(outs), [l.grad for l in leaves])
            ev[4].record()
            torch.cuda.synchronize()
            if it >= 1:
                for k, (a, b) in zip(acc, zip(ev[:-1], ev[1:])):
                    acc[k] += a.elapsed_time(b)
        return {k: round(v / n, 4) for k, v in acc.items()}


def run_reference(args, world, rank, local):
    """Reference arm.  The reference has NO CPU implementation of this path: its rasterizer is a CUDA extension.  When the
    extension compiled here from /root/reference (oracle/_ref) is loadable and a GPU is present, it is what runs
    (kind "reference", device cuda) with the torch-op FK/LBS of oracle/fk_lbs.py on the GPU; otherwise the oracle port on
    the host cores (kind "port").  --ref-device cpu forces the latter."""
    if rank != 0:
        return
    from oracle import ref_ext
    from sk_gs_b200 import scene as S
    use_gpu = args.ref_device != 'cpu' and torch.cuda.is_available() and ref_ext.available()
    K = args.steps
    cfg = S.CONFIGS[args.workload]
    if not use_gpu:
        steps = min(K, args.cpu_steps)
        cpu = cpu_baseline(args.workload, steps=steps)
        line = {'impl': 'reference', 'metric': METRIC,
                'value': cpu['value'], 'unit': 'steps/s', 'n_gpus': world, 'steps': steps, 'warmup': 1,
                'ms_per_step': round(1000.0 / cpu['value'], 3), 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': workload_string(cfg)},
                'cpu_baseline': cpu, 'e2e': {'value': cpu['value'], 'unit': 'steps/s', 'h2d_bytes_per_step': 0,
                                             'd2h_bytes_per_step': 0}}
        print(json.dumps(line), flush=True)
        return
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(runner, e2e, K, Wu, **kw):
        for _ in range(Wu):
            runner.step(e2e, **kw)
        torch.cuda.synchronize()
        return cuda_time_ms(lambda: runner.step(e2e, **kw), K, flush)

    rr = RefRunner(args.workload, dev, 1)
    ms_dev = timed(rr, False, K, args.warmup)
    ms_e2e = timed(rr, True, K, max(args.warmup, 3))
    stages = rr.stages(flush)
    line = {'impl': 'reference', 'metric': METRIC,
            'value': round(K / (ms_dev * 1e-3), 2), 'unit': 'steps/s', 'n_gpus': 1, 'steps': K, 'warmup': args.warmup,
            'ms_per_step': round(ms_dev / K, 4), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_string(cfg), 'views_per_step': 1, 'views_per_rank': 1,
                       'l2_flush': '256 MiB memset between steps, outside the per-step CUDA-event pairs',
                       'launch': 'eager, torch autograd (the reference has no graph-captured path)'},
            'cpu_baseline': {'value': round(K / (ms_dev * 1e-3), 2), 'unit': 'steps/s', 'cores': 0, 'kind': 'reference',
                             'device': 'cuda',
                             'sample': "the reference's own CUDA rasterizer (my_ext/_C/src/nerf/gaussian_*.cu compiled "
                                       'unmodified into oracle/_ref, colmap=True) + torch-op FK/LBS on the GPU; the '
                                       'reference has no CPU implementation of this path'},
            'e2e': {'value': round(K / (ms_e2e * 1e-3), 2), 'unit': 'steps/s', 'ms_per_step': round(ms_e2e / K, 4),
                    'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'stages': stages,
            'raster_only': {'ms_per_step': round(stages['raster_fwd_ms'] + stages['raster_bwd_ms'], 4),
                            'what': "the reference's own rasterizer forward + backward alone (oracle/_ref, unmodified); "
                                    'the FK/LBS stages of this arm are torch ops written for this repo (lietorch / '
                                    'pytorch3d cannot be installed) and are reported separately in `stages`'},
            'dropin': {'value': round(K / (ms_dev * 1e-3), 2), 'e2e': round(K / (ms_e2e * 1e-3), 2), 'unit': 'steps/s',
                       'what': 'the reference arm IS the eager autograd path (same numbers as value / e2e)'}}
    del rr
    torch.cuda.empty_cache()
    if not args.headline_only and not args.no_workloads and args.workload == 'c2':
        wl = {}
        for wname, views in (('ns', 1), ('c3', 8), ('c4', 4)):
            try:
                r2 = RefRunner(wname, dev, views)
                Kw = max(3, min(K, 10))

                def multi(e2e, r2=r2, views=views):
                    for v in range(views):
                        r2.step(e2e, view=v)
                for _ in range(2):
                    multi(False)
                torch.cuda.synchronize()
                ms = cuda_time_ms(lambda: multi(False), Kw, flush)
                wl[wname] = {'workload': workload_string(r2.cfg), 'value': round(Kw / (ms * 1e-3), 2), 'unit': 'steps/s',
                             'ms_per_step': round(ms / Kw, 4), 'steps': Kw, 'views_per_step': views, 'n_gpus': 1,
                             'view_steps_per_s': round(Kw * views / (ms * 1e-3), 2)}
                if wname == 'ns':
                    ms2 = cuda_time_ms(lambda: multi(True), Kw, flush)
                    wl[wname]['e2e'] = {'value': round(Kw / (ms2 * 1e-3), 2), 'unit': 'steps/s',
                                        'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
                del r2
            except Exception as e:  # noqa
                wl[wname] = {'error': f'{type(e).__name__}: {e}'[:300]}
            torch.cuda.empty_cache()
        try:
            r5 = RefRunner('c5', dev, 2)
            with torch.no_grad():
                for v in range(2):
                    r5.step(False, view=v, backward=False)
                torch.cuda.synchronize()
                n = 4
                ms = cuda_time_ms(lambda: [r5.step(False, view=v % 2, backward=False) for v in range(n)], 1)
            wl['c5'] = {'workload': workload_string(r5.cfg), 'value': round(n / (ms * 1e-3), 2), 'unit': 'frames/s',
                        'n_gpus': 1, 'collective': 'none (forward only)'}
            del r5
        except Exception as e:  # noqa
            wl['c5'] = {'error': f'{type(e).__name__}: {e}'[:300]}
        torch.cuda.empty_cache()
        line['workloads'] = wl
        if dev.type == 'cuda':
            try:
                line['widening'] = widening_rows(dev, 'reference', flush)
            except Exception as e:  # noqa
                line['widening'] = {'error': f'{type(e).__name__}: {e}'[:300]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', help='c2 (headline) | ns | c1 | c3 (8 views/step, strong scaling) | '
                                                     'c4 (4 views/step)')
    ap.add_argument('--ref-device', default='auto', choices=['auto', 'cpu', 'cuda'])
    ap.add_argument('--cpu-steps', type=int, default=40,
                    help='bounded CPU sample: full fwd+bwd oracle steps of the same workload (about 10-20 s of host time)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-iteration', action='store_true', help='skip the MLP+loss+Adam full-iteration section')
    ap.add_argument('--no-workloads', action='store_true', help='skip the ns / c3 / c4 / c5 section')
    ap.add_argument('--headline-only', action='store_true', help='only value / e2e / kernels (quick runs)')
    ap.add_argument('--watchdog', type=float, default=360.0,
                    help='seconds after which a run that has not finished prints what it has and exits (0: off)')
    ap.add_argument('--allreduce', default='multimem', choices=['multimem', 'nccl'],
                    help='gradient exchange for N > 1: in-switch multimem kernel over symmetric memory, or NCCL')
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
    if args.watchdog > 0:
        _start_watchdog(args.watchdog, rank)
    try:
        run_ours(args, world, rank, local)
        _STATE['done'] = True
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
