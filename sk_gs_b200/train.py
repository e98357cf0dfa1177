"""One complete optimisation iteration on the hot path: render -> L1 + SSIM loss -> backward -> Adam, as a chain of
libskgs_b200.so kernels with no autograd engine in between, optionally captured into one CUDA graph.

What it stands for in the reference: one pass of the training loop body for the `sk` stage
(/root/reference/networks/sk_gs.py:1206-1242 render, :1517-1529 loss, my_ext/framework.py:308-337 backward + optimizer
step) with the joint MLP left out: joint rotations `sk_r` etc. are leaf parameters here.
Learning rates default to the reference's (networks/gaussian_splatting.py:110-117,447-452 times lr 1e-3,
exps/default.yaml:121-126).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from .optim import adam_hyper, adam_step_raw
from .pipeline import HotPath

# lr = 1e-3 (exps/default.yaml:126) x the per-attribute factors of GaussianSplatting.__init__ (gaussian_splatting.py:110-117)
DEFAULT_LRS = {
    'xyz': 1e-3 * 0.16, 'shs': (1e-3 * 2.5, 1e-3 * 2.5 / 20, 48, 3), 'opacity': 1e-3 * 50., 'scaling': 1e-3 * 5.0,
    'rotation': 1e-3 * 1.0, 'sp_W': 1e-3, 'joints': 1e-3 * 0.1, 'sk_r': 1e-3, 'sk_d_rot': 1e-3, 'sk_d_scale': 1e-3,
    'g_tr': 1e-3, 'theta': 1e-3,  # lr * lr_deform_scale (exps/default.yaml:64)
}


class TrainLoop:
    def __init__(self, hp: HotPath, lrs: Optional[Dict[str, object]] = None, betas=(0.9, 0.999), eps: float = 1e-15,
                 lambda_image: float = 0.8, lambda_ssim: float = 0.2, method: str = 'l1', compact_sp_W: bool = True):
        if 'shs' not in hp.params:
            raise RuntimeError('TrainLoop needs HotPath(..., merged_sh=True)')
        if hp.mode != 'W':
            raise RuntimeError('TrainLoop is built for LBS mode "W" (every shipped config, exps/default.yaml:35)')
        self.hp = hp
        self.names = [n for n in DEFAULT_LRS if n in hp.params]
        self.lrs = dict(DEFAULT_LRS)
        self.lrs.update(lrs or {})
        self.betas, self.eps = tuple(betas), eps
        self.loss = dict(w_image=lambda_image, w_ssim=lambda_ssim, method=method)
        self.compact_sp_W = compact_sp_W
        self.exp_avg = {n: torch.zeros_like(hp.params[n]) for n in self.names}
        self.exp_avg_sq = {n: torch.zeros_like(hp.params[n]) for n in self.names}
        self.iteration = 0
        self._hyper_host = torch.zeros(1 + 2 * len(self.names)).pin_memory()
        self._hyper_dev = torch.zeros(1 + 2 * len(self.names), device=hp.device)
        self.graph = None
        self._capture_args = None
        self.recaptures = 0
        self.after_adam = None  # optional hook(out, grads) launched right after the Adam kernel (densify statistics)
        self.extra_state = None  # optional callable -> tensors the hook mutates (restored after capture's warm-up)

    # ------------------------------------------------------------------------------------------------------ pieces
    def _adam(self, out, grads, grad_scale: float = 1.0, dynamic: bool = False):
        p = self.hp.params
        idx = out['_sk'][-1]
        knn = [idx if (n == 'sp_W' and self.compact_sp_W) else None for n in self.names]
        # device-side guard: if the render that produced these gradients overflowed its (fixed) binning capacity the
        # image was empty and the gradients are garbage - the update is skipped on the device, replay() re-captures
        st = out.get('_raster_state')
        adam_step_raw([p[n].data for n in self.names], [grads[n] for n in self.names],
                      [self.exp_avg[n] for n in self.names], [self.exp_avg_sq[n] for n in self.names],
                      [self.lrs[n] for n in self.names], max(self.iteration, 1), self.betas[0], self.betas[1], self.eps,
                      grad_scale=grad_scale, knn_indices=knn, dynamic_hyper=self._hyper_dev if dynamic else None,
                      skip_flag_ptr=None if st is None else st.overflow_ptr)
        if self.after_adam is not None:
            self.after_adam(out, grads)

    # ------------------------------------------------------------------------------------- densification (8 f-3)
    def replace_gaussians(self, tensors):
        """Install a new Gaussian set: `tensors` = {name: (param, exp_avg, exp_avg_sq)} for every per-Gaussian
        parameter (what the reference's change_optimizer + setattr loop does, gaussian_splatting.py:515-563,571-572)."""
        hp = self.hp
        P_new = None
        for n, (p, m, v) in tensors.items():
            if n not in hp.params:
                raise RuntimeError(f'replace_gaussians: unknown parameter {n!r}')
            P_new = p.shape[0] if P_new is None else P_new
            if p.shape[0] != P_new or p.shape[1:] != hp.params[n].shape[1:]:
                raise RuntimeError(f'replace_gaussians: {n} has shape {tuple(p.shape)}')
            hp.params[n] = p.requires_grad_(hp.params[n].requires_grad)
            self.exp_avg[n], self.exp_avg_sq[n] = m, v
        if 'shs' in tensors:
            shs = hp.params['shs']
            hp.params['f_dc'], hp.params['f_rest'] = shs.detach()[:, :1], shs.detach()[:, 1:]
        missing = [n for n in ('xyz', 'shs', 'scaling', 'rotation', 'opacity', 'sp_W')
                   if n in hp.params and hp.params[n].shape[0] != P_new]
        if missing:
            raise RuntimeError(f'replace_gaussians: {missing} still have the old number of Gaussians')

    def recapture_after_resize(self):
        """The number of Gaussians changed: every buffer of the captured graph has the wrong size - capture again."""
        if self.graph is None or self._capture_args is None:
            return
        torch.cuda.synchronize(self.hp.device)
        self.graph = None
        self.capture(**self._capture_args)

    def _set_hyper(self):
        vals = adam_hyper([self.lrs[n] for n in self.names], self.iteration, *self.betas)
        self._hyper_host.copy_(torch.tensor(vals, dtype=torch.float32))

    # ------------------------------------------------------------------------------------------------------- eager
    def step(self, view: int, target: Tensor):
        """One iteration, eager launches.  Returns the outputs of HotPath.step_grads (with 'loss_terms')."""
        self.iteration += 1
        out, grads = self.hp.step_grads(view, None, compact_sp_W=self.compact_sp_W, target=target, loss=self.loss)
        self._adam(out, grads)
        return out

    # ----------------------------------------------------------------------------------------------------- captured
    def capture(self, view: int, target: Tensor, target_host: Optional[Tensor] = None, uploads=None,
                headroom: float = 1.5):
        """Capture the whole iteration (uploads -> render -> loss -> backward -> Adam) into one CUDA graph.  The warm-up
        iterations capture needs are undone (parameters and moments restored), so `replay()` x n == `step()` x n."""
        self._capture_args = dict(view=view, target=target, target_host=target_host, uploads=uploads, headroom=headroom)
        state = [self.hp.params[n].data for n in self.names] + list(self.exp_avg.values()) + \
            list(self.exp_avg_sq.values()) + (list(self.extra_state()) if self.extra_state is not None else [])
        saved = [t.clone() for t in state]
        it = self.iteration
        self.iteration = max(it, 1)
        self._set_hyper()
        self._hyper_dev.copy_(self._hyper_host)
        ups = list(uploads or []) + [(self._hyper_dev, self._hyper_host)]
        self.graph, self.out, self.grads = self.hp.capture_step(
            view, None, headroom=headroom, compact_sp_W=self.compact_sp_W, uploads=ups,
            epilogue=lambda o, g: self._adam(o, g, dynamic=True), target=target, loss=self.loss,
            target_host=target_host)
        for t, c in zip(state, saved):
            t.copy_(c)
        self.iteration = it
        torch.cuda.synchronize(self.hp.device)
        return self.graph

    def overflowed(self) -> bool:
        """Did the most recent COMPLETED replay exceed the binning capacity the graph was captured with?  (Its Adam
        update was skipped on the device, so parameters and moments are intact.)"""
        return self.graph is not None and int(self.out['_header_words'][3]) != 0

    def _recapture(self):
        """R outgrew the captured capacity (scales / positions moved): capture again with head room over the R that
        overflowed.  The reference re-sizes its buffers on every call (gaussian_rasterizer_forward.cu:211-213)."""
        torch.cuda.synchronize(self.hp.device)
        a = dict(self._capture_args)
        a['headroom'] = max(float(a['headroom']), 1.5)
        self.recaptures += 1
        if self.recaptures > 16:
            raise RuntimeError('TrainLoop: binning capacity overflowed 16 captures in a row; the scene is diverging')
        self.capture(**a)

    def replay(self, wait: bool = True):
        """One captured iteration with this iteration's bias corrections / learning rates.

        Overflow of the graph's fixed binning capacity is never silent: the Adam step of such a replay is skipped on
        the device (skgs_adam_step skip flag) and the host re-captures the graph for the larger R, then repeats the
        iteration.  With `wait` (default) the previous replay is synchronised first, so the check is exact and
        replay() x n == step() x n; without it the check looks at whatever replay completed last (the flag persists
        while the parameters stand still), so detection - and the iteration counter - may lag by the replays in flight:
        fine for throughput measurements, not for exact comparisons."""
        if wait:
            torch.cuda.current_stream(self.hp.device).synchronize()
        if self.overflowed():
            torch.cuda.synchronize(self.hp.device)
            self.iteration -= 1  # the overflowed replay did not update anything
            self.out['_header_words'].zero_()
            self._recapture()
        self.iteration += 1
        self._set_hyper()
        self.graph.replay()
        return self.out
