"""Host-side logic of the loss / optimizer mirrors that needs no GPU: layout detection, hyper-parameter tables, param
group handling, and the 'no CPU fallback' rule (CPU tensors must raise, never silently compute)."""
import ctypes as C

import pytest
import torch

from sk_gs_b200 import _lib
from sk_gs_b200 import losses as LS
from sk_gs_b200.optim import Adam, adam_hyper, adam_step_raw


def test_target_layout_detection():
    chw = torch.rand(3, 10, 12)
    t, s = LS._target(chw)
    assert s == 0 and t.data_ptr() == chw.data_ptr()
    hwc = torch.rand(10, 12, 3)
    t, s = LS._target(hwc)
    assert s == 3 and t.data_ptr() == hwc.data_ptr()
    rgba = torch.rand(10, 12, 4)
    t, s = LS._target(rgba[..., :3])  # what sk_gs.py:1525 hands over: read in place with pixel stride 4
    assert s == 4 and t.data_ptr() == rgba.data_ptr()
    t, s = LS._target(rgba)
    assert s == 4
    t, s = LS._target(hwc[None])
    assert s == 3 and t.shape == (10, 12, 3)
    t, s = LS._target(hwc.double())
    assert s == 3 and t.dtype == torch.float32
    with pytest.raises(RuntimeError):
        LS._target(torch.rand(10, 12, 5))
    with pytest.raises(RuntimeError):
        LS._target(torch.rand(10, 12))


def test_image_layout_detection():
    chw = torch.rand(3, 8, 9)
    assert LS._as_chw(chw).data_ptr() == chw.data_ptr()
    hwc_view = chw.permute(1, 2, 0)  # the rasterizer output permuted to HWC (sk_gs.py:1229): maps back without a copy
    assert LS._as_chw(hwc_view).data_ptr() == chw.data_ptr()
    assert LS._as_chw(hwc_view[None]).data_ptr() == chw.data_ptr()
    assert LS._as_chw(torch.rand(8, 9, 3)).shape == (3, 8, 9)
    with pytest.raises(RuntimeError):
        LS._as_chw(torch.rand(2, 3, 8, 9))
    with pytest.raises(RuntimeError):
        LS._as_chw(torch.rand(8, 9, 2))


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match='no CPU path'):
        LS.image_loss_raw(torch.rand(3, 8, 8), torch.rand(3, 8, 8))
    z = torch.zeros(8)
    with pytest.raises(RuntimeError, match='no CPU path'):
        adam_step_raw([z], [z], [z], [z], [1e-3], 1)
    with pytest.raises(RuntimeError):
        adam_step_raw([z], [z, z], [z], [z], [1e-3], 1)
    with pytest.raises(ValueError):
        LS.ImageLoss(method='huber')
    with pytest.raises(NotImplementedError):
        LS.SSIM_Loss(window_size=7)


def test_adam_hyper_table():
    h = adam_hyper([1e-3, (2e-3, 1e-4, 48, 3)], step=3)
    bc1, bc2 = 1 - 0.9 ** 3, 1 - 0.999 ** 3
    assert len(h) == 5
    assert h[0] == pytest.approx(bc2 ** 0.5) and h[1] == h[2] == pytest.approx(1e-3 / bc1)
    assert h[3] == pytest.approx(2e-3 / bc1) and h[4] == pytest.approx(1e-4 / bc1)


def test_adam_param_groups_like_torch():
    a, b = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))
    opt = Adam([{'params': [a], 'lr': 0.1, 'name': 'xyz'}, {'params': b}], lr=1e-3, eps=1e-15)
    assert [g['lr'] for g in opt.param_groups] == [0.1, 1e-3]
    assert opt.param_groups[0]['name'] == 'xyz' and opt.param_groups[1]['eps'] == 1e-15
    assert opt.param_groups[1]['params'] == [b]
    opt.param_groups[0]['lr'] = 0.05  # schedulers write here (gaussian_splatting.py:466-471)
    opt.step()  # no gradients yet: nothing to do, no library call
    assert opt.state == {}
    a.grad = torch.ones(3)
    opt.zero_grad()
    assert a.grad is None
    with pytest.raises(ValueError):
        Adam([])


def test_adam_tensor_struct_matches_header():
    # struct skgs_adam_tensor (include/skgs_b200.h): 4 pointers, int64, 2 doubles, 4 int32, pointer
    assert C.sizeof(_lib.AdamTensor) == 4 * 8 + 8 + 16 + 16 + 8
    assert _lib.AdamTensor.knn_indices.offset == 72 and _lib.AdamTensor.period.offset == 56


def test_adam_is_a_torch_optimizer():
    """the drop-in must survive what the reference does to its optimizer - isinstance checks and in-place
    surgery of param_groups / state (networks/gaussian_splatting.py:515-563 `change_optimizer`), lr schedulers,
    state_dict round trips (checkpoint save / resume)."""
    a = torch.nn.Parameter(torch.zeros(4, 3))
    opt = Adam([{'params': [a], 'lr': 0.1, 'name': 'xyz'}], lr=1e-3, eps=1e-15)
    assert isinstance(opt, torch.optim.Optimizer)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)
    # state in torch.optim.Adam's format
    opt.state[a] = {'step': torch.tensor(3.0), 'exp_avg': torch.ones_like(a), 'exp_avg_sq': torch.full_like(a, 2.0)}
    sd = opt.state_dict()
    assert set(sd) == {'state', 'param_groups'} and sd['param_groups'][0]['name'] == 'xyz'
    ref = torch.optim.Adam([{'params': [torch.nn.Parameter(torch.zeros(4, 3))], 'lr': 0.1, 'name': 'xyz'}], eps=1e-15)
    ref.load_state_dict({'state': sd['state'], 'param_groups': ref.state_dict()['param_groups']})  # interchangeable
    assert float(next(iter(ref.state.values()))['exp_avg_sq'].mean()) == 2.0
    opt2 = Adam([{'params': [torch.nn.Parameter(torch.zeros(4, 3))], 'lr': 0.1, 'name': 'xyz'}], eps=1e-15)
    opt2.load_state_dict(sd)
    st2 = next(iter(opt2.state.values()))
    assert float(st2['step']) == 3.0 and torch.equal(st2['exp_avg'], torch.ones(4, 3))
    sched.step()
    assert opt.param_groups[0]['lr'] == pytest.approx(0.05)
    # the surgery change_optimizer(op='prune') performs: new Parameter object, state re-keyed with masked moments
    keep = torch.tensor([True, False, True, True])
    group = opt.param_groups[0]
    old = group['params'][0]
    stored = opt.state.get(old, None)
    del opt.state[old]
    group['params'][0] = torch.nn.Parameter(old[keep].requires_grad_(True))
    stored['exp_avg'], stored['exp_avg_sq'] = stored['exp_avg'][keep], stored['exp_avg_sq'][keep]
    opt.state[group['params'][0]] = stored
    assert opt.state[group['params'][0]]['exp_avg'].shape == (3, 3)
    opt.zero_grad()
    opt.step()  # still no gradient: no library call on CPU tensors
    with pytest.raises(ValueError):
        Adam([a], lr=-1.0)


def test_global_transform_forms_of_kinematic():
    """g_tr as 7-vector, 6-vector twist (SE3 exponential) and 4x4 matrix (networks/sk_gs.py:1092-1103): the twist form
    against the matrix exponential of the 4x4 twist matrix, the matrix form against its own rotation, both with autograd."""
    import torch
    from sk_gs_b200.fk_lbs import global_transform
    from oracle import fk_lbs as OF
    g = torch.Generator().manual_seed(3)
    assert global_transform(None) is None
    v7 = torch.randn(7, generator=g).double()
    assert torch.equal(global_transform(v7), v7)
    for scale in (1.0, 1e-9):
        xi = (torch.randn(6, generator=g).double() * scale).requires_grad_()
        T7 = global_transform(xi)
        tau, phi = xi[:3].detach(), xi[3:].detach()
        hat = torch.tensor([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]], dtype=torch.float64)
        A = torch.zeros(4, 4, dtype=torch.float64)
        A[:3, :3], A[:3, 3] = hat, tau
        E = torch.linalg.matrix_exp(A)
        assert (T7[:3].detach() - E[:3, 3]).abs().max() <= 1e-12
        p = torch.randn(5, 3, generator=g).double()
        assert (OF.q_rotate(T7[3:].detach(), p) - p @ E[:3, :3].T).abs().max() <= 1e-12
        T7.sum().backward()
        assert torch.isfinite(xi.grad).all()
        # the matrix form of the same transform gives the same 7-vector (up to the sign of q)
        back = global_transform(E)
        sign = torch.sign((back[3:] * T7[3:].detach()).sum())
        assert (back[:3] - T7[:3].detach()).abs().max() <= 1e-12 and (sign * back[3:] - T7[3:].detach()).abs().max() <= 1e-9
    try:
        global_transform(torch.zeros(5))
        assert False
    except ValueError as e:
        assert 'g_tr got shape' in str(e)
