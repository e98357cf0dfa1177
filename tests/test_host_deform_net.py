"""Host-side logic of the joint-rotation network mirror that needs no GPU: parameter layout of the flat `theta` vector
(skgs_joint_mlp_layout) against the reference's layer shapes, checkpoint key names, no CPU fallback."""
import pytest
import torch

from oracle import deform_net as OD
from sk_gs_b200.deform_net import NetConfig, SimpleDeformationNetwork


@pytest.mark.parametrize('width,depth,skips', [(256, 8, (4,)), (16, 4, (1, 2)), (8, 1, ()), (32, 3, (2,))])
def test_layout_matches_reference_layer_shapes(width, depth, skips):
    cfg = NetConfig(10, 6, width, depth, skips)
    enc, shapes = OD.layer_shapes(width=width, depth=depth, skips=skips)
    assert cfg.enc == enc == 76
    assert cfg.in_dims[:depth] == [s[1] for s in shapes[:depth]]
    assert cfg.in_dims[depth] == shapes[depth][1]  # the three heads share their input width
    total = sum(o * i + o for o, i in shapes)
    assert cfg.param_count == total
    theta = torch.arange(cfg.param_count, dtype=torch.float32)
    views = cfg.views(theta)
    seen = 0
    for i, (w, b) in enumerate(views):
        o = width if i < depth else 11
        assert w.shape == (o, cfg.in_dims[i]) and b.shape == (o,)
        assert w.data_ptr() == theta.data_ptr() + 4 * seen  # dense, in order: W_i then b_i
        seen += w.numel()
        assert b.data_ptr() == theta.data_ptr() + 4 * seen
        seen += b.numel()
    assert seen == cfg.param_count


def test_state_dict_names_and_roundtrip():
    net = SimpleDeformationNetwork(pos_enc_p_cfg=dict(degree=10), pos_enc_t_cfg=dict(degree=6), width=32, depth=8,
                                   skips=(4,))
    sd = net.reference_state_dict()
    want = {f'dynamic_net.net.{i}.{k}' for i in range(8) for k in ('weight', 'bias')} | \
           {f'dynamic_net.last.{j}.{k}' for j in range(3) for k in ('weight', 'bias')}
    assert set(sd) == want  # the keys of the reference module's state_dict (sk_gs.py:149-155, mlp.py:56-66)
    assert sd['dynamic_net.net.0.weight'].shape == (32, 76) and sd['dynamic_net.net.5.weight'].shape == (32, 108)
    assert [sd[f'dynamic_net.last.{j}.weight'].shape[0] for j in range(3)] == [4, 4, 3]
    other = SimpleDeformationNetwork(pos_enc_p_cfg=dict(degree=10), pos_enc_t_cfg=dict(degree=6), width=32, depth=8,
                                     skips=(4,))
    assert not torch.equal(other.theta, net.theta)
    other.load_reference_state_dict(sd)
    assert torch.equal(other.theta, net.theta)
    with pytest.raises(KeyError):
        other.load_reference_state_dict({k: v for k, v in sd.items() if 'last.1' not in k})
    net.reset_heads(1e-6)
    w, b = net.cfg.views(net.theta.detach())[-1]
    assert float(b.abs().max()) == 0 and float(w.abs().max()) < 1e-4


def test_no_cpu_fallback_and_argument_errors():
    net = SimpleDeformationNetwork(pos_enc_p_cfg=dict(degree=10), pos_enc_t_cfg=dict(degree=6), width=8, depth=2, skips=())
    with pytest.raises(RuntimeError, match='no CPU path'):
        net(torch.zeros(4, 3), torch.tensor(0.1))
    with pytest.raises(NotImplementedError):
        SimpleDeformationNetwork(pos_enc_p='hash')
    with pytest.raises(ValueError):
        NetConfig(10, 6, 8, 2, skips=(2,))
    with pytest.raises(RuntimeError):
        NetConfig(10, 6, 8, 40, skips=())  # depth > 32: rejected by the library
