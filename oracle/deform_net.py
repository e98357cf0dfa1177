"""oracle/deform_net.py - pure-torch CPU restatement of the joint-rotation network of the `sk` stage.

TEST INFRASTRUCTURE ONLY (see oracle/raster_oracle.c header): only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg import this module.  The product (sk_gs_b200/) never does.

Reference: `SimpleDeformationNetwork` networks/sk_gs.py:134-164 (two frequency encoders, concatenation, MLP),
`MLP_with_skips` my_ext/blocks/mlp.py:44-85 (ReLU after every layer, skip = cat([x, inputs]) AFTER layer i's ReLU,
one Linear per output head), the head of `kinematic` networks/sk_gs.py:1074-1076
(sk_r = normalize(out_r + (0,0,0,1))), frequency encoder my_ext/_C/src/nerf/freqencoder.cu:7-31 (forward) and :36-62
(backward), configuration exps/default.yaml:48-55 (degree 10 / 6, width 256, depth 8, skips [4]).

Parity status: the network structure (encoder layout, skip wiring, heads) is PINNED against the reference's own
`SimpleDeformationNetwork` imported unmodified with the pure-torch encoder variant ('freq_torch',
networks/encoders/freq_encoder.py:86-134) -> tests/golden/deform_net.npz.  The CUDA encoder variant the configs select
('freq') evaluates cos as sin(y + fl32(pi/2)) in fp32 with the fast `__sinf`; `freq_encode(..., cuda_formula=True)`
restates that argument arithmetic (exact sin of the fp32 argument); it cannot be run here (no GPU): parity unpinned for
that detail.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
from torch import Tensor


class _FreqPair(torch.autograd.Function):
    """(sin(y), sin(shifted)) with the reference's backward, which differentiates through the STORED outputs
    (freqencoder.cu:52-56): d sin-slot / dy = stored cos-slot, d cos-slot / dy = - stored sin-slot."""

    @staticmethod
    def forward(ctx, y, shifted):
        s, c = torch.sin(y), torch.sin(shifted)
        ctx.save_for_backward(s, c)
        return s, c

    @staticmethod
    def backward(ctx, gs, gc):
        s, c = ctx.saved_tensors
        return gs * c - gc * s, None


def freq_encode(x: Tensor, degree: int, cuda_formula: bool = True) -> Tensor:
    """[..., D] -> [..., D + 2 D degree]: (x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...), each block D wide.
    cuda_formula: cos(y) is evaluated as sin(fl32(y + fl32(pi/2))) like freqencoder.cu:26-29 (one fp32 rounding of the
    shifted argument; the sin itself is exact here, `__sinf` there)."""
    out = [x]
    for f in range(degree):
        y = x * (2.0 ** f)  # scalbnf: exact
        if cuda_formula:
            shifted = (y.detach().float() + torch.tensor(math.pi / 2, dtype=torch.float32)).to(x.dtype)
            out += list(_FreqPair.apply(y, shifted))
        else:
            out += [torch.sin(y), torch.cos(y)]
    return torch.cat(out, dim=-1)


def layer_shapes(in_dim: int = 3, degree_p: int = 10, degree_t: int = 6, width: int = 256, depth: int = 8,
                 skips: Sequence[int] = (4,), heads: Sequence[int] = (4, 4, 3)) -> Tuple[int, List[Tuple[int, int]]]:
    """(encoded input width, [(out, in) of every Linear: depth hidden layers then the heads]) - mlp.py:56-66."""
    enc = in_dim * (1 + 2 * degree_p) + (1 + 2 * degree_t)
    shapes, cin = [], enc
    for i in range(depth):
        shapes.append((width, cin))
        cin = width + (enc if i in skips else 0)
    shapes += [(h, cin) for h in heads]
    return enc, shapes


def forward(weights: List[Tensor], biases: List[Tensor], joints: Tensor, t: Tensor, degree_p: int = 10,
            degree_t: int = 6, skips: Sequence[int] = (4,), n_heads: int = 3, cuda_formula: bool = True,
            rotation_head: bool = True):
    """joints [M,3], t scalar tensor -> (sk_r [M,4], d_rot [M,4], d_scale [M,3]) (+ raw head outputs)."""
    depth = len(weights) - n_heads
    p_embed = freq_encode(joints, degree_p, cuda_formula)
    t_embed = freq_encode(t.reshape(1, 1), degree_t, cuda_formula).expand(joints.shape[0], -1)
    x0 = torch.cat([p_embed, t_embed], dim=-1)
    x = x0
    for i in range(depth):
        x = torch.relu(x @ weights[i].t() + biases[i])
        if i in skips:
            x = torch.cat([x, x0], dim=-1)
    outs = [x @ weights[depth + j].t() + biases[depth + j] for j in range(n_heads)]
    if rotation_head:  # sk_gs.py:1075-1076
        q = outs[0] + outs[0].new_tensor([0., 0., 0., 1.])
        outs[0] = torch.nn.functional.normalize(q, dim=-1)
    return outs


def init_params(seed: int = 0, head_std: float = 1e-6, **cfg) -> Tuple[List[Tensor], List[Tensor]]:
    """nn.Linear default init for the hidden layers; heads: zero bias, N(0, 1e-6) weights (sk_gs.py:542-545).
    `head_std` can be raised in tests so that the heads' gradients are not at noise level."""
    g = torch.Generator().manual_seed(seed)
    _, shapes = layer_shapes(**cfg)
    n_heads = len(cfg.get('heads', (4, 4, 3)))
    ws, bs = [], []
    for idx, (o, i) in enumerate(shapes):
        bound = 1.0 / math.sqrt(i)
        if idx < len(shapes) - n_heads:
            ws.append((torch.rand(o, i, generator=g) * 2 - 1) * bound)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), ..)
            bs.append((torch.rand(o, generator=g) * 2 - 1) * bound)
        else:
            ws.append(torch.randn(o, i, generator=g) * head_std)
            bs.append(torch.zeros(o))
    return ws, bs
