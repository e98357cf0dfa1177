"""Time the joint-rotation network alone (forward + backward, M = 32, the exps/default.yaml shape): eager launches and
as a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200.deform_net import NetConfig, joint_mlp_backward_raw, joint_mlp_forward_raw

dev = 'cuda:0'
cfg = NetConfig(10, 6, 256, 8, (4,), rotation_head=True)
g = torch.Generator().manual_seed(0)
theta = (torch.randn(cfg.param_count, generator=g) * 0.05).to(dev)
joints = (torch.randn(32, 3, generator=g) * 0.4).to(dev)
t = torch.tensor([0.3], device=dev)
gs = [torch.randn(32, w, generator=g).to(dev) for w in (4, 4, 3)]
bufs, out = {}, {'theta': torch.empty(cfg.param_count, device=dev), 'joints': torch.empty(32, 3, device=dev)}


def step():
    o, ctx = joint_mlp_forward_raw(cfg, theta, joints, t, out=bufs)
    joint_mlp_backward_raw(ctx, *gs, out=out)


def timeit(fn, n=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


eager = timeit(step)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
graphed = timeit(graph.replay)
print(f'{os.environ.get("SKGS_LIB", "default")[-28:]:28s} joint MLP fwd+bwd (M=32): eager {eager:7.1f} us   graph {graphed:7.1f} us')
