"""Two eager TrainLoop iterations on workload c2 (for ncu captures of the loss / Adam kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import scene as S
from sk_gs_b200.pipeline import HotPath
from sk_gs_b200.train import TrainLoop

cfg = S.CONFIGS['c2']
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', requires_grad=False, merged_sh=True, joint_mlp=True, head_std=0.02)
loop = TrainLoop(hp)
with torch.no_grad():
    target = hp.render(0)['images'].detach().clone()
target = (target + 0.05 * torch.randn_like(target)).clamp_(0, 1)
for _ in range(2):
    out = loop.step(0, target)
torch.cuda.synchronize()
print('loss terms', out['loss_terms'].tolist())
