"""sk_gs_b200 - B200-native (sm_100a) implementation of SK_GS's per-iteration hot path:
skeleton FK -> linear blend skinning -> differentiable tile-based Gaussian splatting (forward + backward).

Public surface (mirrors the reference's entry points):
  sk_gs_b200.diff_gaussian_rasterization   GaussianRasterizationSettings / GaussianRasterizer   (boundary B1)
  sk_gs_b200.renderer.render_gs_offical    networks/renderer/gaussian_render_origin.py:11-68
  sk_gs_b200.fk_lbs.fk_lbs / assemble      networks/sk_gs.py:1109-1150 / :1192,1202-1203        (boundary B3)
  sk_gs_b200.pipeline.HotPath              the whole FK -> LBS -> raster step on a set of parameters
  sk_gs_b200.dist                          view sharding, flat gradient arenas, NVLS multimem all-reduce       (8e)
Steps either side of the path:
  sk_gs_b200.deform_net.SimpleDeformationNetwork   networks/sk_gs.py:134-164 (joint-rotation network)          (f-1)
  sk_gs_b200.losses.ImageLoss / SSIM_Loss / image_ssim_loss   networks/losses/image_loss.py, ssim.py           (f-2)
  sk_gs_b200.optim.Adam                    torch.optim.Adam as configured at networks/gaussian_splatting.py:445-453 (f-3)
  sk_gs_b200.train.TrainLoop               render -> loss -> backward -> Adam as one CUDA graph
The compute lives in libskgs_b200.so (include/skgs_b200.h); importing this package does not load it, calling does.
"""
__all__ = ['diff_gaussian_rasterization', 'renderer', 'fk_lbs', 'pipeline', 'scene', 'dist', 'deform_net', 'losses',
           'optim', 'train']
