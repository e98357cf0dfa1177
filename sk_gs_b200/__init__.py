"""sk_gs_b200 - B200-native (sm_100a) implementation of SK_GS's per-iteration hot path:
skeleton FK -> linear blend skinning -> differentiable tile-based Gaussian splatting (forward + backward).

Public surface (mirrors the reference's entry points):
  sk_gs_b200.diff_gaussian_rasterization   GaussianRasterizationSettings / GaussianRasterizer   (boundary B1)
  sk_gs_b200.renderer.render_gs_offical    networks/renderer/gaussian_render_origin.py:11-68
  sk_gs_b200.fk_lbs.fk_lbs / assemble      networks/sk_gs.py:1109-1150 / :1192,1202-1203        (boundary B3)
  sk_gs_b200.pipeline.HotPath              the whole FK -> LBS -> raster step on a set of parameters
The compute lives in libskgs_b200.so (include/skgs_b200.h); importing this package does not load it, calling does.
"""
__all__ = ['diff_gaussian_rasterization', 'renderer', 'fk_lbs', 'pipeline', 'scene']
