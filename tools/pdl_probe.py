"""What does a kernel boundary cost, with and without programmatic dependent launch, eagerly and in a CUDA graph?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import _lib
L = _lib.lib()
L.skgs_debug_pdl_chain.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
p = torch.zeros(4, dtype=torch.int32, device='cuda')
N = 200
for grid in (1, 148, 1184):
    for pdl in (0, 1):
        st = torch.cuda.current_stream().cuda_stream
        L.skgs_debug_pdl_chain(p.data_ptr(), 20, pdl, grid, st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); L.skgs_debug_pdl_chain(p.data_ptr(), N, pdl, grid, st); b.record(); torch.cuda.synchronize()
        eager = a.elapsed_time(b) / N * 1e3
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            L.skgs_debug_pdl_chain(p.data_ptr(), N, pdl, grid, torch.cuda.current_stream().cuda_stream)
        g.replay(); torch.cuda.synchronize()
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        print(f'grid {grid:5d} pdl={pdl}: eager {eager:6.2f} us/kernel   graph {a.elapsed_time(b) / N * 1e3:6.2f} us/kernel', flush=True)
print('count', int(p[0]))
