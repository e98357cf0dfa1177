"""Build tuning variants of libskgs_b200.so (compile-time knobs) into sk_gs_b200/variants/ for A/B runs on the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
VARIANTS = {   # compile-time knobs that exist in csrc/ today
    'base': [],
    'pb256x2': ['-DSKGS_PB_THREADS=256', '-DSKGS_PB_MINBLOCKS=2'],
    'pb128x4': ['-DSKGS_PB_THREADS=128', '-DSKGS_PB_MINBLOCKS=4'],
    'pb128x5': ['-DSKGS_PB_THREADS=128', '-DSKGS_PB_MINBLOCKS=5'],
    'pb128x6': ['-DSKGS_PB_THREADS=128', '-DSKGS_PB_MINBLOCKS=6'],
    'pb256x3': ['-DSKGS_PB_THREADS=256', '-DSKGS_PB_MINBLOCKS=3'],
    'pb64x10': ['-DSKGS_PB_THREADS=64', '-DSKGS_PB_MINBLOCKS=10'],
    'pb64x12': ['-DSKGS_PB_THREADS=64', '-DSKGS_PB_MINBLOCKS=12'],
}
out_dir = os.path.join(ROOT, 'sk_gs_b200', 'variants')
os.makedirs(out_dir, exist_ok=True)
procs = []
for name, flags in VARIANTS.items():
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    srcs = [os.path.join(ge.CSRC, f) for f in ge.FILES]
    objs = []
    for f, extra in ge.FILES.items():
        obj = os.path.join(ROOT, 'build', 'obj', f'{name}_{f}.o')
        objs.append(obj)
        cmd = [ge.NVCC, *ge.ARCH, '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC,-fvisibility=hidden',
               '--expt-relaxed-constexpr', *extra, *flags, '-c', os.path.join(ge.CSRC, f), '-o', obj]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    VARIANTS[name] = objs
for name, p in procs:
    out, _ = p.communicate()
    if p.returncode:
        print(name, 'FAILED\n', out)
        sys.exit(1)
for name, objs in VARIANTS.items():
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    lib = os.path.join(out_dir, f'libskgs_{name}.so')
    subprocess.run([ge.NVCC, *ge.ARCH, '-shared', '-Xcompiler', '-fPIC', *objs, '-o', lib, '-lcudart'], check=True)
    print('built', lib)
