"""Debug tool: per-phase cycle breakdown of the fused kernels (build with -DLSR_PHASE_TIMING into a
side .so, run one bench-shaped step, read lsr_phase_cycles).  Not part of the product path."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, 'loopy_slam_b200', 'liblsr_phase.so')


def build():
    csrc = os.path.join(ROOT, 'loopy_slam_b200', 'csrc')
    srcs = [os.path.join(csrc, f) for f in ('lsr_grid.cu', 'lsr_sample.cu', 'lsr_loss.cu', 'lsr_render_fwd.cu', 'lsr_render_bwd.cu', 'lsr_render_bwd_umma.cu', 'lsr_geo_bwd_umma.cu')]
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-rdc=true',
                           '-DLSR_PHASE_TIMING', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-shared', '-cudart', 'static',
                           '-o', SO] + srcs)


if __name__ == '__main__':
    if not os.path.exists(SO) or '--build' in sys.argv or '--build-only' in sys.argv:
        build()
    if '--build-only' in sys.argv:
        sys.exit(0)
    os.environ['LSR_LIB'] = SO
    import torch
    import bench
    from loopy_slam_b200 import _lib
    sys.argv = ['bench.py', '--steps', '10', '--warmup', '3', '--no-cpu-baseline', '--no-extra'] + \
        [a for a in sys.argv[1:] if a not in ('--build', '--build-only')]
    bench.main()
    torch.cuda.synchronize()
    rt = ctypes.CDLL('libcudart.so.12') if False else None
    # read the device symbol through the library's own runtime
    L = _lib.lib()
    buf = (ctypes.c_ulonglong * 32)()
    get = L.lsr_debug_phase_cycles
    get.argtypes = [ctypes.c_void_p]
    get(buf)
    names = [['load k-NN lists', 'gather+fourier', 'geo MLP', 'relpos MLP', "e'+colour trunk", 'colour head', 'compositing'],
             ['state+composite bwd', 'head bwd+setup', 'colour trunk bwd', 'fourier bwd+dC', 'relpos bwd', 'geo bwd+scatter']]
    for k in range(2):
        tot = sum(buf[k * 16 + i] for i in range(16)) or 1
        print(['render_fwd', 'render_bwd'][k], 'phase share of CTA-cycles:')
        for i, n in enumerate(names[k]):
            print(f'   {n:24s} {100.0 * buf[k * 16 + i] / tot:5.1f}%   {buf[k * 16 + i] / 1e6:9.2f} Mcycles')
