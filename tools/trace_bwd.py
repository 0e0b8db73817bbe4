"""Timeline of one tile of trunk_bwd_umma_kernel (bring-up; builds liblsr_trace.so with -DLSR_TRACE, runs the bench
step once and prints the clock64 stamps of CTA 0: issuer per ring op, compute warp 0 per layer phase)."""
import ctypes
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
so = os.path.join(ROOT, 'loopy_slam_b200', 'liblsr_trace.so')
if '--build' in sys.argv:
    entry.build(out=so, defines=('LSR_TRACE',))
    sys.exit(0)
os.environ['LSR_LIB'] = so
import torch
from loopy_slam_b200 import _lib
_lib.LIB_PATH = so
import bench
sys.argv = ['bench.py', '--steps', '3', '--warmup', '3', '--no-cpu-baseline', '--no-extra']
import io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
lib = _lib.lib()
out = (ctypes.c_longlong * (3 * 512))()
lib.lsr_debug_trace(out)
T = [list(out[r * 512:(r + 1) * 512]) for r in range(3)]
t0 = T[2][0]
print('tile: start 0, state loaded %d, compositing %d, head act %d, out_linear FMA %d, layers done %d, end %d' % (T[2][4] - t0, T[2][5] - t0, T[2][6] - t0, T[2][1] - t0, T[2][2] - t0, T[2][3] - t0))
names = ['A0', 'A1', 'B-waited-d1', 'B-loaded', 'C-signalled-a', 'flush-issued', 'E-start(ZT written)', 'E-waited-d0', 'E-read', 'F-signalled-b']
for li in range(5):
    print('layer', 4 - li, ' '.join(f'{n}:{T[1][li * 10 + k] - t0}' for k, n in enumerate(names)))
print('rel-pos per neighbour k: start, a-signalled, d1-waited, b-signalled, d0-waited, end')
for k in range(8):
    print(' k', k, [T[2][16 + 6 * k + j] - t0 for j in range(6)])
print('issuer ops (i: wait-start conv-ready issued) relative to tile start:')
i = 0
while T[0][3 * i] != 0 and i < 170:
    print(i, T[0][3 * i] - t0, T[0][3 * i + 1] - t0, T[0][3 * i + 2] - t0)
    i += 1
