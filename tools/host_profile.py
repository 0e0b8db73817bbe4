"""cProfile of the literal mapping iteration (host side of the public API).  Bring-up tool."""
import cProfile, pstats, sys, os, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device('cuda:0')
it = bench.MapperIteration('replica', 200000, dev)
for _ in range(5):
    it.step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(30):
    it.step()
torch.cuda.synchronize()
print('literal ms/step', (time.perf_counter() - t0) / 30 * 1e3)
t0 = time.perf_counter()
for _ in range(30):
    it.sample()
torch.cuda.synchronize()
print('sample() only ms', (time.perf_counter() - t0) / 30 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    it.step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22)
print(s.getvalue()[:6000])
