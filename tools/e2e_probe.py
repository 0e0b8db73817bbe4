"""Wall-clock breakdown of the literal mapping iteration (bring-up tool): each phase with a synchronize after it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device('cuda:0')
it = bench.MapperIteration('replica', 200000, dev)
L = it.L
for _ in range(5):
    it.step()
torch.cuda.synchronize()
N = 40


def wall(fn, n=N):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print('literal step, no per-step sync      %.3f ms' % wall(lambda: it.step()))
print('literal step + loss.item()          %.3f ms' % wall(lambda: it.step().item()))
poses_host = torch.stack([f[2] for f in it.sc['frames']]).pin_memory()


def e2e():
    poses = poses_host.to(dev, non_blocking=True)
    return it.step(c2ws=[poses[f] for f in range(it.n_frames)]).item()


print('e2e step (poses H2D + item)         %.3f ms' % wall(e2e))
print('extension step + item               %.3f ms' % wall(lambda: it.step(literal=False).item()))
ph = {}


def phased():
    def tick(name, t0):
        torch.cuda.synchronize()
        ph[name] = ph.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()
    t = time.perf_counter()
    for p in it.train_params:
        p.grad = None
    o, d, g, c, rq, fid = it.sample()
    t = tick('sample', t)
    it.npc_geo[it.indices] = it.geo_leaf
    it.npc_col[it.indices] = it.col_leaf
    t = tick('index_put', t)
    depth, var, color, valid = it.rend.render_batch_ray(it.npc, it.model, d, o, dev, 'color', gt_depth=g, npc_geo_feats=it.npc_geo,
                                                        npc_col_feats=it.npc_col, is_tracker=False, cloud_pos=it.cloud)
    t = tick('render_fwd', t)
    loss = bench.mapper_loss_reference(depth, color, valid, g, c, 'color', it.w_color)
    t = tick('loss', t)
    loss.backward()
    t = tick('backward', t)
    it.npc_geo, it.npc_col = it.npc_geo.detach(), it.npc_col.detach()


for _ in range(N):
    phased()
print({k: round(v / N * 1e3, 3) for k, v in ph.items()})
# get_samples alone
c2w = it.c2ws[0]
room = it.room
def gs():
    L.get_samples(0, room.H, 0, room.W, it.pix_per_image, room.H, room.W, room.fx, room.fy, room.cx, room.cy, c2w, it.depths[0],
                  it.colors[0], dev, depth_filter=True, return_index=True)
print('one get_samples call                %.3f ms' % wall(gs, 200))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        it.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=35, max_name_column_width=60)[:9000])
