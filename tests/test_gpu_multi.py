"""2-GPU tests (NCCL, one process per GPU; skipped on a single-GPU box -- run with `gpurun --gpus 2`):
  * sharded render_batch_ray + ONE all-reduce of the fused backward's gradient buffer == the
    single-GPU gradient of the whole batch (feature leaf blocks, decoder weights) to 1e-5 L2;
  * render_img sharded over the ranks (parallel.render_img_sharded) == render_img on one GPU, bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import loopy_slam_b200 as L
    from loopy_slam_b200 import parallel
    from helpers import Golden, rel_l2
    from parity import cfg_from_ocfg, build_model, SlamLike
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    dev = f'cuda:{rank}'
    try:
        g = Golden('replica_color_mapper')
        cfg = cfg_from_ocfg(g.ocfg)
        H, W, fx, fy, cx, cy = g.raw['intrinsics']

        class NPC:
            def get_radius_query(self):
                return g.ocfg.radius_query

        def run(sl):
            model = build_model(cfg, g.weights, dev)
            for p in model.parameters():
                p.requires_grad_(True)
            rend = L.Renderer(cfg, None, SlamLike(H, W, fx, fy, cx, cy))
            rend.sigmoid_coefficient = g.ocfg.sigmoid_coef
            geo = g.t('geo_feats').to(dev).requires_grad_(True)
            col = g.t('col_feats').to(dev).requires_grad_(True)
            depth, var, rgb, valid = rend.render_batch_ray(NPC(), model, g.t('rays_d').to(dev)[sl], g.t('rays_o').to(dev)[sl], dev, 'color',
                                                           gt_depth=g.t('gt_depth').to(dev)[sl], npc_geo_feats=geo, npc_col_feats=col,
                                                           cloud_pos=g.t('cloud').to(dev))
            ((g.t('up_depth').to(dev)[sl] * depth).sum() + (g.t('up_rgb').to(dev)[sl] * rgb).sum()).backward()
            return model, rend, geo, col
        R = g.t('rays_o').shape[0]
        lo, hi = parallel.shard_bounds(R, rank, world)
        model, rend, geo, col = run(slice(lo, hi))
        params = [p for p in model.parameters() if p.grad is not None] + [geo, col]
        nbytes = parallel.GradAllReducer(params).allreduce_(rend.last_grad_buffer)
        ref_model, _, rgeo, rcol = run(slice(0, R))                  # the whole batch on this GPU
        worst = max(rel_l2(geo.grad, rgeo.grad), rel_l2(col.grad, rcol.grad))
        for (k, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
            if q.grad is not None and float(q.grad.abs().max()) > 0:
                worst = max(worst, rel_l2(p.grad, q.grad))
        # render_img, sharded
        g2 = Golden('replica_color_sparse_zero_depth')
        rend2 = L.Renderer(cfg_from_ocfg(g2.ocfg), None, SlamLike(24, 30, 20.0, 20.0, 14.5, 11.5), ray_batch_size=100)
        rend2.sigmoid_coefficient = g2.ocfg.sigmoid_coef
        model2 = build_model(cfg_from_ocfg(g2.ocfg), g2.weights, dev)
        c2w = torch.eye(4, device=dev)
        c2w[:3, 3] = g.t('rays_o')[0].to(dev)
        gen = torch.Generator().manual_seed(21)
        gt = 0.4 + torch.rand(24, 30, generator=gen) * 0.8
        gt[::4, ::3] = 0.0
        kw = dict(gt_depth=gt.to(dev), npc_geo_feats=g.t('geo_feats').to(dev), npc_col_feats=g.t('col_feats').to(dev),
                  cloud_pos=g.t('cloud').to(dev))
        a = parallel.render_img_sharded(rend2, NPC(), model2, c2w, dev, 'color', **kw)
        b = rend2.render_img(NPC(), model2, c2w, dev, 'color', **kw)
        same = all(torch.equal(x, y) for x, y in zip(a, b))
        out[rank] = (float(worst), bool(same), int(nbytes))
    finally:
        dist.destroy_process_group()


def test_sharded_render_allreduce_equals_single_gpu():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        worst, same, nbytes = out[r]
        assert worst < 1e-5, (r, worst)      # sums of the same per-ray terms in a different order
        assert same, 'sharded render_img differs from the single-GPU image'
        assert nbytes > 0
