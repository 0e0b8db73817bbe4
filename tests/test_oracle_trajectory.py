"""CPU: the oracle restatement replays the golden TRAJECTORIES of the real reference (tests/golden/traj_*.npz) with the
same loop code (tests/trajectory_loops.py).  Pins the oracle at caller level (optimiser groups, index_put flow, stage
switch, exposure slices) and MEASURES how fast two fp32 implementations of the same tracker loop drift apart -- the
yardstick the GPU replay (tests/test_gpu_trajectory.py) is held to."""
import os
import types

import numpy as np
import pytest
import torch

import loopy_slam_b200 as L
import trajectory_loops as TL
from oracle import render as orc
from oracle import sampling as osm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class _OracleRenderer:
    def __init__(self, model, ocfg):
        self.model, self.ocfg = model, ocfg
        self.sigmoid_coefficient = ocfg.sigmoid_coef

    def weights(self):
        W = dict(self.model.named_parameters())
        W['color_decoder.embedder._B'] = self.model.color_decoder.embedder._B
        return W

    def render_batch_ray(self, npc, decoders, rays_d, rays_o, device, stage, gt_depth=None, npc_geo_feats=None,
                         npc_col_feats=None, is_tracker=False, cloud_pos=None, dynamic_r_query=None, exposure_feat=None):
        import dataclasses
        ocfg = dataclasses.replace(self.ocfg, sigmoid_coef=self.sigmoid_coefficient)
        depth, var, rgb, valid, _ = orc.render_rays(self.weights(), ocfg, rays_o, rays_d, gt_depth, npc_geo_feats, npc_col_feats,
                                                    cloud_pos, stage, is_tracker=is_tracker, dynamic_r=dynamic_r_query,
                                                    exposure_feat=exposure_feat)
        return depth, var, rgb, valid


def _mods(g):
    yaml = str(g['yaml'])
    fam = 'replica' if 'Replica' in yaml else 'tum' if 'TUM' in yaml else 'scannet'
    cfg = L.default_cfg(fam)
    torch.manual_seed(0)
    model = L.get_model(cfg)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('w0/') and k != 'w0/color_decoder.embedder._B'}
    model.load_state_dict(sd, strict=True)
    model.color_decoder.embedder._B = torch.from_numpy(g['w0/color_decoder.embedder._B']).clone()
    rend = _OracleRenderer(model, orc.OracleCfg.from_cfg(cfg))

    def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, device, depth_filter=False, return_index=False,
                    depth_limit=None):
        return osm.get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, depth_filter=depth_filter,
                               depth_limit=depth_limit)
    return cfg, model, types.SimpleNamespace(get_samples=get_samples, get_camera_from_tensor=osm.get_camera_from_tensor,
                                             renderer=rend, decoders=model, npc=None)


def _frames(g):
    out, f = [], 0
    while f'color{f}' in g:
        out.append((torch.from_numpy(g[f'color{f}']), torch.from_numpy(g[f'depth{f}']), torch.from_numpy(g[f'c2w{f}'])))
        f += 1
    return out


@pytest.mark.parametrize('name', ['traj_replica_tracker', 'traj_replica_mapper'])
def test_oracle_replays_reference_trajectory(name):
    z = np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    cfg, model, mods = _mods(g)
    frames = _frames(g)
    H, W, fx, fy, cx, cy = g['intr']
    intr = (int(H), int(W), float(fx), float(fy), float(cx), float(cy))
    cloud, geo, col = [torch.from_numpy(g[k]) for k in ('cloud', 'geo', 'col')]
    picks = [torch.from_numpy(g[f'pick{i}']) for i in range(int(g['n_picks']))]
    n_it = 6      # enough to cross the geometry -> colour switch of the mapper loop; keeps the CPU suite short
    with TL.Picks(stored=picks, device='cpu'):
        if str(g['kind']) == 'tracker':
            mods.renderer.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_tracker']
            losses, cam, _ = TL.tracker_loop(mods, torch.from_numpy(g['cam0']), frames[-1][0], frames[-1][1], intr, cloud, geo, col, n_it,
                                             pixels=200, edge=4, cam_lr=cfg['tracking']['lr'], w_color=cfg['tracking']['w_color_loss'])
        else:
            mods.renderer.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_mapper']
            lrs = {'geometry': tuple(g['lrs'][0]), 'color': tuple(g['lrs'][1])}
            losses, gl, cl, _ = TL.mapper_loop(mods, frames, intr, cloud, geo, col, torch.from_numpy(g['indices']), n_it, geo_iters=3,
                                               pixels=240, lrs=lrs, w_color=cfg['mapping']['w_color_loss'])
    rel = np.abs(np.array(losses) - g['losses'][:n_it]) / np.abs(g['losses'][:n_it])
    print(name, 'oracle vs reference, loss rel err per iteration:', rel)
    assert rel[:4].max() < 2e-4, rel                 # iterations before fp32 noise has been amplified
    assert rel.max() < (5e-2 if str(g['kind']) == 'tracker' else 2e-3), rel
