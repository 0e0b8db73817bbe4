"""Mint golden TRAJECTORIES with the REAL reference modules (build container only): 10 Adam iterations of a tracker-shaped
loop (src/Tracker.py:102-197,361-368) and of a mapper-shaped loop (src/Mapper.py:498-541,576-735: index_put flow,
parameter groups, stage switch, and -- ScanNet -- the per-frame exposure slices :697-715).  The loop bodies live in
tests/trajectory_loops.py and are the SAME code the GPU test replays with loopy_slam_b200.

    python tests/golden/make_golden_trajectory.py      # writes tests/golden/traj_*.npz
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

from oracle import ref_import  # noqa: E402
from oracle.knn import ExactKNNPointCloud  # noqa: E402
from loopy_slam_b200.stream import SyntheticRoom, build_point_cloud  # noqa: E402
import make_golden as mg  # noqa: E402
import trajectory_loops as TL  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def c2w_to_cam7(c2w):
    R = c2w[:3, :3].double()
    w = torch.sqrt(torch.clamp(1 + R[0, 0] + R[1, 1] + R[2, 2], min=1e-12)) / 2
    q = torch.stack([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
    return torch.cat([q, c2w[:3, 3].double()]).float()


def scene(seed, n_points=2500, n_frames=3):
    room = SyntheticRoom(H=64, W=64, fx=40.0, fy=40.0, cx=31.5, cy=31.5, seed=seed, n_frames=200, half=(0.9, 0.7, 0.5))
    ids = [0, 10, 20][:n_frames]
    cloud, geo, col = build_point_cloud(room, n_points, pixels_per_frame=2500, seed=seed, frame_ids=ids, max_frames=60)
    frames = [room.frame(i) for i in ids]
    return room, cloud, geo, col, frames


def ref_mods(cfg, model, cloud, room):
    common, decoder_mod, renderer_mod, _ = ref_import.import_reference()
    renderer = renderer_mod.Renderer(cfg, None, mg._SlamLike(room))
    _gd = torch.Tensor.get_device

    def get_camera_from_tensor(t):          # common.py:314 needs a device index: CPU fix as in make_golden_sampling.py
        torch.Tensor.get_device = lambda x: 'cpu'
        try:
            return common.get_camera_from_tensor(t)
        finally:
            torch.Tensor.get_device = _gd
    return types.SimpleNamespace(get_samples=common.get_samples, get_camera_from_tensor=get_camera_from_tensor, renderer=renderer,
                                 decoders=model, npc=ExactKNNPointCloud(cloud, radius_query=cfg['pointcloud']['radius_query']))


def dyn_maps(cfg, frames, seed):
    if not cfg['use_dynamic_radius']:
        return None
    g = torch.Generator().manual_seed(seed + 3)      # smooth per-pixel radii in [0.04, 0.16] (Tracker.py:243-258 produces float64 maps)
    return [0.04 + 0.12 * torch.rand(f[1].shape, generator=g, dtype=torch.float64) for f in frames]


def run(name, yaml, kind, seed, n_iters=10):
    cfg = ref_import.load_cfg(yaml)
    torch.manual_seed(1219)
    model = ref_import.build_model(cfg)
    room, cloud, geo, col, frames = scene(seed)
    mods = ref_mods(cfg, model, cloud, room)
    intr = (room.H, room.W, room.fx, room.fy, room.cx, room.cy)
    out = {'yaml': yaml, 'kind': kind, 'seed': seed, 'n_iters': n_iters, 'cloud': cloud.numpy(), 'geo': geo.numpy(), 'col': col.numpy(),
           'intr': np.array(intr, np.float64)}
    for f, (c, d, m) in enumerate(frames):
        out[f'color{f}'], out[f'depth{f}'], out[f'c2w{f}'] = c.numpy(), d.numpy(), m.numpy()
    wts = {k: v.numpy().copy() for k, v in model.state_dict().items()}
    wts['color_decoder.embedder._B'] = model.color_decoder.embedder._B.numpy().copy()
    for k, v in wts.items():
        out['w0/' + k] = v
    maps = dyn_maps(cfg, frames, seed)
    if maps is not None:
        for f, m in enumerate(maps):
            out[f'rmap{f}'] = m.numpy()
    exposure = cfg['model']['encode_exposure']
    g = torch.Generator().manual_seed(seed + 5)
    with mg._ZeroNoise(), TL.Picks() as picks:
        if kind == 'tracker':
            mods.renderer.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_tracker']
            cam_gt = c2w_to_cam7(frames[-1][2])
            cam0 = cam_gt + torch.tensor([0.004, -0.003, 0.002, 0.003, 0.01, -0.008, 0.006])     # perturbed initial pose
            ef = (torch.randn(cfg['model']['exposure_dim'], generator=g) * 0.01) if exposure else None
            losses, cam, ef_out = TL.tracker_loop(mods, cam0, frames[-1][0], frames[-1][1], intr, cloud, geo, col, n_iters, pixels=200,
                                                  edge=4, cam_lr=cfg['tracking']['lr'], w_color=cfg['tracking']['w_color_loss'],
                                                  dynamic_r_map=None if maps is None else maps[-1], exposure_feat=ef)
            out.update(cam0=cam0.numpy(), cam_final=cam.numpy(), losses=np.array(losses))
            if ef is not None:
                out.update(ef0=ef.numpy(), ef_final=ef_out.numpy())
        else:
            mods.renderer.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_mapper']
            gi = torch.Generator().manual_seed(seed + 7)
            indices = torch.nonzero(torch.rand(cloud.shape[0], generator=gi) < 0.7, as_tuple=True)[0]    # frustum-selected rows
            st = cfg['mapping']['stage']
            lrs = {s: (st[s]['decoders_lr'], st[s]['geometry_lr'], st[s]['color_lr']) for s in ('geometry', 'color')}
            efs = [torch.randn(cfg['model']['exposure_dim'], generator=g) * 0.01 for _ in frames] if exposure else None
            losses, gl, cl, efo = TL.mapper_loop(mods, frames, intr, cloud, geo, col, indices, n_iters, geo_iters=3, pixels=240, lrs=lrs,
                                                 w_color=cfg['mapping']['w_color_loss'], dynamic_r_maps=maps, exposure_feats=efs)
            out.update(indices=indices.numpy(), losses=np.array(losses), geo_leaf=gl.numpy(), col_leaf=cl.numpy(),
                       lrs=np.array([lrs['geometry'], lrs['color']], np.float64))
            if efs is not None:
                out.update(ef0=torch.stack(efs).numpy(), ef_final=torch.stack(efo).numpy())
    for k, v in model.state_dict().items():      # decoder weights after the run (they train in the mapper loop)
        if not np.array_equal(v.numpy(), wts[k]):
            out['w1/' + k] = v.numpy()
    for i, p in enumerate(picks.log):
        out[f'pick{i}'] = p.numpy()
    out['n_picks'] = len(picks.log)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    print(name, 'losses', [round(x, 4) for x in losses], '->', round(os.path.getsize(path) / 1e6, 2), 'MB')


if __name__ == '__main__':
    _, decoder_mod, _, _ = ref_import.import_reference()
    mg._patch_reference(decoder_mod)
    run('traj_replica_tracker', 'configs/Replica/room0.yaml', 'tracker', 31)
    run('traj_replica_mapper', 'configs/Replica/room0.yaml', 'mapper', 32)
    run('traj_scannet_mapper_exposure', 'configs/ScanNet/scene0000.yaml', 'mapper', 33)
    run('traj_tum_tracker_dynr', 'configs/TUM_RGBD/freiburg1_desk.yaml', 'tracker', 34)
