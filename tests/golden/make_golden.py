"""Mint golden vectors by running the REAL reference (imported from /root/reference) on
seeded synthetic scenes.  Build-container only; the .npz files it writes are committed and
travel to the GPU box, this script's inputs do not.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

What is run: the reference's own ``Renderer.render_batch_ray`` (src/utils/Renderer.py:71-201)
+ ``NICER`` (src/conv_onet/models/decoder.py:549-626) + ``raw2outputs_nerf_color``
(src/common.py:382-422) on CPU, float32, with
  * the FAISS-GPU search replaced by ``oracle.knn.ExactKNNPointCloud`` (exact k-NN, same
    output contract -- faiss-gpu 1.7.2 is third-party, approximate and not installable),
  * the N(0,0.01) random fill for neighbour-less samples patched to zeros,
  * ``NICER.forward`` stage 'geometry' device string fixed for CPU (decoder.py:591,597-598
    hard-code 'cuda:<id>'; the 4 lines :593-600 are re-issued with device=p.device).
Loss used for the gradients: L = sum(a*depth) + sum(b*rgb) with seeded random a, b
(stored), so upstream gradients are known, dense and non-trivial.
"""
import dataclasses
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle.knn import ExactKNNPointCloud  # noqa: E402
from oracle.render import OracleCfg  # noqa: E402
from loopy_slam_b200.stream import SyntheticRoom, build_point_cloud, sample_batch  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _SlamLike:
    def __init__(self, room):
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = (room.H, room.W, room.fx, room.fy,
                                                               room.cx, room.cy)


def _patch_reference(decoder_mod):
    """CPU fix for stage 'geometry' + zero noise fill."""
    orig_forward = decoder_mod.NICER.forward

    def forward(self, p, npc, stage, npc_geo_feats, npc_col_feats, pts_num=16, is_tracker=False,
                cloud_pos=None, pts_views_d=None, dynamic_r_query=None, exposure_feat=None):
        if stage == 'geometry':                     # decoder.py:593-600 with device=p.device
            occ, ray_mask, point_mask = self.geo_decoder(
                p, npc, npc_geo_feats, npc_col_feats, pts_num=pts_num, is_tracker=is_tracker,
                cloud_pos=cloud_pos, dynamic_r_query=dynamic_r_query)
            raw = torch.zeros(occ.shape[0], 4, device=p.device, dtype=torch.float)
            raw[..., -1] = occ
            return raw, ray_mask, point_mask
        _gd = torch.Tensor.get_device
        torch.Tensor.get_device = lambda t: 0       # only feeds an unused f-string in 'color'
        try:
            return orig_forward(self, p, npc, stage, npc_geo_feats, npc_col_feats, pts_num,
                                is_tracker, cloud_pos, pts_views_d, dynamic_r_query, exposure_feat)
        finally:
            torch.Tensor.get_device = _gd

    decoder_mod.NICER.forward = forward


class _ZeroNoise:
    """normal_(std=0.01) -> zeros while active (the reference's only use of std=0.01 on this
    path is the neighbour-less fill)."""

    def __enter__(self):
        self._orig = torch.Tensor.normal_

        def normal_(t, mean=0, std=1, *, generator=None):
            if std == 0.01:
                return t.zero_()
            return self._orig(t, mean, std, generator=generator)
        torch.Tensor.normal_ = normal_

    def __exit__(self, *a):
        torch.Tensor.normal_ = self._orig


def make_scene(seed, n_points, n_rays, sparse=False, with_zero_depth=False, dynamic=False):
    room = SyntheticRoom(H=64, W=64, fx=40.0, fy=40.0, cx=31.5, cy=31.5, seed=seed, n_frames=200,
                         half=(0.9, 0.7, 0.5))
    cloud, geo, col = build_point_cloud(room, n_points, pixels_per_frame=1500, seed=seed,
                                        frame_ids=[0, 40], max_frames=40)
    if sparse:   # thin the cloud so some samples have 0/1 neighbours and some rays are invalid
        g = torch.Generator().manual_seed(seed + 1)
        keep = torch.rand(cloud.shape[0], generator=g) < 0.12
        cloud, geo, col = cloud[keep], geo[keep], col[keep]
    o, d, gdepth, gcol = sample_batch(room, [0, 40], n_rays // 2, seed + 2)
    if with_zero_depth:
        gdepth = gdepth.clone()
        gdepth[::7] = 0.0
    dyn = None
    if dynamic:   # per-ray float64 radii in [0.04, 0.16] (Tracker.py:255-258)
        g = torch.Generator().manual_seed(seed + 3)
        dyn = 0.04 + 0.12 * torch.rand(o.shape[0], generator=g, dtype=torch.float64)
    return room, cloud, geo, col, o, d, gdepth, gcol, dyn


def run_case(name, yaml, stage, is_tracker, seed, n_points=2000, n_rays=64, sparse=False,
             with_zero_depth=False, exposure=None, save_param_grads=True):
    common, decoder_mod, renderer_mod, config = ref_import.import_reference()
    cfg = ref_import.load_cfg(yaml)
    torch.manual_seed(1219)                         # one weight set per config family
    model = ref_import.build_model(cfg)
    dynamic = cfg['use_dynamic_radius']
    room, cloud, geo, col, o, d, gdepth, gcol, dyn = make_scene(
        seed, n_points, n_rays, sparse, with_zero_depth, dynamic)
    renderer = renderer_mod.Renderer(cfg, None, _SlamLike(room))
    renderer.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_mapper']
    npc = ExactKNNPointCloud(cloud, radius_query=cfg['pointcloud']['radius_query'])

    geo = geo.clone().requires_grad_(True)
    col = col.clone().requires_grad_(True)
    o = o.clone().requires_grad_(is_tracker)
    d = d.clone().requires_grad_(is_tracker)
    exposure_feat = None
    if cfg['model']['encode_exposure'] and exposure == 'feat':
        g = torch.Generator().manual_seed(seed + 5)
        exposure_feat = (torch.randn(cfg['model']['exposure_dim'], generator=g) * 0.5
                         ).requires_grad_(True)
        # the reference initialises mlp_exposure with std 0.01 -> affine ~ 0; make it matter
        with torch.no_grad():
            model.color_decoder.mlp_exposure.linear2.weight.mul_(30.0)
            model.color_decoder.mlp_exposure.linear2.bias.copy_(
                torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, 1, 0.05, -0.05, 0.02]))

    with _ZeroNoise():
        depth, var, rgb, valid = renderer.render_batch_ray(
            npc, model, d, o, 'cpu', stage, gt_depth=gdepth, npc_geo_feats=geo,
            npc_col_feats=col, is_tracker=is_tracker, cloud_pos=cloud,
            dynamic_r_query=dyn, exposure_feat=exposure_feat)
    g = torch.Generator().manual_seed(seed + 9)
    a = torch.randn(depth.shape, generator=g)
    b = torch.randn(rgb.shape, generator=g)
    loss = (a * depth).sum() + (b * rgb).sum()
    loss.backward()

    out = {
        'stage': stage, 'is_tracker': int(is_tracker), 'yaml': yaml,
        'ocfg': json.dumps(dataclasses.asdict(OracleCfg.from_cfg(cfg))),
        'intrinsics': np.array([room.H, room.W, room.fx, room.fy, room.cx, room.cy], np.float64),
        'cloud': cloud.numpy(), 'geo_feats': geo.detach().numpy(), 'col_feats': col.detach().numpy(),
        'rays_o': o.detach().numpy(), 'rays_d': d.detach().numpy(), 'gt_depth': gdepth.numpy(),
        'up_depth': a.numpy(), 'up_rgb': b.numpy(),
        'depth': depth.detach().numpy(), 'var': var.detach().numpy(), 'rgb': rgb.detach().numpy(),
        'valid': valid.numpy(),
        'g_geo_feats': geo.grad.numpy() if geo.grad is not None else np.zeros_like(geo.detach().numpy()),
        'g_col_feats': col.grad.numpy() if col.grad is not None else np.zeros_like(col.detach().numpy()),
    }
    if dyn is not None:
        out['dynamic_r'] = dyn.numpy()
    if is_tracker:
        out['g_rays_o'] = o.grad.numpy()
        out['g_rays_d'] = d.grad.numpy()
    if exposure_feat is not None:
        out['exposure_feat'] = exposure_feat.detach().numpy()
        out['g_exposure_feat'] = exposure_feat.grad.numpy()
    family = yaml.split('/')[1].lower()
    wts = {k: v.numpy() for k, v in model.state_dict().items()}
    wts['color_decoder.embedder._B'] = model.color_decoder.embedder._B.numpy()
    if exposure_feat is not None:
        family += '_exposure'                       # mlp_exposure was rescaled above
    np.savez_compressed(os.path.join(OUT, f'weights_{family}.npz'), **wts)
    out['weights_file'] = f'weights_{family}.npz'
    if save_param_grads:
        for k, p_ in model.named_parameters():
            if p_.grad is not None:
                out['gw/' + k] = p_.grad.numpy()
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: R={depth.shape[0]} N={cloud.shape[0]} valid={int(valid.sum())} '
          f'loss={loss.item():.6f} -> {os.path.getsize(path) / 1e6:.2f} MB')


CASES = [
    dict(name='replica_color_mapper', yaml='configs/Replica/room0.yaml', stage='color', is_tracker=False, seed=11),
    dict(name='replica_geometry_mapper', yaml='configs/Replica/room0.yaml', stage='geometry', is_tracker=False, seed=12),
    dict(name='replica_color_tracker', yaml='configs/Replica/room0.yaml', stage='color', is_tracker=True, seed=13,
         save_param_grads=False),
    dict(name='tum_color_mapper_dynr', yaml='configs/TUM_RGBD/freiburg1_desk.yaml', stage='color', is_tracker=False,
         seed=14, save_param_grads=False),
    dict(name='tum_color_tracker_dynr', yaml='configs/TUM_RGBD/freiburg1_desk.yaml', stage='color', is_tracker=True,
         seed=15, save_param_grads=False),
    dict(name='scannet_color_tracker_exposure', yaml='configs/ScanNet/scene0000.yaml', stage='color', is_tracker=True,
         seed=16, exposure='feat'),
    dict(name='scannet_color_mapper_presigmoid', yaml='configs/ScanNet/scene0000.yaml', stage='color', is_tracker=False,
         seed=17, save_param_grads=False),
    dict(name='replica_color_sparse_zero_depth', yaml='configs/Replica/office0.yaml', stage='color', is_tracker=False,
         seed=18, sparse=True, with_zero_depth=True, save_param_grads=False),
]

if __name__ == '__main__':
    _, decoder_mod, _, _ = ref_import.import_reference()
    _patch_reference(decoder_mod)
    only = set(sys.argv[1:])
    for case in CASES:
        if only and case['name'] not in only:
            continue
        run_case(**case)
