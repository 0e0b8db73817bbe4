"""GPU parity tests: the fused sm_100a path (through the C ABI) vs the oracle and the
reference-generated golden vectors.  Tolerances: forward <= 1e-4 relative vs the
fp32 oracle, masks exact, neighbour index sets bit-exact, gradients <= max(1e-4, 2x the fp32
oracle's own error) vs the fp64 restatement."""
import numpy as np
import pytest
import torch

import loopy_slam_b200 as L
from helpers import GOLDEN_CASES, Golden, rel_l2
from oracle.knn import exact_knn, neighbor_num, radius_sq, FLT_MAX
from parity import run_case_cuda_vs_oracle, run_cuda, cfg_from_ocfg, build_model, SlamLike

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _inradius(D, I, r2):
    if r2.dtype == torch.float64:
        keep = ~(D.to(torch.float64) > r2) & (I >= 0)
    else:
        keep = ~(D > r2) & (I >= 0)
    return torch.where(keep, D, torch.full_like(D, FLT_MAX)), torch.where(keep, I, torch.full_like(I, -1))


@pytest.mark.parametrize('n,p,cell,maxc', [(5000, 3000, 0.08, 1 << 22), (7, 200, 0.08, 1 << 22), (0, 50, 0.08, 1 << 10),
                                            (20000, 4000, 0.05, 1 << 12), (3000, 2000, 0.16, 1 << 22)])
def test_knn_bit_exact(n, p, cell, maxc):
    gen = torch.Generator().manual_seed(n + p)
    cloud = torch.rand(n, 3, generator=gen) * torch.tensor([2.0, 1.5, 1.0]) - 0.5
    if n > 100:
        cloud[10] = cloud[3]
        cloud[11] = cloud[3]                       # duplicates: tie-break by lower id
    q = torch.rand(p, 3, generator=gen) * torch.tensor([2.2, 1.7, 1.2]) - 0.6
    if n > 100:
        q[0] = cloud[3]
        q[1] = cloud[50] + torch.tensor([0.08, 0.0, 0.0])   # a point (nearly) exactly at r
    from loopy_slam_b200.renderer import GridIndex
    grid = GridIndex(cloud.to(DEV), cell, max_cells=maxc)
    for radius, dyn in ((0.08, None), (0.04, None), (0.08, 'dyn')):
        dr = None
        if dyn:
            dr = 0.04 + 0.12 * torch.rand(p, generator=gen, dtype=torch.float64)
        D, I, nn = grid.query(q.to(DEV), radius, None if dr is None else dr.to(DEV))
        D, I, nn = D.cpu(), I.cpu(), nn.cpu()
        Dr, Ir = exact_knn(q, cloud, 8)
        r2 = radius_sq(radius, dr)
        De, Ie = _inradius(Dr, Ir, r2)
        assert torch.equal(I, Ie), (radius, dyn)
        assert torch.equal(D, De)
        assert torch.equal(nn, neighbor_num(Dr, r2))


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_render_matches_oracle_and_fp64_truth(name):
    res = run_case_cuda_vs_oracle(name, DEV, verbose=True)
    assert res['ok'], res


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_render_matches_reference_golden(name):
    """Directly against the numbers the real reference produced (tests/golden/make_golden.py)."""
    g = Golden(name)
    ours = run_cuda(g, DEV)
    assert torch.equal(ours['valid'].bool(), g.t('valid'))
    torch.testing.assert_close(ours['depth'], g.t('depth'), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(ours['rgb'], g.t('rgb'), rtol=1e-4, atol=1e-5)
    # var = sum w (z - depth)^2 cancels two nearly equal 1e-2 m terms: 1e-4 relative on var needs ~1e-6 on depth, so the
    # bound is rtol 1e-4 plus an absolute floor of 3e-8 m^2 (= (0.17 mm)^2, fp32 rounding of z - depth at 3 m)
    torch.testing.assert_close(ours['var'], g.t('var'), rtol=1e-4, atol=3e-8)
    assert rel_l2(ours['g_geo'], g.t('g_geo_feats')) < 5e-4
    if g.stage == 'color':
        assert rel_l2(ours['g_col'], g.t('g_col_feats')) < 5e-4
    if g.is_tracker:
        assert rel_l2(ours['g_o'], g.t('g_rays_o')) < 2e-3
        assert rel_l2(ours['g_d'], g.t('g_rays_d')) < 2e-3
    if g.has('g_exposure_feat'):
        assert rel_l2(ours['g_ef'], g.t('g_exposure_feat')) < 5e-4
    for k, ref in g.param_grads.items():
        if k not in ours['g_w']:
            assert ref.abs().max() == 0, k
            continue
        assert rel_l2(ours['g_w'][k], ref) < 1e-3, k


def test_grad_pruning_flags():
    """needs_input_grad pruning: with frozen decoders and constant rays only feature grads come back."""
    g = Golden('replica_color_mapper')
    full = run_cuda(g, DEV, param_grads=True)
    pruned = run_cuda(g, DEV, param_grads=False)
    assert pruned['g_w'] == {}
    assert rel_l2(pruned['g_geo'], full['g_geo']) < 1e-5   # atomics reorder only
    assert rel_l2(pruned['g_col'], full['g_col']) < 1e-5
    torch.testing.assert_close(pruned['depth'], full['depth'], rtol=0, atol=0)


def test_sample_rays_and_pose_match_torch():
    gen = torch.Generator().manual_seed(3)
    H, W, fx, fy, cx, cy = 68, 120, 60.0, 61.0, 59.5, 33.5
    depth = torch.rand(H, W, generator=gen) + 0.5
    depth[::5, ::7] = 0
    color = torch.rand(H, W, 3, generator=gen)
    cam = torch.tensor([0.9, 0.1, -0.2, 0.3, 0.5, -0.4, 1.2], requires_grad=True)
    cam_g = cam.detach().clone().to(DEV).requires_grad_(True)
    # ours
    torch.manual_seed(5)
    c2w_g = L.get_camera_from_tensor(cam_g)
    ro, rd, sd, sc, i, j = L.get_samples(4, H - 4, 6, W - 6, 500, H, W, fx, fy, cx, cy, c2w_g, depth.to(DEV),
                                         color.to(DEV), DEV, depth_filter=True, return_index=True)
    up = torch.randn(rd.shape, generator=torch.Generator().manual_seed(9))
    ((rd * up.to(DEV)).sum() + (ro * up.to(DEV) * 0.5).sum()).backward()
    # torch restatement of common.py:104-120,301-343 on the same pixels
    c2w = torch.cat([L.quad2rotation(cam[None, :4])[0], cam[4:, None]], 1)
    ii, jj = i.cpu().float(), j.cpu().float()
    dirs = torch.stack([(ii - cx) / fx, -(jj - cy) / fy, -torch.ones_like(ii)], -1)
    rd_t = torch.sum(dirs[:, None, :] * c2w[:3, :3], -1)
    ro_t = c2w[:3, -1].expand(rd_t.shape)
    ((rd_t * up).sum() + (ro_t * up * 0.5).sum()).backward()
    torch.testing.assert_close(rd.detach().cpu(), rd_t.detach(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(ro.detach().cpu(), ro_t.detach(), rtol=0, atol=1e-7)
    assert (sd > 0).all() and torch.equal(sd.cpu(), depth[j.cpu(), i.cpu()])
    torch.testing.assert_close(sc.cpu(), color[j.cpu(), i.cpu()])
    assert (i >= 6).all() and (i < W - 6).all() and (j >= 4).all() and (j < H - 4).all()
    assert rel_l2(cam_g.grad, cam.grad) < 1e-5


def test_get_samples_matches_reference_golden():
    """L.get_samples / get_camera_from_tensor through the C ABI vs the vectors of the REAL reference
    (tests/golden/sampling.npz): same pixel picks (global CUDA generator replaced by the stored indices through
    torch.randint monkeypatching is not possible -- the CUDA generator differs from the CPU one -- so the picks are
    injected: the kernel consumes `pix`), rays within 1e-6, gathered depth / colour and (i, j) exact."""
    import os
    from loopy_slam_b200.common import _SampleRaysFn
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'sampling.npz'))
    depth, color = torch.from_numpy(z['depth']).to(DEV), torch.from_numpy(z['color']).to(DEV)
    H, W, fx, fy, cx, cy = z['intr']
    c2w = L.get_camera_from_tensor(torch.from_numpy(z['cam']).to(DEV))
    torch.testing.assert_close(c2w.cpu(), torch.from_numpy(z['c2w']), rtol=1e-6, atol=1e-7)
    for k in range(3):
        H0, H1, W0, W1, n, filt, lim = z[f'case{k}']
        pix = torch.from_numpy(z[f'idx{k}']).to(DEV)
        geom = (int(H), int(W), float(fx), float(fy), float(cx), float(cy), int(H0), int(H1), int(W0), int(W1))
        o, d, sd, sc, i, j = _SampleRaysFn.apply(torch.from_numpy(z['c2w']).to(DEV), depth, color, pix, geom)
        mask = torch.ones_like(sd, dtype=torch.bool)
        if filt:
            mask = sd > 0
            if lim > 0:
                mask = mask & (sd < lim)
        for got, name, tol in ((o[mask], 'o', 1e-6), (d[mask], 'd', 1e-6)):
            torch.testing.assert_close(got.cpu(), torch.from_numpy(z[f'{name}{k}']), rtol=tol, atol=tol)
        assert torch.equal(sd[mask].cpu(), torch.from_numpy(z[f'sd{k}']))
        assert torch.equal(sc[mask].cpu(), torch.from_numpy(z[f'sc{k}']))
        assert torch.equal(i[mask].cpu(), torch.from_numpy(z[f'i{k}']))
        assert torch.equal(j[mask].cpu(), torch.from_numpy(z[f'j{k}']))
        if filt:   # the one-launch filter + order-preserving compaction the public get_samples uses
            from loopy_slam_b200.common import _SampleRaysFilteredFn
            o2, d2, sd2, sc2, i2, j2 = _SampleRaysFilteredFn.apply(torch.from_numpy(z['c2w']).to(DEV), depth, color, pix, geom,
                                                                  None if lim < 0 else float(lim))
            for got, ref in ((o2, o[mask]), (d2, d[mask]), (sd2, sd[mask]), (sc2, sc[mask]), (i2, i[mask]), (j2, j[mask])):
                assert torch.equal(got, ref)


def test_frustum_selection_kernel_matches_reference_restatement():
    """lsr_frustum_mask (two-pass CUDA kernel) vs oracle/frustum.py = the reference's numpy + cv2.remap lines
    (src/Mapper.py:165-217).  cv2.remap interpolates with 1/32-pixel fixed-point weights, so a point whose camera depth
    sits within 1 mm of `lookup + 0.5` or whose projection sits within 1e-3 px of the crop edge may flip: at most 0.2 %
    of the points may differ, and none that is not such a borderline case."""
    from loopy_slam_b200.frustum import get_mask_from_c2w
    from loopy_slam_b200.stream import SyntheticRoom, build_point_cloud
    from oracle.frustum import get_mask_from_c2w as ref_mask
    room = SyntheticRoom(H=96, W=128, fx=80., fy=80., cx=63.5, cy=47.5, n_frames=100, half=(1.0, 0.8, 0.6), hole_frac=0.02)
    cloud, _, _ = build_point_cloud(room, 20000, pixels_per_frame=4000, frame_ids=[0, 10, 20, 30, 50, 70], max_frames=60)
    for fid, edge in ((5, -4), (40, 0), (75, 10)):
        color, depth, c2w = room.frame(fid)
        got = get_mask_from_c2w(cloud.to(DEV), c2w, depth.to(DEV), room.H, room.W, room.fx, room.fy, room.cx, room.cy, edge=edge)
        ref = ref_mask(cloud.numpy(), c2w.numpy(), depth.numpy(), room.H, room.W, room.fx, room.fy, room.cx, room.cy, edge)
        a, b = set(got.cpu().tolist()), set(ref.tolist())
        assert len(b) > 200
        assert len(a ^ b) <= 0.002 * cloud.shape[0], (fid, len(a), len(b), len(a ^ b))
        assert got.dtype == torch.int64 and bool((got[1:] > got[:-1]).all())     # sorted row ids, like np.where


def test_frustum_kernel_matches_reference_golden():
    """lsr_frustum_mask vs row ids minted from the REAL Mapper.get_mask_from_c2w (tests/golden/make_golden_frustum.py): the
    only admissible differences are points whose bilinear depth test sits on the 0.5 m threshold within fp32 rounding."""
    import os
    from loopy_slam_b200.frustum import get_mask_from_c2w, filter_point_before_add
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'frustum.npz'))
    H, W, fx, fy, cx, cy = G['intr']
    cloud = torch.from_numpy(G['cloud']).to(DEV)
    for k, (fid, edge) in enumerate(G['cases']):
        got = get_mask_from_c2w(cloud, torch.from_numpy(G[f'c2w{k}']), torch.from_numpy(G[f'depth{k}']).to(DEV), int(H), int(W),
                                fx, fy, cx, cy, edge=int(edge))
        a, b = set(got.cpu().tolist()), set(G[f'idx{k}'].tolist())
        assert len(b) > 300 and len(a ^ b) <= 0.002 * cloud.shape[0], (int(fid), len(a), len(b), len(a ^ b))
    m = filter_point_before_add(torch.from_numpy(G['f_o']).to(DEV), torch.from_numpy(G['f_d']).to(DEV),
                                torch.from_numpy(G['f_g']).to(DEV), torch.from_numpy(G['f_prev']), int(H), int(W), fx, fy, cx, cy)
    assert torch.equal(m.cpu(), torch.from_numpy(G['f_mask']))


def test_sample_near_pcl_matches_reference_golden():
    """NeuralPointCloud.sample_near_pcl on the device (grid k-NN, torch glue) vs the REAL reference method on the same cloud
    (tests/golden/make_golden_point_store.py): same invalid rays, sample depths to fp32 rounding of the float64 linspace."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'point_store.npz'))
    cfg = L.default_cfg('replica')
    cfg['mapping']['device'] = DEV
    npc = L.NeuralPointCloud(cfg, device=DEV)
    npc.set_cloud(torch.from_numpy(G['snp_cloud']).to(DEV))
    z, invalid = npc.sample_near_pcl(torch.from_numpy(G['snp_o']).to(DEV), torch.from_numpy(G['snp_d']).to(DEV), 0.3,
                                     float(G['snp_far']), 5)
    assert torch.equal(invalid.cpu(), torch.from_numpy(G['snp_invalid']))
    torch.testing.assert_close(z.cpu(), torch.from_numpy(G['snp_z']), rtol=0, atol=2e-7)


def test_render_img_matches_tiled_oracle():
    """render_img (one fused launch, per-3000-ray far statistics) vs the oracle run tile by tile."""
    from oracle import render as orc
    g = Golden('replica_color_sparse_zero_depth')
    cfg = cfg_from_ocfg(g.ocfg)
    gen = torch.Generator().manual_seed(21)
    H, W = 24, 30
    fx = fy = 20.0
    cx, cy = 14.5, 11.5
    model = build_model(cfg, g.weights, DEV)
    rend = L.Renderer(cfg, None, SlamLike(H, W, fx, fy, cx, cy), ray_batch_size=100)
    rend.sigmoid_coefficient = g.ocfg.sigmoid_coef
    gdc = Golden('replica_color_mapper')     # dense cloud, camera from its first ray origin
    cloud, geo, col = gdc.t('cloud'), gdc.t('geo_feats'), gdc.t('col_feats')
    c2w = torch.eye(4)
    c2w[:3, 3] = gdc.t('rays_o')[0]
    dvec = gdc.t('rays_d')[0]
    zc = -dvec / dvec.norm()
    xc = torch.linalg.cross(torch.tensor([0., 0., 1.]), zc)
    xc = xc / xc.norm()
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2] = xc, torch.linalg.cross(zc, xc), zc
    gt = 0.4 + torch.rand(H, W, generator=gen) * 0.8
    gt[::4, ::3] = 0.0

    class NPC:
        def get_radius_query(self):
            return g.ocfg.radius_query
    d_img, v_img, c_img = rend.render_img(NPC(), model, c2w.to(DEV), DEV, 'color', gt_depth=gt.to(DEV),
                                          npc_geo_feats=geo.to(DEV), npc_col_feats=col.to(DEV), cloud_pos=cloud.to(DEV))
    assert d_img.dtype == torch.float64 and d_img.shape == (H, W) and c_img.shape == (H, W, 3)
    ro, rd = L.get_rays(H, W, fx, fy, cx, cy, c2w.to(DEV), DEV)
    ro, rd = ro.reshape(-1, 3).cpu(), rd.reshape(-1, 3).cpu()
    W_ = {k: v for k, v in g.weights.items()}
    outs = []
    for s in range(0, H * W, 100):
        dd, vv, cc, _, _ = orc.render_rays(W_, g.ocfg, ro[s:s + 100], rd[s:s + 100], gt.reshape(-1)[s:s + 100], geo, col,
                                           cloud, 'color')
        outs.append((dd, vv, cc))
    d_ref = torch.cat([o[0] for o in outs]).reshape(H, W)
    c_ref = torch.cat([o[2] for o in outs]).reshape(H, W, 3)
    assert (d_img.cpu()[gt == 0] == 0).all()
    # SURVEY 8c forward contract (rtol 1e-4); atol = 2e-5 colour units / 1e-5 m for values that cancel to ~0
    torch.testing.assert_close(d_img.cpu().float(), d_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(c_img.cpu(), c_ref, rtol=1e-4, atol=2e-5)


def test_eval_points_matches_oracle_decode():
    from oracle import render as orc
    g = Golden('replica_color_mapper')
    cfg = cfg_from_ocfg(g.ocfg)
    model = build_model(cfg, g.weights, DEV)
    H, W, fx, fy, cx, cy = g.raw['intrinsics']
    rend = L.Renderer(cfg, None, SlamLike(H, W, fx, fy, cx, cy))
    o32 = __import__('parity').run_oracle(g, torch.float32)
    pts = o32['aux']['p']

    class NPC:
        def get_radius_query(self):
            return g.ocfg.radius_query
    raw, ray_mask, point_mask = rend.eval_points(pts.to(DEV), model, NPC(), 'color', DEV, g.t('geo_feats').to(DEV),
                                                 g.t('col_feats').to(DEV), False, g.t('cloud').to(DEV), None,
                                                 ray_pts_num=5)
    ref_raw = o32['aux']['raw'].reshape(-1, 4)
    has = o32['aux']['has']
    assert torch.equal(point_mask.cpu(), has)
    assert torch.equal(ray_mask.cpu(), o32['valid'])
    torch.testing.assert_close(raw.cpu()[has][:, :3], ref_raw[has][:, :3], rtol=1e-4, atol=1e-5)
    # occupancy logits: the pretrained geometry decoder produces logits of magnitude 1e2..1e3 from cancelling terms, so
    # 1e-4 relative + an absolute floor of 2e-3 logit units (sigmoid(0.1 * x) changes by < 5e-5 for that)
    torch.testing.assert_close(raw.cpu()[has][:, 3], ref_raw[has][:, 3], rtol=1e-4, atol=2e-3)


def test_nicer_forward_mesh_and_color_only_stages():
    """NICER.forward stages 'mesh' (decoder.py:611-620: the mesher's radius, no ray mask) and 'color_only' (:621-626)."""
    g = Golden('replica_color_mapper')
    cfg = cfg_from_ocfg(g.ocfg)
    model = build_model(cfg, g.weights, DEV)
    H, W, fx, fy, cx, cy = g.raw['intrinsics']
    rend = L.Renderer(cfg, None, SlamLike(H, W, fx, fy, cx, cy))
    pts = (g.t('rays_o') + g.t('rays_d') * g.t('gt_depth').reshape(-1, 1)).to(DEV)
    geo, col, cloud = g.t('geo_feats').to(DEV), g.t('col_feats').to(DEV), g.t('cloud').to(DEV)

    class NPC:
        radius_mesh = 0.6 * g.ocfg.radius_query

        def __init__(self, r):
            self.r = r

        def get_radius_query(self):
            return self.r
    model._lsr_renderer = rend
    full, ray_mask, pm = model(pts, NPC(g.ocfg.radius_query), 'color', geo, col, pts_num=1, cloud_pos=cloud)
    only = model(pts, NPC(g.ocfg.radius_query), 'color_only', geo, col, pts_num=1, cloud_pos=cloud)
    assert only.shape == (pts.shape[0], 3) and torch.equal(only, full[:, :3])
    mesh, rm, pm_mesh = model(pts, NPC(g.ocfg.radius_query), 'mesh', geo, col, pts_num=1, cloud_pos=cloud)
    small, _, pm_small = model(pts, NPC(NPC.radius_mesh), 'color', geo, col, pts_num=1, cloud_pos=cloud)
    assert rm is None and torch.equal(mesh, small) and torch.equal(pm_mesh, pm_small)
    assert not torch.equal(mesh, full)                      # the smaller radius really changed the neighbourhoods


def test_neural_point_cloud_insert_and_query():
    cfg = L.default_cfg('replica')
    cfg['mapping']['device'] = DEV
    npc = L.NeuralPointCloud(cfg)
    gen = torch.Generator().manual_seed(4)
    o = torch.zeros(400, 3)
    d = torch.randn(400, 3, generator=gen)
    d = d / d.norm(dim=1, keepdim=True)
    dep = 1.0 + 0.2 * torch.rand(400, generator=gen)
    col = torch.rand(400, 3, generator=gen)
    n1 = int(npc.add_neural_points(o.to(DEV), d.to(DEV), dep.to(DEV), col.to(DEV)))
    assert n1 == 400 and npc.pts_num() == 1200
    n2 = int(npc.add_neural_points(o.to(DEV), d.to(DEV), dep.to(DEV), col.to(DEV)))   # same locations again
    assert n2 == 0 and npc.pts_num() == 1200
    D, I, nn = npc.find_neighbors_faiss((o + d * dep[:, None]).to(DEV), step='query')
    Dr, Ir = exact_knn(o + d * dep[:, None], npc.cloud_pos_tensor().cpu(), 8)
    r2 = radius_sq(0.08)
    _, Ie = _inradius(Dr, Ir, r2)
    assert torch.equal(I.cpu(), Ie) and (nn >= 1).all()
    assert npc.get_geo_feats().shape == (1200, 32) and abs(float(npc.get_geo_feats().std()) - 0.1) < 0.01


def test_segmented_point_store_matches_oracle():
    """f1: a stream of frames through NeuralPointCloud.add_neural_points on the device vs the CPU
    restatement of the reference store (oracle/point_store.py: add_neural_points :1557-1631, check_index /
    init_segment :1220-1315, update_fragments): same kept samples, same inserted positions, same segment boundaries and
    inheritance masks, same merged end-of-run cloud and features."""
    from loopy_slam_b200.stream import SyntheticRoom, sample_batch
    from oracle.point_store import PointStoreOracle
    room = SyntheticRoom(H=64, W=64, fx=40., fy=40., cx=31.5, cy=31.5, n_frames=60, half=(0.9, 0.7, 0.5), hole_frac=0.02)
    cfg = L.default_cfg('replica')
    cfg['mapping']['device'] = DEV
    cfg['mapping']['segment_rot_cos'] = 0.985        # ~10 degrees: several segments within the short stream
    cfg['mapping']['segment_rel_trans'] = 0.12

    class Slam:
        H, W, fx, fy, cx, cy = room.H, room.W, room.fx, room.fy, room.cx, room.cy
    npc = L.NeuralPointCloud(cfg, Slam)
    orc_ = PointStoreOracle(room.H, room.W, room.fx, room.fy, room.cx, room.cy, segment_rot_cos=0.985, segment_rel_trans=0.12)
    feat = lambda p, which: torch.sin(p.cpu().double() @ torch.arange(1, 97, dtype=torch.float64).reshape(3, 32) * (1 + which)).float()
    for fid in range(0, 60, 3):
        o, d, g, c = sample_batch(room, [fid], 500, seed=100 + fid)
        c2w = room.frame(fid)[2]
        n_before = npc.fragments[-1].n if npc.fragments else 0
        segs_before = len(npc.fragments)
        kept = npc.add_neural_points(o.to(DEV), d.to(DEV), g.to(DEV), c.to(DEV), idx=torch.tensor(fid), cur_c2w=c2w)
        seg = npc.fragments[-1]
        first_new = seg.n_inherited if len(npc.fragments) != segs_before else n_before
        new_pos = seg.pos[first_new:seg.n]
        seg.geo[first_new:seg.n] = feat(new_pos, 0).to(DEV)          # deterministic features instead of N(0, 0.1) draws
        seg.col[first_new:seg.n] = feat(new_pos, 1).to(DEV)
        kept_ref = orc_.add_neural_points(o, d, g, fid, c2w, feat_fn=feat)
        assert int(kept) == kept_ref, fid
    keys = list(orc_.fragments_dict.keys())
    assert len(npc.fragments) == len(keys) >= 3
    for seg, k in zip(npc.fragments, keys):
        f = orc_.fragments_dict[k]
        assert seg.n == len(f['npc']) and seg.n_inherited == f['idx_start_segment_features']
        torch.testing.assert_close(seg.pos[:seg.n].cpu(), torch.tensor(f['npc'], dtype=torch.float32), rtol=0, atol=1e-6)
        if f['mask'] is not None:
            assert torch.equal(seg.mask.cpu(), torch.from_numpy(f['mask']))
    np.testing.assert_allclose(npc.get_cloud_pos(True).cpu().numpy(), orc_.merged('npc'), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npc.get_geo_feats(True).cpu().numpy(), orc_.merged('geo_feats'), rtol=1e-5, atol=1e-6)
    # the active index serves queries with row ids of the ACTIVE segment
    D, I, nn = npc.find_neighbors_faiss(npc.get_cloud_pos()[:50], step='query')
    assert bool((D[:, 0] == 0).all()) and torch.equal(npc.get_cloud_pos()[I[:, 0]], npc.get_cloud_pos()[:50])   # (a batch may hold one pixel twice)
    npc.train_index_global()
    assert npc.index_ntotal() == orc_.merged('npc').shape[0]


@pytest.mark.parametrize('strategy', ['rot_trans', 'fixed'])
def test_point_store_matches_reference_golden(strategy):
    """The device point store against vectors minted from the REAL reference methods (tests/golden/make_golden_point_store.py):
    same kept samples, segment boundaries, inherited-point masks, inserted positions and merged end-of-run cloud."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'point_store.npz'))
    H, W, fx, fy, cx, cy = G['intr']
    rel_trans, rot_cos, fixed = G[f'{strategy}_cfg']
    cfg = L.default_cfg('replica')
    cfg['mapping'].update(device=DEV, segment_strategy=strategy, fixed_segment_size=int(fixed), segment_rel_trans=float(rel_trans),
                          segment_rot_cos=float(rot_cos))

    class Slam:
        pass
    Slam.H, Slam.W, Slam.fx, Slam.fy, Slam.cx, Slam.cy = int(H), int(W), float(fx), float(fy), float(cx), float(cy)
    npc = L.NeuralPointCloud(cfg, Slam)
    kept = []
    for fid in G[f'{strategy}_fids']:
        t = lambda k: torch.from_numpy(G[f'{strategy}_{k}{fid}'])
        o = t('o').to(DEV)
        kept.append(int(npc.add_neural_points(o, t('d').to(DEV), t('g').to(DEV), torch.zeros_like(o), idx=torch.tensor(int(fid)),
                                              cur_c2w=t('c2w'))))
    assert kept == G[f'{strategy}_kept'].tolist()
    nseg = int(G[f'{strategy}_nseg'])
    assert len(npc.fragments) == nseg
    for k, seg in enumerate(npc.fragments):
        ref = G[f'{strategy}_seg{k}_npc']
        assert [seg.start_idx, seg.n_inherited] == G[f'{strategy}_seg{k}_start'].tolist() and seg.n == ref.shape[0]
        torch.testing.assert_close(seg.pos[:seg.n].cpu().double(), torch.from_numpy(ref), rtol=0, atol=1e-6)
        m = G[f'{strategy}_seg{k}_mask']
        if k < nseg - 1:
            assert torch.equal(seg.mask.cpu(), torch.from_numpy(m))
        else:
            assert seg.mask is None and m.size == 0
    np.testing.assert_allclose(npc.get_cloud_pos(True).cpu().numpy(), G[f'{strategy}_end_pos'], rtol=1e-5, atol=1e-6)


def test_dynamic_radius_map_matches_numpy_restatement():
    """vs a scipy/numpy restatement of skimage 0.19.3 rgb2gray + sobel_h/sobel_v (reflect) + interp1d
    (src/Tracker.py:243-258).  skimage itself is not installed: pinned to its documented kernels."""
    from scipy import ndimage as ndi
    from loopy_slam_b200.radius_map import dynamic_radius_maps
    cfg = L.default_cfg('tum')
    gen = torch.Generator().manual_seed(8)
    img = torch.rand(37, 53, 3, generator=gen, dtype=torch.float64)
    img[10:20, 5:30] = 0.3                                  # flat patch -> r_max
    gray = img.numpy() @ np.array([0.2125, 0.7154, 0.0721])
    smooth, edge = np.array([1., 2., 1.]) / 4.0, np.array([1., 0., -1.])
    gy = ndi.convolve(gray, np.outer(edge, smooth), mode='reflect')
    gx = ndi.convolve(gray, np.outer(smooth, edge), mode='reflect')
    mag = np.clip(np.sqrt(gx ** 2 + gy ** 2), 0.0, 0.15)
    ref_add = np.interp(mag, [0, 0.01, 0.15], [0.08, 0.08, 0.02])
    ref_q = np.interp(mag, [0, 0.01, 0.15], [0.16, 0.16, 0.04])
    for dt in (torch.float64, torch.float32):
        ra, rq = dynamic_radius_maps(img.to(DEV, dt), cfg)
        assert ra.dtype == torch.float64 and ra.shape == (37, 53)
        tol = 1e-12 if dt == torch.float64 else 1e-6
        np.testing.assert_allclose(ra.cpu().numpy(), ref_add, rtol=0, atol=tol)
        np.testing.assert_allclose(rq.cpu().numpy(), ref_q, rtol=0, atol=2 * tol)
    assert float(ra[12, 10]) == 0.08


def test_edge_cases_empty_and_tiny():
    """R = 0 rays, empty cloud, N < 8 points, gt_depth=None, N_surface != 5."""
    cfg = L.default_cfg('replica')
    torch.manual_seed(3)
    model = L.get_model(cfg).to(DEV)
    rend = L.Renderer(cfg, None, SlamLike(64, 64, 40., 40., 31.5, 31.5))
    rend.sigmoid_coefficient = 0.1

    class NPC:
        def get_radius_query(self):
            return 0.08
    gen = torch.Generator().manual_seed(1)
    cloud5 = (torch.rand(5, 3, generator=gen) * 0.05).to(DEV)
    feats5 = torch.randn(5, 32, generator=gen).to(DEV)
    o = torch.zeros(7, 3, device=DEV)
    d = torch.tensor([[0.01, 0.01, 1.0]], device=DEV).repeat(7, 1) * 0.02
    g = torch.ones(7, device=DEV)
    # N < 8: every neighbour list is padded, all five points are within the radius of the first samples
    dep, var, rgb, valid = rend.render_batch_ray(NPC(), model, d, o, DEV, 'color', gt_depth=g, npc_geo_feats=feats5,
                                                 npc_col_feats=feats5, cloud_pos=cloud5)
    assert torch.isfinite(dep).all() and torch.isfinite(rgb).all() and valid.dtype == torch.bool and valid.all()
    # empty cloud: nothing has neighbours -> invalid rays, finite outputs
    dep, var, rgb, valid = rend.render_batch_ray(NPC(), model, d, o, DEV, 'geometry', gt_depth=g,
                                                 npc_geo_feats=torch.zeros(0, 32, device=DEV),
                                                 npc_col_feats=torch.zeros(0, 32, device=DEV),
                                                 cloud_pos=torch.zeros(0, 3, device=DEV))
    assert (~valid).all() and torch.isfinite(dep).all()
    # R = 0
    dep, var, rgb, valid = rend.render_batch_ray(NPC(), model, d[:0], o[:0], DEV, 'color', gt_depth=g[:0],
                                                 npc_geo_feats=feats5, npc_col_feats=feats5, cloud_pos=cloud5)
    assert dep.shape == (0,) and rgb.shape == (0, 3)
    # gt_depth None: z in [near_end, 10], depth forced to 0 (Renderer.py:107-113,197-198)
    dep, var, rgb, valid = rend.render_batch_ray(NPC(), model, d, o, DEV, 'color', gt_depth=None, npc_geo_feats=feats5,
                                                 npc_col_feats=feats5, cloud_pos=cloud5)
    assert (dep == 0).all() and torch.isfinite(var).all()
    # N_surface = 3 (tile = 21 rays x 3 samples)
    cfg3 = L.default_cfg('replica')
    cfg3['rendering']['N_surface'] = 3
    model3 = L.get_model(cfg3).to(DEV)
    rend3 = L.Renderer(cfg3, None, SlamLike(64, 64, 40., 40., 31.5, 31.5))
    rend3.sigmoid_coefficient = 0.1
    geo = feats5.clone().requires_grad_(True)
    dep, var, rgb, valid = rend3.render_batch_ray(NPC(), model3, d, o, DEV, 'color', gt_depth=g, npc_geo_feats=geo,
                                                  npc_col_feats=feats5, cloud_pos=cloud5)
    (dep.sum() + rgb.sum()).backward()
    assert torch.isfinite(geo.grad).all() and geo.grad.abs().sum() > 0
    assert ((dep >= 0.98 * g - 1e-5) & (dep <= 1.02 * g + 1e-5)).all()


@pytest.mark.parametrize('name,S', [('replica_color_mapper', 3), ('replica_color_tracker', 3), ('tum_color_mapper_dynr', 7)])
def test_other_n_surface_parity(name, S):
    """N_surface != 5 (tiles of floor(128 / S) rays): full forward + gradient parity against the fp32 / fp64 oracle, not
    just finiteness (goldens only cover N_surface = 5)."""
    import dataclasses
    g = Golden(name)
    g.ocfg = dataclasses.replace(g.ocfg, N_surface=S)
    res = run_case_cuda_vs_oracle(g, DEV, verbose=True)
    assert res['ok'], res


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['replica_color_mapper', 'tum_color_mapper_dynr'])
def test_feature_subset_equals_index_put_flow(name):
    """row_remap / leaf blocks (FeatureSubset) == the reference's table[indices] = leaf flow
    (src/Mapper.py:581-582): identical forward, leaf gradients == gathered table gradients."""
    if name not in GOLDEN_CASES:
        pytest.skip('golden case not present')
    g = Golden(name)
    dev = torch.device('cuda:0')
    cfg = cfg_from_ocfg(g.ocfg)
    H, W, fx, fy, cx, cy = g.raw['intrinsics']
    model = build_model(cfg, g.weights, dev)
    rend = L.Renderer(cfg, None, SlamLike(H, W, fx, fy, cx, cy))
    rend.sigmoid_coefficient = g.ocfg.sigmoid_coef
    cloud = g.t('cloud').to(dev)
    N = cloud.shape[0]
    gen = torch.Generator().manual_seed(5)
    indices = torch.randperm(N, generator=gen)[: (2 * N) // 3].to(dev)          # a third of the rows is frozen
    geo_tab, col_tab = g.t('geo_feats').to(dev), g.t('col_feats').to(dev)
    dyn = g.t('dynamic_r').to(dev) if g.has('dynamic_r') else None

    class NPC:
        def get_radius_query(self_inner):
            return g.ocfg.radius_query

    def run(use_subset):
        for p in model.parameters():
            p.grad = None
        geo_leaf = (geo_tab[indices] + 0.01).clone().requires_grad_(True)        # leaf differs from the stale table rows
        col_leaf = (col_tab[indices] - 0.02).clone().requires_grad_(True)
        kw = dict(gt_depth=g.t('gt_depth').to(dev), is_tracker=False, cloud_pos=cloud, dynamic_r_query=dyn)
        if use_subset:
            sub = L.FeatureSubset(indices, N)
            out = rend.render_batch_ray(NPC(), model, g.t('rays_d').to(dev), g.t('rays_o').to(dev), dev, g.stage,
                                        npc_geo_feats=geo_tab, npc_col_feats=col_tab,
                                        feat_subset=(sub, geo_leaf, col_leaf), **kw)
        else:
            gt_, ct_ = geo_tab.index_put((indices,), geo_leaf), col_tab.index_put((indices,), col_leaf)
            out = rend.render_batch_ray(NPC(), model, g.t('rays_d').to(dev), g.t('rays_o').to(dev), dev, g.stage,
                                        npc_geo_feats=gt_, npc_col_feats=ct_, **kw)
        depth, var, rgb, valid = out
        loss = (g.t('up_depth').to(dev) * depth).sum() + (g.t('up_rgb').to(dev) * rgb).sum()
        loss.backward()
        w = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        return depth.detach(), var.detach(), rgb.detach(), valid, geo_leaf.grad, col_leaf.grad, w

    a, b = run(False), run(True)
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    for x, y in zip(a[4:6], b[4:6]):
        assert x is not None and y is not None and float(x.abs().max()) > 0
        assert rel_l2(y.cpu(), x.cpu()) < 1e-5          # atomics order only
    assert a[6].keys() == b[6].keys()
    for k in a[6]:
        assert rel_l2(b[6][k].cpu(), a[6][k].cpu()) < 1e-4, k


def test_render_img_with_sample_near_pcl_matches_tile_loop():
    """rendering.sample_near_pcl=True through render_img (Renderer.py:243-266 calls render_batch_ray per tile, so
    zero-depth pixels are sampled near the cloud with THAT tile's far bound): one fused launch vs our own tile loop over
    render_batch_ray (checked against the oracle in test_sample_near_pcl_on_the_fused_path); also gt_depth=None."""
    g = Golden('replica_color_sparse_zero_depth')
    cfg = cfg_from_ocfg(g.ocfg)
    cfg['rendering']['sample_near_pcl'] = True
    gen = torch.Generator().manual_seed(22)
    H, W = 24, 30
    model = build_model(cfg, g.weights, DEV)
    rend = L.Renderer(cfg, None, SlamLike(H, W, 20.0, 20.0, 14.5, 11.5), ray_batch_size=100)
    rend.sigmoid_coefficient = g.ocfg.sigmoid_coef
    gdc = Golden('replica_color_mapper')
    cloud, geo, col = gdc.t('cloud').to(DEV), gdc.t('geo_feats').to(DEV), gdc.t('col_feats').to(DEV)
    c2w = torch.eye(4)
    c2w[:3, 3] = gdc.t('rays_o')[0]
    dvec = gdc.t('rays_d')[0]
    zc = -dvec / dvec.norm()
    xc = torch.linalg.cross(torch.tensor([0., 0., 1.]), zc)
    xc = xc / xc.norm()
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2] = xc, torch.linalg.cross(zc, xc), zc
    gt = 0.4 + torch.rand(H, W, generator=gen) * 0.8
    gt[::4, ::3] = 0.0
    gt[20:, :] = 0.0                                         # whole tiles without any sensor depth
    qcfg = L.default_cfg('replica')
    qcfg['pointcloud']['radius_query'] = g.ocfg.radius_query
    qcfg['mapping']['device'] = DEV
    npc = L.NeuralPointCloud(qcfg, device=DEV)
    npc.set_cloud(cloud, geo, col)
    ro, rd = L.get_rays(H, W, 20.0, 20.0, 14.5, 11.5, c2w.to(DEV), DEV)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    for gt_img in (gt.to(DEV), None):
        d_img, v_img, c_img = rend.render_img(npc, model, c2w.to(DEV), DEV, 'color', gt_depth=gt_img, npc_geo_feats=geo,
                                              npc_col_feats=col, cloud_pos=cloud)
        outs = []
        with torch.no_grad():
            for s0 in range(0, H * W, 100):
                gg = gt_img.reshape(-1)[s0:s0 + 100] if gt_img is not None else None
                outs.append(rend.render_batch_ray(npc, model, rd[s0:s0 + 100], ro[s0:s0 + 100], DEV, 'color', gt_depth=gg,
                                                  npc_geo_feats=geo, npc_col_feats=col, cloud_pos=cloud))
        d_ref = torch.cat([o[0] for o in outs]).reshape(H, W).double()
        c_ref = torch.cat([o[2] for o in outs]).reshape(H, W, 3)
        assert rel_l2(d_img.cpu(), d_ref.cpu()) < 1e-6 and rel_l2(c_img.cpu(), c_ref.cpu()) < 1e-6
        assert float(d_img.abs().max()) > 0
        if gt_img is not None:
            assert (d_img[gt_img == 0] != 0).any()           # zero-depth pixels keep their rendered depth on this path


def test_sample_near_pcl_on_the_fused_path():
    """rendering.sample_near_pcl=True (configs/point_slam.yaml:127; Renderer.py:150-158,191-198): zero-depth rays take
    their samples from NeuralPointCloud.sample_near_pcl, keep their rendered depth, and are invalid when the cloud is
    not near.  CUDA (mirror npc on the grid k-NN + z override in the fused kernels) vs the oracle restatement."""
    import torch
    from helpers import Golden, rel_l2
    from oracle import render as orc
    from parity import run_cuda, run_oracle, grad_ok
    g = Golden('replica_color_sparse_zero_depth')
    gt = g.t('gt_depth').reshape(-1)
    zero = ~(gt > 0)
    assert zero.any() and (~zero).any()
    S = g.ocfg.N_surface
    far = float(orc.far_for_zero_depth(gt))
    z0, invalid = orc.sample_near_pcl(g.t('rays_o')[zero], g.t('rays_d')[zero], g.ocfg.near_end, far, S, g.t('cloud'),
                                      g.ocfg.radius_query)
    z_zero = torch.zeros(gt.shape[0], S)
    z_zero[zero] = z0
    o32, o64 = run_oracle(g, torch.float32, z_zero=z_zero), run_oracle(g, torch.float64, z_zero=z_zero)
    valid_ref = o32['valid'].clone()
    valid_ref[torch.nonzero(zero, as_tuple=True)[0][invalid]] = False
    ours = run_cuda(g, near_pcl=True)
    assert torch.equal(ours['valid'].bool(), valid_ref)
    assert rel_l2(ours['depth'], o32['depth']) < 1e-4 and rel_l2(ours['rgb'], o32['rgb']) < 1e-4
    assert rel_l2(ours['var'], o32['var']) < 1e-4
    # zero-depth rays keep a (non-zero) rendered depth on this path, and its gradient flows
    assert (ours['depth'][zero] != 0).any()
    for key in ('g_geo', 'g_col'):
        ok, e, eo = grad_ok(ours[key], o64[key], o32[key])
        assert ok, (key, e, eo)
