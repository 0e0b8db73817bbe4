"""CPU: pin the oracle restatement (oracle/render.py, oracle/knn.py) against golden vectors
produced by the REAL reference code (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from oracle import render as orc
from oracle.knn import exact_knn, squared_dist_f32
from helpers import GOLDEN_CASES, Golden, rel_l2


def _run(g, dtype=torch.float32):
    geo = g.t('geo_feats').clone().requires_grad_(True)
    col = g.t('col_feats').clone().requires_grad_(True)
    o = g.t('rays_o').clone().requires_grad_(g.is_tracker)
    d = g.t('rays_d').clone().requires_grad_(g.is_tracker)
    W = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in g.weights.items()}
    ef = g.t('exposure_feat').clone().requires_grad_(True) if g.has('exposure_feat') else None
    dyn = g.t('dynamic_r') if g.has('dynamic_r') else None
    depth, var, rgb, valid, aux = orc.render_rays(
        W, g.ocfg, o, d, g.t('gt_depth'), geo, col, g.t('cloud'), g.stage,
        is_tracker=g.is_tracker, dynamic_r=dyn, exposure_feat=ef, dtype=dtype)
    loss = (g.t('up_depth').to(dtype) * depth).sum() + (g.t('up_rgb').to(dtype) * rgb).sum()
    loss.backward()
    return dict(depth=depth, var=var, rgb=rgb, valid=valid, geo=geo, col=col, o=o, d=d, W=W, ef=ef)


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_forward_matches_reference(name):
    g = Golden(name)
    r = _run(g)
    assert torch.equal(r['valid'], g.t('valid'))
    torch.testing.assert_close(r['depth'].detach(), g.t('depth'), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(r['rgb'].detach(), g.t('rgb'), rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(r['var'].detach(), g.t('var'), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_gradients_match_reference(name):
    g = Golden(name)
    r = _run(g)
    # two fp32 formulations of the same math agree to ~1e-4
    assert rel_l2(r['geo'].grad, g.t('g_geo_feats')) < 2e-4
    if g.stage == 'color':
        assert rel_l2(r['col'].grad, g.t('g_col_feats')) < 3e-4
    if g.is_tracker:
        assert rel_l2(r['o'].grad, g.t('g_rays_o')) < 1e-3
        assert rel_l2(r['d'].grad, g.t('g_rays_d')) < 1e-3
    if g.has('g_exposure_feat'):
        assert rel_l2(r['ef'].grad, g.t('g_exposure_feat')) < 1e-4
    for k, gref in g.param_grads.items():
        got = r['W'][k].grad
        if got is None:
            assert gref.abs().max() == 0, k
            continue
        assert rel_l2(got, gref) < 5e-4, k


def test_fp64_truth_close_to_fp32():
    g = Golden('replica_color_mapper')
    r32, r64 = _run(g, torch.float32), _run(g, torch.float64)
    assert rel_l2(r32['depth'], r64['depth']) < 1e-6
    assert rel_l2(r32['rgb'], r64['rgb']) < 1e-4
    assert rel_l2(r32['geo'].grad, r64['geo'].grad) < 5e-4


def test_exact_knn_properties():
    gen = torch.Generator().manual_seed(0)
    cloud = torch.rand(500, 3, generator=gen)
    cloud[10] = cloud[3]                      # duplicate point -> tie broken by lower id
    q = torch.rand(64, 3, generator=gen)
    q[0] = cloud[3]
    D, I = exact_knn(q, cloud, 8)
    full = squared_dist_f32(q, cloud)
    assert (D[:, 1:] >= D[:, :-1]).all()
    assert torch.equal(D, torch.gather(full, 1, I))
    srt = torch.sort(full, dim=1, stable=True)
    assert torch.equal(srt.values[:, :8], D)
    assert torch.equal(srt.indices[:, :8], I)
    assert I[0, 0] == 3 and I[0, 1] == 10
    D2, I2 = exact_knn(q, cloud[:5], 8)       # N < K -> padded
    assert (I2[:, 5:] == -1).all() and (D2[:, 5:] > 1e38).all()


def test_oracle_sample_near_pcl_restatement():
    """oracle.render.sample_near_pcl (neural_point.py:1734-1786): first-two-hits rule, invalid rays, float64 linspace."""
    import numpy as np
    import torch
    from oracle import render as orc
    # a wall of points at x = 1.0, rays along +x from the origin; near 0.2, far 2.2 -> coarse step 2.0 / 24
    ys, zs = torch.meshgrid(torch.linspace(-0.05, 0.05, 5), torch.linspace(-0.05, 0.05, 5), indexing='ij')
    cloud = torch.stack([torch.full_like(ys, 1.0), ys, zs], -1).reshape(-1, 3)
    o = torch.zeros(3, 3)
    d = torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [1.0, 0, 0]])
    z, invalid = orc.sample_near_pcl(o, d, 0.2, 2.2, 5, cloud, 0.08)
    assert invalid.tolist() == [False, True, False]
    sec = np.linspace(0.2, 2.2, 25)
    hits = [i for i, zc in enumerate(sec) if abs(zc - 1.0) < 0.08 - 1e-6]
    assert len(hits) >= 2
    expect = np.linspace(sec[hits[0]], sec[hits[1]], 5).astype(np.float32)
    assert np.array_equal(z[0].numpy(), expect) and np.array_equal(z[2].numpy(), expect)
    assert np.array_equal(z[1].numpy(), np.linspace(0.2, 2.2, 5).astype(np.float32))


def test_oracle_sampling_matches_reference_golden():
    """oracle/sampling.py vs the vectors the REAL reference produced (tests/golden/make_golden_sampling.py):
    get_samples with / without the depth filter and depth_limit, get_camera_from_tensor."""
    import os
    import numpy as np
    from oracle import sampling as osm
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'sampling.npz'))
    depth, color = torch.from_numpy(z['depth']), torch.from_numpy(z['color'])
    H, W, fx, fy, cx, cy = z['intr']
    c2w = osm.get_camera_from_tensor(torch.from_numpy(z['cam']))
    assert torch.equal(c2w, torch.from_numpy(z['c2w']))
    for k in range(3):
        H0, H1, W0, W1, n, filt, lim = z[f'case{k}']
        torch.manual_seed(100 + k)   # the picks come from the global generator, exactly as in the reference
        o, d, sd, sc, i, j = osm.get_samples(int(H0), int(H1), int(W0), int(W1), int(n), int(H), int(W), fx, fy, cx, cy,
                                             c2w, depth, color, depth_filter=bool(filt), depth_limit=None if lim < 0 else lim)
        for got, name in ((o, 'o'), (d, 'd'), (sd, 'sd'), (sc, 'sc'), (i, 'i'), (j, 'j')):
            assert torch.equal(got, torch.from_numpy(z[f'{name}{k}'])), (k, name)


@pytest.mark.parametrize('strategy', ['rot_trans', 'fixed'])
def test_oracle_point_store_matches_reference_golden(strategy):
    """oracle/point_store.py vs the REAL reference methods (add_neural_points, check_index, init_segment, update_fragments,
    get_*(end=True) of /root/reference/src/neural_point.py) run on a 20-frame stream with an exhaustive-search index
    (tests/golden/make_golden_point_store.py): same kept samples, segment boundaries, inherited-point masks, positions,
    features (same generator draws) and merged end-of-run tables."""
    import numpy as np
    import torch
    from oracle.point_store import PointStoreOracle
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'point_store.npz'))
    H, W, fx, fy, cx, cy = G['intr']
    rel_trans, rot_cos, fixed = G[f'{strategy}_cfg']
    st = PointStoreOracle(int(H), int(W), fx, fy, cx, cy, segment_strategy=strategy, fixed_segment_size=int(fixed),
                          segment_rel_trans=rel_trans, segment_rot_cos=rot_cos, c_dim=4)
    torch.manual_seed(77)
    feat = lambda p, which: torch.zeros([p.shape[0], 4]).normal_(mean=0, std=0.1)     # neural_point.py:1614-1617, geo then col
    kept = []
    for fid in G[f'{strategy}_fids']:
        t = lambda k: torch.from_numpy(G[f'{strategy}_{k}{fid}'])
        kept.append(st.add_neural_points(t('o'), t('d'), t('g'), int(fid), t('c2w'), feat_fn=feat))
    assert kept == G[f'{strategy}_kept'].tolist()
    keys = list(st.fragments_dict.keys())
    assert len(keys) == int(G[f'{strategy}_nseg']) >= 5
    for k, key in enumerate(keys):
        f = st.fragments_dict[key]
        assert int(key.split('_')[-1]) == int(G[f'{strategy}_seg{k}_name'])
        assert [f['start_idx'], f['idx_start_segment_features']] == G[f'{strategy}_seg{k}_start'].tolist()
        np.testing.assert_array_equal(np.array(f['npc'], dtype=np.float64), G[f'{strategy}_seg{k}_npc'])
        np.testing.assert_array_equal(f['geo_feats'].numpy(), G[f'{strategy}_seg{k}_geo'])
        np.testing.assert_array_equal(f['col_feats'].numpy(), G[f'{strategy}_seg{k}_col'])
        m = G[f'{strategy}_seg{k}_mask']
        if k == len(keys) - 1:
            assert f['mask'] is None and m.size == 0
        else:
            np.testing.assert_array_equal(np.asarray(f['mask']), m)
    np.testing.assert_allclose(st.merged('npc'), G[f'{strategy}_end_pos'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(st.merged('geo_feats'), G[f'{strategy}_end_geo'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(st.merged('col_feats'), G[f'{strategy}_end_col'], rtol=1e-6, atol=1e-7)


def test_oracle_frustum_matches_reference_golden():
    """oracle/frustum.py and the package's host path vs the REAL Mapper.get_mask_from_c2w / filter_point_before_add
    (/root/reference/src/Mapper.py:137-217, numpy + cv2.remap; tests/golden/make_golden_frustum.py)."""
    import numpy as np
    import torch
    from oracle.frustum import get_mask_from_c2w as oracle_mask
    from loopy_slam_b200.frustum import get_mask_from_c2w, filter_point_before_add
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'frustum.npz'))
    H, W, fx, fy, cx, cy = G['intr']
    H, W = int(H), int(W)
    cloud = G['cloud']
    for k, (fid, edge) in enumerate(G['cases']):
        ref = G[f'idx{k}']
        assert ref.size > 300
        got = oracle_mask(cloud, G[f'c2w{k}'], G[f'depth{k}'], H, W, fx, fy, cx, cy, int(edge))
        np.testing.assert_array_equal(np.asarray(got), ref)                      # the restatement: bit-exact row ids
        host = get_mask_from_c2w(torch.from_numpy(cloud), torch.from_numpy(G[f'c2w{k}']), torch.from_numpy(G[f'depth{k}']), H, W,
                                 fx, fy, cx, cy, edge=int(edge))                 # torch path for CPU tensors
        a, b = set(host.tolist()), set(ref.tolist())
        assert len(a ^ b) <= 0.002 * cloud.shape[0], (int(fid), len(a), len(b), len(a ^ b))
    m = filter_point_before_add(torch.from_numpy(G['f_o']), torch.from_numpy(G['f_d']), torch.from_numpy(G['f_g']),
                                torch.from_numpy(G['f_prev']), H, W, fx, fy, cx, cy)
    assert torch.equal(m.cpu(), torch.from_numpy(G['f_mask']))


def test_oracle_sample_near_pcl_matches_reference_golden():
    """oracle.render.sample_near_pcl vs the REAL NeuralPointCloud.sample_near_pcl (neural_point.py:1734-1786) run on the active
    index of the golden stream (exhaustive search): sample depths bit for bit, same invalid rays."""
    import numpy as np
    import torch
    from oracle import render as orc
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'point_store.npz'))
    z, invalid = orc.sample_near_pcl(torch.from_numpy(G['snp_o']), torch.from_numpy(G['snp_d']), 0.3, float(G['snp_far']), 5,
                                     torch.from_numpy(G['snp_cloud']), 0.08)
    assert 0 < int(G['snp_invalid'].sum()) < G['snp_invalid'].size
    np.testing.assert_array_equal(invalid.numpy(), G['snp_invalid'])
    np.testing.assert_array_equal(z.numpy(), G['snp_z'])
