"""CPU: host-side logic that needs no GPU -- weight-blob aliasing, config mirror, stream generator,
frustum selection, decoder/state_dict compatibility with the reference layout."""
import numpy as np
import pytest
import torch

import loopy_slam_b200 as L
from loopy_slam_b200.frustum import get_mask_from_c2w
from loopy_slam_b200.stream import SyntheticRoom, build_point_cloud, sample_batch
from helpers import Golden


def test_state_dict_matches_reference_layout():
    cfg = L.default_cfg('replica')
    m = L.get_model(cfg)
    g = Golden('replica_color_mapper')
    ref = {k: v.shape for k, v in g.weights.items() if k != 'color_decoder.embedder._B'}
    ours = {k: v.shape for k, v in m.state_dict().items()}
    assert list(ours.keys()) == list(ref.keys())          # same names, same ORDER (Appendix B offsets)
    assert ours == ref
    assert sum(p.numel() for p in m.parameters()) == 127447
    assert m.color_decoder.embedder._B.shape == (3, 20) and not isinstance(m.color_decoder.embedder._B, torch.nn.Parameter)
    cfg2 = L.default_cfg('scannet')
    m2 = L.get_model(cfg2)
    assert 'color_decoder.mlp_exposure.linear2.weight' in m2.state_dict()


def test_weight_blob_aliases_parameters_and_survives_updates():
    cfg = L.default_cfg('replica')
    m = L.get_model(cfg)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat, W = m.blob.ensure('cpu')
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k                 # values preserved
    for t, off in zip(m.blob.tensors(), m.blob.offsets):
        assert off % 4 == 0 and t.data_ptr() == flat.data_ptr() + 4 * off
    assert W.n_elems == m.blob.n_elems and W.c_lin_w[3] == m.blob.offsets[[e[0] + str(e[1]) for e in m.blob.entries].index('c_lin_w3')]
    # an in-place optimiser step is visible through the blob without re-packing
    opt = torch.optim.SGD([m.color_decoder.pts_linears[1].weight], lr=1.0)
    m.color_decoder.pts_linears[1].weight.grad = torch.ones_like(m.color_decoder.pts_linears[1].weight)
    opt.step()
    off = W.c_lin_w[1]
    assert torch.equal(flat[off:off + 128 * 128].view(128, 128), m.color_decoder.pts_linears[1].weight.detach())
    flat2, _ = m.blob.ensure('cpu')
    assert flat2.data_ptr() == flat.data_ptr()              # no rebuild when aliasing is intact
    # load_state_dict copies in place -> aliasing intact; replacing .data breaks it -> rebuilt
    m.load_state_dict(before)
    assert m.blob.ensure('cpu')[0].data_ptr() == flat.data_ptr()
    m.geo_decoder.output_linear.weight.data = torch.zeros(1, 32)
    flat3, _ = m.blob.ensure('cpu')
    assert flat3.data_ptr() != flat.data_ptr()
    assert float(flat3[m.blob.struct.g_out_w:m.blob.struct.g_out_w + 32].abs().sum()) == 0


def test_config_mirror_inherit(tmp_path):
    base = tmp_path / 'base.yaml'
    mid = tmp_path / 'mid.yaml'
    leaf = tmp_path / 'leaf.yaml'
    base.write_text('a: 1\nsub:\n  x: 1\n  y: 2\n')
    mid.write_text('sub:\n  y: 3\n')
    leaf.write_text(f'inherit_from: {mid}\nsub:\n  z: 4\n')
    cfg = L.load_config(str(leaf), str(base))
    assert cfg['a'] == 1 and cfg['sub'] == {'x': 1, 'y': 3, 'z': 4}
    d = L.default_cfg('scannet')
    assert d['model']['encode_exposure'] and d['rendering']['near_end_surface'] == 0.96 and d['use_dynamic_radius']
    assert not L.default_cfg('replica')['use_dynamic_radius']


def test_synthetic_stream_geometry():
    room = SyntheticRoom(H=48, W=64, fx=40., fy=40., cx=31.5, cy=23.5, n_frames=100, half=(1.0, 0.8, 0.6))
    color, depth, c2w = room.frame(3)
    assert color.shape == (48, 64, 3) and depth.shape == (48, 64) and c2w.shape == (4, 4)
    assert 0 < float((depth == 0).float().mean()) < 0.05 and float(color.min()) >= 0 and float(color.max()) <= 1
    R = c2w[:3, :3].double()
    assert torch.allclose(R @ R.t(), torch.eye(3, dtype=torch.float64), atol=1e-6)
    o, d, g, c = sample_batch(room, [3], 200, seed=1)
    hit = o + d * g[:, None]
    on_wall = ((hit.abs() - torch.tensor([1.0, 0.8, 0.6])).abs() < 1e-4).any(1)
    assert on_wall.all()
    cloud, geo, col = build_point_cloud(room, 900, pixels_per_frame=500, frame_ids=[3], max_frames=10)
    assert cloud.shape[0] <= 900 and cloud.shape[0] % 3 == 0 and geo.shape == (cloud.shape[0], 32)
    assert abs(float(geo.std()) - 0.1) < 0.02


def test_frustum_selection_cpu():
    room = SyntheticRoom(H=48, W=64, fx=40., fy=40., cx=31.5, cy=23.5, n_frames=100, half=(1.0, 0.8, 0.6), hole_frac=0.0)
    color, depth, c2w = room.frame(0)
    o, d, g, _ = sample_batch(room, [0], 300, seed=2)
    surf = o + d * g[:, None]
    behind = o - d * g[:, None]                  # behind the camera
    pts = torch.cat([surf, behind, surf + 3.0 * d])   # visible | behind | far beyond the wall
    idx = get_mask_from_c2w(pts, c2w, depth, room.H, room.W, room.fx, room.fy, room.cx, room.cy, edge=0)
    sel = torch.zeros(pts.shape[0], dtype=torch.bool)
    sel[idx] = True
    assert sel[:surf.shape[0]].float().mean() > 0.9
    assert not sel[surf.shape[0]:2 * surf.shape[0]].any()
    assert sel[2 * surf.shape[0]:].float().mean() < 0.1


def test_forward_gemm_program_invariants():
    """The static tensor-core program of a forward tile (csrc/lsr_render_fwd.cu:build_program) must match what the
    epilogue warps of render_fwd_kernel expect: one operand hand-over per GEMM group, one completion per epilogue
    wait, chunks that fit a ring stage, packed weights that fit the scratch."""
    import ctypes
    import numpy as np
    import __graft_entry__ as entry
    entry.build()
    from loopy_slam_b200 import _lib
    L = _lib.lib()
    W = _lib.LsrWeights()
    # realistic offsets (state_dict order, every tensor 16-byte aligned); the blob itself is never dereferenced here
    W.blob = 16
    W.n_elems = 1 << 20
    off = 0
    sizes = {'g_fc_w': [1024] * 5, 'g_fc_b': [32] * 5, 'g_B': 279, 'g_lin_w': [2976, 1024, 1024, 4000, 1024], 'g_lin_b': [32] * 5,
             'g_out_w': 32, 'g_out_b': 1, 'c_fc_w': [4096] * 5, 'c_fc_b': [128] * 5, 'c_B': 60, 'c_Brel': 30,
             'c_nb1_w': 6656, 'c_nb1_b': 128, 'c_nb2_w': 4096, 'c_nb2_b': 32,
             'c_lin_w': [5120, 16384, 16384, 21504, 16384], 'c_lin_b': [128] * 5, 'c_out_w': 384, 'c_out_b': 3}
    for name, sz in sizes.items():
        if isinstance(sz, list):
            for i, n in enumerate(sz):
                getattr(W, name)[i] = off
                off += (n + 3) // 4 * 4
        else:
            setattr(W, name, off)
            off += (sz + 3) // 4 * 4
    out = (ctypes.c_int64 * 8)()
    L.lsr_debug_program_stats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    expect = {  # (stage, flags) -> (operand hand-overs, completions on barrier 0, on barrier 1)
        (0, 0): (5, 5, 0),              # geometry: 5 layers
        (1, 0): (5 + 6, 5 + 6, 0),      # + colour trunk (5 layers) + head
        (1, 1): (5 + 9 + 6, 5 + 4 + 1 + 6, 4),   # + rel-pos MLP: 8 neighbour passes on ping-pong accumulators + V2
    }
    for (stage, flags), (waits, c0, c1) in expect.items():
        assert L.lsr_debug_program_stats(ctypes.byref(W), stage, flags, out) == 0
        n_ops, n_jobs, packed, w_, a_, b_, maxb, cap = list(out)
        assert (w_, a_, b_) == (waits, c0, c1), (stage, flags, list(out))
        assert 0 < n_ops <= 96 and 0 < n_jobs <= 40
        assert maxb <= 32768 and maxb % 32 == 0          # a chunk (hi + lo) fits one ring stage, 16-byte halves
        assert packed <= cap
    assert L.lsr_debug_program_stats(None, 1, 1, out) == 1


def test_frustum_helpers_match_numpy_restatement():
    """frustum.filter_point_before_add / keyframe_overlap_percent vs a line-by-line numpy restatement of
    /root/reference/src/Mapper.py:137-163 and :252-274 (pure projection logic, runs on any device)."""
    import numpy as np
    import torch
    from loopy_slam_b200.frustum import filter_point_before_add, keyframe_overlap_percent
    rng = np.random.default_rng(3)
    H, W, fx, fy, cx, cy = 68, 120, 60.0, 60.0, 59.5, 33.5

    def pose(yaw, t):
        c, s = np.cos(yaw), np.sin(yaw)
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
        m[:3, 3] = t
        return m
    n = 400
    o = np.tile(np.array([[0.1, -0.05, 0.2]], dtype=np.float32), (n, 1))
    d = np.concatenate([rng.uniform(-1.2, 1.2, (n, 2)), -np.ones((n, 1))], 1).astype(np.float32)
    g = rng.uniform(0.5, 4.0, n).astype(np.float32)
    prev = pose(0.35, [0.3, 0.0, 0.1])

    def ref_filter(o, d, g, prev_c2w):        # Mapper.py:139-163
        points = (o[..., None, :] + d[..., None, :] * g[..., None, None]).reshape(-1, 3)
        w2c = np.linalg.inv(prev_c2w)
        homo = np.concatenate([points, np.ones_like(points[:, 0]).reshape(-1, 1)], axis=1).reshape(-1, 4, 1)
        cam = (w2c @ homo)[:, :3]
        K = np.array([[fx, .0, cx], [.0, fy, cy], [.0, .0, 1.0]]).reshape(3, 3)
        cam[:, 0] *= -1
        uv = K @ cam
        z = uv[:, -1:] + 1e-5
        uv = (uv[:, :2] / z).astype(np.float32)
        mask = (uv[:, 0] < W) * (uv[:, 0] > 0) * (uv[:, 1] < H) * (uv[:, 1] > 0)
        return ~mask.reshape(-1)
    ours = filter_point_before_add(torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(g), torch.from_numpy(prev),
                                   H, W, fx, fy, cx, cy)
    ref = ref_filter(o, d, g, prev)
    assert ours.dtype == torch.bool and 0 < ref.sum() < n
    assert np.array_equal(ours.numpy(), ref)

    verts = (o[:, None, :] + d[:, None, :] * np.linspace(0.8, 1.2, 4, dtype=np.float32)[None, :, None] * g[:, None, None]).reshape(-1, 3)
    kfs = [pose(0.0, [0, 0, 0]), pose(0.6, [0.5, 0, 0]), pose(3.0, [0, 0, 1.0]), prev]

    def ref_percent(vertices, c2w, edge=20):  # Mapper.py:253-272
        w2c = np.linalg.inv(c2w)
        homo = np.concatenate([vertices, np.ones_like(vertices[:, 0]).reshape(-1, 1)], axis=1).reshape(-1, 4, 1)
        cam = (w2c @ homo)[:, :3]
        K = np.array([[fx, .0, cx], [.0, fy, cy], [.0, .0, 1.0]]).reshape(3, 3)
        uv = K @ cam
        z = uv[:, -1:] + 1e-5
        uv = (uv[:, :2] / z).astype(np.float32)
        mask = (uv[:, 0] < W - edge) * (uv[:, 0] > edge) * (uv[:, 1] < H - edge) * (uv[:, 1] > edge)
        mask = mask & (z[:, :, 0] < 0)
        return mask.reshape(-1).sum() / uv.shape[0]
    pct = keyframe_overlap_percent(torch.from_numpy(verts), [torch.from_numpy(k) for k in kfs], H, W, fx, fy, cx, cy)
    refp = np.array([ref_percent(verts, k) for k in kfs])
    assert pct.shape == (4,) and refp.max() > 0.05 and refp.min() == 0.0
    assert np.allclose(pct.numpy(), refp, atol=2.0 / verts.shape[0])      # at most a borderline vertex or two


# ---------------------------------------------------------------------------------------- round-2 host fixes
def test_config_loader_inherit_chain(tmp_path):
    """load_config: inherit_from chain + default file, nested override semantics of the reference loader
    (/root/reference/src/config.py:10-57)."""
    import yaml
    from loopy_slam_b200.config import load_config
    (tmp_path / 'default.yaml').write_text(yaml.dump({'a': 1, 'm': {'x': 1, 'y': {'p': 1, 'q': 2}}, 'only_default': 7}))
    (tmp_path / 'base.yaml').write_text(yaml.dump({'a': 2, 'm': {'y': {'q': 3}}, 'b': {'k': 1}}))
    (tmp_path / 'scene.yaml').write_text(yaml.dump({'inherit_from': str(tmp_path / 'base.yaml'), 'm': {'x': 5}, 'b': 9}))
    cfg = load_config(str(tmp_path / 'scene.yaml'), str(tmp_path / 'default.yaml'))
    assert cfg['a'] == 2 and cfg['only_default'] == 7
    assert cfg['m'] == {'x': 5, 'y': {'p': 1, 'q': 3}}
    assert cfg['b'] == 9                      # a scalar replaces a mapping
    import os
    if os.path.isdir('/root/reference/configs'):   # build container only: identical to the real loader on the real files
        import glob
        from oracle import ref_import
        _, _, _, rc = ref_import.import_reference()
        cwd = os.getcwd()
        os.chdir('/root/reference')
        try:
            for y in glob.glob('configs/*/*.yaml'):
                assert rc.load_config(y, 'configs/point_slam.yaml') == load_config(y, 'configs/point_slam.yaml'), y
        finally:
            os.chdir(cwd)


def test_npc_accessors_take_the_reference_bool():
    """src/Mapper.py:490-493 calls get_geo_feats(self.end) / get_cloud_pos(self.end) with a BOOLEAN and
    :773-777 update_*_feats(..., end=self.end): `end` must never be used as a row count."""
    import loopy_slam_b200 as L
    cfg = L.default_cfg('replica')
    npc = L.NeuralPointCloud(cfg, device='cpu')
    npc.set_cloud(torch.arange(30, dtype=torch.float32).reshape(10, 3), torch.zeros(10, 32), torch.zeros(10, 32))
    for end in (False, True):
        assert npc.get_cloud_pos(end).shape == (10, 3)
        assert npc.get_geo_feats(end).shape == (10, 32) and npc.get_col_feats(end).shape == (10, 32)
    npc.update_geo_feats(torch.ones(3, 32), indices=[1, 4, 5], end=False)
    npc.update_col_feats(torch.ones(10, 32) * 2, end=False)
    assert npc.get_geo_feats(False)[[1, 4, 5]].eq(1).all() and npc.get_geo_feats(False)[0].eq(0).all()
    assert npc.get_col_feats(False).eq(2).all()


def _child_step(model, q):
    """runs in a spawned process: the blob must ADOPT the shared storage, and an optimiser step must land in it"""
    import torch
    flat, _ = model.blob.ensure('cpu')
    p = model.color_decoder.output_linear.bias
    aliased = p.data_ptr() >= flat.data_ptr() and p.data_ptr() < flat.data_ptr() + 4 * flat.numel()
    opt = torch.optim.SGD([p], lr=1.0)
    p.grad = torch.ones_like(p)
    opt.step()
    q.put((bool(aliased), bool(flat.is_shared())))


def test_shared_decoders_stay_shared_across_processes():
    """src/Point_SLAM.py shares `shared_decoders` between the tracker and mapper processes.  After
    share_memory() + spawn, WeightBlob.ensure() in a child must adopt the shared flat buffer (not re-allocate a private
    one), so that the mapper's optimiser steps reach the weights the tracker renders with."""
    import torch.multiprocessing as mp
    import loopy_slam_b200 as L
    cfg = L.default_cfg('replica')
    torch.manual_seed(0)
    model = L.get_model(cfg)
    model.share_memory()
    before = model.color_decoder.output_linear.bias.detach().clone()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    pr = ctx.Process(target=_child_step, args=(model, q))
    pr.start()
    aliased, shared = q.get(timeout=120)
    pr.join(60)
    assert aliased and shared
    after = model.color_decoder.output_linear.bias.detach()
    assert torch.allclose(after, before - 1.0), 'the child\'s update did not reach the shared weights'
    flat, _ = model.blob.ensure('cpu')      # and the parent's own blob still aliases the same memory
    assert model.color_decoder.output_linear.bias.data_ptr() >= flat.data_ptr()


def test_segment_store_merge_matches_reference_restatement():
    """The end-of-run merge of the segmented point store (inherited points averaged over the segments that carried them,
    /root/reference/src/neural_point.py:1252-1281,1435-1510) on hand-built segments vs oracle/point_store.py."""
    import numpy as np
    import loopy_slam_b200 as L
    from loopy_slam_b200.neural_point import _Segment
    from oracle.point_store import PointStoreOracle
    gen = torch.Generator().manual_seed(3)
    cfg = L.default_cfg('replica')
    npc = L.NeuralPointCloud(cfg, device='cpu')
    orc = PointStoreOracle(64, 64, 40., 40., 31.5, 31.5)
    orc.fragments_dict = {}
    prev_keep = None
    prev = None
    for si, (n_new, frac) in enumerate(((40, 0.5), (30, 0.4), (25, None))):
        inh_pos = prev['pos'][prev_keep] + 0.01 * (si) if prev is not None else torch.zeros(0, 3)   # inherited rows may have moved (PGO)
        inh_geo = prev['geo'][prev_keep] if prev is not None else torch.zeros(0, 32)
        inh_col = prev['col'][prev_keep] if prev is not None else torch.zeros(0, 32)
        pos = torch.cat([inh_pos, torch.rand(n_new, 3, generator=gen)])
        geo = torch.cat([inh_geo, torch.randn(n_new, 32, generator=gen)])
        col = torch.cat([inh_col, torch.randn(n_new, 32, generator=gen)])
        seg = _Segment('cpu', 32, torch.eye(4), si, inh_pos.shape[0])
        seg.append(pos, geo, col)
        mask = (torch.rand(pos.shape[0], generator=gen) < frac) if frac is not None else None
        seg.mask = mask
        npc.fragments.append(seg)
        orc.fragments_dict[f'segment_{si}'] = {'npc': pos.tolist(), 'geo_feats': geo, 'col_feats': col,
                                               'idx_start_segment_features': inh_pos.shape[0],
                                               'mask': None if mask is None else mask.numpy()}
        prev, prev_keep = {'pos': pos, 'geo': geo, 'col': col}, mask
    np.testing.assert_allclose(npc.get_cloud_pos(True).numpy(), orc.merged('npc'), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(npc.get_geo_feats(True).numpy(), orc.merged('geo_feats'), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(npc.get_col_feats(True).numpy(), orc.merged('col_feats'), rtol=1e-6, atol=1e-7)
    assert npc.get_cloud_pos(False).shape[0] == npc.fragments[-1].n
