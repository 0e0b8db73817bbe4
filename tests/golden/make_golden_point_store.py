"""Mint the golden vectors of the segmented point store by running the REAL reference methods
(/root/reference/src/neural_point.py: add_neural_points :1557-1631, check_index :1283-1315, init_segment :1220-1250,
update_fragments :1138-1218, get_cloud_pos / get_geo_feats / get_col_feats with end=True :1252-1281,1435-1510) as unbound
functions on a stand-in object.  The two things of NeuralPointCloud.__init__ that cannot exist here are replaced: the
FAISS GPU index by an exhaustive-search index with the same interface (is_trained / train / add / search: exact 8-NN,
squared L2 ascending), and the ORB/DBoW3 hooks by no-ops.  Build-container only; writes tests/golden/point_store.npz.

    python tests/golden/make_golden_point_store.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'point_store.npz')


class ExactIndex:
    """faiss index interface, exhaustive search (what an IVF index returns when every list is probed)."""

    def __init__(self):
        self.is_trained = False
        self.x = torch.zeros(0, 3)
        self.nprobe = 0

    def train(self, x):
        self.is_trained = True

    def add(self, x):
        self.x = torch.cat([self.x, torch.as_tensor(x, dtype=torch.float32).reshape(-1, 3)], 0)

    @property
    def ntotal(self):
        return self.x.shape[0]

    def search(self, q, k):
        q = q.float()
        n = self.x.shape[0]
        D = torch.full((q.shape[0], k), 3.4028234663852886e38)
        I = torch.full((q.shape[0], k), -1, dtype=torch.int64)
        if n:
            d = ((q[:, None, :] - self.x[None, :, :]) ** 2).sum(-1)
            kk = min(k, n)
            dv, di = torch.topk(d, kk, dim=1, largest=False)
            D[:, :kk], I[:, :kk] = dv, di
        return D, I


def make_store(npm, strategy, H, W, fx, fy, cx, cy, rel_trans, rot_cos, fixed_size):
    s = types.SimpleNamespace()
    s.device = 'cpu'
    s.c_dim, s.nn_num, s.N_add = 4, 8, 3          # the bookkeeping does not depend on the feature width: 4 keeps the fixture small
    s.radius_add, s.radius_min, s.radius_query, s.radius_mesh = 0.04, 0.02, 0.08, 0.08
    s.near_end_surface, s.far_end_surface = 0.98, 1.02
    s.fix_interval_when_add_along_ray = False
    s.segment_strategy, s.fixed_segment_size = strategy, fixed_size
    s.segment_rel_trans, s.segment_rot_cos = rel_trans, rot_cos
    s.H, s.W, s.fx, s.fy, s.cx, s.cy = H, W, fx, fy, cx, cy
    s.fragments_dict, s.new_segment = None, False
    s._input_pos, s._input_rgb, s._pts_num = [], [], 0
    s.end_geo_feats = s.end_col_feats = None
    s.index = ExactIndex()
    s.resource, s.cuda_id, s.nlist, s.nprobe = None, 0, 400, 4
    s.extract_orb_features = lambda *a, **k: None
    s.add_orb_features = lambda *a, **k: None
    C = npm.NeuralPointCloud
    for name in ('check_index', 'init_segment', 'update_fragments', 'find_neighbors_faiss', 'get_cloud_pos', 'get_geo_feats',
                 'get_col_feats', 'add_neural_points', 'sample_near_pcl'):
        setattr(s, name, types.MethodType(getattr(C, name), s))
    return s


def main():
    ref_import.import_reference()
    import src.neural_point as npm
    npm.faiss.index_cpu_to_gpu = lambda *a, **k: ExactIndex()           # check_index :1292-1296 re-creates the index
    from loopy_slam_b200.stream import SyntheticRoom, sample_batch
    room = SyntheticRoom(H=64, W=64, fx=40., fy=40., cx=31.5, cy=31.5, n_frames=60, half=(0.9, 0.7, 0.5), hole_frac=0.02)
    out = dict(intr=np.array([room.H, room.W, room.fx, room.fy, room.cx, room.cy]))
    for strategy, rel_trans, rot_cos, fixed in (('rot_trans', 0.30, 0.90, 50), ('fixed', 0.3, 0.94, 9)):
        torch.manual_seed(77)
        s = make_store(npm, strategy, room.H, room.W, room.fx, room.fy, room.cx, room.cy, rel_trans, rot_cos, fixed)
        fids, kept = list(range(0, 60, 3)), []
        for fid in fids:
            o, d, g, c = sample_batch(room, [fid], 200, seed=100 + fid)
            c2w = room.frame(fid)[2]
            n = s.add_neural_points(o, d, g, c, idx=torch.tensor(fid), cur_c2w=c2w, gt_color=torch.zeros(2, 2, 3),
                                    gt_depth=torch.zeros(2, 2), gt_camera=c2w)
            kept.append(int(n))
            out[f'{strategy}_o{fid}'], out[f'{strategy}_d{fid}'] = o.numpy(), d.numpy()
            out[f'{strategy}_g{fid}'], out[f'{strategy}_c2w{fid}'] = g.numpy(), c2w.numpy()
        keys = list(s.fragments_dict.keys())
        out[f'{strategy}_cfg'] = np.array([rel_trans, rot_cos, fixed], dtype=np.float64)
        out[f'{strategy}_fids'], out[f'{strategy}_kept'] = np.array(fids), np.array(kept)
        out[f'{strategy}_nseg'] = np.array(len(keys))
        for k, key in enumerate(keys):
            f = s.fragments_dict[key]
            out[f'{strategy}_seg{k}_name'] = np.array(int(key.split('_')[-1]))
            out[f'{strategy}_seg{k}_npc'] = np.array(f['npc'], dtype=np.float64)
            # copies: on CPU tensors `.detach().cpu()` (:1442,1448) is not a copy, so the end=True merge below adds into the
            # fragments' own storage (on the GPU it works on copies); the merged result is the same either way
            out[f'{strategy}_seg{k}_geo'] = f['geo_feats'].detach().clone().numpy()
            out[f'{strategy}_seg{k}_col'] = f['col_feats'].detach().clone().numpy()
            out[f'{strategy}_seg{k}_start'] = np.array([f['start_idx'], f['idx_start_segment_features']])
            out[f'{strategy}_seg{k}_mask'] = np.zeros(0, bool) if f['mask'] is None else np.asarray(f['mask'])
        out[f'{strategy}_end_pos'] = np.array(s.get_cloud_pos(True), dtype=np.float64)
        _cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t                   # :1464 / :1500 move the merged tables to the GPU
        try:
            out[f'{strategy}_end_geo'] = s.get_geo_feats(True).numpy()
            out[f'{strategy}_end_col'] = s.get_col_feats(True).numpy()
        finally:
            torch.Tensor.cuda = _cuda
        print(strategy, 'segments', len(keys), 'kept', kept, 'end points', out[f'{strategy}_end_pos'].shape)
        if strategy == 'rot_trans':
            # sample_near_pcl :1734-1786 on the active index (what the renderer asks for the zero-depth rays of a batch)
            o, d, g, _ = sample_batch(room, [57, 30], 40, seed=9)
            far = torch.max(g * 1.2)
            z, invalid = s.sample_near_pcl(o, d, 0.3, far, 5)
            out.update(snp_o=o.numpy(), snp_d=d.numpy(), snp_far=far.numpy(), snp_z=z.numpy(), snp_invalid=invalid.numpy(),
                       snp_cloud=s.index.x.numpy())
            print('sample_near_pcl: invalid', int(invalid.sum()), 'of', invalid.numel())
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
