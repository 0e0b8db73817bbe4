"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the reference's segmented neural point
store, /root/reference/src/neural_point.py -- add_neural_points :1557-1631, check_index :1283-1315, init_segment
:1220-1250, update_fragments :1138-1218 (point / feature bookkeeping only), get_cloud_pos :1252-1281, get_geo_feats /
get_col_feats :1435-1510 -- with the FAISS index replaced by the exact k-NN of oracle/knn.py (the index holds the active
segment's points: inherited ones first, :1247-1248, then every insertion, :1627).  Python lists as in the reference.
PINNED: tests/golden/make_golden_point_store.py runs the REAL reference methods (as unbound functions on a stand-in object
with an exhaustive-search index) over a 20-frame stream for both segment strategies; tests/test_oracle_golden.py checks this
restatement against those vectors bit for bit (positions, masks, segment boundaries, features, merged tables).  Only the
approximate FAISS IVF search itself stays "parity unpinned" (third-party, absent)."""
import numpy as np
import torch

from .knn import exact_knn, neighbor_num, radius_sq


class PointStoreOracle:
    def __init__(self, H, W, fx, fy, cx, cy, radius_add=0.04, radius_min=0.02, N_add=3, near_end_surface=0.98,
                 far_end_surface=1.02, segment_strategy='rot_trans', fixed_segment_size=50, segment_rel_trans=0.30,
                 segment_rot_cos=0.94, c_dim=32):
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = H, W, fx, fy, cx, cy
        self.radius_add, self.radius_min, self.N_add = radius_add, radius_min, N_add
        self.near_end_surface, self.far_end_surface = near_end_surface, far_end_surface
        self.segment_strategy, self.fixed_segment_size = segment_strategy, fixed_segment_size
        self.segment_rel_trans, self.segment_rot_cos = segment_rel_trans, segment_rot_cos
        self.c_dim = c_dim
        self.fragments_dict = None
        self.index = np.zeros((0, 3), np.float32)          # what the FAISS index holds

    # :1283-1315
    def check_index(self, method, idx, cur_c2w):
        if self.fragments_dict is None:
            return None
        if method == 'fixed':
            index_pc, new = idx // self.fixed_segment_size, idx % self.fixed_segment_size
            if new == 0 and f'segment_{index_pc}' not in self.fragments_dict:
                self.index = np.zeros((0, 3), np.float32)
                return self.init_segment(cur_c2w)
            return None
        last = list(self.fragments_dict.keys())[-1]
        kf = self.fragments_dict[last]['keyframe']
        rel_trans = (cur_c2w[:3, -1] - kf[:3, -1]).norm(2)                                   # common.py:772-777
        ax = torch.tensor([0., 0., 1.])
        cos = torch.dot(kf[:3, :3] @ ax, cur_c2w[:3, :3] @ ax)                               # common.py:759-769
        if rel_trans > self.segment_rel_trans or cos < self.segment_rot_cos:
            self.index = np.zeros((0, 3), np.float32)
            return self.init_segment(cur_c2w)
        return None

    # :1220-1250
    def init_segment(self, cur_c2w):
        last = self.fragments_dict[list(self.fragments_dict.keys())[-1]]
        c2w = cur_c2w.cpu().numpy()
        w2c = np.linalg.inv(c2w)
        pts = np.asarray(last['npc'])
        ones = np.ones_like(pts[:, 0]).reshape(-1, 1)
        homo = np.concatenate([pts, ones], axis=1).reshape(-1, 4, 1)
        cam = (w2c @ homo)[:, :3]
        K = np.array([[self.fx, .0, self.cx], [.0, self.fy, self.cy], [.0, .0, 1.0]]).reshape(3, 3)
        uv = K @ cam
        z = uv[:, -1:] + 1e-5
        uv = (uv[:, :2] / z).astype(np.float32)
        edge = 20
        mask = ((uv[:, 0] < self.W - edge) * (uv[:, 0] > edge) * (uv[:, 1] < self.H - edge) * (uv[:, 1] > edge)).reshape(-1)
        pts = pts[mask]
        m = torch.from_numpy(mask)
        self.index = np.asarray(pts.tolist(), dtype=np.float32).reshape(-1, 3)              # index.train / add (:1247-1248)
        return {'npc': pts, 'geo_feats': last['geo_feats'].detach().clone()[m], 'col_feats': last['col_feats'].detach().clone()[m],
                'mask': mask}

    # :1138-1218 (bookkeeping of points and features)
    def update_fragments(self, method, idx, cur_c2w, npc, geo_feats, col_feats, init):
        if method == 'fixed':
            name = f'segment_{idx // self.fixed_segment_size}'
        if self.fragments_dict is None:
            name = name if method == 'fixed' else 'segment_0'
            self.fragments_dict = {name: {'keyframe': cur_c2w.detach().clone().cpu(), 'npc': npc.tolist(), 'geo_feats': geo_feats,
                                          'col_feats': col_feats, 'start_idx': idx, 'idx_start_segment_features': 0, 'mask': None}}
            return
        last = list(self.fragments_dict.keys())[-1]
        if init is not None:
            init_npc = init['npc'].tolist()
            self.fragments_dict[last]['mask'] = init['mask']
            name = name if method == 'fixed' else f'segment_{int(last.split("_")[-1]) + 1}'
            self.fragments_dict[name] = {'keyframe': cur_c2w.detach().clone().cpu(), 'npc': init_npc + npc.tolist(),
                                         'geo_feats': torch.cat([init['geo_feats'], geo_feats], 0),
                                         'col_feats': torch.cat([init['col_feats'], col_feats], 0), 'start_idx': idx,
                                         'idx_start_segment_features': len(init_npc), 'mask': None}
        else:
            f = self.fragments_dict[last]
            f['npc'] += npc.tolist()
            f['geo_feats'] = torch.cat([f['geo_feats'], geo_feats], 0)
            f['col_feats'] = torch.cat([f['col_feats'], col_feats], 0)

    # :1557-1631
    def add_neural_points(self, rays_o, rays_d, gt_depth, idx, cur_c2w, is_pts_grad=False, feat_fn=None):
        if rays_o.shape[0] == 0:
            return 0
        init = self.check_index(self.segment_strategy, idx, cur_c2w)
        mask = gt_depth > 0
        rays_o, rays_d, gt_depth = rays_o[mask], rays_d[mask], gt_depth[mask]
        pts_gt = (rays_o[..., None, :] + rays_d[..., None, :] * gt_depth[..., None, None]).reshape(-1, 3)
        keep = torch.ones(pts_gt.shape[0], dtype=torch.bool)
        if self.index.shape[0] > 0:
            D, _ = exact_knn(pts_gt, torch.from_numpy(self.index), 8)
            r2 = radius_sq(self.radius_add if not is_pts_grad else self.radius_min)
            keep = neighbor_num(D, r2) == 0
        gs = gt_depth.unsqueeze(-1).repeat(1, self.N_add)
        t = torch.linspace(0.0, 1.0, steps=self.N_add)
        z = self.near_end_surface * gs * (1. - t) + self.far_end_surface * gs * t
        pts = (rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None])[keep].reshape(-1, 3)
        geo = feat_fn(pts, 0) if feat_fn else torch.zeros(pts.shape[0], self.c_dim)
        col = feat_fn(pts, 1) if feat_fn else torch.zeros(pts.shape[0], self.c_dim)
        self.update_fragments(self.segment_strategy, idx, cur_c2w, pts, geo, col, init)
        self.index = np.concatenate([self.index, pts.numpy().astype(np.float32)], 0)           # index.add(pts) (:1627)
        return int(keep.sum())

    # :1252-1281 / :1435-1510
    def merged(self, key):
        keys = list(self.fragments_dict.keys())
        out = []
        get = lambda f: (np.array(self.fragments_dict[f]['npc'], dtype=np.float64) if key == 'npc'
                         else self.fragments_dict[f][key].detach().cpu().numpy().astype(np.float64))
        x_old = get(keys[0])
        mask_old = np.array([False] * len(x_old))
        counter_old = np.array([0] * len(x_old))
        for frag in keys[:-1]:
            x_s = get(frag).copy()
            counter_s = np.array([1] * len(x_s))
            mask_s = self.fragments_dict[frag]['mask']
            i0 = self.fragments_dict[frag]['idx_start_segment_features']
            counter_s[:i0] += counter_old[mask_old]
            x_s[:i0] += x_old[mask_old]
            out.append(x_s[~mask_s] / counter_s[..., np.newaxis][~mask_s])
            x_old, mask_old, counter_old = x_s, mask_s, counter_s
        x_last = get(keys[-1]).copy()
        counter_last = np.array([1] * len(x_last))
        i0 = self.fragments_dict[keys[-1]]['idx_start_segment_features']
        counter_last[:i0] += counter_old[mask_old]
        x_last[:i0] += x_old[mask_old]
        out.append(x_last / counter_last[..., np.newaxis])
        return np.concatenate(out, 0)
