"""Host-side mirror of the point store + neighbour query half of ``NeuralPointCloud``
(/root/reference/src/neural_point.py): state :29-124, segment lifecycle (check_index :1283-1315,
init_segment :1220-1250, update_fragments :1138-1218), merged end-of-run views (get_cloud_pos :1252-1281,
get_geo_feats / get_col_feats :1435-1510, update_*_feats :1512-1546), train_index_global :1382-1392,
add_neural_points :1557-1631, find_neighbors_faiss :1659-1708, sample_near_pcl :1734-1786.

Differences in *mechanism*, not behaviour:
  * every segment ("fragment") keeps its points and both feature tables as DEVICE tensors with amortised-doubling
    capacity instead of Python lists of 3-float lists (neural_point.py:1144-1216) -- no list<->tensor round trips per
    mapped frame;
  * the neighbour index of the ACTIVE segment is the exact sm_100a uniform grid (lsr_grid_build / lsr_knn_query) instead
    of faiss GpuIndexIVFFlat (approximate: nprobe 4 of nlist 400).  It is rebuilt lazily: insertions only mark it stale,
    the next query (at most one per mapped frame in the reference's call pattern) rebuilds it;
  * the in-view test of init_segment is a device projection (same formula, float64).
The loop-closure half of the reference class (ORB/DBoW3, registration, PGO, TSDF) is out of scope;
`transform_segment` is the hook a pose-graph correction uses to move a segment's points (:144-232,1125-1131).
"""
import numpy as np
import torch

from .renderer import GridIndex


class _Segment:
    """One fragment: inherited points first (``n_inherited`` of them), then the points added while it was active."""

    def __init__(self, device, c_dim, keyframe, start_idx, n_inherited):
        self.device, self.c_dim = device, c_dim
        self.keyframe = keyframe            # (4,4) CPU tensor, the camera that opened the segment
        self.start_idx = int(start_idx)
        self.n_inherited = int(n_inherited)  # "idx_start_segment_features"
        self.index_pc = None                # 'fixed' strategy: start_idx // fixed_segment_size (the reference's segment name)
        self.mask = None                    # bool over this segment's points: inherited by the NEXT segment (set when that one opens)
        self.n = 0
        self.cap = 0
        self.pos = torch.zeros(0, 3, device=device)
        self.geo = torch.zeros(0, c_dim, device=device)
        self.col = torch.zeros(0, c_dim, device=device)

    def append(self, pts, geo, col):
        m = pts.shape[0]
        if self.n + m > self.cap:
            cap = max(2 * self.cap, self.n + m, 1024)
            for name, width in (('pos', 3), ('geo', self.c_dim), ('col', self.c_dim)):
                old = getattr(self, name)
                buf = torch.zeros(cap, width, device=self.device)
                buf[:self.n] = old[:self.n]
                setattr(self, name, buf)
            self.cap = cap
        self.pos[self.n:self.n + m] = pts
        self.geo[self.n:self.n + m] = geo
        self.col[self.n:self.n + m] = col
        self.n += m


class NeuralPointCloud(object):
    def __init__(self, cfg, slam=None, args=None, device=None):
        self.cfg = cfg
        self.c_dim = cfg['model']['c_dim']
        self.device = device or cfg['mapping']['device']
        self.use_dynamic_radius = cfg['use_dynamic_radius']
        pc = cfg['pointcloud']
        self.nn_num = pc['nn_num']
        self.radius_add = pc['radius_add']
        self.radius_min = pc['radius_min']
        self.radius_query = pc['radius_query']
        self.radius_mesh = pc['radius_mesh']
        self.radius_add_max = pc.get('radius_add_max', 0.08)
        self.radius_query_ratio = pc.get('radius_query_ratio', 2)
        self.fix_interval_when_add_along_ray = pc['fix_interval_when_add_along_ray']
        self.N_surface = cfg['rendering']['N_surface']
        self.N_add = pc['N_add']
        self.near_end_surface = pc['near_end_surface']
        self.far_end_surface = pc['far_end_surface']
        mp_ = cfg.get('mapping', {})
        self.fixed_segment_size = mp_.get('fixed_segment_size', 50)           # configs/point_slam.yaml:77-80
        self.segment_strategy = mp_.get('segment_strategy', 'rot_trans')
        self.segment_rel_trans = mp_.get('segment_rel_trans', 0.30)
        self.segment_rot_cos = mp_.get('segment_rot_cos', 0.94)
        if slam is not None:
            self.H, self.W, self.fx, self.fy, self.cx, self.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy
        else:
            cam = cfg['cam']
            ce = cam.get('crop_edge') or 0
            self.H, self.W, self.fx, self.fy = cam['H'] - 2 * ce, cam['W'] - 2 * ce, cam['fx'], cam['fy']
            self.cx, self.cy = cam['cx'] - ce, cam['cy'] - ce
        if self.nn_num != 8:
            raise NotImplementedError('lsr kernels are specialised for pointcloud.nn_num == 8')
        self.fragments = []                 # list of _Segment (the reference's fragments_dict, in insertion order)
        self.new_segment = False
        self.end_geo_feats = None
        self.end_col_feats = None
        self._pts_num = 0
        self._input_pos = torch.zeros(0, 3, device=self.device)
        self._input_rgb = torch.zeros(0, 3, device=self.device)
        self._grid = None
        self._grid_n = -1

    # ------------------------------------------------------------------ small accessors (neural_point.py:1341-1433)
    @property
    def _active(self):
        return self.fragments[-1] if self.fragments else None

    def cloud_pos_tensor(self):
        s = self._active
        return s.pos[:s.n] if s is not None else torch.zeros(0, 3, device=self.device)

    def input_pos(self):
        return self._input_pos

    def input_rgb(self):
        return self._input_rgb

    def pts_num(self):
        return self._pts_num

    def index_ntotal(self):
        s = self._active
        return s.n if s is not None else 0

    def get_device(self):
        return self.device

    def get_c_dim(self):
        return self.c_dim

    def get_radius_query(self):
        return self.radius_query

    def get_radius_add(self):
        return self.radius_add

    # ------------------------------------------------------------------ merged / active views
    def _merged(self, key):
        """The end-of-run cloud (neural_point.py:1252-1281 for positions, :1435-1510 for features): points inherited
        from segment to segment are AVERAGED over all the segments that carried them and emitted once, by the last
        segment that holds them."""
        out = []
        old = None                           # (running sums, counts) of the previous segment
        mask_old = None
        last = len(self.fragments) - 1
        for si, seg in enumerate(self.fragments):
            x = getattr(seg, key)[:seg.n].detach().clone()
            cnt = torch.ones(seg.n, device=x.device)
            if old is not None and seg.n_inherited > 0:
                x[:seg.n_inherited] += old[0][mask_old]
                cnt[:seg.n_inherited] += old[1][mask_old]
            if si == last:
                out.append(x / cnt[:, None])
            else:
                keep = ~seg.mask
                out.append(x[keep] / cnt[keep][:, None])
                old, mask_old = (x, cnt), seg.mask
        return torch.cat(out, 0) if out else torch.zeros(0, 3 if key == 'pos' else self.c_dim, device=self.device)

    def get_cloud_pos(self, end=False):
        """Positions of the active segment, or (end=True, the reference's BOOLEAN, src/Mapper.py:492) of the merged
        end-of-run cloud.  The reference returns a Python list that callers immediately wrap in torch.tensor(...)
        (src/Mapper.py:492-493, src/Tracker.py:209-210); a device tensor works with both."""
        if end:
            return self._merged('pos')
        return self.cloud_pos_tensor()

    def get_geo_feats(self, end=False):
        if end:
            if self.end_geo_feats is None:
                self.end_geo_feats = self._merged('geo')
            return self.end_geo_feats
        s = self._active
        return s.geo[:s.n] if s is not None else torch.zeros(0, self.c_dim, device=self.device)

    def get_col_feats(self, end=False):
        if end:
            if self.end_col_feats is None:
                self.end_col_feats = self._merged('col')
            return self.end_col_feats
        s = self._active
        return s.col[:s.n] if s is not None else torch.zeros(0, self.c_dim, device=self.device)

    def _update(self, name, feats, indices, end):
        assert torch.is_tensor(feats), 'use tensor to update features'
        if end:                              # neural_point.py:1514-1519
            cur = getattr(self, 'end_' + name + '_feats')
            if indices is not None:
                cur[indices] = feats.clone().detach()
            else:
                assert feats.shape[0] == cur.shape[0], 'feature shape[0] mismatch'
                setattr(self, 'end_' + name + '_feats', feats.clone().detach())
            return
        s = self._active
        tab = getattr(s, name)
        if indices is not None:
            tab[:s.n][indices] = feats.clone().detach().to(tab.dtype)
        else:
            assert feats.shape[0] == s.n, 'feature shape[0] mismatch'
            tab[:s.n] = feats.clone().detach()

    def update_geo_feats(self, feats, indices=None, end=False):
        self._update('geo', feats, indices, end)

    def update_col_feats(self, feats, indices=None, end=False):
        self._update('col', feats, indices, end)

    # ------------------------------------------------------------------ index
    def _cell(self):
        return float(self.radius_add_max * self.radius_query_ratio) if self.use_dynamic_radius else float(self.radius_query)

    def grid_index(self):
        """Exact grid index over the ACTIVE segment (row id == row of get_cloud_pos() / get_*_feats()).  Rebuilt only
        when points were added since the last query."""
        s = self._active
        if s is None or s.n == 0:
            return None
        if self._grid is None or self._grid_n != s.n:
            self._grid = GridIndex(s.pos[:s.n], self._cell())
            self._grid_n = s.n
        return self._grid

    def train_index_global(self):
        """neural_point.py:1382-1392: one index over the merged end-of-run cloud.  The merged cloud becomes a single
        segment (what the reference's callers use from here on: get_*(end=True))."""
        pos, geo, col = self._merged('pos'), self._merged('geo'), self._merged('col')
        self.set_cloud(pos, geo, col)

    def set_cloud(self, pos, geo_feats=None, col_feats=None):
        """Replace the store by ONE segment holding `pos` (tests, benches, train_index_global)."""
        pos = pos.to(self.device).float().reshape(-1, 3).contiguous()
        n = pos.shape[0]
        seg = _Segment(self.device, self.c_dim, torch.eye(4), 0, 0)
        geo = geo_feats.to(self.device).float() if geo_feats is not None else torch.zeros(n, self.c_dim, device=self.device)
        col = col_feats.to(self.device).float() if col_feats is not None else torch.zeros(n, self.c_dim, device=self.device)
        seg.append(pos, geo, col)
        self.fragments = [seg]
        self._pts_num = n
        self._grid = None

    def transform_segment(self, k, T):
        """Rigidly move the points of segment k by the 4x4 transform T (what a pose-graph correction does to a
        fragment, neural_point.py:144-232) and invalidate the index if it is the active one."""
        seg = self.fragments[k]
        T = torch.as_tensor(T, dtype=torch.float32, device=self.device)
        seg.pos[:seg.n] = seg.pos[:seg.n] @ T[:3, :3].t() + T[:3, 3]
        if seg is self._active:
            self._grid = None

    def find_neighbors_faiss(self, pos, step='add', retrain=False, is_pts_grad=False, dynamic_radius=None):
        """Same contract as neural_point.py:1659-1708 -- D (P,8) f32 squared distances ascending,
        I (P,8) i64 row ids, neighbor_num (P,) i32 = #{D < r^2} -- but exact and limited to the query
        radius: entries beyond it (which the decoders weight by zero) are I = -1, D = FLT_MAX."""
        assert step in ['add', 'query', 'mesh']
        if step == 'query':
            radius = self.radius_query
        elif step == 'add':
            radius = self.radius_add if not is_pts_grad else self.radius_min
        else:
            radius = self.radius_mesh
        pos = pos.reshape(-1, 3)
        dyn = None
        if dynamic_radius is not None and dynamic_radius.numel() == pos.shape[0]:   # :1698-1704 (else: fixed)
            dyn = dynamic_radius
        grid = self.grid_index()
        if grid is None:
            P = pos.shape[0]
            return (torch.full((P, 8), torch.finfo(torch.float32).max, device=pos.device),
                    torch.full((P, 8), -1, dtype=torch.int64, device=pos.device),
                    torch.zeros(P, dtype=torch.int32, device=pos.device))
        return grid.query(pos, radius, dyn)

    # ------------------------------------------------------------------ segment lifecycle
    def _wants_new_segment(self, method, idx, cur_c2w):
        if not self.fragments:
            return False
        if method == 'fixed':                # :1285-1298
            i = int(idx.item() if torch.is_tensor(idx) else idx)
            pc = i // self.fixed_segment_size
            return i % self.fixed_segment_size == 0 and all(sg.index_pc != pc for sg in self.fragments)
        if method == 'rot_trans':            # :1299-1312, src/common.py:759-777
            kf = self.fragments[-1].keyframe
            c = torch.as_tensor(cur_c2w).detach().to('cpu', torch.float32)
            rel_trans = (c[:3, -1] - kf[:3, -1]).norm(2)
            cos = torch.dot(kf[:3, :3] @ torch.tensor([0., 0., 1.]), c[:3, :3] @ torch.tensor([0., 0., 1.]))
            return bool(rel_trans > self.segment_rel_trans or cos < self.segment_rot_cos)
        raise NotImplementedError

    def check_index(self, method, idx, cur_c2w=None):
        """neural_point.py:1283-1315: when a new segment starts, drop the index and seed the new one with the points of
        the last segment that are in view of the current camera (init_segment).  -> init dict or None."""
        if not self._wants_new_segment(method, idx, cur_c2w):
            return None
        return self.init_segment(cur_c2w)

    def init_segment(self, cur_c2w):
        """neural_point.py:1220-1250: points of the last segment whose projection lies inside the image (20 px border)."""
        last = self.fragments[-1]
        c2w = torch.as_tensor(cur_c2w).detach().to('cpu', torch.float32)
        w2c = torch.linalg.inv(c2w).double().to(self.device)
        pts = last.pos[:last.n].double()
        cam = pts @ w2c[:3, :3].t() + w2c[:3, 3]
        z = cam[:, 2] + 1e-5                                                      # uv = K @ cam, no axis flip here (:1233-1236)
        u = ((self.fx * cam[:, 0] + self.cx * cam[:, 2]) / z).float()
        v = ((self.fy * cam[:, 1] + self.cy * cam[:, 2]) / z).float()
        edge = 20
        mask = (u < self.W - edge) & (u > edge) & (v < self.H - edge) & (v > edge)
        init = {'npc': last.pos[:last.n][mask].clone(), 'geo_feats': last.geo[:last.n][mask].detach().clone(),
                'col_feats': last.col[:last.n][mask].detach().clone(), 'mask': mask}
        return init

    def _update_fragments(self, idx, cur_c2w, npc, geo_feats, col_feats, init):
        """neural_point.py:1138-1218 (the bookkeeping of points / features; the RGB-D keyframe copies and ORB features
        feed the loop-closure stack, which is out of scope)."""
        kf = torch.as_tensor(cur_c2w).detach().to('cpu', torch.float32).clone() if cur_c2w is not None else torch.eye(4)
        i = int(idx.item() if torch.is_tensor(idx) else (idx or 0))
        if not self.fragments:
            seg = _Segment(self.device, self.c_dim, kf, i, 0)
            seg.index_pc = i // self.fixed_segment_size
            seg.append(npc, geo_feats, col_feats)
            self.fragments.append(seg)
        elif init is not None:
            self.fragments[-1].mask = init['mask']
            seg = _Segment(self.device, self.c_dim, kf, i, init['npc'].shape[0])
            seg.index_pc = i // self.fixed_segment_size
            seg.append(init['npc'], init['geo_feats'], init['col_feats'])
            seg.append(npc, geo_feats, col_feats)
            self.fragments.append(seg)
            self.new_segment = True
            self._grid = None
        else:
            self.fragments[-1].append(npc, geo_feats, col_feats)

    # ------------------------------------------------------------------ insertion
    def add_neural_points(self, batch_rays_o, batch_rays_d, batch_gt_depth, batch_gt_color, train=False,
                          is_pts_grad=False, dynamic_radius=None, idx=None, gt_color=None, gt_depth=None,
                          cur_c2w=None, gt_camera=None):
        """neural_point.py:1557-1631: (maybe) open a new segment, keep sampled surface locations with no existing point
        of the ACTIVE index within the add-radius, insert N_add points per location along the ray, features ~ N(0, 0.1)."""
        if batch_rays_o.shape[0] == 0:
            return 0
        init = None
        if idx is not None and cur_c2w is not None:
            init = self.check_index(self.segment_strategy, idx, cur_c2w)
        if init is not None:                 # the index now holds exactly the inherited points (:1247-1248)
            probe = GridIndex(init['npc'], self._cell()) if init['npc'].shape[0] > 0 else None
        else:
            probe = self.grid_index()
        mask = batch_gt_depth > 0
        batch_gt_color = batch_gt_color * 255
        o, d, g, c = batch_rays_o[mask], batch_rays_d[mask], batch_gt_depth[mask], batch_gt_color[mask]
        if dynamic_radius is not None:
            dynamic_radius = dynamic_radius[mask]
        pts_gt = (o[..., None, :] + d[..., None, :] * g[..., None, None]).reshape(-1, 3)
        keep = torch.ones(pts_gt.shape[0], dtype=torch.bool, device=pts_gt.device)
        if probe is not None:
            radius = self.radius_add if not is_pts_grad else self.radius_min
            dyn = dynamic_radius if (dynamic_radius is not None and dynamic_radius.numel() == pts_gt.shape[0]) else None
            _, _, nn = probe.query(pts_gt, radius, dyn)
            keep = nn == 0
        self._input_pos = torch.cat([self._input_pos, pts_gt[keep]], 0)
        self._input_rgb = torch.cat([self._input_rgb, c[keep].float()], 0)
        gs = g.unsqueeze(-1).repeat(1, self.N_add)
        t = torch.linspace(0.0, 1.0, steps=self.N_add, device=g.device)
        if self.fix_interval_when_add_along_ray:
            z = gs + torch.linspace(-0.04, 0.04, steps=self.N_add, device=g.device).unsqueeze(0)
        else:
            z = self.near_end_surface * gs * (1. - t) + self.far_end_surface * gs * t
        pts = (o[..., None, :] + d[..., None, :] * z[..., :, None])[keep].reshape(-1, 3)
        self._pts_num += pts.shape[0]
        geo = torch.zeros([pts.shape[0], self.c_dim], device=pts.device).normal_(mean=0, std=0.1)
        col = torch.zeros([pts.shape[0], self.c_dim], device=pts.device).normal_(mean=0, std=0.1)
        self._update_fragments(idx, cur_c2w, pts.float(), geo, col, init)
        self.end_geo_feats = self.end_col_feats = None
        return torch.sum(keep)

    # ------------------------------------------------------------------ zero-depth ray sampling
    def sample_near_pcl(self, rays_o, rays_d, near, far, num):
        """neural_point.py:1734-1786: for rays without sensor depth, place the `num` samples between the
        first TWO of 25 coarse steps that have any neighbour (`item[0]`, `item[1]`, :1781-1782); rays with < 2 such
        steps are invalid and keep linspace(near, far, num)."""
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        n_rays = rays_d.shape[0]
        intervals = 25
        far_f = float(far)
        zc = torch.linspace(near, far_f, steps=intervals, device=rays_o.device)
        pts = (rays_o[..., None, :] + rays_d[..., None, :] * zc[..., :, None]).reshape(-1, 3)
        _, _, nn = self.find_neighbors_faiss(pts, step='query')
        hit = nn.reshape(n_rays, intervals) > 0
        invalid = hit.sum(-1) < 2
        z_sec = torch.from_numpy(np.linspace(near, far_f, intervals)).to(rays_o.device)
        first = torch.argmax(hit.int(), dim=1)
        second = torch.argmax((hit & (torch.arange(intervals, device=hit.device)[None, :] > first[:, None])).int(), dim=1)
        # np.linspace(a, b, num) in float64: a + arange(num) * ((b - a) / (num - 1)), last element = b exactly
        za, zb = z_sec[first], z_sec[second]
        steps = torch.arange(num, device=rays_o.device, dtype=torch.float64)
        z_valid = za[:, None] + steps[None, :] * ((zb - za) / max(num - 1, 1))[:, None]
        z_valid[:, -1] = zb
        z_def = torch.from_numpy(np.linspace(near, far_f, num)).to(rays_o.device)[None, :].repeat(n_rays, 1)
        z = torch.where(invalid[:, None], z_def, z_valid)
        return z.float(), invalid
