"""Host-side mirror of the point store + neighbour query half of ``NeuralPointCloud``
(/root/reference/src/neural_point.py:29-124 state, :1659-1708 find_neighbors_faiss,
:1557-1631 add_neural_points, :1252-1281 / :1435-1546 accessors, :1734-1786 sample_near_pcl).

Differences in *mechanism*, not behaviour:
  * the cloud is ONE device tensor with amortised-doubling capacity instead of a Python list of
    3-float lists (neural_point.py:1144-1216) -- no list<->tensor round trips per mapped frame;
  * the neighbour index is the exact sm_100a uniform grid (lsr_grid_build / lsr_knn_query) instead of
    faiss GpuIndexIVFFlat (approximate: nprobe 4 of nlist 400); it is rebuilt after each insertion.
The loop-closure half of the reference class (ORB/DBoW3, registration, PGO, TSDF, fragments) is out
of scope and not mirrored.
"""
import numpy as np
import torch

from .renderer import GridIndex


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
        if self.nn_num != 8:
            raise NotImplementedError('lsr kernels are specialised for pointcloud.nn_num == 8')
        self._cap = 0
        self._n = 0
        self._pos = torch.zeros(0, 3, device=self.device)
        self.geo_feats = torch.zeros(0, self.c_dim, device=self.device)
        self.col_feats = torch.zeros(0, self.c_dim, device=self.device)
        self._input_pos = torch.zeros(0, 3, device=self.device)
        self._input_rgb = torch.zeros(0, 3, device=self.device)
        self._grid = None

    # ------------------------------------------------------------------ accessors
    def cloud_pos_tensor(self):
        return self._pos[:self._n]

    def get_cloud_pos(self, end=False):
        """Positions of the active cloud.  `end` is the reference's BOOLEAN (src/neural_point.py:1252: True = the merged
        end-of-run cloud of all segments); it is not a row count.  The reference returns a Python list that callers
        immediately wrap in torch.tensor(...) (src/Mapper.py:492-493, src/Tracker.py:209-210); a device tensor works
        with both."""
        return self._pos[:self._n]

    def input_pos(self):
        return self._input_pos

    def input_rgb(self):
        return self._input_rgb

    def pts_num(self):
        return self._n

    def index_ntotal(self):
        return self._n

    def get_device(self):
        return self.device

    def get_c_dim(self):
        return self.c_dim

    def get_radius_query(self):
        return self.radius_query

    def get_radius_add(self):
        return self.radius_add

    def get_geo_feats(self, end=False):
        return self.geo_feats[:self._n]

    def get_col_feats(self, end=False):
        return self.col_feats[:self._n]

    def update_geo_feats(self, feats, indices=None, end=False):
        assert torch.is_tensor(feats), 'use tensor to update features'
        if indices is not None:
            self.geo_feats[indices] = feats.clone().detach()
        else:
            assert feats.shape[0] == self._n, 'feature shape[0] mismatch'
            self.geo_feats[:self._n] = feats.clone().detach()

    def update_col_feats(self, feats, indices=None, end=False):
        assert torch.is_tensor(feats), 'use tensor to update features'
        if indices is not None:
            self.col_feats[indices] = feats.clone().detach()
        else:
            assert feats.shape[0] == self._n, 'feature shape[0] mismatch'
            self.col_feats[:self._n] = feats.clone().detach()

    # ------------------------------------------------------------------ index
    def _cell(self):
        return float(self.radius_add_max * self.radius_query_ratio) if self.use_dynamic_radius else float(self.radius_query)

    def grid_index(self):
        if self._grid is None and self._n > 0:
            self._grid = GridIndex(self._pos[:self._n], self._cell())
        return self._grid

    def set_cloud(self, pos, geo_feats=None, col_feats=None):
        """Replace the whole cloud (e.g. after a pose-graph correction moved the points,
        neural_point.py:1125-1131) and rebuild the index."""
        pos = pos.to(self.device).float().reshape(-1, 3).contiguous()
        self._pos, self._n, self._cap = pos, pos.shape[0], pos.shape[0]
        if geo_feats is not None:
            self.geo_feats = geo_feats.to(self.device).float().contiguous()
        if col_feats is not None:
            self.col_feats = col_feats.to(self.device).float().contiguous()
        self._grid = None

    def _append(self, pts, geo, col):
        n_new = pts.shape[0]
        if self._n + n_new > self._cap:
            cap = max(2 * self._cap, self._n + n_new, 1024)
            for name, width in (('_pos', 3), ('geo_feats', self.c_dim), ('col_feats', self.c_dim)):
                old = getattr(self, name)
                buf = torch.zeros(cap, width, device=self.device)
                buf[:self._n] = old[:self._n]
                setattr(self, name, buf)
            self._cap = cap
        self._pos[self._n:self._n + n_new] = pts
        self.geo_feats[self._n:self._n + n_new] = geo
        self.col_feats[self._n:self._n + n_new] = col
        self._n += n_new
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

    # ------------------------------------------------------------------ insertion
    def add_neural_points(self, batch_rays_o, batch_rays_d, batch_gt_depth, batch_gt_color, train=False,
                          is_pts_grad=False, dynamic_radius=None, idx=None, gt_color=None, gt_depth=None,
                          cur_c2w=None, gt_camera=None):
        """neural_point.py:1557-1631: keep sampled surface locations with no existing point within the
        add-radius, insert N_add points per location along the ray, features ~ N(0, 0.1)."""
        if batch_rays_o.shape[0] == 0:
            return 0
        mask = batch_gt_depth > 0
        batch_gt_color = batch_gt_color * 255
        o, d, g, c = batch_rays_o[mask], batch_rays_d[mask], batch_gt_depth[mask], batch_gt_color[mask]
        if dynamic_radius is not None:
            dynamic_radius = dynamic_radius[mask]
        pts_gt = (o[..., None, :] + d[..., None, :] * g[..., None, None]).reshape(-1, 3)
        keep = torch.ones(pts_gt.shape[0], dtype=torch.bool, device=pts_gt.device)
        if self._n > 0:
            _, _, nn = self.find_neighbors_faiss(pts_gt, step='add', is_pts_grad=is_pts_grad,
                                                 dynamic_radius=dynamic_radius)
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
        geo = torch.zeros([pts.shape[0], self.c_dim], device=pts.device).normal_(mean=0, std=0.1)
        col = torch.zeros([pts.shape[0], self.c_dim], device=pts.device).normal_(mean=0, std=0.1)
        self._append(pts.float(), geo, col)
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
