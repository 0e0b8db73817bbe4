"""Synthetic Replica-room0-shaped RGB-D stream + neural point cloud (host-side data
generation for benches and tests).  Replaces the dataset readers
(/root/reference/src/utils/datasets.py:87-121) with an analytic scene so no files are needed.

Scene: axis-aligned box room 6 x 4 x 2.8 m seen from inside; depth = analytic ray/box
intersection measured along the camera -z axis (same convention as the reference's
un-normalised ray directions, /root/reference/src/common.py:104-120); 1 % of pixels are
holes (depth 0); colour = smooth procedural texture in [0,1]; poses on a closed loop in the
reference's OpenGL-style c2w convention (datasets.py:143-144).

Point cloud: the reference insertion rule (/root/reference/src/neural_point.py:1557-1631):
per mapped frame sample pixels, keep those with no existing point within ``radius_add``,
insert N_add=3 points at depth*{0.98,1.0,1.02}; features ~ N(0, 0.1).
"""
import math

import numpy as np
import torch


class SyntheticRoom:
    def __init__(self, H=680, W=1200, fx=600.0, fy=600.0, cx=599.5, cy=339.5, seed=1219,
                 n_frames=2000, half=(3.0, 2.0, 1.4), hole_frac=0.01):
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = H, W, fx, fy, cx, cy
        self.seed = seed
        self.n_frames = n_frames
        self.half = np.asarray(half, dtype=np.float64)
        self.hole_frac = hole_frac
        jj, ii = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64),
                             indexing='ij')
        self._dirs = np.stack([(ii - cx) / fx, -(jj - cy) / fy, -np.ones_like(ii)], -1)

    # ------------------------------------------------------------------ poses
    def c2w(self, idx):
        """Closed-loop trajectory: ~0.5 cm and ~0.3 deg per frame."""
        t = 2 * math.pi * (idx % self.n_frames) / self.n_frames
        hx, hy, hz = self.half
        pos = np.array([0.4 * hx * math.cos(t), 0.4 * hy * math.sin(t),
                        (0.15 / 1.4) * hz * math.sin(2 * t)])
        yaw = t * 1.0 + 0.5 * math.pi          # look roughly along the direction of travel
        pitch = 0.15 * math.sin(3 * t)
        # camera looks along -z; build R = Rz(yaw) * Rx(pi/2 + pitch)
        cy_, sy_ = math.cos(yaw), math.sin(yaw)
        a = 0.5 * math.pi + pitch
        ca, sa = math.cos(a), math.sin(a)
        Rz = np.array([[cy_, -sy_, 0], [sy_, cy_, 0], [0, 0, 1.0]])
        Rx = np.array([[1.0, 0, 0], [0, ca, -sa], [0, sa, ca]])
        M = np.eye(4)
        M[:3, :3] = Rz @ Rx
        M[:3, 3] = pos
        return M

    # ----------------------------------------------------------------- frames
    def _depth_color(self, M):
        R, o = M[:3, :3], M[:3, 3]
        d = self._dirs @ R.T                                   # (H,W,3) world dirs, |z_cam|=1
        with np.errstate(divide='ignore', invalid='ignore'):
            t_hi = (self.half - o) / d
            t_lo = (-self.half - o) / d
        t = np.where(d > 0, t_hi, t_lo)
        t = np.where(np.isfinite(t) & (t > 0), t, np.inf)
        depth = t.min(-1)                                      # z-depth (dirs have z=-1)
        axis = t.argmin(-1)
        hit = o + d * depth[..., None]
        u = hit / self.half                                    # [-1,1]^3
        col = np.stack([
            0.5 + 0.35 * np.sin(3.1 * u[..., 0] + 1.7 * u[..., 1]) + 0.1 * (axis == 0),
            0.5 + 0.35 * np.sin(2.3 * u[..., 1] - 2.9 * u[..., 2]) + 0.1 * (axis == 1),
            0.5 + 0.35 * np.cos(2.7 * u[..., 2] + 1.3 * u[..., 0]) + 0.1 * (axis == 2)], -1)
        return depth, np.clip(col, 0.0, 1.0)

    def frame(self, idx):
        """-> (color (H,W,3) f32 in [0,1], depth (H,W) f32 metres with holes, c2w (4,4) f32)."""
        M = self.c2w(idx)
        depth, col = self._depth_color(M)
        rng = np.random.default_rng(self.seed * 7919 + idx)
        holes = rng.random(depth.shape) < self.hole_frac
        depth = np.where(holes, 0.0, depth)
        return (torch.from_numpy(col.astype(np.float32)),
                torch.from_numpy(depth.astype(np.float32)),
                torch.from_numpy(M.astype(np.float32)))

    def rays(self, M, jj, ii):
        """Pixel (row jj, col ii) -> world ray (o, d) as the reference computes them
        (common.py:113-119), numpy float32."""
        R, o = M[:3, :3].astype(np.float32), M[:3, 3].astype(np.float32)
        dirs = np.stack([(ii.astype(np.float32) - np.float32(self.cx)) / np.float32(self.fx),
                         -(jj.astype(np.float32) - np.float32(self.cy)) / np.float32(self.fy),
                         -np.ones(len(ii), np.float32)], -1)
        d = (dirs[:, None, :] * R[None, :, :]).sum(-1)
        return np.broadcast_to(o, d.shape).copy(), d


def build_point_cloud(room, n_target, frame_stride=20, pixels_per_frame=7000, radius_add=0.04,
                      n_add=3, near=0.98, far=1.02, c_dim=32, seed=1219, max_frames=400,
                      frame_ids=None):
    """Reference insertion rule on the synthetic stream.  Returns
    (cloud_pos (N,3) f32, geo_feats (N,C) f32, col_feats (N,C) f32) CPU tensors, N <= n_target."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    pts_all = np.zeros((0, 3), np.float32)
    H, W = room.H, room.W
    for f in range(max_frames):
        if pts_all.shape[0] >= n_target:
            break
        if frame_ids is not None:
            idx = frame_ids[f % len(frame_ids)]
        else:
            idx = (f * frame_stride) % room.n_frames
        _, depth, M = room.frame(idx)
        M = M.numpy().astype(np.float64)
        depth = depth.numpy()
        pix = rng.integers(0, H * W, size=pixels_per_frame)
        jj, ii = pix // W, pix % W
        g = depth[jj, ii]
        keep = g > 0
        jj, ii, g = jj[keep], ii[keep], g[keep]
        o, d = room.rays(M, jj, ii)
        p_gt = o + d * g[:, None]
        if pts_all.shape[0]:
            tree = cKDTree(pts_all)
            nn = tree.query_ball_point(p_gt, r=radius_add, return_length=True)
            new = nn == 0
        else:
            new = np.ones(len(g), bool)
        o, d, g = o[new], d[new], g[new]
        tvals = np.linspace(0.0, 1.0, n_add, dtype=np.float32)
        z = near * g[:, None] * (1 - tvals) + far * g[:, None] * tvals
        pts = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3).astype(np.float32)
        pts_all = np.concatenate([pts_all, pts], 0)
    pts_all = pts_all[:n_target]
    gen = torch.Generator().manual_seed(seed)
    N = pts_all.shape[0]
    geo = torch.randn(N, c_dim, generator=gen) * 0.1
    col = torch.randn(N, c_dim, generator=gen) * 0.1
    return torch.from_numpy(pts_all), geo, col


def sample_batch(room, frame_ids, n_per_frame, seed, edge=0):
    """A mapper/tracker-shaped ray batch: n_per_frame uniform pixels (with replacement, inside an
    ``edge`` border) from each frame, depth>0 only.  Returns CPU float32 tensors
    (rays_o, rays_d, gt_depth, gt_color) -- the inputs of render_batch_ray."""
    rng = np.random.default_rng(seed)
    O, Dv, G, C = [], [], [], []
    for fid in frame_ids:
        color, depth, M = room.frame(fid)
        jj = rng.integers(edge, room.H - edge, size=n_per_frame)
        ii = rng.integers(edge, room.W - edge, size=n_per_frame)
        g = depth.numpy()[jj, ii]
        keep = g > 0
        jj, ii, g = jj[keep], ii[keep], g[keep]
        o, d = room.rays(M.numpy(), jj, ii)
        O.append(o), Dv.append(d), G.append(g), C.append(color.numpy()[jj, ii])
    cat = lambda xs: torch.from_numpy(np.concatenate(xs, 0).astype(np.float32))
    return cat(O), cat(Dv), cat(G), cat(C)
