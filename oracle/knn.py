"""Exact k-NN oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates the *contract* of ``NeuralPointCloud.find_neighbors_faiss``
(/root/reference/src/neural_point.py:1659-1708): for every query return the K=8 nearest
cloud points by squared L2 distance, ascending, as ``D (P,K) float32`` / ``I (P,K) int64``
plus ``neighbor_num (P,) int32 = #{D < r^2}`` (strict, :1698-1706).

The reference delegates the search itself to faiss-gpu==1.7.2 ``GpuIndexIVFFlat``
(env.yaml:95; nlist 400, nprobe 4 -- approximate, not vendored, not installable here).
The oracle is therefore *exact* brute force:

* ``D = (dx*dx + dy*dy) + dz*dz`` evaluated in float32 with individually rounded
  operations (no FMA contraction) -- the same formula and order the reference uses when
  it recomputes D in tracker mode (decoder.py:194-195);
* ties are broken by the lower row id (a (D, id) lexicographic order);
* fewer than K points: padded with ``I = -1`` and ``D = FLT_MAX``.
"""
import numpy as np
import torch

FLT_MAX = float(np.finfo(np.float32).max)


def squared_dist_f32(q, cloud):
    """(P,3),(N,3) float32 -> (P,N) float32, op order ((dx^2+dy^2)+dz^2), no FMA."""
    q = q.to(torch.float32)
    cloud = cloud.to(torch.float32)
    dx = cloud[None, :, 0] - q[:, None, 0]
    dy = cloud[None, :, 1] - q[:, None, 1]
    dz = cloud[None, :, 2] - q[:, None, 2]
    return (dx * dx + dy * dy) + dz * dz


def exact_knn(q, cloud, K=8, chunk=2048):
    """Brute-force exact K-NN with (D, id) lexicographic tie-break.

    Returns D (P,K) float32 ascending, I (P,K) int64."""
    q = q.detach().to(torch.float32).cpu().reshape(-1, 3)
    cloud = cloud.detach().to(torch.float32).cpu().reshape(-1, 3)
    P, N = q.shape[0], cloud.shape[0]
    D_out = torch.full((P, K), FLT_MAX, dtype=torch.float32)
    I_out = torch.full((P, K), -1, dtype=torch.int64)
    if N == 0 or P == 0:
        return D_out, I_out
    kk = min(K, N)
    ids = torch.arange(N, dtype=torch.int64)
    for s in range(0, P, chunk):
        d = squared_dist_f32(q[s:s + chunk], cloud)               # (p,N) >= 0
        # non-negative float32 bit patterns are monotone as integers: build (D,id) keys
        key = (d.view(torch.int32).to(torch.int64) << 32) | ids[None, :]
        top = torch.topk(key, kk, dim=1, largest=False, sorted=True).values
        I_out[s:s + chunk, :kk] = top & 0xFFFFFFFF
        D_out[s:s + chunk, :kk] = (top >> 32).to(torch.int32).view(torch.float32)
    return D_out, I_out


def radius_sq(radius, dynamic_radius=None):
    """r^2 with the reference's dtypes: a python-float radius is squared in double and
    compared against float32 D as a float32 scalar; a per-sample float64 tensor stays
    float64."""
    if dynamic_radius is not None:
        return dynamic_radius.reshape(-1, 1).to(torch.float64) ** 2
    return torch.tensor(float(radius) ** 2, dtype=torch.float32)


def neighbor_num(D, r2):
    """#{D < r^2}, strict (src/neural_point.py:1701-1706)."""
    if r2.dtype == torch.float64:
        return (D.to(torch.float64) < r2).sum(-1).to(torch.int32)
    return (D < r2).sum(-1).to(torch.int32)


class ExactKNNPointCloud:
    """Minimal stand-in for the reference NeuralPointCloud on the hot path: the three
    methods the decoders / renderer call (decoder.py:186-190, Renderer.py:153)."""

    def __init__(self, cloud_pos, radius_query=0.08, radius_add=0.04, radius_min=0.02,
                 radius_mesh=0.08, nn_num=8):
        self.cloud = cloud_pos.detach().to(torch.float32).cpu().reshape(-1, 3)
        self.radius_query = radius_query
        self.radius_add = radius_add
        self.radius_min = radius_min
        self.radius_mesh = radius_mesh
        self.nn_num = nn_num

    def get_radius_query(self):
        return self.radius_query

    def find_neighbors_faiss(self, pos, step='add', retrain=False, is_pts_grad=False,
                             dynamic_radius=None):
        assert step in ('add', 'query', 'mesh')
        D, I = exact_knn(pos, self.cloud, self.nn_num)
        if step == 'query':
            radius = self.radius_query
        elif step == 'add':
            radius = self.radius_min if is_pts_grad else self.radius_add
        else:
            radius = self.radius_mesh
        n = neighbor_num(D, radius_sq(radius, dynamic_radius))
        return D, I, n
