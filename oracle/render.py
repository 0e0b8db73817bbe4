"""Torch-CPU restatement of the Loopy-SLAM render hot path (TEST INFRASTRUCTURE).

Functional, dtype-generic (float32 = "the fp32 oracle", float64 = "fp64 truth" used to
judge gradients) re-statement of

* z-sampling + point generation      /root/reference/src/utils/Renderer.py:98-176
* neighbour weights / IDW features   /root/reference/src/conv_onet/models/decoder.py:180-231, 431-492
* Gaussian Fourier features          decoder.py:34-43
* geometry MLP                       decoder.py:233-288
* colour MLP (+ rel-pos neighbour MLP, exposure affine)  decoder.py:307-342, 494-546
* stage dispatch                     decoder.py:573-626
* point mask + alpha compositing     Renderer.py:184-201, /root/reference/src/common.py:382-422

Gradients come from torch autograd on this restatement.  The only deliberate deviation
from the reference: the N(0, 0.01) random fill for samples with < min_nn_num neighbours
(decoder.py:202-203, 228-229, 453-454, 489-490) is replaced by zeros --
those samples are forced to alpha = sigmoid(-10) by Renderer.py:184-186 anyway.

Weights are a flat dict keyed like ``NICER.state_dict()`` plus the non-persistent
``color_decoder.embedder._B`` (decoder.py:394-395).
"""
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from .knn import exact_knn, neighbor_num, radius_sq

TWO_PI = 2 * math.pi


@dataclass
class OracleCfg:
    N_surface: int = 5               # configs/point_slam.yaml:121
    near_end: float = 0.3            # :122
    near_end_surface: float = 0.98   # :123
    far_end_surface: float = 1.02    # :124
    sigmoid_coef: float = 0.1        # :125-126
    nn_num: int = 8                  # :136
    min_nn_num: int = 2              # :137
    radius_query: float = 0.08       # :142
    use_dynamic_radius: bool = False
    encode_rel_pos_in_col: bool = True
    encode_exposure: bool = False
    skip_zero_depth_pixel: bool = False

    @staticmethod
    def from_cfg(cfg):
        return OracleCfg(
            N_surface=cfg['rendering']['N_surface'], near_end=cfg['rendering']['near_end'],
            near_end_surface=cfg['rendering']['near_end_surface'],
            far_end_surface=cfg['rendering']['far_end_surface'],
            sigmoid_coef=cfg['rendering']['sigmoid_coef_mapper'],
            nn_num=cfg['pointcloud']['nn_num'], min_nn_num=cfg['pointcloud']['min_nn_num'],
            radius_query=cfg['pointcloud']['radius_query'],
            use_dynamic_radius=cfg['use_dynamic_radius'],
            encode_rel_pos_in_col=cfg['model']['encode_rel_pos_in_col'],
            encode_exposure=cfg['model']['encode_exposure'],
            skip_zero_depth_pixel=cfg['rendering']['skip_zero_depth_pixel'])


# --------------------------------------------------------------------------- sampling
def sample_z(gt_depth, ocfg, far_zero=None):
    """Renderer.py:136-163 (sample_near_pcl off).  Float32, reference op order:
    z = (near*g)*(1-t) + (far*g)*t,  t = linspace(0,1,S).  Zero-depth rays get
    linspace(near_end, far_zero, S)."""
    g = gt_depth.reshape(-1, 1).to(torch.float32)
    S = ocfg.N_surface
    t = torch.linspace(0.0, 1.0, steps=S, dtype=torch.float32)
    gs = g.repeat(1, S)
    z = ocfg.near_end_surface * gs * (1. - t) + ocfg.far_end_surface * gs * t
    zero = (g <= 0).squeeze(-1)
    if zero.any():
        assert far_zero is not None
        z = z.clone()
        z[zero] = torch.linspace(ocfg.near_end, float(far_zero), steps=S, dtype=torch.float32)
    return z


def sample_near_pcl(rays_o, rays_d, near, far, num, cloud_pos, radius_query, nn_num=8):
    """NeuralPointCloud.sample_near_pcl (src/neural_point.py:1734-1786) on the exact k-NN: 25 coarse steps in
    [near, far] per ray; a ray with >= 2 steps that have a neighbour within radius_query takes its `num` samples
    between the FIRST TWO such steps (np.linspace in float64, cast to float32), the others keep
    linspace(near, far, num) and are reported invalid.  -> (z (n,num) f32, invalid (n,) bool)"""
    o, d = rays_o.reshape(-1, 3).float(), rays_d.reshape(-1, 3).float()
    n_rays = d.shape[0]
    intervals = 25
    z_vals = torch.linspace(near, far, steps=intervals)                              # :1748
    pts = (o[..., None, :] + d[..., None, :] * z_vals[..., :, None]).reshape(-1, 3)  # :1749-1751
    z_sec = np.linspace(near, far, intervals)                                        # :1755
    z_total = np.tile(np.linspace(near, far, num), (n_rays, 1))                      # :1756-1757
    D, _ = exact_knn(pts, cloud_pos, nn_num)
    nn = neighbor_num(D, radius_sq(radius_query, None)).numpy().reshape(n_rays, -1)  # :1762-1772
    invalid = nn.astype(bool).sum(axis=-1) < 2                                       # :1773-1774
    if invalid.sum() < n_rays:
        r, c = np.where(nn[~invalid].astype(bool))
        idx = np.concatenate(([0], np.flatnonzero(r[1:] != r[:-1]) + 1, [r.size]))
        out = [c[idx[i]:idx[i + 1]] for i in range(len(idx) - 1)]
        z_total[~invalid] = np.asarray([np.linspace(z_sec[it[0]], z_sec[it[1]], num=num) for it in out])   # :1781-1783
    return torch.from_numpy(z_total).float(), torch.from_numpy(invalid)


def far_for_zero_depth(gt_depth):
    """Renderer.py:102-121: far = clamp(min(5*mean, max(1.2*g)), 0, max(1.2*g))."""
    g = gt_depth.reshape(-1).to(torch.float32)
    far_bb = torch.minimum(5 * g.mean(), torch.max(g * 1.2))
    if torch.max(g) > 0:
        return torch.clamp(far_bb, 0, torch.max(g * 1.2))
    return far_bb


# ------------------------------------------------------------------------ decoder bits
def fourier(x, B, concat):
    """decoder.py:34-43.  x (P,3)."""
    y = (TWO_PI * x) @ B
    if concat:
        return torch.cat((torch.sin(y), torch.cos(y)), dim=-1)
    return torch.sin(y)


def softplus100(x):
    return F.softplus(x, beta=100)


def idw_weights(D, r2):
    """decoder.py:206-220: w = 1/(D+1e-10); w[D > r^2] = 0 (no grad); L1-normalise."""
    w = 1.0 / (D + 1e-10)
    with torch.no_grad():
        if r2.dtype == torch.float64 and D.dtype != torch.float64:
            out = D.to(torch.float64) > r2
        else:
            out = D > r2.to(D.dtype)
    w = torch.where(out, torch.zeros_like(w), w)
    return F.normalize(w, p=1, dim=1)


def _linear(W, prefix, x):
    return x @ W[prefix + '.weight'].t() + W[prefix + '.bias']


def geo_mlp(W, p, c):
    """decoder.py:263-288 (ReLU although self.actvn is Softplus, :277)."""
    pre = 'geo_decoder.'
    emb = fourier(p, W[pre + 'embedder._B'], concat=False)
    h = emb
    for i in range(5):
        h = torch.relu(_linear(W, f'{pre}pts_linears.{i}', h))
        h = h + _linear(W, f'{pre}fc_c.{i}', c)
        if i == 2:
            h = torch.cat([emb, h], -1)
    return _linear(W, pre + 'output_linear', h).squeeze(-1)


def col_trunk(W, p, c):
    """decoder.py:513-533 -- returns the pre-activation (P,3)."""
    pre = 'color_decoder.'
    emb = fourier(p, W[pre + 'embedder._B'], concat=True)
    h = emb
    for i in range(5):
        h = softplus100(_linear(W, f'{pre}pts_linears.{i}', h))
        h = h + _linear(W, f'{pre}fc_c.{i}', c)
        if i == 2:
            h = torch.cat([emb, h], -1)
    return _linear(W, pre + 'output_linear', h)


def exposure_affine(W, exposure_feat):
    """decoder.py:326-342,536-538: 12 numbers = [A (3x3 row-major) | t (3)]."""
    pre = 'color_decoder.mlp_exposure.'
    h = softplus100(_linear(W, pre + 'linear1', exposure_feat))
    return _linear(W, pre + 'linear2', h)


# --------------------------------------------------------------------------- the path
def decode_points(W, ocfg, p, D_search, I, n_search, geo_feats, col_feats, cloud_pos,
                  stage, is_tracker, r2, exposure_feat=None):
    """NICER.forward for stages 'geometry' / 'color' (decoder.py:573-610) given the
    neighbour search result.  p (P,3) in the compute dtype.  Returns
    raw (P,4) [r,g,b,occ], has_neighbors (P,) bool, aux dict."""
    dt = p.dtype
    valid = I >= 0
    Isafe = torch.where(valid, I, torch.zeros_like(I))
    pos_n = cloud_pos[Isafe]                                   # (P,K,3)
    if is_tracker or dt == torch.float64:
        # tracker: D recomputed differentiably from cloud_pos[I]-p (decoder.py:191-198);
        # fp64 truth: recompute so the weights are fp64-accurate.
        d = pos_n - p[:, None, :]
        if not is_tracker:
            d = d.detach()
        Dw = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        out = (D_search.to(torch.float64) > r2.to(torch.float64)) | ~valid
        Dw = torch.where(out, torch.full_like(Dw, 1e4), Dw)    # :198
    else:
        out = ~valid
        Dw = torch.where(out, torch.full_like(D_search, 1e4), D_search).to(dt)
    # radius test always on the fp32 search distances (has/neighbor_num come from the
    # search, decoder.py:186-204), so fp32 and fp64 runs share one neighbour set
    r2c = r2.to(torch.float64)
    outside = (D_search.to(torch.float64) > r2c) | ~valid
    w = 1.0 / (Dw + 1e-10)
    w = torch.where(outside, torch.zeros_like(w), w)
    w = F.normalize(w, p=1, dim=1)                             # (P,K)
    has = n_search > (ocfg.min_nn_num - 1)                     # :204

    aux = {'w': w, 'has': has}
    # ---- geometry
    cg = (w.unsqueeze(-1) * geo_feats[Isafe]).sum(1)
    cg = torch.where(has[:, None], cg, torch.zeros_like(cg))   # noise -> 0 (8c)
    occ = geo_mlp(W, p, cg)
    aux['cg'] = cg
    if stage == 'geometry':
        raw = torch.cat([torch.zeros(p.shape[0], 3, dtype=dt), occ[:, None]], -1)
        return raw, has, aux
    # ---- colour
    nf = col_feats[Isafe]                                      # (P,K,C)
    if ocfg.encode_rel_pos_in_col:                             # decoder.py:477-485
        pre = 'color_decoder.'
        rel = pos_n - p[:, None, :]
        emb = fourier(rel.reshape(-1, 3), W[pre + 'embedder_rel_pos._B'], concat=True)
        x = torch.cat([emb.reshape(p.shape[0], -1, emb.shape[-1]), nf], -1)
        x = softplus100(_linear(W, pre + 'mlp_col_neighbor.linear1', x))
        nf = _linear(W, pre + 'mlp_col_neighbor.linear2', x)
    cc = (w.unsqueeze(-1) * nf).sum(1)
    cc = torch.where(has[:, None], cc, torch.zeros_like(cc))
    aux['cc'] = cc
    out3 = col_trunk(W, p, cc)
    if ocfg.encode_exposure:
        if exposure_feat is not None:                          # decoder.py:535-540
            aff = exposure_affine(W, exposure_feat)
            out3 = torch.sigmoid(out3 @ aff[:9].reshape(3, 3) + aff[-3:])
        # else: pre-sigmoid output, exposure applied after compositing (:541-542)
    else:
        out3 = torch.sigmoid(out3)
    raw = torch.cat([out3, occ[:, None]], -1)
    return raw, has, aux


def composite(raw, z, coef):
    """common.py:402-422 on raw (R,S,4), z (R,S)."""
    rgb = raw[..., :3]
    alpha = torch.sigmoid(coef * raw[..., 3])
    ones = torch.ones(alpha.shape[0], 1, dtype=alpha.dtype)
    T = torch.cumprod(torch.cat([ones, 1. - alpha + 1e-10], -1), -1)[:, :-1]
    wts = alpha * T
    wsum = wts.sum(-1, keepdim=True) + 1e-10
    rgb_map = (wts[..., None] * rgb).sum(-2) / wsum
    depth = (wts * z).sum(-1) / wsum.squeeze(-1)
    tmp = z - depth.unsqueeze(-1)
    var = (wts * tmp * tmp).sum(1)
    return depth, var, rgb_map, wts


def render_rays(W, ocfg, rays_o, rays_d, gt_depth, geo_feats, col_feats, cloud_pos, stage,
                is_tracker=False, dynamic_r=None, exposure_feat=None, dtype=torch.float32,
                knn=None, z_zero=None):
    """Renderer.render_batch_ray (Renderer.py:71-201) for gt_depth given.  sample_near_pcl: pass
    ``z_zero`` = (R,S) sample depths whose zero-depth rows come from sample_near_pcl() (:150-158); those rays
    then keep their rendered depth (:197-198) -- the caller clears their valid bit where sample_near_pcl said so.

    Returns (depth, var, rgb, valid_mask, aux).  ``knn`` = optional precomputed
    (D, I, n) from a previous call (to share neighbour sets between fp32 / fp64 runs)."""
    S = ocfg.N_surface
    R = rays_o.shape[0]
    g32 = gt_depth.reshape(-1).to(torch.float32)
    far_zero = far_for_zero_depth(g32) if (g32 <= 0).any() else None
    z32 = sample_z(g32, ocfg, far_zero)                        # (R,S) fp32, exact op order
    if z_zero is not None:
        z32 = torch.where((g32 <= 0)[:, None], z_zero.to(torch.float32), z32)
    o32, d32 = rays_o.detach().to(torch.float32), rays_d.detach().to(torch.float32)
    p32 = (o32[:, None, :] + d32[:, None, :] * z32[:, :, None]).reshape(-1, 3)

    r_pts = None
    if ocfg.use_dynamic_radius:
        r_pts = dynamic_r.reshape(-1, 1).repeat_interleave(S, dim=0)   # :174-176
    r2 = radius_sq(ocfg.radius_query, r_pts)
    if knn is None:
        D, I = exact_knn(p32, cloud_pos, ocfg.nn_num)
        n = neighbor_num(D, r2)
        knn = (D, I, n)
    D, I, n = knn

    z = z32.to(dtype)
    p = (rays_o.to(dtype)[:, None, :] + rays_d.to(dtype)[:, None, :] * z[:, :, None]).reshape(-1, 3)
    Wd = {k: v.to(dtype) for k, v in W.items()}
    raw, has, aux = decode_points(Wd, ocfg, p, D, I, n, geo_feats.to(dtype), col_feats.to(dtype),
                                  cloud_pos.to(dtype), stage, is_tracker, r2,
                                  None if exposure_feat is None else exposure_feat.to(dtype))
    # Renderer.py:184-186: occupancy of neighbour-less samples := -100, written under
    # no_grad (value replaced, gradient still passes through as identity)
    occ = raw[:, 3]
    occ = occ + (torch.where(has, occ, torch.full_like(occ, -100.0)) - occ).detach()
    raw = torch.cat([raw[:, :3], occ[:, None]], -1).reshape(R, S, 4)
    depth, var, rgb, wts = composite(raw, z, ocfg.sigmoid_coef)
    valid = has.view(R, S).sum(1) >= int(S / 2 + 1)            # decoder.py:259-260
    nz = g32 > 0
    if z_zero is None:
        depth = torch.where(nz, depth, torch.zeros_like(depth))    # Renderer.py:197-198 (not with sample_near_pcl)
    if ocfg.skip_zero_depth_pixel:
        rgb = torch.where(nz[:, None], rgb, torch.zeros_like(rgb))
    aux.update({'knn': knn, 'z': z32, 'p': p32, 'raw': raw, 'weights': wts})
    return depth, var, rgb, valid, aux
