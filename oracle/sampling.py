"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the reference's pixel/ray
sampling on the hot path, /root/reference/src/common.py:

    get_rays_from_uv   :104-120      dirs = ((i-cx)/fx, -(j-cy)/fy, -1);  rays_d = sum(dirs * c2w[:3,:3], -1);  rays_o = c2w[:3,-1]
    select_uv          :123-138      indices = torch.randint(n_window_pixels, (n,)); gathers i, j, depth, colour
    get_sample_uv      :160-172      window [H0,H1) x [W0,W1), i = column (W0..W1-1), j = row (H0..H1-1), row-major flattening
    get_samples        :237-259      rays of n window pixels (with replacement), optional depth > 0 (< depth_limit) filter
    get_camera_from_tensor / quad2rotation :301-343   [quat (w,x,y,z; unnormalised) | T] -> 3x4 c2w, two_s = 2 / (q.q)

Pinned against the real reference (imported in the build container) by tests/golden/make_golden_sampling.py ->
tests/golden/sampling.npz, checked in tests/test_oracle_golden.py.  The pixel picks come from the global torch
generator exactly as in the reference (same call, same order), so seeding torch reproduces the reference's pixels.
"""
import torch


def get_rays_from_uv(i, j, c2w, fx, fy, cx, cy):
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)      # :113-115
    dirs = dirs.reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)                                          # :117
    rays_o = c2w[:3, -1].expand(rays_d.shape)                                           # :119
    return rays_o, rays_d


def get_sample_uv(H0, H1, W0, W1, n, depth, color, indices=None):
    depth = depth[H0:H1, W0:W1]                                                         # :165-166
    color = color[H0:H1, W0:W1]
    i, j = torch.meshgrid(torch.linspace(W0, W1 - 1, W1 - W0), torch.linspace(H0, H1 - 1, H1 - H0), indexing='ij')
    i, j = i.t().reshape(-1), j.t().reshape(-1)                                         # :169-170, select_uv :128-129
    if indices is None:
        indices = torch.randint(i.shape[0], (n,))                                       # :130
    indices = indices.clamp(0, i.shape[0])                                              # :131
    return i[indices], j[indices], depth.reshape(-1)[indices], color.reshape(-1, 3)[indices]


def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, depth_filter=False, depth_limit=None,
                indices=None):
    """-> rays_o, rays_d, sample_depth, sample_color, i (int64), j (int64)   (the return_index=True form)."""
    i, j, sample_depth, sample_color = get_sample_uv(H0, H1, W0, W1, n, depth, color, indices)
    rays_o, rays_d = get_rays_from_uv(i, j, c2w, fx, fy, cx, cy)
    if depth_filter:                                                                    # :249-255
        mask = sample_depth > 0
        if depth_limit is not None:
            mask = mask & (sample_depth < depth_limit)
        rays_o, rays_d, sample_depth, sample_color = rays_o[mask], rays_d[mask], sample_depth[mask], sample_color[mask]
        i, j = i[mask], j[mask]
    return rays_o, rays_d, sample_depth, sample_color, i.to(torch.int64), j.to(torch.int64)


def quad2rotation(quad):
    two_s = 2.0 / (quad * quad).sum(-1)                                                 # :311
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    rot = torch.zeros(quad.shape[0], 3, 3, dtype=quad.dtype)
    rot[:, 0, 0] = 1 - two_s * (qj ** 2 + qk ** 2)
    rot[:, 0, 1] = two_s * (qi * qj - qk * qr)
    rot[:, 0, 2] = two_s * (qi * qk + qj * qr)
    rot[:, 1, 0] = two_s * (qi * qj + qk * qr)
    rot[:, 1, 1] = 1 - two_s * (qi ** 2 + qk ** 2)
    rot[:, 1, 2] = two_s * (qj * qk - qi * qr)
    rot[:, 2, 0] = two_s * (qi * qk - qj * qr)
    rot[:, 2, 1] = two_s * (qj * qk + qi * qr)
    rot[:, 2, 2] = 1 - two_s * (qi ** 2 + qj ** 2)
    return rot


def get_camera_from_tensor(inputs):
    N = len(inputs.shape)                                                               # :331-343
    if N == 1:
        inputs = inputs.unsqueeze(0)
    quad, T = inputs[:, :4], inputs[:, 4:]
    RT = torch.cat([quad2rotation(quad), T[:, :, None]], 2)
    return RT[0] if N == 1 else RT
