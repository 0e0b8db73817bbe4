"""Frustum feature selection on the device: which neural points may be
optimised for the current frame.  Restates Mapper.get_mask_from_c2w
(/root/reference/src/Mapper.py:165-217: CPU numpy projection of ALL points + cv2.remap per mapped
frame, returning a Python list) as a handful of device tensor ops returning an index tensor, plus the two
other projection helpers of the mapper (filter_point_before_add :137-163, keyframe_selection_overlap :219-282).
These are the callers either side of the render hot path, not the path itself: plain device tensor ops."""
import torch
import torch.nn.functional as F


def get_mask_from_c2w(cloud_pos, c2w, depth, H, W, fx, fy, cx, cy, edge=-4):
    """-> int64 indices of the cloud rows the mapper may optimise for this frame (src/Mapper.py:165-217): the points
    that project inside the (edge-cropped) image and lie in front of the camera no deeper than the bilinearly looked-up
    sensor depth + 0.5 m (zero lookups -> the maximum lookup).  CUDA tensors run the two-pass lsr_frustum_mask
    kernels (float64 projection like the reference); CPU tensors (host-logic tests) take the torch restatement."""
    if cloud_pos.is_cuda:
        import ctypes
        from . import _lib
        from ._lib import lib, check, ptr, stream_ptr
        dev = cloud_pos.device
        cloud = cloud_pos.detach()
        if cloud.dtype != torch.float32 or not cloud.is_contiguous():
            cloud = cloud.float().contiguous()
        c = torch.as_tensor(c2w).detach().to('cpu', torch.float32)
        if c.shape[0] == 3:
            c = torch.cat([c, torch.tensor([[0., 0., 0., 1.]])], 0)
        w2c = torch.linalg.inv(c).double()[:3].contiguous()        # np.linalg.inv(c2w) on the float32 pose, promoted (:180-181)
        arr = (ctypes.c_double * 12)(*w2c.reshape(-1).tolist())
        d = depth.detach().to(dev, torch.float32).contiguous()
        n = cloud.shape[0]
        nb = ctypes.c_size_t()
        check(lib().lsr_frustum_scratch_bytes(n, ctypes.byref(nb)), 'lsr_frustum_scratch_bytes')
        scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        with _lib.on_device(dev):
            check(lib().lsr_frustum_mask(ptr(cloud), n, arr, ptr(d), int(H), int(W), float(fx), float(fy), float(cx), float(cy),
                                         int(edge), ptr(scratch), ptr(mask), stream_ptr(dev)), 'lsr_frustum_mask')
        return torch.nonzero(mask, as_tuple=True)[0]
    return _get_mask_from_c2w_torch(cloud_pos, c2w, depth, H, W, fx, fy, cx, cy, edge)


def _get_mask_from_c2w_torch(cloud_pos, c2w, depth, H, W, fx, fy, cx, cy, edge=-4):
    """The same selection as plain tensor ops (any device)."""
    dev = cloud_pos.device
    c2w = c2w.to(device=dev, dtype=torch.float32)
    if c2w.shape[0] == 3:
        c2w = torch.cat([c2w, torch.tensor([[0., 0., 0., 1.]], device=dev)], 0)
    w2c = torch.linalg.inv(c2w)
    cam = cloud_pos.float() @ w2c[:3, :3].t() + w2c[:3, 3]
    z = cam[:, 2] + 1e-5                      # negative in front of the camera
    u = (fx * (-cam[:, 0]) + cx * cam[:, 2]) / z
    v = (fy * cam[:, 1] + cy * cam[:, 2]) / z
    # bilinear lookup of the sensor depth at (u, v), pixel centres at integer coordinates
    gx = (u / (W - 1)) * 2 - 1
    gy = (v / (H - 1)) * 2 - 1
    grid = torch.stack([gx, gy], -1).reshape(1, 1, -1, 2)
    d = F.grid_sample(depth.float().reshape(1, 1, H, W), grid, mode='bilinear', padding_mode='zeros',
                      align_corners=True).reshape(-1)
    d = torch.where(d == 0, d.max(), d)
    mask = (u < W - edge) & (u > edge) & (v < H - edge) & (v > edge) & (-z >= 0) & (-z <= d + 0.5)
    return torch.nonzero(mask, as_tuple=True)[0]


def _w2c(c2w, dev, dtype):
    c2w = torch.as_tensor(c2w).to(device=dev, dtype=dtype)
    if c2w.shape[0] == 3:
        c2w = torch.cat([c2w, torch.tensor([[0., 0., 0., 1.]], device=dev, dtype=dtype)], 0)
    return torch.linalg.inv(c2w)


def filter_point_before_add(rays_o, rays_d, gt_depth, prev_c2w, H, W, fx, fy, cx, cy):
    """Mapper.filter_point_before_add (/root/reference/src/Mapper.py:137-163, CPU numpy per mapped frame) on the
    device: True for rays whose surface point does NOT project into the previous frame (edge 0) -- the candidates
    for new neural points."""
    dev = rays_o.device
    with torch.no_grad():
        pts = (rays_o + rays_d * gt_depth.reshape(-1, 1)).to(torch.float32)
        w2c = _w2c(prev_c2w, dev, torch.float32)
        cam = pts @ w2c[:3, :3].t() + w2c[:3, 3]
        cam = cam.double()
        z = cam[:, 2] + 1e-5
        u = ((fx * (-cam[:, 0]) + cx * cam[:, 2]) / z).float()      # cam_cord[:, 0] *= -1 (:154)
        v = ((fy * cam[:, 1] + cy * cam[:, 2]) / z).float()
        mask = (u < W) & (u > 0) & (v < H) & (v > 0)
    return ~mask


def keyframe_overlap_percent(vertices, keyframe_c2ws, H, W, fx, fy, cx, cy, edge=20):
    """The deterministic core of Mapper.keyframe_selection_overlap (/root/reference/src/Mapper.py:252-274): the
    fraction of `vertices` (P,3) that project inside each keyframe (edge-cropped, in front of the camera) -- all
    keyframes in one batched projection instead of a Python loop over numpy matmuls.  -> (n_keyframes,) float64"""
    dev = vertices.device
    if len(keyframe_c2ws) == 0:
        return torch.zeros(0, dtype=torch.float64, device=dev)
    with torch.no_grad():
        w2c = torch.stack([_w2c(c, dev, torch.float32) for c in keyframe_c2ws])          # (F,4,4)
        cam = torch.einsum('fij,pj->fpi', w2c[:, :3, :3], vertices.float()) + w2c[:, None, :3, 3]
        cam = cam.double()
        z = cam[..., 2] + 1e-5
        u = ((fx * cam[..., 0] + cx * cam[..., 2]) / z).float()       # no x flip here (:261 is commented out)
        v = ((fy * cam[..., 1] + cy * cam[..., 2]) / z).float()
        mask = (u < W - edge) & (u > edge) & (v < H - edge) & (v > edge) & (z < 0)
        return mask.double().mean(dim=1)


def keyframe_selection_overlap(gt_color, gt_depth, c2w, keyframe_dict, k, H, W, fx, fy, cx, cy, device,
                               N_samples=8, pixels=200):
    """Mapper.keyframe_selection_overlap (/root/reference/src/Mapper.py:219-282): sample `pixels` rays of the current
    frame, `N_samples` points around the sensor depth on each, rank the keyframes by the fraction of points they see
    and return a random subset of k of those that see any.  Same RNG use as the reference (get_samples' randint on
    the device, numpy permutation on the host)."""
    import numpy as np
    from .common import get_samples
    rays_o, rays_d, depth, _ = get_samples(0, H, 0, W, pixels, H, W, fx, fy, cx, cy, c2w, gt_depth, gt_color, device,
                                           depth_filter=True)
    depth = depth.reshape(-1, 1).repeat(1, N_samples)
    t_vals = torch.linspace(0., 1., steps=N_samples, device=depth.device)
    z_vals = depth * 0.8 * (1. - t_vals) + (depth + 0.5) * t_vals
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]).reshape(-1, 3)
    pct = keyframe_overlap_percent(pts, [kf['est_c2w'] for kf in keyframe_dict], H, W, fx, fy, cx, cy).cpu().numpy()
    order = sorted(range(len(pct)), key=lambda i: pct[i], reverse=True)
    selected = [i for i in order if pct[i] > 0.00]
    return list(np.random.permutation(np.array(selected))[:k])
