"""Host-side mirror of the hot-path functions of /root/reference/src/common.py (same names,
arguments and return values), backed by the lsr CUDA kernels:

    get_samples (:237-259) -> get_sample_uv (:160-172) -> select_uv (:123-138) -> get_rays_from_uv (:104-120)
    get_rays (:425-442), quad2rotation (:301-324), get_camera_from_tensor (:327-343),
    raw2outputs_nerf_color (:382-422)

The reference rebuilds an H x W meshgrid and launches ~10 tiny kernels per get_samples call; here
it is torch.randint (kept so the pixel sequence follows the same global generator, Appendix D of
SURVEY.md) plus ONE fused kernel, with a hand-written backward to the camera matrix.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr


class _SampleRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c2w, depth_img, color_img, pix, geom):
        H, W, fx, fy, cx, cy, H0, H1, W0, W1 = geom
        dev = pix.device
        n = pix.shape[0]
        c2w_f = c2w.to(torch.float32)
        if not c2w_f.is_contiguous():
            c2w_f = c2w_f.contiguous()
        rays_o = torch.empty(n, 3, dtype=torch.float32, device=dev)
        rays_d = torch.empty(n, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        color = torch.empty(n, 3, dtype=torch.float32, device=dev) if color_img is not None else None
        i = torch.empty(n, dtype=torch.int64, device=dev)
        j = torch.empty(n, dtype=torch.int64, device=dev)
        if c2w_f.device != dev:
            c2w_f = c2w_f.to(dev)
        with _lib.on_device(dev):
            check(lib().lsr_sample_rays(ptr(depth_img), ptr(color_img), H, W, fx, fy, cx, cy, ptr(c2w_f),
                                        c2w_f.shape[-1], ptr(pix), n, H0, H1, W0, W1, ptr(rays_o), ptr(rays_d),
                                        ptr(depth), ptr(color), ptr(i), ptr(j), stream_ptr(dev)), 'lsr_sample_rays')
        ctx.geom = geom
        ctx.c2w_shape = c2w.shape
        ctx.c2w_dtype = c2w.dtype
        ctx.save_for_backward(i, j)
        ctx.mark_non_differentiable(depth, i, j)
        if color is not None:
            ctx.mark_non_differentiable(color)
            return rays_o, rays_d, depth, color, i, j
        return rays_o, rays_d, depth, i, j

    @staticmethod
    def backward(ctx, g_o, g_d, *_):
        i, j = ctx.saved_tensors
        H, W, fx, fy, cx, cy, H0, H1, W0, W1 = ctx.geom
        dev = i.device
        d12 = torch.empty(12, dtype=torch.float32, device=dev)
        g_o = g_o.contiguous().float() if g_o is not None else None
        g_d = g_d.contiguous().float() if g_d is not None else None
        with _lib.on_device(dev):
            check(lib().lsr_sample_rays_bwd(ptr(g_o), ptr(g_d), ptr(i), ptr(j), i.shape[0], fx, fy, cx, cy, ptr(d12),
                                            stream_ptr(dev)), 'lsr_sample_rays_bwd')
        g = torch.zeros(ctx.c2w_shape, dtype=torch.float32, device=dev)
        g[:3, :4] = d12.view(3, 4)
        return g.to(ctx.c2w_dtype), None, None, None, None


def _sample_rays_filtered(c2w_f, depth_img, color_img, pix, geom, depth_limit):
    """lsr_sample_rays_filtered_sync: one allocation for all outputs, one launch; the number of kept pixels (the size of the
    returned tensors) comes back through a host-mapped word the C side spins on -- no D2H copy, no stream synchronise."""
    H, W, fx, fy, cx, cy, H0, H1, W0, W1 = geom
    dev = pix.device
    n = pix.shape[0]
    # [rays_o 3n | rays_d 3n | depth n | colour 3n] f32 and [i n | j n] i64 in two buffers
    fbuf = torch.empty(10 * n, dtype=torch.float32, device=dev)
    ibuf = torch.empty(2 * n, dtype=torch.int64, device=dev)
    base_f, base_i = fbuf.data_ptr(), ibuf.data_ptr()
    cnt = ctypes.c_int32(0)
    check(lib().lsr_sample_rays_filtered_sync(depth_img.data_ptr(), color_img.data_ptr(), H, W, fx, fy, cx, cy, c2w_f.data_ptr(),
                                              c2w_f.shape[-1], pix.data_ptr(), n, H0, H1, W0, W1,
                                              float(depth_limit) if depth_limit is not None else 0.0, base_f, base_f + 12 * n,
                                              base_f + 24 * n, base_f + 28 * n, base_i, base_i + 8 * n, ctypes.byref(cnt),
                                              stream_ptr(dev)), 'lsr_sample_rays_filtered_sync')
    m = cnt.value
    return (fbuf[:3 * m].view(m, 3), fbuf[3 * n:3 * n + 3 * m].view(m, 3), fbuf[6 * n:6 * n + m],
            fbuf[7 * n:7 * n + 3 * m].view(m, 3), ibuf[:m], ibuf[n:n + m])


class _SampleRaysFilteredFn(torch.autograd.Function):
    """get_samples with depth_filter=True in one launch, differentiable w.r.t. the camera matrix."""

    @staticmethod
    def forward(ctx, c2w, depth_img, color_img, pix, geom, depth_limit):
        c2w_f = c2w.to(torch.float32)
        if not c2w_f.is_contiguous():
            c2w_f = c2w_f.contiguous()
        with _lib.on_device(pix.device):
            rays_o, rays_d, depth, color, i, j = _sample_rays_filtered(c2w_f, depth_img, color_img, pix, geom, depth_limit)
        ctx.geom = geom
        ctx.c2w_shape = c2w.shape
        ctx.c2w_dtype = c2w.dtype
        ctx.save_for_backward(i, j)
        ctx.mark_non_differentiable(depth, color, i, j)
        return rays_o, rays_d, depth, color, i, j

    @staticmethod
    def backward(ctx, g_o, g_d, *_):
        i, j = ctx.saved_tensors
        H, W, fx, fy, cx, cy, H0, H1, W0, W1 = ctx.geom
        dev = i.device
        d12 = torch.empty(12, dtype=torch.float32, device=dev)
        g_o = g_o.contiguous().float() if g_o is not None else None
        g_d = g_d.contiguous().float() if g_d is not None else None
        with _lib.on_device(dev):
            check(lib().lsr_sample_rays_bwd(ptr(g_o), ptr(g_d), ptr(i), ptr(j), i.shape[0], fx, fy, cx, cy, ptr(d12),
                                            stream_ptr(dev)), 'lsr_sample_rays_bwd')
        g = torch.zeros(ctx.c2w_shape, dtype=torch.float32, device=dev)
        g[:3, :4] = d12.view(3, 4)
        return g.to(ctx.c2w_dtype), None, None, None, None, None


def _as_c2w_tensor(c2w, device):
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w).to(device)
    return c2w


def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, device, depth_filter=False,
                return_index=False, depth_limit=None):
    """n rays from the image window [H0,H1) x [W0,W1) (with replacement), as the reference."""
    c2w = _as_c2w_tensor(c2w, device)
    _lib.require_cuda(depth, 'depth')
    win = (H1 - H0) * (W1 - W0)
    pix = torch.randint(win, (n,), device=device)                       # common.py:130
    depth_f = depth if (depth.dtype == torch.float32 and depth.is_contiguous()) else depth.float().contiguous()
    fused_color = color.dtype == torch.float32 and color.is_contiguous()
    geom = (int(H), int(W), float(fx), float(fy), float(cx), float(cy), int(H0), int(H1), int(W0), int(W1))
    if not c2w.is_cuda:
        c2w = c2w.to(device)
    if depth_filter and fused_color and depth.dtype == torch.float32:
        # the common case of both callers (src/Mapper.py:652-655, src/Tracker.py:138-141): sampling, depth filter and
        # order-preserving compaction in one kernel, one host sync for the output size
        if c2w.requires_grad and torch.is_grad_enabled():
            out = _SampleRaysFilteredFn.apply(c2w, depth_f, color, pix, geom, depth_limit)
        else:   # mapper without BA: constant poses, no autograd node
            c2w_f = c2w.detach()
            if c2w_f.dtype != torch.float32 or not c2w_f.is_contiguous():
                c2w_f = c2w_f.to(torch.float32).contiguous()
            with _lib.on_device(pix.device):
                out = _sample_rays_filtered(c2w_f, depth_f, color, pix, geom, depth_limit)
        return out if return_index else out[:4]
    out = _SampleRaysFn.apply(c2w, depth_f, color if fused_color else None, pix, geom)
    if fused_color:
        rays_o, rays_d, sample_depth, sample_color, i, j = out
    else:   # e.g. float64 colour images (datasets.py:101): gather with torch to keep the dtype
        rays_o, rays_d, sample_depth, i, j = out
        sample_color = color[j, i]
    if depth.dtype != torch.float32:
        sample_depth = depth[j, i]
    if depth_filter:                                                     # common.py:249-255
        mask = sample_depth > 0
        if depth_limit is not None:
            mask = mask & (sample_depth < depth_limit)
        keep = torch.nonzero(mask, as_tuple=True)[0]      # ONE host sync for the output size instead of one per masked gather
        rays_o, rays_d = rays_o.index_select(0, keep), rays_d.index_select(0, keep)
        sample_depth, sample_color = sample_depth.index_select(0, keep), sample_color.index_select(0, keep)
        i, j = i.index_select(0, keep), j.index_select(0, keep)
    if return_index:
        return rays_o, rays_d, sample_depth, sample_color, i, j
    return rays_o, rays_d, sample_depth, sample_color


def get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device):
    """Rays for given pixel coordinates (i = column, j = row; float or int tensors)."""
    c2w = _as_c2w_tensor(c2w, device)
    ii, jj = i.to(torch.int64).reshape(-1), j.to(torch.int64).reshape(-1)
    pix = jj * int(W) + ii
    geom = (int(H), int(W), float(fx), float(fy), float(cx), float(cy), 0, int(H), 0, int(W))
    rays_o, rays_d, _, _, _ = _SampleRaysFn.apply(c2w, None, None, pix, geom)
    return rays_o, rays_d


def get_rays(H, W, fx, fy, cx, cy, c2w, device, crop_edge=0):
    """Rays of a whole image, (H-2c, W-2c, 3) each (common.py:425-442)."""
    c2w = _as_c2w_tensor(c2w, device)
    if not c2w.is_cuda:
        c2w = c2w.to(device)
    hh, ww = H - 2 * crop_edge, W - 2 * crop_edge
    pix = torch.arange(hh * ww, device=c2w.device, dtype=torch.int64)
    geom = (int(H), int(W), float(fx), float(fy), float(cx), float(cy), crop_edge, H - crop_edge, crop_edge,
            W - crop_edge)
    rays_o, rays_d, _, _, _ = _SampleRaysFn.apply(c2w, None, None, pix, geom)
    return rays_o.reshape(hh, ww, 3), rays_d.reshape(hh, ww, 3)


class _PoseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam):
        cam_f = cam.detach().to(torch.float32).contiguous()
        out = torch.empty(3, 4, dtype=torch.float32, device=cam.device)
        with _lib.on_device(cam.device):
            check(lib().lsr_pose_fwd(ptr(cam_f), ptr(out), stream_ptr(cam.device)), 'lsr_pose_fwd')
        ctx.save_for_backward(cam_f)
        ctx.dtype = cam.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        cam_f, = ctx.saved_tensors
        d = torch.empty(7, dtype=torch.float32, device=cam_f.device)
        with _lib.on_device(cam_f.device):
            check(lib().lsr_pose_bwd(ptr(cam_f), ptr(g.contiguous().float()), ptr(d), stream_ptr(cam_f.device)),
                  'lsr_pose_bwd')
        return d.to(ctx.dtype)


def quad2rotation(quad):
    """Batch quaternion (w,x,y,z; unnormalised) -> rotation, differentiable (common.py:301-324).
    Torch restatement for batched input (not on the per-iteration path)."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [
        torch.stack([1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr)], -1),
        torch.stack([two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr)], -1),
        torch.stack([two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)], -1)]
    return torch.stack(rows, 1)


def get_camera_from_tensor(inputs):
    """[quat(4) | T(3)] -> 3x4 c2w (common.py:327-343).  The single-pose case (every call on the hot
    path: src/Tracker.py:122, src/Mapper.py:633,643) is one fused kernel forward + one backward."""
    if inputs.dim() == 1 and inputs.is_cuda:
        return _PoseFn.apply(inputs)
    squeeze = inputs.dim() == 1
    x = inputs.unsqueeze(0) if squeeze else inputs
    RT = torch.cat([quad2rotation(x[:, :4]), x[:, 4:, None]], 2)
    return RT[0] if squeeze else RT


def raw2outputs_nerf_color(raw, z_vals, rays_d, device='cuda:0', coef=0.1):
    """Alpha compositing of already-decoded samples (common.py:382-422).  Kept for API completeness
    (eval_points users); render_batch_ray composites inside the fused kernel."""
    rgb = raw[..., :-1]
    alpha = torch.sigmoid(coef * raw[..., -1])
    ones = torch.ones((alpha.shape[0], 1), device=alpha.device, dtype=alpha.dtype)
    weights = alpha * torch.cumprod(torch.cat([ones, (1. - alpha + 1e-10)], -1), -1)[:, :-1]
    wsum = weights.sum(-1, keepdim=True) + 1e-10
    rgb_map = (weights[..., None] * rgb).sum(-2) / wsum
    depth_map = (weights * z_vals).sum(-1) / wsum.squeeze(-1)
    tmp = z_vals - depth_map.unsqueeze(-1)
    depth_var = (weights * tmp * tmp).sum(1)
    return depth_map, depth_var, rgb_map, weights
