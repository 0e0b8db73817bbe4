"""Per-frame dynamic add/query radius maps on the device.

The reference computes them on the CPU per frame with skimage + scipy
(/root/reference/src/Tracker.py:243-258, /root/reference/src/Mapper.py:854-869): rgb2gray, Sobel
magnitude, clip to [0, color_grad_threshold], interp1d([0, 0.01, thr] -> [r_max, r_max, r_min]) and
uploads two float64 (H,W) maps.  Here: one fused CUDA kernel, float64 arithmetic, no host round trip.
"""
import torch

from ._lib import lib, check, ptr, stream_ptr, require_cuda


def dynamic_radius_maps(gt_color, cfg):
    """gt_color (H,W,3) float32/float64 CUDA tensor -> (dynamic_r_add, dynamic_r_query), each (H,W) float64."""
    require_cuda(gt_color, 'gt_color')
    pc = cfg['pointcloud']
    H, W = gt_color.shape[0], gt_color.shape[1]
    c = gt_color.contiguous()
    r_add = torch.empty(H, W, dtype=torch.float64, device=c.device)
    r_query = torch.empty(H, W, dtype=torch.float64, device=c.device)
    f32 = c if c.dtype == torch.float32 else None
    f64 = c if c.dtype == torch.float64 else None
    if f32 is None and f64 is None:
        f32 = c.float()
    check(lib().lsr_dynamic_radius(ptr(f32), ptr(f64), H, W, float(pc['color_grad_threshold']),
                                   float(pc['radius_add_max']), float(pc['radius_add_min']),
                                   float(pc['radius_query_ratio']), ptr(r_add), ptr(r_query),
                                   stream_ptr(c.device)), 'lsr_dynamic_radius')
    return r_add, r_query
