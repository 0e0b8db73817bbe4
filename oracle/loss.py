"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's inline loss
expressions, dtype-generic (run in float64 for the gradient oracle).

  mapper_loss   src/Mapper.py:689-693, 713-720 (the encode_exposure branch transforms `color` before the
                loss; that transform belongs to the renderer's rgb_mode, not to the loss)
  tracker_loss  src/Tracker.py:171-191
"""
import torch


def mapper_loss(depth, color, valid, gt_depth, gt_color, stage, w_color):
    depth_mask = (gt_depth > 0) & valid
    depth_mask = depth_mask & (~torch.isnan(depth))
    geo = torch.abs(gt_depth[depth_mask] - depth[depth_mask]).sum()
    loss = geo.clone()
    col = torch.zeros((), dtype=depth.dtype)
    if stage == 'color':
        col = torch.abs(gt_color[depth_mask] - color[depth_mask]).sum()
        loss = loss + w_color * col
    return loss, geo, col


def tracker_loss(depth, uncertainty, color, gt_depth, gt_color, handle_dynamic, use_color, w_color):
    uncertainty = uncertainty.detach()
    nan_mask = (~torch.isnan(depth)) & (~torch.isnan(uncertainty))
    if handle_dynamic:
        tmp = torch.abs(gt_depth - depth) / torch.sqrt(uncertainty + 1e-10)
        mask = (tmp < 10 * tmp.mean()) & (gt_depth > 0)
    else:
        tmp = torch.abs(gt_depth - depth)
        mask = (tmp < 10 * tmp.median()) & (gt_depth > 0)
    mask = mask & nan_mask
    geo = torch.clamp(torch.abs(gt_depth - depth) / torch.sqrt(uncertainty + 1e-10), min=0.0, max=1e3)[mask].sum()
    loss = geo
    col = torch.abs(gt_color - color)[mask].sum()
    if use_color:
        loss = loss + w_color * col
    return loss, geo, col, mask
