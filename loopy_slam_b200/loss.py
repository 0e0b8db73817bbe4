"""Fused mapper / tracker losses behind the expressions the reference writes
inline: src/Mapper.py:689-693,713-720 and src/Tracker.py:171-191.

    loss, geo_loss, color_loss = mapper_loss(depth, color, valid_ray_mask, batch_gt_depth, batch_gt_color,
                                             stage, w_color_loss)
    loss, geo_loss, color_loss, mask = tracker_loss(depth, uncertainty, color, batch_gt_depth, batch_gt_color,
                                                    handle_dynamic, use_color_in_tracking, w_color_loss)

One lsr kernel computes the masked L1 sums AND dloss/d(depth, color); autograd's backward only scales
those by grad_output.  No CPU path: CUDA tensors only.
"""
import ctypes

import torch

from . import _lib
from ._lib import LSR_STAGE, check, lib, ptr, require_cuda, stream_ptr

_SCRATCH_BYTES = None


def _scratch(dev):
    global _SCRATCH_BYTES
    if _SCRATCH_BYTES is None:
        n = ctypes.c_size_t()
        check(lib().lsr_loss_scratch_bytes(ctypes.byref(n)), 'lsr_loss_scratch_bytes')
        _SCRATCH_BYTES = int(n.value)
    return torch.empty(_SCRATCH_BYTES, dtype=torch.uint8, device=dev)


def _f32(t, name):
    require_cuda(t, name)
    return t.detach().to(torch.float32).contiguous()


class _MapperLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, color, valid, gt_depth, gt_color, stage, w_color):
        dev = depth.device
        R = depth.shape[0]
        use_color = stage == 'color'
        d = _f32(depth, 'depth')
        g = _f32(gt_depth, 'gt_depth')
        c = _f32(color, 'color') if use_color else None
        gc = _f32(gt_color, 'gt_color') if use_color else None
        v = None if valid is None else valid.detach().to(torch.uint8).contiguous() if valid.dtype != torch.bool \
            else valid.detach().contiguous().view(torch.uint8)
        loss3 = torch.empty(3, dtype=torch.float32, device=dev)
        d_depth = torch.empty(R, dtype=torch.float32, device=dev)
        d_color = torch.empty(R, 3, dtype=torch.float32, device=dev) if use_color else None
        scratch = _scratch(dev)
        with _lib.on_device(dev):
            check(lib().lsr_mapper_loss(ptr(d), ptr(c), ptr(v), ptr(g), ptr(gc), R, LSR_STAGE[stage], float(w_color),
                                        ptr(scratch), ptr(loss3), ptr(d_depth), ptr(d_color), stream_ptr(dev)),
                  'lsr_mapper_loss')
        ctx.save_for_backward(d_depth, d_color)
        ctx.mark_non_differentiable(loss3)
        return loss3[0], loss3

    @staticmethod
    def backward(ctx, g_loss, _g3):
        d_depth, d_color = ctx.saved_tensors
        if g_loss is None:
            return (None,) * 7
        gd = d_depth * g_loss if ctx.needs_input_grad[0] else None
        gc = d_color * g_loss if (d_color is not None and ctx.needs_input_grad[1]) else None
        return gd, gc, None, None, None, None, None


def mapper_loss(depth, color, valid_ray_mask, gt_depth, gt_color, stage, w_color_loss):
    """src/Mapper.py:689-693 + 713-720.  Returns (loss, geo_loss, color_loss); loss carries the graph,
    the two terms are detached 0-dim tensors (what the reference logs)."""
    loss, l3 = _MapperLossFn.apply(depth, color, valid_ray_mask, gt_depth, gt_color, stage, w_color_loss)
    return loss, l3[1], l3[2]


class _TrackerLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, var, color, gt_depth, gt_color, handle_dynamic, use_color, w_color):
        dev = depth.device
        R = depth.shape[0]
        d, u, c = _f32(depth, 'depth'), _f32(var, 'uncertainty'), _f32(color, 'color')
        g, gc = _f32(gt_depth, 'gt_depth'), _f32(gt_color, 'gt_color')
        loss3 = torch.empty(3, dtype=torch.float32, device=dev)
        tmp = torch.empty(R, dtype=torch.float32, device=dev)
        d_depth = torch.empty(R, dtype=torch.float32, device=dev)
        d_color = torch.empty(R, 3, dtype=torch.float32, device=dev)
        mask = torch.empty(R, dtype=torch.uint8, device=dev)
        scratch = _scratch(dev)
        with _lib.on_device(dev):
            st = stream_ptr(dev)
            check(lib().lsr_tracker_resid(ptr(d), ptr(u), ptr(g), R, int(bool(handle_dynamic)), ptr(scratch), ptr(tmp),
                                          st), 'lsr_tracker_resid')
            # src/Tracker.py:179-180: the static-scene branch thresholds at 10 * (lower) median
            thr = None if handle_dynamic else (10 * tmp.median()).reshape(1).contiguous()
            check(lib().lsr_tracker_loss(ptr(d), ptr(u), ptr(c), ptr(g), ptr(gc), ptr(tmp), R, ptr(thr),
                                         int(bool(use_color)), float(w_color), ptr(scratch), ptr(loss3), ptr(d_depth),
                                         ptr(d_color), ptr(mask), st), 'lsr_tracker_loss')
        ctx.save_for_backward(d_depth, d_color)
        mask = mask.view(torch.bool)
        ctx.mark_non_differentiable(loss3, mask)
        return loss3[0], loss3, mask

    @staticmethod
    def backward(ctx, g_loss, _g3, _gm):
        d_depth, d_color = ctx.saved_tensors
        if g_loss is None:
            return (None,) * 8
        gd = d_depth * g_loss if ctx.needs_input_grad[0] else None
        gc = d_color * g_loss if ctx.needs_input_grad[2] else None
        return gd, None, gc, None, None, None, None, None


def tracker_loss(depth, uncertainty, color, gt_depth, gt_color, handle_dynamic, use_color_in_tracking,
                 w_color_loss):
    """src/Tracker.py:171-191 (uncertainty is detached there, so it gets no gradient here either).
    Returns (loss, geo_loss, color_loss, mask)."""
    loss, l3, mask = _TrackerLossFn.apply(depth, uncertainty, color, gt_depth, gt_color, handle_dynamic,
                                          use_color_in_tracking, w_color_loss)
    return loss, l3[1], l3[2], mask
