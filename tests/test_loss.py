"""Fused mapper / tracker losses: oracle restatement sanity on CPU, CUDA kernels vs
the oracle in fp64 on the GPU (values, gradients, masks; NaN / zero-depth / invalid rays)."""
import pytest
import torch

import loopy_slam_b200 as L
from oracle import loss as oloss


def _batch(R, seed, nan=True):
    g = torch.Generator().manual_seed(seed)
    gt_d = torch.rand(R, generator=g) * 3 + 0.3
    gt_d[torch.rand(R, generator=g) < 0.1] = 0.0                       # missing depth
    depth = gt_d + 0.2 * torch.randn(R, generator=g)
    var = torch.rand(R, generator=g) * 0.05 + 1e-4
    color = torch.rand(R, 3, generator=g)
    gt_c = torch.rand(R, 3, generator=g)
    valid = torch.rand(R, generator=g) > 0.15
    if nan and R > 8:
        depth[3] = float('nan')
        var[5] = float('nan')
        depth[7] = gt_d[7]                                               # sign(0) = 0
        color[7, 1] = gt_c[7, 1]
    depth[R // 2:R // 2 + 3] += 40.0                                     # outliers for the tracker mask
    return depth, var, color, gt_d, gt_c, valid


def test_oracle_mapper_loss_equals_masked_sum_formulation():
    depth, var, color, gt_d, gt_c, valid = _batch(500, 1)
    for stage in ('geometry', 'color'):
        loss, geo, col = oloss.mapper_loss(depth.double(), color.double(), valid, gt_d.double(), gt_c.double(), stage, 0.1)
        m = (gt_d > 0) & valid & ~torch.isnan(depth)
        geo2 = torch.where(m, (gt_d.double() - depth.double()).abs(), torch.zeros((), dtype=torch.float64)).sum()
        col2 = torch.where(m[:, None], (gt_c.double() - color.double()).abs(), torch.zeros((), dtype=torch.float64)).sum()
        assert torch.allclose(geo, geo2, rtol=1e-12)
        assert torch.allclose(loss, geo2 + (0.1 * col2 if stage == 'color' else 0.0), rtol=1e-12)


def test_oracle_tracker_loss_masks_outliers_and_nans():
    depth, var, color, gt_d, gt_c, valid = _batch(400, 2)
    for hd in (True, False):
        d = depth.clone()
        # a NaN in tmp makes mean() / median() NaN and empties the mask (reference behaviour)
        loss, geo, col, mask = oloss.tracker_loss(d, var, color, gt_d, gt_c, hd, True, 0.5)
        assert not mask.any() and float(loss) == 0.0
        d[3] = gt_d[3]
        v = var.clone()
        if hd:
            v[5] = 0.01       # the static branch's tmp does not involve var: its NaN is caught by nan_mask alone
        loss, geo, col, mask = oloss.tracker_loss(d, v, color, gt_d, gt_c, hd, True, 0.5)
        assert mask.any() and not mask[200:203].any() and not mask[gt_d <= 0].any()
        assert torch.isfinite(loss)


@pytest.mark.gpu
@pytest.mark.parametrize('stage', ['geometry', 'color'])
@pytest.mark.parametrize('R', [1, 37, 4936, 70001])
def test_mapper_loss_matches_oracle(stage, R):
    dev = torch.device('cuda:0')
    depth, var, color, gt_d, gt_c, valid = _batch(R, 10 + R)
    d64 = depth.double().requires_grad_(True)
    c64 = color.double().requires_grad_(True)
    lo, geo_o, col_o = oloss.mapper_loss(d64, c64, valid, gt_d.double(), gt_c.double(), stage, 0.1)
    (3.0 * lo).backward()
    dd = depth.to(dev).requires_grad_(True)
    cc = color.to(dev).requires_grad_(True)
    loss, geo, col = L.mapper_loss(dd, cc, valid.to(dev), gt_d.to(dev), gt_c.to(dev), stage, 0.1)
    (3.0 * loss).backward()
    assert abs(float(loss) - float(lo)) <= 2e-6 * max(1.0, abs(float(lo)))       # fp32 rounding of an fp64 sum
    assert abs(float(geo) - float(geo_o)) <= 2e-6 * max(1.0, abs(float(geo_o)))
    if stage == 'color':
        assert abs(float(col) - float(col_o)) <= 2e-6 * max(1.0, abs(float(col_o)))
        assert torch.equal(cc.grad.cpu().double(), torch.nan_to_num(c64.grad).to(torch.float32).double())
    else:
        assert cc.grad is None or not cc.grad.any()
    assert torch.equal(dd.grad.cpu(), d64.grad.to(torch.float32))                 # +-3 / 0: exact


@pytest.mark.gpu
@pytest.mark.parametrize('handle_dynamic', [True, False])
@pytest.mark.parametrize('use_color', [True, False])
def test_tracker_loss_matches_oracle(handle_dynamic, use_color):
    dev = torch.device('cuda:0')
    R = 1500
    depth, var, color, gt_d, gt_c, valid = _batch(R, 77, nan=not handle_dynamic)
    if not handle_dynamic:
        var[5] = 0.02          # keep var finite where depth is finite; depth[3] stays NaN -> median path sees NaN
        depth[3] = gt_d[3] + 0.1
        var[9] = float('nan')  # masked by nan_mask only (tmp does not involve var here)
    d64 = depth.double().requires_grad_(True)
    c64 = color.double().requires_grad_(True)
    lo, geo_o, col_o, mask_o = oloss.tracker_loss(d64, var.double(), c64, gt_d.double(), gt_c.double(),
                                                  handle_dynamic, use_color, 0.5)
    lo.backward()
    dd = depth.to(dev).requires_grad_(True)
    cc = color.to(dev).requires_grad_(True)
    loss, geo, col, mask = L.tracker_loss(dd, var.to(dev), cc, gt_d.to(dev), gt_c.to(dev), handle_dynamic, use_color, 0.5)
    loss.backward()
    assert mask_o.any()
    assert torch.equal(mask.cpu(), mask_o)
    assert abs(float(loss) - float(lo)) <= 1e-5 * max(1.0, abs(float(lo)))
    assert abs(float(geo) - float(geo_o)) <= 1e-5 * max(1.0, abs(float(geo_o)))
    assert abs(float(col) - float(col_o)) <= 1e-5 * max(1.0, abs(float(col_o)))
    gd_o = torch.nan_to_num(d64.grad)
    assert torch.allclose(dd.grad.cpu().double(), gd_o, rtol=2e-6, atol=0)
    if use_color:
        assert torch.equal(cc.grad.cpu().double(), c64.grad)
    else:
        assert cc.grad is None or not cc.grad.any()


@pytest.mark.gpu
def test_losses_drive_the_render_backward_like_the_inline_expressions():
    """mapper_loss(...) and the reference's inline expression give the same feature gradients through
    render_batch_ray's backward."""
    from helpers import Golden
    from parity import run_cuda
    g = Golden('replica_color_mapper')

    def inline(depth, var, color, valid, gt_d, gt_c):
        m = (gt_d > 0) & valid & ~torch.isnan(depth)
        return torch.abs(gt_d[m] - depth[m]).sum() + 0.1 * torch.abs(gt_c[m] - color[m]).sum()

    def fused(depth, var, color, valid, gt_d, gt_c):
        return L.mapper_loss(depth, color, valid, gt_d, gt_c, 'color', 0.1)[0]

    a = run_cuda(g, loss_fn=inline)
    b = run_cuda(g, loss_fn=fused)
    assert abs(a['loss'] - b['loss']) <= 1e-5 * abs(a['loss'])
    for k in a['grads']:
        assert torch.allclose(a['grads'][k], b['grads'][k], rtol=1e-4, atol=1e-6 * float(a['grads'][k].abs().max())), k
