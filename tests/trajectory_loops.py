"""The caller-side optimisation loops of the reference, restated ONCE and run with two different sets of modules:

  * tests/golden/make_golden_trajectory.py runs them with the REAL reference modules (src.common, src.utils.Renderer,
    NICER from src.conv_onet) on CPU and stores the trajectories;
  * tests/test_gpu_trajectory.py replays them with loopy_slam_b200 on the GPU.

    tracker_loop : src/Tracker.py:361-368 (optimiser over the 7-vector camera tensor) + optimize_cam_in_batch :102-197
    mapper_loop  : src/Mapper.py:498-541 (leaf feature blocks, parameter groups), :576-735 (index_put, per-frame
                   get_samples, inside mask, render, loss incl. the per-frame exposure slices :697-715, Adam step, write-back)

`mods` supplies: get_samples, get_camera_from_tensor, renderer (render_batch_ray), decoders, npc.  The pixel picks
(torch.randint inside get_samples) are recorded when minting and injected when replaying, because the CPU and CUDA
generators differ.
"""
import torch


class Picks:
    """Record (mint) or replay (test) the results of torch.randint in call order."""

    def __init__(self, stored=None, device=None):
        self.stored, self.device, self.log, self.k = stored, device, [], 0
        self._orig = torch.randint

    def __enter__(self):
        def randint(high, size, **kw):
            if self.stored is None:
                kw.pop('device', None)
                v = self._orig(high, size)
                self.log.append(v.clone())
                return v
            v = self.stored[self.k]
            self.k += 1
            assert v.shape == tuple(size) and int(v.max()) < high
            return v.to(self.device)
        torch.randint = randint
        return self

    def __exit__(self, *a):
        torch.randint = self._orig


def inside_mask_of(batch_gt_depth):
    with torch.no_grad():
        return batch_gt_depth <= torch.minimum(10 * batch_gt_depth.median(), 1.2 * torch.max(batch_gt_depth))


def tracker_loop(mods, cam0, gt_color, gt_depth, intr, cloud, geo, col, n_iters, pixels, edge, cam_lr, w_color,
                 dynamic_r_map=None, exposure_feat=None, device='cpu'):
    """-> losses (list of float), final camera tensor, final exposure_feat (or None)."""
    H, W, fx, fy, cx, cy = intr
    camera_tensor = cam0.clone().to(device).requires_grad_(True)
    groups = [{'params': [camera_tensor], 'lr': cam_lr}]
    if exposure_feat is not None:
        exposure_feat = exposure_feat.clone().to(device).requires_grad_(True)
        groups.append({'params': [exposure_feat], 'lr': 0.001})
        groups.append({'params': list(mods.decoders.color_decoder.mlp_exposure.parameters()), 'lr': 0.001})
    optimizer = torch.optim.Adam(groups)
    losses = []
    for _ in range(n_iters):
        optimizer.zero_grad()
        c2w = mods.get_camera_from_tensor(camera_tensor)
        o, d, g, c, i, j = mods.get_samples(edge, H - edge, edge, W - edge, pixels, H, W, fx, fy, cx, cy, c2w, gt_depth, gt_color,
                                            device, depth_filter=True, return_index=True, depth_limit=None)
        rq = dynamic_r_map[j, i] if dynamic_r_map is not None else None
        inside = inside_mask_of(g)
        d, o, g, c = d[inside], o[inside], g[inside], c[inside]
        rq = rq[inside] if rq is not None else None
        depth, uncertainty, color, _ = mods.renderer.render_batch_ray(
            mods.npc, mods.decoders, d, o, device, stage='color', gt_depth=g, npc_geo_feats=geo, npc_col_feats=col,
            is_tracker=True, cloud_pos=cloud, dynamic_r_query=rq, exposure_feat=exposure_feat)
        uncertainty = uncertainty.detach()
        nan_mask = (~torch.isnan(depth)) & (~torch.isnan(uncertainty))
        tmp = torch.abs(g - depth) / torch.sqrt(uncertainty + 1e-10)          # handle_dynamic: True (point_slam.yaml:45)
        mask = (tmp < 10 * tmp.mean()) & (g > 0)
        mask = mask & nan_mask
        geo_loss = torch.clamp(torch.abs(g - depth) / torch.sqrt(uncertainty + 1e-10), min=0.0, max=1e3)[mask].sum()
        loss = geo_loss
        color_loss = torch.abs(c - color)[mask].sum()
        loss = loss + w_color * color_loss
        loss.backward()
        optimizer.step()
        optimizer.zero_grad()
        losses.append(float(loss.item()))
    return losses, camera_tensor.detach().cpu(), None if exposure_feat is None else exposure_feat.detach().cpu()


def mapper_loop(mods, frames, intr, cloud, geo_table, col_table, indices, n_iters, geo_iters, pixels, lrs, w_color,
                dynamic_r_maps=None, exposure_feats=None, device='cpu'):
    """frames: list of (color, depth, c2w).  lrs: {'geometry': (decoders_lr, geo_lr, col_lr), 'color': (...)}.
    -> losses, final (geo_leaf, col_leaf), exposure feats."""
    H, W, fx, fy, cx, cy = intr
    n_frames = len(frames)
    pix_per_image = pixels // n_frames
    npc_geo_feats, npc_col_feats = geo_table.clone().to(device), col_table.clone().to(device)
    geo_pcl_grad = npc_geo_feats[indices].clone().detach().requires_grad_(True)          # Mapper.py:502-505
    color_pcl_grad = npc_col_feats[indices].clone().detach().requires_grad_(True)
    dec = mods.decoders
    decoders_para_list = list(dec.color_decoder.parameters())                            # fix_color_decoder: False
    decoders_para_list += list(dec.geo_decoder.embedder.parameters())                    # fix_geo_decoder: True (:537-541)
    decoders_para_list += list(dec.geo_decoder.embedder_rel_pos.parameters())
    groups = [{'params': decoders_para_list, 'lr': 0}, {'params': [geo_pcl_grad], 'lr': 0}, {'params': [color_pcl_grad], 'lr': 0}]
    if exposure_feats is not None:
        exposure_feats = [e.clone().to(device).requires_grad_(True) for e in exposure_feats]
        groups.append({'params': exposure_feats, 'lr': 0.001})
    optimizer = torch.optim.Adam(groups)
    losses = []
    for joint_iter in range(n_iters):
        npc_geo_feats[indices] = geo_pcl_grad                                            # :581-582
        npc_col_feats[indices] = color_pcl_grad
        stage = 'geometry' if joint_iter <= geo_iters else 'color'                       # :588-591
        for gi in range(3):
            optimizer.param_groups[gi]['lr'] = lrs[stage][gi]
        optimizer.zero_grad()
        O, D, G, C, RQ, FI = [], [], [], [], [], []
        for f, (gt_color, gt_depth, c2w) in enumerate(frames):
            o, d, g, c, i, j = mods.get_samples(0, H, 0, W, pix_per_image, H, W, fx, fy, cx, cy, c2w, gt_depth, gt_color, device,
                                                depth_filter=True, return_index=True)
            O.append(o.float()); D.append(d.float()); G.append(g.float()); C.append(c.float())
            if dynamic_r_maps is not None:
                RQ.append(dynamic_r_maps[f][j, i])
            if exposure_feats is not None:
                FI.append(torch.full((i.shape[0],), f, dtype=torch.long, device=device))
        d, o, g, c = torch.cat(D), torch.cat(O), torch.cat(G), torch.cat(C)
        rq = torch.cat(RQ) if dynamic_r_maps is not None else None
        inside = inside_mask_of(g)
        d, o, g, c = d[inside], o[inside], g[inside], c[inside]
        if rq is not None:
            rq = rq[inside]
        depth, uncertainty, color, valid = mods.renderer.render_batch_ray(
            mods.npc, dec, d, o, device, stage, gt_depth=g, npc_geo_feats=npc_geo_feats, npc_col_feats=npc_col_feats,
            is_tracker=False, cloud_pos=cloud, dynamic_r_query=rq, exposure_feat=None)
        depth_mask = (g > 0) & valid
        depth_mask = depth_mask & (~torch.isnan(depth))
        geo_loss = torch.abs(g[depth_mask] - depth[depth_mask]).sum()
        loss = geo_loss.clone()
        if stage == 'color':
            if exposure_feats is not None:                                               # :697-715
                indices_tensor = torch.cat(FI, dim=0)[inside]
                start_end = []
                for fid in torch.unique_consecutive(indices_tensor, return_counts=False):
                    match = torch.where(indices_tensor == fid)[0]
                    start_end.append((match[0].item(), match[-1].item() + 1))
                color = color.clone()
                for fidx, ef in enumerate(exposure_feats):
                    start, end = start_end[fidx]
                    aff = dec.color_decoder.mlp_exposure(ef)
                    rot, trans = aff[:9].reshape(3, 3), aff[-3:]
                    color_slice = color[start:end].clone()
                    color_slice = torch.matmul(color_slice, rot) + trans
                    color[start:end] = color_slice
                color = torch.sigmoid(color)
            color_loss = torch.abs(c[depth_mask] - color[depth_mask]).sum()
            loss = loss + w_color * color_loss
        loss.backward(retain_graph=False)
        optimizer.step()
        optimizer.zero_grad()
        npc_geo_feats, npc_col_feats = npc_geo_feats.detach(), npc_col_feats.detach()    # :727-735
        npc_geo_feats[indices], npc_col_feats[indices] = geo_pcl_grad.clone().detach(), color_pcl_grad.clone().detach()
        losses.append(float(loss.item()))
    return losses, geo_pcl_grad.detach().cpu(), color_pcl_grad.detach().cpu(), \
        None if exposure_feats is None else [e.detach().cpu() for e in exposure_feats]
